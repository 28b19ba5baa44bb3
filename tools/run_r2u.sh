timeout -k 10 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or toy or cfg2" --tb=short 2>&1 | tail -5
timeout -k 10 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout -k 10 120 python tools/quick_bench.py 2>&1 | tail -8
