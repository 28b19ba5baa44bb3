timeout 600 python -m pytest tests/test_gpu_range.py -q -m gpu --tb=line 2>&1 | tail -12
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_range.py 2>&1 | tail -3
timeout 120 python tools/quick_bench.py 2>&1 | tail -8
