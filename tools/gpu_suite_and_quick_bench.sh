timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 120 python tools/quick_bench.py 2>&1 | tail -8
