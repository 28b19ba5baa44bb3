timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "large" --tb=short 2>&1 | tail -15
timeout 120 python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | cut -c1-400
