timeout 600 python -m pytest tests/test_gpu_logits_viterbi.py -x -q -m gpu --tb=short 2>&1 | tail -15
