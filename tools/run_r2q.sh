for p in torch-asg_b200 _ab/exp1 _ab/exp2; do echo "== $p"; ASG_PKG=$PWD/$p timeout 120 python tools/quick_bench.py 2>&1 | grep -E "backward|fwd\+bwd"; done
