for i in 1 2; do for p in torch-asg_b200 _ab/exp3; do echo "== $p"; ASG_PKG=$PWD/$p timeout 120 python tools/quick_bench.py 2>&1 | grep -E "fcc_fwd|fac_fwd|forward\(|fwd\+bwd"; done; done
