timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python tools/quick_bench.py 2>&1 | tail -8
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:fac_scan_kernel -s 2 -c 1 -f -o gpurun_out/fac_r2a python tools/quick_bench.py > gpurun_out/ncu_fac_r2a.log 2>&1
