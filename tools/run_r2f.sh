NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:fcc_scan_kernel -s 2 -c 1 -f -o gpurun_out/fcc_r2a python tools/quick_bench.py > gpurun_out/ncu_fcc_r2a.log 2>&1
tail -3 gpurun_out/ncu_fcc_r2a.log
