NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:fac_scan_kernel -s 4 -c 1 -f -o gpurun_out/fac_r2c python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api eager > gpurun_out/ncu_a.log 2>&1
tail -2 gpurun_out/ncu_a.log | cut -c1-200
