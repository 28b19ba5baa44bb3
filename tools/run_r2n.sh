NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:grad_kernel -s 4 -c 1 -f -o gpurun_out/grad_r2a python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api eager > gpurun_out/ncu_g.log 2>&1
tail -3 gpurun_out/ncu_g.log
