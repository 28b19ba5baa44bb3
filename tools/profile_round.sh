set -x
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:grad_kernel -s 4 -c 1 -f -o gpurun_out/grad_r1k python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g.log 2>&1
timeout 400 $NCU -k regex:fcc_scan_kernel -s 4 -c 1 -f -o gpurun_out/fcc_r1n python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
timeout 400 $NCU -k regex:fac_scan_kernel -s 4 -c 1 -f -o gpurun_out/fac_r1k python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
cat gpurun_out/bench_r1g.json
