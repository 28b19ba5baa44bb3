set -x
R=${1:-r2}
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg3_$R.json 2> gpurun_out/bench_cfg3_$R.err
python bench.py --workload cfg4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_$R.json 2> gpurun_out/bench_cfg4_$R.err
NCU="ncu --set full --clock-control none --import-source on"
timeout 400 $NCU -k regex:grad_kernel -s 4 -c 1 -f -o gpurun_out/grad_$R python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api eager > gpurun_out/ncu_g.log 2>&1
timeout 400 $NCU -k regex:fcc_scan_kernel -s 4 -c 1 -f -o gpurun_out/fcc_$R python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api eager > gpurun_out/ncu_c.log 2>&1
timeout 400 $NCU -k regex:fac_scan_kernel -s 4 -c 1 -f -o gpurun_out/fac_$R python bench.py --steps 2 --warmup 3 --no-cpu-baseline --api eager > gpurun_out/ncu_a.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --api eager-ws > gpurun_out/ncu_l.log 2>&1
tail -c 600 gpurun_out/bench_$R.err
cat gpurun_out/bench_$R.json | cut -c1-1500
cat gpurun_out/bench_cfg3_$R.json | cut -c1-900
cat gpurun_out/bench_cfg4_$R.json | cut -c1-900
