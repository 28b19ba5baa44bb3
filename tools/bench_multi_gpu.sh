for n in 8 4; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 20 > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
tail -c 400 gpurun_out/bench_${n}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${n}gpu.json').read().strip().splitlines()[-1])
print($n, d['value'], d['ms_per_step'], d['host_issue_ms_per_step'], d['parity_check'], d['kernels'].get('peer_allreduce_kernel'), d['e2e']['value'], d['clocks'])
PY
done
PEER_TEST_DUMP_AFTER=100 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/multi/peer_exchange.py 2>&1 | grep PEER_OK
