for ex in peer nccl; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$((RANDOM%10)) bench.py --gpus 8 --steps 200 --warmup 20 --exchange $ex > gpurun_out/bench_8gpu_$ex.json 2> gpurun_out/bench_8gpu_$ex.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_8gpu_$ex.json').read().strip().splitlines()[-1])
print("$ex", d['value'], d['ms_per_step'], d['host_issue_ms_per_step'], d['parity_check']['ok'], d['kernels'].get('peer_allreduce_kernel'), d['config']['parallelism'][:90])
PY
done
