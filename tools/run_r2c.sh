NG=$1
for api in step graph; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 --api $api > gpurun_out/r2c_${NG}gpu_$api.json 2> gpurun_out/r2c_${NG}gpu_$api.err
echo "rc=$?"; tail -2 gpurun_out/r2c_${NG}gpu_$api.err
python -c "
import json,sys; d=json.loads(open('gpurun_out/r2c_${NG}gpu_$api.json').read().strip().splitlines()[-1]); print('$api', 'value', round(d['value']), 'ms/step', round(d['ms_per_step'],3), 'host', round(d['host_issue_ms_per_step'],3), d['config']['parallelism'], d.get('parity_check'), {k:(round(v['ms_per_step'],3), round(v['host_issue_ms_per_step'],3)) for k,v in d['config']['other_routes'].items()}, d['e2e']['value'])"
done
