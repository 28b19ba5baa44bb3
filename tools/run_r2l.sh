timeout 200 python -m pytest tests/test_gpu_peer.py -q -m gpu --tb=short 2>&1 | tail -15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/multi/peer_exchange.py 2>&1 | tail -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 2500 gpurun_out/bench_2gpu.json | cut -c1-1200; tail -3 gpurun_out/bench_2gpu.err
