"""Developer diagnostic (2+ GPUs, torchrun): host cost of graph replays with / without a captured NCCL all-reduce."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-asg_b200")]
import torch, torch.distributed as dist, torch_asg
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T, N, L, S = 1000, 64, 40, 200
torch.manual_seed(rank)
x = torch.randn(T, N, L, device=dev).log_softmax(-1).requires_grad_(True)
y = torch.randint(0, L, (N, S), device=dev)
il = torch.full((N,), T, dtype=torch.long, device=dev); tl = torch.full((N,), S, dtype=torch.long, device=dev)
crit = torch_asg.ASGLoss(L).to(dev)
def timeit(fn, n=40):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    h = time.perf_counter() - t0; torch.cuda.synchronize(); return 1e3 * h / n, 1e3 * (time.perf_counter() - t0) / n
res = {}
g1 = torch_asg.GraphedASGLoss(crit, group=None, global_batch=N * world)          # no collective in the graph
e1 = g1.prepare(x, y, il, tl, kinds=("step",))
res["step graph, no NCCL"] = timeit(lambda i: e1.step.replay())
g2 = torch_asg.GraphedASGLoss(crit, group=dist.group.WORLD, global_batch=N * world)
e2 = g2.prepare(x, y, il, tl, kinds=("step",))
res["step graph, NCCL captured (%s)" % g2.nccl_in_graph] = timeit(lambda i: e2.step.replay())
buf = torch.zeros(L * L + 2, device=dev)
res["eager all_reduce alone"] = timeit(lambda i: dist.all_reduce(buf))
def mixed(i):
    e1.step.replay(); dist.all_reduce(e1.ws.fused)
res["step graph + eager all_reduce"] = timeit(mixed)
ga = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    dist.all_reduce(buf)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
with torch.cuda.graph(ga, stream=s, capture_error_mode="thread_local"):
    dist.all_reduce(buf)
res["all_reduce-only graph"] = timeit(lambda i: ga.replay())
if rank == 0:
    for k, v in res.items(): print("%-45s host %.3f ms  total %.3f ms" % (k, v[0], v[1]), flush=True)
import threading
threading.Timer(20.0, lambda: os._exit(0)).start()
del e1, e2, g1, g2, ga
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group(); os._exit(0)
