"""Developer diagnostic: host cost of graph replays in a multi-process job.  DIAG_NODIST=1: no process group at all."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-asg_b200")]
import torch, torch.distributed as dist, torch_asg
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
nodist = bool(os.environ.get("DIAG_NODIST"))
if not nodist and world > 1:
    dist.init_process_group("nccl", device_id=dev)
T, N, L, S = 1000, 64, 40, 200
torch.manual_seed(rank)
x = torch.randn(T, N, L, device=dev).log_softmax(-1).requires_grad_(True)
y = torch.randint(0, L, (N, S), device=dev)
il = torch.full((N,), T, dtype=torch.long, device=dev); tl = torch.full((N,), S, dtype=torch.long, device=dev)
crit = torch_asg.ASGLoss(L).to(dev)
def timeit(fn, n=40):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    h = time.perf_counter() - t0; torch.cuda.synchronize(); return 1e3 * h / n, 1e3 * (time.perf_counter() - t0) / n
res = {}
g1 = torch_asg.GraphedASGLoss(crit, group=None, global_batch=N * world)
e1 = g1.prepare(x, y, il, tl, kinds=("fwd", "bwd", "step"))
res["step graph (no NCCL)"] = timeit(lambda i: e1.step.replay())
res["fwd graph"] = timeit(lambda i: e1.fwd.replay())
res["bwd graph"] = timeit(lambda i: e1.bwd.replay())
def fb(i):
    e1.fwd.replay(); e1.bwd.replay()
res["fwd+bwd graphs"] = timeit(fb)
res["step graph again"] = timeit(lambda i: e1.step.replay())
ws = e1.ws
A = crit.transition.detach()
res["eager ws.step"] = timeit(lambda i: ws.step(x.detach(), A, y, il, tl, 1.0 / N))
if rank == 0:
    print("world", world, "nodist", nodist, {k: os.environ[k] for k in os.environ if k.startswith("NCCL")})
    for k, v in res.items(): print("  %-28s host %.3f ms  total %.3f ms" % (k, v[0], v[1]), flush=True)
os._exit(0)
