"""Developer diagnostic (GPU box): where does the host time of the graphed autograd route go?"""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-asg_b200")]
import torch, torch_asg
T, N, L, S = 1000, 64, 40, 200
dev = "cuda"
torch.manual_seed(0)
nsets = int(sys.argv[1]) if len(sys.argv) > 1 else 3
xs = [torch.randn(T, N, L, device=dev).log_softmax(-1).requires_grad_(True) for _ in range(nsets)]
ys = [torch.randint(0, L, (N, S), device=dev) for _ in range(nsets)]
il = torch.full((N,), T, dtype=torch.long, device=dev); tl = torch.full((N,), S, dtype=torch.long, device=dev)
crit = torch_asg.ASGLoss(L).to(dev)
g = torch_asg.GraphedASGLoss(crit, max_graphs=4 * nsets)
for i in range(nsets):
    g.prepare(xs[i], ys[i], il, tl, kinds=("fwd", "bwd", "step"))
def step(i):
    x = xs[i % nsets]; x.grad = None; crit.transition.grad = None
    loss = g(x, ys[i % nsets], il, tl); loss.backward(); return loss
for i in range(5): step(i)
torch.cuda.synchronize()
def timeit(fn, n=50):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    h = time.perf_counter() - t0; torch.cuda.synchronize(); return 1e3 * h / n, 1e3 * (time.perf_counter() - t0) / n
print("public step: host %.3f ms, total %.3f ms" % timeit(step))
def fwd_only(i):
    x = xs[i % nsets]
    return g(x, ys[i % nsets], il, tl)
print("fwd apply only: host %.3f ms, total %.3f ms" % timeit(fwd_only))
e = g.prepare(xs[0], ys[0], il, tl)
print("raw fwd replay: host %.3f ms, total %.3f ms" % timeit(lambda i: e.fwd.replay()))
print("raw bwd replay: host %.3f ms, total %.3f ms" % timeit(lambda i: e.bwd.replay()))
print("raw step replay: host %.3f ms, total %.3f ms" % timeit(lambda i: e.step.replay()))
def fb(i):
    e.fwd.replay(); e.bwd.replay()
print("raw fwd+bwd replay: host %.3f ms, total %.3f ms" % timeit(fb))
pr = cProfile.Profile(); pr.enable()
for i in range(50): step(i)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:6000])
import ctypes
lib = ctypes.CDLL(os.path.join(ROOT, "torch-asg_b200", "libasg_b200.so"))
def rot(i):
    ee = g.prepare(xs[i % nsets], ys[i % nsets], il, tl)
    ee.fwd.replay(); ee.bwd.replay()
print("rotating raw fwd+bwd replay: host %.3f ms, total %.3f ms" % timeit(rot))
def rot1(i):
    ee = g.prepare(xs[i % nsets], ys[i % nsets], il, tl)
    ee.step.replay()
print("rotating raw step replay: host %.3f ms, total %.3f ms" % timeit(rot1))
print("public step again: host %.3f ms, total %.3f ms" % timeit(step))
def eager_ws(i):
    g.use_graph = False
    step(i)
    g.use_graph = True
print("eager-ws step (no profiler): host %.3f ms, total %.3f ms" % timeit(eager_ws))
lib.asg_profile_select(ctypes.c_uint(4)); lib.asg_profile_stride(ctypes.c_uint(1)); lib.asg_profile_enable(1)
print("eager-ws step (profiler on, grad only): host %.3f ms, total %.3f ms" % timeit(eager_ws))
print("public step (profiler on): host %.3f ms, total %.3f ms" % timeit(step))
def mixed(i):
    if i % 8 == 7: eager_ws(i)
    else: step(i)
print("mixed 1/8 eager (profiler on): host %.3f ms, total %.3f ms" % timeit(mixed, 96))
lib.asg_profile_enable(0)
print("public step (profiler off): host %.3f ms, total %.3f ms" % timeit(step))
