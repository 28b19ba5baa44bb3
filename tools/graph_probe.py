#!/usr/bin/env python
"""Developer probe (GPU box): eager vs CUDA-graph replay of the public-API training step."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "torch-asg_b200"))
import torch
import torch_asg
from bench import synth

T, N, L, S = 1000, 64, 40, 200
dev = torch.device("cuda")
nsets = 17
xs, ys = [], []
for i in range(nsets):
    x, A, y = synth(T, N, L, S, i, torch.float32)
    xs.append(x.to(dev).requires_grad_(True)); ys.append(y.to(dev))
il = torch.full((N,), T, dtype=torch.long, device=dev); tl = torch.full((N,), S, dtype=torch.long, device=dev)
crit = torch_asg.ASGLoss(L, reduction="mean").to(dev)
with torch.no_grad(): crit.transition.copy_(A.to(dev))

def step(i):
    x = xs[i % nsets]; x.grad = None; crit.transition.grad = None
    loss = crit(x, ys[i % nsets], il, tl)
    loss.backward()
    return loss

for i in range(5): step(i)
torch.cuda.synchronize()
K = 200
t0 = time.perf_counter(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K): step(i)
t_enq = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print("eager: cpu enqueue %.1f us/step, gpu %.1f us/step" % (1e6 * t_enq / K, 1e3 * e0.elapsed_time(e1) / K))

# graphs: one per input set, shared pool, replayed in capture order
graphs = []
pool = None
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for i in range(3): step(i)
torch.cuda.current_stream().wait_stream(s)
losses = []
for i in range(nsets):
    g = torch.cuda.CUDAGraph()
    xs[i].grad = None; crit.transition.grad = None
    with torch.cuda.graph(g, pool=pool):
        loss = crit(xs[i], ys[i], il, tl)
        loss.backward()
    pool = g.pool()
    graphs.append(g); losses.append(loss)
torch.cuda.synchronize()
for i in range(nsets): graphs[i].replay()
torch.cuda.synchronize()
t0 = time.perf_counter(); e0.record()
for i in range(K): graphs[i % nsets].replay()
t_enq = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print("graph: cpu enqueue %.1f us/step, gpu %.1f us/step, loss %.4f" % (1e6 * t_enq / K, 1e3 * e0.elapsed_time(e1) / K, float(losses[0])))
