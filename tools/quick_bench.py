#!/usr/bin/env python
"""Developer timing script (GPU box): per-phase CUDA-event timings of the native entry points."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.environ.get("ASG_PKG", os.path.join(ROOT, "torch-asg_b200")))
import torch, numpy as np
import torch_asg_native as nat

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2], ts[0]

def main():
    T, N, L, S = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (1000, 64, 40, 200))]
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    x = torch.randn(T, N, L, generator=g).log_softmax(-1).to(dev)
    A = (0.1 * torch.randn(L, L, generator=g)).to(dev)
    y = torch.randint(0, L, (N, S), generator=g).to(dev)
    il = torch.full((N,), T, dtype=torch.long, device=dev); tl = torch.full((N,), S, dtype=torch.long, device=dev)
    gf = torch.full((N,), 1.0 / N, device=dev); ga = -gf
    res = {}
    def fwd():
        res["f"] = nat.fast_asg_gpu_forward(x, y, A, il, tl, T, N, L, S)
    def bwd():
        f = res["f"]
        res["b"] = nat.fast_asg_gpu_backward(gf, ga, f[2], f[3], f[4], f[5], y, il, tl, T, N, L, S)
    def fwdonly():
        res["fo"] = nat.fast_asg_gpu_forward_only(x, y, A, il, tl, T, N, L, S)
    def fcc():
        res["fcc"] = nat.fully_connected_forward(x, A, il, T, N, L)
    def fac():
        res["fac"] = nat.force_aligned_forward(x, y, A, il, tl, T, N, L, S)
    for name, fn in (("fcc_fwd(a+b)", fcc), ("fac_fwd(a+b)", fac), ("forward(4 scans)", fwd), ("backward(grad)", bwd), ("forward_only", fwdonly)):
        med, best = timeit(fn)
        print(f"{name:18s} median {med*1e3:9.1f} us   best {best*1e3:9.1f} us", flush=True)
    def both():
        fwd(); bwd()
    med, best = timeit(both)
    print(f"fwd+bwd            median {med*1e3:9.1f} us   best {best*1e3:9.1f} us  -> {N/(med*1e-3):.0f} utt/s")
    alg = 16*T*N*L + 8*T*N*S + 8*N*S + 8*L*L
    bm, bb = timeit(bwd)
    print(f"grad assembly algorithmic bytes {alg/1e6:.1f} MB -> {alg/(bm*1e-3)/1e9:.0f} GB/s (median)")

if __name__ == "__main__":
    main()
