"""Developer diagnostic (GPU): peaked emissions, fp32 kernels vs oracle, FCC and FAC gradients separately, per frame."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-asg_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import torch_asg_native as nat
from oracle import asg_oracle as O
T, N, L, S = [int(v) for v in sys.argv[1:5]]
scale = float(sys.argv[5])
rng = np.random.default_rng(int(scale) * 100 + T)
x = scale * rng.standard_normal((T, N, L))
A = 0.3 * rng.standard_normal((L, L))
y = rng.integers(0, L, (N, S))
il = rng.integers(max(S, T // 2), T + 1, N); il[0] = T
tl = rng.integers(1, S + 1, N); tl[0] = S
g = rng.uniform(0.5, 1.5, N)
xq = torch.tensor(x, dtype=torch.float32).double().numpy()
Aq = torch.tensor(A, dtype=torch.float32).double().numpy()
ref = O.asg_loss(xq, y, Aq, il, tl, g=g)
dev = "cuda"
for dtype in (torch.float32, torch.float64):
    xt = torch.tensor(xq, dtype=dtype, device=dev); At = torch.tensor(Aq, dtype=dtype, device=dev)
    yt, ilt, tlt = torch.tensor(y, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev)
    full, aligned, gf, ga, cf, ca = nat.fast_asg_gpu_forward(xt, yt, At, ilt, tlt, T, N, L, S)
    gt = torch.tensor(g, dtype=dtype, device=dev); z = torch.zeros_like(gt)
    gA_f, gx_f = nat.fast_asg_gpu_backward(gt, z, gf, ga, cf, ca, yt, ilt, tlt, T, N, L, S)
    gA_a, gx_a = nat.fast_asg_gpu_backward(z, gt, gf, ga, cf, ca, yt, ilt, tlt, T, N, L, S)
    flags = nat.range_flags(ca, cf, dtype == torch.float64, T, N, L, S).cpu().numpy()
    print(dtype, "flags", flags, "full", full.cpu().numpy(), ref["full"], "aligned", aligned.cpu().numpy(), ref["aligned"])
    for name, got, want in (("fcc", gx_f, ref["gx_full"]), ("fac", gx_a, ref["gx_aligned"])):
        got = got.double().cpu().numpy()
        for n in range(N):
            e = np.abs(got[:, n] - want[:, n]).max(-1)
            print(" ", name, "utt", n, "il", il[n], "tl", tl[n], "max err", e.max(), "ref max", np.abs(want[:, n]).max(), "bad frames", np.argwhere(e > 1e-4).ravel()[:20].tolist())
            bad = np.argwhere(e > 1e-4).ravel()
            for t in bad[:2]:
                print("    t", t, "got", np.round(got[t, n], 4).tolist(), "want", np.round(want[t, n], 4).tolist(), "x", np.round(xq[t, n], 1).tolist())
    for name, got, want in (("gA fcc", gA_f, ref["gA_full"]), ("gA fac", gA_a, ref["gA_aligned"])):
        got = got.double().cpu().numpy()
        print(" ", name, "err", np.abs(got - want).max(), "ref max", np.abs(want).max())
