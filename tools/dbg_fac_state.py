"""Developer diagnostic (GPU): decode the FAC states (mantissa * 2^exponent) of the fp32 and fp64 kernels and report where
they first disagree."""
import ctypes, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-asg_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import torch_asg_native as nat
from test_cabi_symbols import Layout
lib = ctypes.CDLL(os.path.join(ROOT, "torch-asg_b200", "libasg_b200.so"))
T, N, L, S = [int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (1000, 4, 40, 200))]
g = torch.Generator().manual_seed(2)
x = torch.randn(T, N, L, generator=g, dtype=torch.float64).log_softmax(-1)
A = 0.1 * torch.randn(L, L, generator=g, dtype=torch.float64)
y = torch.randint(0, L, (N, S), generator=g)
il = torch.full((N,), T, dtype=torch.long); tl = torch.full((N,), S, dtype=torch.long)
if len(sys.argv) > 5 and sys.argv[5] == "varlen":
    tl = torch.randint(1, S + 1, (N,), generator=g)
    il = torch.tensor([int(torch.randint(max(int(tl[n]), T // 2), T + 1, (1,), generator=g)) for n in range(N)])
print("il", il.tolist(), "tl", tl.tolist())
def run(dtype):
    dt = 0 if dtype == torch.float32 else 1
    lay = Layout()
    assert lib.asg_layout(dt, ctypes.c_int64(T), ctypes.c_int64(N), ctypes.c_int64(L), ctypes.c_int64(S), ctypes.byref(lay)) == 0
    out = nat.fast_asg_gpu_forward(x.to(dtype).cuda(), y.cuda(), A.to(dtype).cuda(), il.cuda(), tl.cuda(), T, N, L, S)
    torch.cuda.synchronize()
    state = out[5].cpu().numpy()
    rec = state.reshape(N, lay.tp, lay.rec_bytes)
    eb = lay.elem_bytes
    ft = np.float32 if eb == 4 else np.float64
    res = {}
    for name, om, oe in (("alpha", lay.rec_fac_alpha, lay.rec_fac_ea), ("beta", lay.rec_fac_beta, lay.rec_fac_eb)):
        m = rec[:, :T, om:om + eb * lay.sp].copy().view(ft).reshape(N, T, lay.sp).astype(np.float64)
        w = rec[:, :T, oe:oe + 4 * lay.nblk].copy().view(np.int32).reshape(N, T, lay.nblk)
        e = ((w >> 8) + ((w & 0xff) - 8)).astype(np.float64)          # packed word: (E << 8) | (d + 8); raw exponent = E + d
        with np.errstate(divide="ignore"):
            res[name] = np.log(m) + np.repeat(e, 8, axis=2) * np.log(2.0)
        res[name + "_m"], res[name + "_e"] = m, e
    res["aligned"] = out[1].double().cpu().numpy()
    return res
a32, a64 = run(torch.float32), run(torch.float64)
print("aligned scores f32", a32["aligned"], "f64", a64["aligned"])
for name in ("alpha", "beta"):
    u, v = a32[name].copy(), a64[name].copy()
    for n_ in range(N):
        u[n_, int(il[n_]):] = -np.inf; v[n_, int(il[n_]):] = -np.inf
    # compare relative to each frame's maximum (posteriors are shift invariant per frame)
    u = u - np.max(np.where(np.isfinite(u), u, -np.inf), axis=2, keepdims=True)
    v = v - np.max(np.where(np.isfinite(v), v, -np.inf), axis=2, keepdims=True)
    both = np.isfinite(u) & np.isfinite(v) & (v > -60)
    d = np.where(both, np.abs(u - v), 0.0)
    miss = np.isfinite(v) & (v > -60) & ~np.isfinite(u)
    print(name, "max |log diff| on live entries", d.max(), "at", np.unravel_index(d.argmax(), d.shape), "entries missing in f32:", int(miss.sum()))
    if miss.any():
        idx = np.argwhere(miss)[:5]
        for n_, t_, s_ in idx:
            print("  missing n,t,s", n_, t_, s_, "f64 rel", v[n_, t_, s_], "block mant f32", a32[name + "_m"][n_, t_, (s_ // 8) * 8:(s_ // 8) * 8 + 8], "exp", a32[name + "_e"][n_, t_, s_ // 8])
    if d.max() > 1e-3:
        n_, t_, s_ = np.unravel_index(d.argmax(), d.shape)
        print("  worst: f32", u[n_, t_, s_], "f64", v[n_, t_, s_], "mant", a32[name + "_m"][n_, t_, (s_ // 8) * 8:(s_ // 8) * 8 + 8], "exp", a32[name + "_e"][n_, t_, s_ // 8],
              "f64 mant", a64[name + "_m"][n_, t_, (s_ // 8) * 8:(s_ // 8) * 8 + 8], "exp", a64[name + "_e"][n_, t_, s_ // 8])
        bad = np.argwhere(d > 1e-3)
        print("  first bad (n,t,s):", bad[:3].tolist(), " count", len(bad), " frames:", sorted(set(bad[:, 1].tolist()))[:10])

# ---- backward: rows whose gradient does not sum to zero
def grads(dtype):
    xx = x.to(dtype).cuda(); AA = A.to(dtype).cuda()
    out = nat.fast_asg_gpu_forward(xx, y.cuda(), AA, il.cuda(), tl.cuda(), T, N, L, S)
    gf = torch.ones(N, dtype=dtype, device="cuda"); ga = -gf
    gA, gx = nat.fast_asg_gpu_backward(gf, ga, out[2], out[3], out[4], out[5], y.cuda(), il.cuda(), tl.cuda(), T, N, L, S)
    return gx.double().cpu().numpy(), gA.double().cpu().numpy()
gx32, gA32 = grads(torch.float32)
gx64, gA64 = grads(torch.float64)
rows = np.abs(gx32.sum(-1))
bad = np.argwhere(rows > 1e-4)
print("max mantissa alpha", np.nanmax(a32["alpha_m"]), "beta", np.nanmax(a32["beta_m"]))
print("bad rows (t,n):", bad[:20].tolist(), "count", len(bad))
print("gx rel err", np.abs(gx32 - gx64).max() / np.abs(gx64).max(), "gA rel err", np.abs(gA32 - gA64).max() / np.abs(gA64).max())
for t_, n_ in bad[:3]:
    print("frame", t_, "utt", n_)
    am, ae = a32["alpha_m"][n_, t_], a32["alpha_e"][n_, t_]
    bm, be = a32["beta_m"][n_, t_], a32["beta_e"][n_, t_]
    with np.errstate(divide="ignore"):
        la = np.log2(am) + np.repeat(ae, 8); lb = np.log2(bm) + np.repeat(be, 8)
    s = la + lb
    k = int(np.nanargmax(np.where(np.isfinite(s), s, -np.inf)))
    print("  peak position", k, "log2 alpha*beta", s[k], "block", k // 8, "alpha exp", ae[k // 8], "beta exp", be[k // 8])
    print("  alpha mant block", am[(k // 8) * 8:(k // 8) * 8 + 8], "beta mant block", bm[(k // 8) * 8:(k // 8) * 8 + 8])
    blk = np.arange(len(ae))
    pmax = np.array([np.max(am[b * 8:b * 8 + 8] * bm[b * 8:b * 8 + 8]) for b in blk])
    print("  per block ea+eb", (ae + be).astype(int).tolist())
    with np.errstate(divide="ignore"):
        print("  per block log2 max product", np.round(np.log2(pmax), 1).tolist())

# ---- posteriors gamma = alpha * beta (normalised per frame) from the decoded states
def gamma(a):
    s = a["alpha"] + a["beta"]
    s = np.where(np.isfinite(s), s, -np.inf)
    m = s.max(axis=2, keepdims=True)
    with np.errstate(invalid="ignore"):
        p = np.exp(s - m)
    p = np.where(np.isfinite(p), p, 0.0)
    return p / np.maximum(p.sum(axis=2, keepdims=True), 1e-300)
g32, g64 = gamma(a32), gamma(a64)
for n_ in range(N):
    g32[n_, int(il[n_]):] = 0; g64[n_, int(il[n_]):] = 0
dg = np.abs(g32 - g64).max(axis=2)
print("posterior (from states) max abs diff", dg.max(), "at (n,t)", np.unravel_index(dg.argmax(), dg.shape))
badf = np.argwhere(dg > 1e-3)
print("frames with posterior diff > 1e-3:", len(badf), badf[:10].tolist())
# gx difference by frame
dgx = np.abs(gx32 - gx64).max(axis=2)   # [T,N]
bb = np.argwhere(dgx > 1e-3)
print("gx diff > 1e-3 at (t,n):", len(bb), bb[:12].tolist())
if len(bb):
    t_, n_ = bb[0]
    print(" frame", t_, "utt", n_, "il", int(il[n_]), "tl", int(tl[n_]))
    print(" gx32", np.round(gx32[t_, n_], 4).tolist())
    print(" gx64", np.round(gx64[t_, n_], 4).tolist())
    print(" gamma32 nonzero", [(int(i), round(float(v), 4)) for i, v in enumerate(g32[n_, t_]) if v > 1e-3])
    print(" gamma64 nonzero", [(int(i), round(float(v), 4)) for i, v in enumerate(g64[n_, t_]) if v > 1e-3])
    print(" alpha exps", a32["alpha_e"][n_, t_].astype(int).tolist())
    print(" beta exps", a32["beta_e"][n_, t_].astype(int).tolist())
if len(bb):
    t_, n_ = bb[0]
    yy = y[n_].numpy()
    print(" labels[0:40]", yy[:40].tolist())
    fac32 = gx32[t_, n_] - gx64[t_, n_]     # difference must come from the FAC scatter (FCC parts agree to 1e-6)
    print(" gx32-gx64 by label", [(int(l), round(float(v), 4)) for l, v in enumerate(fac32) if abs(v) > 1e-3])
    for dt in (-1, 0, 1):
        gam = g64[n_, t_ + dt]
        exp_l = np.zeros(L)
        for s_, p_ in enumerate(gam):
            if s_ < int(tl[n_]): exp_l[yy[s_]] -= p_
        print(" expected FAC part from frame t%+d" % dt, [(int(l), round(float(v), 4)) for l, v in enumerate(exp_l) if abs(v) > 1e-3])
if len(bb):
    t_, n_ = bb[0]
    am = a32["alpha_m"][n_, t_].astype(np.float32); bm = a32["beta_m"][n_, t_].astype(np.float32)
    ae = a32["alpha_e"][n_, t_].astype(np.int64); be = a32["beta_e"][n_, t_].astype(np.int64)
    prod = (am * (bm * np.float32(2.0 ** 110))).astype(np.float32).reshape(-1, 8)
    pm = prod.max(axis=1)
    pexp = (pm.view(np.int32) >> 23) & 0xff
    c = ae + be
    ce = np.where(pm > 0, c + pexp, -2 ** 31)
    cmax = ce.max()
    sexp = c - cmax + 227
    print(" emulation: block pm", pm.tolist())
    print(" pexp", pexp.tolist(), "c", c.tolist())
    print(" ce", ce.tolist(), "cmax", int(cmax), "s", sexp.tolist())
    print(" alpha mant blocks 2,3", am[16:32].tolist())
    print(" beta mant blocks 2,3", bm[16:32].tolist())
