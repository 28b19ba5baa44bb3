"""Developer diagnostic (GPU): re-evaluate the range-report criterion on the host from the stored FAC states and list the
(frame, block) pairs that raise RF_POSTERIOR_THIN."""
import ctypes, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "torch-asg_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import torch_asg_native as nat
from test_cabi_symbols import Layout
lib = ctypes.CDLL(os.path.join(ROOT, "torch-asg_b200", "libasg_b200.so"))
T, N, L, S = [int(v) for v in sys.argv[1:5]]
scale = float(sys.argv[5]); dtype = torch.float64 if sys.argv[6] == "f64" else torch.float32
rng = np.random.default_rng(int(scale) * 100 + T)
x = scale * rng.standard_normal((T, N, L))
A = 0.3 * rng.standard_normal((L, L))
y = rng.integers(0, L, (N, S))
il = rng.integers(max(S, T // 2), T + 1, N); il[0] = T
tl = rng.integers(1, S + 1, N); tl[0] = S
dt = 0 if dtype == torch.float32 else 1
lay = Layout()
assert lib.asg_layout(dt, ctypes.c_int64(T), ctypes.c_int64(N), ctypes.c_int64(L), ctypes.c_int64(S), ctypes.byref(lay)) == 0
dev = "cuda"
out = nat.fast_asg_gpu_forward(torch.tensor(x, dtype=dtype, device=dev), torch.tensor(y, device=dev), torch.tensor(A, dtype=dtype, device=dev),
                               torch.tensor(il, device=dev), torch.tensor(tl, device=dev), T, N, L, S)
torch.cuda.synchronize()
rec = out[5].cpu().numpy().reshape(N, lay.tp, lay.rec_bytes)
eb_ = lay.elem_bytes
ft = np.float32 if eb_ == 4 else np.float64
HALF, MAXS, MANT, BIAS = (55, 126, 24, 127) if eb_ == 4 else (480, 1022, 53, 1023)
MINN = np.finfo(ft).tiny
def dec(om, oe):
    m = rec[:, :T, om:om + eb_ * lay.sp].copy().view(ft).reshape(N, T, lay.nblk if False else lay.sp // 8, 8).astype(np.float64)
    w = rec[:, :T, oe:oe + 4 * lay.nblk].copy().view(np.int32).reshape(N, T, lay.nblk)
    return m[:, :, :lay.nblk], w
am, aw = dec(lay.rec_fac_alpha, lay.rec_fac_ea)
bm, bw = dec(lay.rec_fac_beta, lay.rec_fac_eb)
def bexp(v):
    return np.floor(np.log2(v)).astype(int) + BIAS
for n in range(N):
    hits = []
    for t in range(int(il[n])):
        E = {}
        ces = []
        info = []
        for b in range(lay.nblk):
            wa, wb = int(aw[n, t, b]), int(bw[n, t, b])
            da, db = (wa & 0xff) - 8, (wb & 0xff) - 8
            ea, eb = wa >> 8, wb >> 8
            dca, dcb = min(da, MAXS + 1 - HALF), min(db, MAXS + 1 - HALF)
            ea += da - dca; eb += db - dcb
            an = am[n, t, b] * 2.0 ** (dca + HALF); bn = bm[n, t, b] * 2.0 ** (dcb + HALF)
            p = an * bn
            c = ea + eb
            ce = c + bexp(p.max()) if p.max() > 0 else None
            info.append((b, wa, wb, da, db, c, ce, an, bn))
            if ce is not None: ces.append(ce)
        if not ces: continue
        cmax = max(ces)
        for (b, wa, wb, da, db, c, ce, an, bn) in info:
            live = (wa & 0xff) != 0 and (wb & 0xff) != 0
            gap = BIAS + 2 * HALF - MAXS + 5 + MANT + max(da, db, 0)
            if live and cmax - c < gap:
                k = HALF - MAXS + 3 + MANT
                za = am[n, t, b] < MINN; zb = bm[n, t, b] < MINN
                b0 = bn[za].max() if za.any() else 0.0
                a0 = an[zb].max() if zb.any() else 0.0
                r1 = b0 > 0 and cmax - c < bexp(b0) + k + max(da, 0)
                r2 = a0 > 0 and cmax - c < bexp(a0) + k + max(db, 0)
                if r1 or r2:
                    hits.append((t, b, da, db, cmax - c, r1, r2, np.log2(np.maximum(am[n, t, b], 1e-320)).round(0).tolist(), np.log2(np.maximum(bm[n, t, b], 1e-320)).round(0).tolist()))
    print("utt", n, "il", il[n], "tl", tl[n], "hits", len(hits))
    for h in hits[:4]:
        print("   t,b,da,db,cmax-c,r1,r2:", h[:7], "\n      log2 alpha", h[7], "\n      log2 beta ", h[8])
