#!/usr/bin/env python
"""Developer diagnostic (GPU box): runs every golden case through the CUDA extension in the three
code paths and both dtypes and prints per-tensor errors without stopping at the first failure."""
import os, sys, time, traceback
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "torch-asg_b200"))
import numpy as np
import torch
import torch_asg
from torch_asg.asg import FCC, FAC, ASGLoss

def rel(a, ref, floor):
    a = np.asarray(a, dtype=np.float64); fin = np.isfinite(ref)
    if not (np.isfinite(a) == fin).all(): return float("nan")
    if not fin.any(): return 0.0
    if not (a[~fin] == ref[~fin]).all(): return float("inf")
    return float(np.abs(a[fin] - ref[fin]).max() / max(np.abs(ref[fin]).max(), floor))

def run_case(c, dtype, path):
    dev = "cuda"
    x = torch.tensor(c["x"], dtype=dtype, device=dev, requires_grad=True)
    y = torch.tensor(c["y"], device=dev); il = torch.tensor(c["il"], device=dev); tl = torch.tensor(c["tl"], device=dev)
    g = torch.tensor(c["g"], dtype=dtype, device=dev)
    L = x.shape[2]
    m = ASGLoss(L, reduction="none", gpu_no_stream_impl=(path == "serial")).to(dev).to(dtype)
    with torch.no_grad(): m.transition.copy_(torch.tensor(c["A"], dtype=dtype))
    out = {}
    if path == "fwdonly":
        m.eval()
        with torch.no_grad(): out["loss"] = m(x, y, il, tl).cpu().numpy()
        return out
    if path == "serial":
        full = FCC.apply(m.transition, x, y, il, tl); al = FAC.apply(m.transition, x, y, il, tl)
        out["full"] = full.detach().cpu().numpy(); out["aligned"] = al.detach().cpu().numpy()
        loss = full - al
        (full * g).sum().backward(); (al * (-g)).sum().backward()
    else:
        loss = m(x, y, il, tl)
        fin = torch.isfinite(loss)
        (torch.where(fin, loss, torch.zeros_like(loss)) * g).sum().backward() if bool(fin.all()) else None
    out["loss"] = loss.detach().cpu().numpy()
    if x.grad is not None:
        out["gx"] = x.grad.cpu().numpy(); out["gA"] = m.transition.grad.cpu().numpy()
    return out

def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_cases.npz"))
    cases = {}
    for k in z.files:
        n, f = k.split("/"); cases.setdefault(n, {})[f] = z[k]
    only = sys.argv[1:] 
    worst = 0.0
    for name, c in cases.items():
        if only and name not in only: continue
        for dtype in (torch.float64, torch.float32):
            for path in ("fast", "serial", "fwdonly"):
                try:
                    o = run_case(c, dtype, path)
                    torch.cuda.synchronize()
                    errs = []
                    for k, v in o.items():
                        e = rel(v, c["ref_" + k], 1.0 if k in ("loss", "full", "aligned") else 1e-30)
                        errs.append(f"{k}:{e:.1e}")
                        tol = 1e-9 if dtype == torch.float64 else 1e-4
                        if not (e <= tol): errs[-1] += "<<<FAIL"
                        if np.isfinite(e): worst = max(worst, e if dtype == torch.float32 else 0)
                    print(f"{name:22s} {str(dtype)[6:]:8s} {path:8s}", *errs, flush=True)
                except Exception as ex:
                    print(f"{name:22s} {str(dtype)[6:]:8s} {path:8s} EXC {type(ex).__name__}: {str(ex)[:200]}", flush=True)
                    if "CUDA error" in str(ex) or "illegal" in str(ex): raise
    print("worst fp32 err", worst)

if __name__ == "__main__":
    main()
