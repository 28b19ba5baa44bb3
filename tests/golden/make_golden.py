#!/usr/bin/env python
"""Regenerates ``tests/golden/ref_cases.npz`` by running the compiled, UNMODIFIED reference
(``oracle/_ref`` built from /root/reference by ``oracle/build_ref.py``) on CPU in float64.

Only runs where /root/reference exists (the build container).  The committed .npz is what the
CPU and GPU test tiers read; nothing at test time touches /root/reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import ref_driver  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "ref_cases.npz")


def rand_case(seed, T, N, L, S, varlen=False, feasible=True, g=False, x_kind="randn",
              a_kind="uniform", tl_fixed=None, il_fixed=None):
    r = np.random.default_rng(seed)
    if x_kind == "randn":
        x = r.standard_normal((T, N, L))
    elif x_kind == "logsoftmax":
        x = r.standard_normal((T, N, L))
        x = x - np.log(np.exp(x).sum(-1, keepdims=True))
    elif x_kind == "logprob_zeros":          # log of probabilities with exact zeros -> -inf emissions
        p = r.uniform(size=(T, N, L)) * (r.uniform(size=(T, N, L)) > 0.35)
        p[..., 0] += 0.05                    # keep every frame non-degenerate
        with np.errstate(divide="ignore"):
            x = np.log(p / p.sum(-1, keepdims=True))
    A = {"uniform": lambda: r.uniform(size=(L, L)),
         "zeros": lambda: np.zeros((L, L)),
         "randn": lambda: 0.5 * r.standard_normal((L, L)),
         "forbid": lambda: np.where(r.uniform(size=(L, L)) < 0.2, -np.inf, r.standard_normal((L, L)))}[a_kind]()
    y = r.integers(0, L, (N, S))
    if x_kind == "logprob_zeros":
        # make targets follow labels that are possible: label 0 always has mass
        y[:, ::2] = 0
    il = np.full(N, T)
    tl = np.full(N, S)
    if varlen:
        tl = r.integers(1, S + 1, N)
        lo = np.maximum(tl, 1) if feasible else np.ones(N, dtype=np.int64)
        il = np.array([r.integers(lo[n], T + 1) for n in range(N)])
    if tl_fixed is not None:
        tl = np.array(tl_fixed)
    if il_fixed is not None:
        il = np.array(il_fixed)
    gv = r.uniform(0.5, 1.5, N) if g else np.ones(N)
    return dict(x=x, A=A, y=y, il=il.astype(np.int64), tl=tl.astype(np.int64), g=gv)


CASES = {
    # README / test_run shapes (cfg-1)
    "toy_full": dict(seed=1, T=6, N=2, L=7, S=5),
    "toy_varlen": dict(seed=2, T=6, N=2, L=7, S=5, varlen=True, g=True),
    "toy_infeasible": dict(seed=3, T=6, N=3, L=7, S=5, il_fixed=[6, 2, 4], tl_fixed=[5, 4, 4]),
    "neginf_emissions": dict(seed=4, T=9, N=3, L=6, S=4, x_kind="logprob_zeros", a_kind="zeros", g=True),
    "forbidden_transitions": dict(seed=5, T=10, N=3, L=6, S=3, a_kind="forbid", varlen=True),
    "s1": dict(seed=6, T=7, N=3, L=5, S=1, varlen=True),
    "t1s1": dict(seed=7, T=1, N=2, L=4, S=1),
    "t2s2": dict(seed=8, T=2, N=2, L=3, S=2),
    "s_equals_t": dict(seed=9, T=40, N=2, L=5, S=40),
    "labels2": dict(seed=10, T=12, N=3, L=2, S=6, varlen=True, g=True),
    "labels8": dict(seed=11, T=15, N=2, L=8, S=5, varlen=True),
    "labels33": dict(seed=12, T=11, N=2, L=33, S=6, g=True),
    "labels64": dict(seed=13, T=20, N=3, L=64, S=7, varlen=True, g=True),
    "long_target": dict(seed=14, T=320, N=2, L=9, S=300, x_kind="logsoftmax", a_kind="randn",
                        il_fixed=[320, 310], tl_fixed=[300, 270]),
    "mid40": dict(seed=15, T=120, N=4, L=40, S=30, varlen=True, g=True, x_kind="logsoftmax", a_kind="randn"),
    "mid32": dict(seed=16, T=100, N=3, L=32, S=40, varlen=True, x_kind="logsoftmax", a_kind="randn"),
}


def main():
    out = {}
    for name, kw in CASES.items():
        c = rand_case(**kw)
        r = ref_driver.ref_asg(c["x"], c["y"], c["A"], c["il"], c["tl"], c["g"])
        for k, v in c.items():
            out[f"{name}/{k}"] = v
        for k, v in r.items():
            out[f"{name}/ref_{k}"] = v
        print(f"{name:24s} loss={np.array2string(r['loss'], precision=4, threshold=4)}")
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
