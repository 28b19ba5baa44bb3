"""TEST INFRASTRUCTURE ONLY.  Drives the compiled, unmodified reference CPU implementation
(``oracle/_ref``, see ``build_ref.py``) through the same Python glue the product uses
(``torch-asg_b200/torch_asg/_glue.py`` -- a mirror of the reference's ``asg.py``; the reference's
own ``asg.py`` lives under /root/reference and cannot travel to the GPU box).

Used to (a) pin ``asg_oracle.py``, (b) generate ``tests/golden/*.npz``, (c) time the reference's
CPU path for ``bench.py --impl reference`` and the ``cpu_baseline`` leg.
"""
import importlib.util
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_GLUE = os.path.join(_HERE, "..", "torch-asg_b200", "torch_asg", "_glue.py")
_cache = {}


def available():
    from . import build_ref  # noqa
    return build_ref.build() is not None


def surface():
    """dict(FAC, FCC, ASGLoss, ...) bound to the reference's native CPU module."""
    if "s" not in _cache:
        try:
            from . import build_ref
        except ImportError:  # executed as a script / plain module
            import build_ref
        native = build_ref.load()
        spec = importlib.util.spec_from_file_location("_asg_glue_for_ref", _GLUE)
        glue = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(glue)
        _cache["s"] = glue.bind_native(native)
        _cache["native"] = native
    return _cache["s"]


def ref_asg(x, y, A, il=None, tl=None, g=None, dtype=torch.float64):
    """Run the reference (CPU) on tensors/arrays; returns dict of numpy arrays
    loss[N] (reduction none), full[N], aligned[N], gx[T,N,L], gA[L,L] for upstream grad g[N]."""
    s = surface()
    x = torch.as_tensor(x).detach().to(dtype).clone().requires_grad_(True)
    A = torch.as_tensor(A).detach().to(dtype).clone().requires_grad_(True)
    y = torch.as_tensor(y).long()
    T, N, L = x.shape
    S = y.shape[1]
    tl = torch.full((N,), S, dtype=torch.long) if tl is None else torch.as_tensor(tl).long()
    il = torch.full((N,), T, dtype=torch.long) if il is None else torch.as_tensor(il).long()
    if S > T:
        y = y[:, :T]
        tl = torch.clamp(tl, max=T)
    full = s["FCC"].apply(A, x, y, il, tl)
    aligned = s["FAC"].apply(A, x, y, il, tl)
    loss = full - aligned
    g = torch.ones(N, dtype=dtype) if g is None else torch.as_tensor(g).to(dtype)
    # two separate backward calls: an infeasible utterance has aligned=-inf (loss=+inf) but the
    # reference's backward only looks at the upstream gradient, so its grads stay finite.
    (full * g).sum().backward()
    (aligned * (-g)).sum().backward()
    return {"loss": loss.detach().numpy(), "full": full.detach().numpy(),
            "aligned": aligned.detach().numpy(), "gx": x.grad.numpy(), "gA": A.grad.numpy()}
