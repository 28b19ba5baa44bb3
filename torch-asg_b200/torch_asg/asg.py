"""``torch_asg.asg`` -- the reference's public module name, bound to the B200 extension.

``import torch_asg_native`` resolves to the product's own sm_100a CUDA extension (built in-tree by
``__graft_entry__.build()`` next to this package).  There is no CPU or PyTorch fallback: if the
extension is missing this import raises, loudly, by design.
Reference counterpart: ``torch_asg/asg.py:4`` (the import) and ``:7-142`` (the classes).
"""
try:
    import torch_asg_native
except ImportError as exc:  # pragma: no cover - exercised only on a broken install
    raise ImportError(
        "torch_asg: the B200 CUDA extension 'torch_asg_native' is not built. "
        "Run `python -c 'import __graft_entry__ as g; g.build()'` at the repo root "
        "(no CPU fallback exists).") from exc

from ._glue import bind_native

_surface = bind_native(torch_asg_native)
FAC = _surface["FAC"]
FCC = _surface["FCC"]
ASGGPUFastForwardOnly = _surface["ASGGPUFastForwardOnly"]
ASGGPUFast = _surface["ASGGPUFast"]
ASGLoss = _surface["ASGLoss"]

__all__ = ["FAC", "FCC", "ASGGPUFastForwardOnly", "ASGGPUFast", "ASGLoss"]
