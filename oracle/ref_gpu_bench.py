#!/usr/bin/env python
"""TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Times the UNMODIFIED reference's own CUDA path on this GPU:
the reference's ``asg.py`` (``oracle/_ref/ref_py``) on top of the reference's CUDA extension compiled for
sm_100a from the reference sources where they lie (``oracle/_ref/ref_cuda``, recipe: ``build_ref.build_cuda``:
``setup.py:13-32`` flags with the arch list swapped).  Nothing of the product is on ``sys.path`` here.

Run as a subprocess by ``bench.py`` (``gpu_reference`` leg) and prints one JSON line:
    python oracle/ref_gpu_bench.py T N L S steps warmup [forward_only]
The reference's CUDA path is the 4-stream orchestration of ``streamlined_fast_gpu.cpp:104-297``: a host loop
of ATen ops per frame (about 4e4 launches per forward at T=1000) -- the "recompiled kernels" baseline.
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "_ref", "ref_py"), os.path.join(HERE, "_ref", "ref_cuda")]


def main():
    T, N, L, S, steps, warmup = [int(v) for v in sys.argv[1:7]]
    fwd_only = len(sys.argv) > 7 and sys.argv[7] == "1"
    import torch
    import torch_asg
    import torch_asg_native
    assert not hasattr(torch_asg_native, "IS_B200_BUILD"), "this leg must run the reference's extension"
    assert torch_asg.__file__.startswith(os.path.join(HERE, "_ref", "ref_py"))
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = torch.randn(T, N, L, generator=g).log_softmax(-1).to(dev).requires_grad_(not fwd_only)
    A = 0.1 * torch.randn(L, L, generator=g)
    y = torch.randint(0, L, (N, S), generator=g).to(dev)
    il = torch.full((N,), T, dtype=torch.long, device=dev)
    tl = torch.full((N,), S, dtype=torch.long, device=dev)
    crit = torch_asg.ASGLoss(L, reduction="mean", forward_only=fwd_only).to(dev)
    with torch.no_grad():
        crit.transition.copy_(A.to(dev))

    def step():
        x.grad = None
        crit.transition.grad = None
        if fwd_only:
            with torch.no_grad():
                return crit(x, y, il, tl)
        loss = crit(x, y, il, tl)
        loss.backward()
        return loss.detach()

    for _ in range(warmup):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / steps
    print(json.dumps({"value": N / (ms * 1e-3), "unit": "utterances/s", "ms_per_step": ms, "wall_ms_per_step": 1e3 * wall / steps,
                      "steps": steps, "warmup": warmup, "loss": float(loss), "kind": "reference CUDA path (asg.py + native/*.cpp,.cu "
                      "unmodified, compiled for sm_100a), 4-stream ATen-op-per-frame orchestration",
                      "peak_mem_mb": torch.cuda.max_memory_allocated() / 1e6}))


if __name__ == "__main__":
    main()
