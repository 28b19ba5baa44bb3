"""GPU tier: dynamic range (severity high).

The kernels run the recursions in the real domain with power-of-two rescaling instead of the reference's log domain.  That is
exact as long as the entries that share a scale (a frame's emissions; the 8 target positions of a block) lie within the
format's range of each other -- ~87 nats in fp32, ~708 in fp64.  These tests drive emissions scaled by 10-50 (hundreds of nats
of spread) through both dtypes and assert the contract:
  * float64 kernels match the float64 oracle (1e-9) without any flag up to scale 10, and are accurate-or-flagged beyond;
  * float32 kernels NEVER return a silently wrong result: every utterance is either within 1e-4 of the oracle or carries a
    range flag (``asg_range_flags``), no output is NaN, and mass is flushed, not clamped (the aligned score of an unflagged
    utterance is exact);
  * normalised inputs never raise a flag.
"""
import numpy as np
import pytest
import torch

from oracle import asg_oracle as O

pytestmark = pytest.mark.gpu


def run(x, A, y, il, tl, g, dtype):
    import torch_asg_native as nat
    dev = "cuda"
    T, N, L = x.shape
    S = y.shape[1]
    xt = torch.tensor(x, dtype=dtype, device=dev)
    At = torch.tensor(A, dtype=dtype, device=dev)
    yt, ilt, tlt = torch.tensor(y, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev)
    full, aligned, gf, ga, cf, ca = nat.fast_asg_gpu_forward(xt, yt, At, ilt, tlt, T, N, L, S)
    gt = torch.tensor(g, dtype=dtype, device=dev)
    gA, gx = nat.fast_asg_gpu_backward(gt, -gt, gf, ga, cf, ca, yt, ilt, tlt, T, N, L, S)
    flags = nat.range_flags(ca, cf, dtype == torch.float64, T, N, L, S)
    return {"loss": (full - aligned).double().cpu().numpy(), "aligned": aligned.double().cpu().numpy(), "gx": gx.double().cpu().numpy(),
            "gA": gA.double().cpu().numpy(), "flags": flags.cpu().numpy()}


def per_utt_err(got, ref, n):
    e_loss = abs(got["loss"][n] - ref["loss"][n]) / max(abs(ref["loss"][n]), 1.0) if np.isfinite(got["loss"][n]) else np.inf
    e_gx = np.abs(got["gx"][:, n] - ref["gx"][:, n]).max() / max(np.abs(ref["gx"][:, n]).max(), 1e-30)
    return max(e_loss, e_gx)


@pytest.mark.parametrize("scale", [1.0, 10.0, 20.0, 50.0])
@pytest.mark.parametrize("shape", [(60, 2, 12, 10), (200, 6, 40, 60)])
def test_peaked_emissions(scale, shape):
    T, N, L, S = shape
    rng = np.random.default_rng(int(scale) * 100 + T)
    x = scale * rng.standard_normal((T, N, L))
    A = 0.3 * rng.standard_normal((L, L))
    y = rng.integers(0, L, (N, S))
    il = rng.integers(max(S, T // 2), T + 1, N); il[0] = T
    tl = rng.integers(1, S + 1, N); tl[0] = S
    g = rng.uniform(0.5, 1.5, N)
    ref = O.asg_loss(x, y, A, il, tl, g=g)
    # float64 kernels: exact, and without a flag up to scale 10 (their range is ~708 nats); beyond that, accurate or flagged
    d = run(x, A, y, il, tl, g, torch.float64)
    assert not np.isnan(d["gx"]).any() and not np.isnan(d["gA"]).any()
    if scale <= 10.0:
        assert (d["flags"] == 0).all()
    for n in range(N):
        if d["flags"][n] == 0:
            assert per_utt_err(d, ref, n) <= (1e-9 if scale <= 20.0 else 1e-8), ("f64", scale, n)
    if (d["flags"] == 0).all():
        assert np.abs(d["gA"] - ref["gA"]).max() / np.abs(ref["gA"]).max() <= 1e-9
    # float32 kernels: accurate or flagged, never NaN
    xq = torch.tensor(x, dtype=torch.float32).double().numpy()          # the oracle sees the values the fp32 kernels see
    Aq = torch.tensor(A, dtype=torch.float32).double().numpy()
    ref32 = O.asg_loss(xq, y, Aq, il, tl, g=g)
    f = run(xq, Aq, y, il, tl, g, torch.float32)
    assert not np.isnan(f["gx"]).any() and not np.isnan(f["gA"]).any() and not np.isnan(f["loss"]).any()
    clean = f["flags"] == 0
    for n in range(N):
        if clean[n]:
            assert per_utt_err(f, ref32, n) <= 1e-4, ("f32", scale, n, int(f["flags"][n]))
    if clean.all():
        assert np.abs(f["gA"] - ref32["gA"]).max() / np.abs(ref32["gA"]).max() <= 1e-4
    if scale == 1.0:
        assert clean.all()                                               # unit-scale inputs never raise a flag


def test_normalised_inputs_raise_no_flag():
    rng = np.random.default_rng(3)
    T, N, L, S = 1000, 16, 40, 200
    z = rng.standard_normal((T, N, L))
    x = z - np.log(np.exp(z).sum(-1, keepdims=True))
    A = 0.1 * rng.standard_normal((L, L))
    y = rng.integers(0, L, (N, S))
    f = run(x, A, y, np.full(N, T), np.full(N, S), np.full(N, 1.0 / N), torch.float32)
    assert (f["flags"] == 0).all()


def test_out_of_range_labels_are_reported_not_silently_clamped():
    """The reference gathers out of bounds for a label outside [0, L) (force_aligned_lattice.cpp:36-60).  Here: bit 8 of the
    utterance's range flags (the label itself is clamped so nothing faults), GraphedASGLoss.check() raises, and
    ASGLoss(validate_targets=True) raises before the call.  Labels beyond target_lengths are padding and never looked at."""
    import torch_asg
    import torch_asg_native as nat
    dev = "cuda"
    T, N, L, S = 50, 4, 10, 8
    g = torch.Generator().manual_seed(0)
    x = torch.randn(T, N, L, generator=g).log_softmax(-1).to(dev)
    y = torch.randint(0, L, (N, S), generator=g)
    tl = torch.tensor([8, 5, 8, 3])
    y[1, 6] = 99          # padding (beyond target_lengths[1] = 5): fine
    y[3, 7] = -4          # padding
    y[2, 4] = L           # live and out of range
    il = torch.full((N,), T)
    A = torch.zeros(L, L, device=dev)
    out = nat.fast_asg_gpu_forward(x, y.to(dev), A, il.to(dev), tl.to(dev), T, N, L, S)
    flags = nat.range_flags(out[5], out[4], False, T, N, L, S).cpu()
    assert flags.tolist() == [0, 0, 8, 0]
    crit = torch_asg.ASGLoss(L).to(dev)
    gc = torch_asg.GraphedASGLoss(crit)
    xg = x.clone().requires_grad_(True)
    gc.step(xg, y.to(dev), il.to(dev), tl.to(dev))
    with pytest.raises(IndexError, match="utterance 2"):
        gc.check()
    y[2, 4] = 0
    gc.step(xg, y.to(dev).clone(), il.to(dev), tl.to(dev))
    assert gc.check().tolist() == [0, 0, 0, 0]
    y[2, 4] = -1
    strict = torch_asg.ASGLoss(L, validate_targets=True).to(dev)
    with pytest.raises(IndexError, match="utterance 2"):
        strict(x, y.to(dev), il.to(dev), tl.to(dev))
