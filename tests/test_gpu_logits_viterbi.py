"""GPU tier: the two rows SURVEY 8(f) ranks first after the hot path.

* fused log_softmax + ASG (``torch_asg.ASGLossFromLogits``): loss and gradients w.r.t. the RAW activations equal those of
  ``log_softmax`` -> oracle (chain rule applied on the host in float64), including an infeasible utterance, whose input gradient
  needs the softmax term; fp32 1e-4, fp64 1e-9.
* Viterbi forced alignment (``torch_asg.viterbi.forced_align`` / C ABI ``asg_viterbi``): scores and paths against the numpy
  oracle (itself pinned by brute force in tests/test_oracle.py) -- paths are integers: exact.
"""
import numpy as np
import pytest
import torch

from oracle import asg_oracle as O

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.float64: 1e-9}


def rel(a, ref, floor=1e-30):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    fin = np.isfinite(ref)
    assert (np.isfinite(a) == fin).all()
    return float(np.abs(a[fin] - ref[fin]).max() / max(np.abs(ref[fin]).max(), floor)) if fin.any() else 0.0


def log_softmax(z):
    m = z.max(-1, keepdims=True)
    return z - m - np.log(np.exp(z - m).sum(-1, keepdims=True))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("reduction", ["mean", "none"])
def test_fused_log_softmax_matches_oracle(dtype, reduction):
    import torch_asg
    rng = np.random.default_rng(11)
    T, N, L, S = 70, 5, 12, 9
    z = 3.0 * rng.standard_normal((T, N, L))                  # raw activations, not normalised
    A = 0.3 * rng.standard_normal((L, L))
    y = rng.integers(0, L, (N, S))
    il = np.array([70, 61, 8, 70, 55])                         # utterance 2: 8 frames for 9 labels -> infeasible
    tl = np.array([9, 7, 9, 1, 4])
    g = np.full(N, 1.0 / N) if reduction == "mean" else rng.uniform(0.5, 1.5, N)
    x = log_softmax(z)
    ref = O.asg_loss(x, y, A, il, tl, g=g)
    # chain rule through log_softmax: dz = gx - softmax(z) * sum_l gx
    sm = np.exp(x)
    want_gz = ref["gx"] - sm * ref["gx"].sum(-1, keepdims=True)
    for t in range(T):
        for n in range(N):
            if t >= il[n]:
                want_gz[t, n] = 0.0
    crit = torch_asg.ASGLossFromLogits(L, reduction=reduction).cuda().to(dtype)
    with torch.no_grad():
        crit.transition.copy_(torch.tensor(A, dtype=dtype))
    zt = torch.tensor(z, dtype=dtype, device="cuda", requires_grad=True)
    loss = crit(zt, torch.tensor(y, device="cuda"), torch.tensor(il, device="cuda"), torch.tensor(tl, device="cuda"))
    if reduction == "mean":
        # the infeasible utterance makes the mean +inf (as in the reference); its gradients stay finite
        assert torch.isinf(loss)
        loss.backward()
    else:
        assert rel(loss.detach().cpu().numpy(), ref["loss"], 1.0) <= TOL[dtype]
        loss.backward(torch.tensor(g, dtype=dtype, device="cuda"))
    assert rel(zt.grad.cpu().numpy(), want_gz) <= TOL[dtype]
    assert rel(crit.transition.grad.cpu().numpy(), ref["gA"]) <= TOL[dtype]
    # and it is the same number as the unfused route on the GPU
    plain = torch_asg.ASGLoss(L, reduction="none").cuda().to(dtype)
    with torch.no_grad():
        plain.transition.copy_(torch.tensor(A, dtype=dtype))
    zt2 = torch.tensor(z, dtype=dtype, device="cuda", requires_grad=True)
    l2 = plain(zt2.log_softmax(-1), torch.tensor(y, device="cuda"), torch.tensor(il, device="cuda"), torch.tensor(tl, device="cuda"))
    l2.backward(torch.tensor(g, dtype=dtype, device="cuda"))
    assert rel(zt.grad.cpu().numpy(), zt2.grad.cpu().numpy()) <= 10 * TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(40, 6, 10, 7), (300, 4, 40, 130), (260, 3, 33, 257), (1000, 8, 40, 200)])
def test_viterbi_matches_oracle(dtype, shape):
    from torch_asg.viterbi import forced_align
    T, N, L, S = shape
    rng = np.random.default_rng(T + S)
    x = log_softmax(rng.standard_normal((T, N, L)))
    A = 0.5 * rng.standard_normal((L, L))
    y = rng.integers(0, L, (N, S))
    il = rng.integers(max(S, T // 2), T + 1, N)
    tl = rng.integers(1, S + 1, N)
    il[0], tl[0] = T, S
    if N > 2:
        il[1], tl[1] = S // 2, S                               # infeasible: more labels than frames
    xq = torch.tensor(x, dtype=dtype)                          # the oracle sees exactly the values the kernel sees
    Aq = torch.tensor(A, dtype=dtype)
    want_s, want_p = O.viterbi_align(xq.double().numpy(), y, Aq.double().numpy(), il, tl)
    scores, pos, labels = forced_align(xq.cuda(), torch.tensor(y, device="cuda"), Aq.cuda(), torch.tensor(il, device="cuda"), torch.tensor(tl, device="cuda"))
    assert rel(scores.cpu().numpy(), want_s, 1.0) <= (1e-5 if dtype == torch.float32 else 1e-12)
    got_p = pos.cpu().numpy()
    if dtype == torch.float64:
        assert (got_p == want_p).all()
    else:
        # fp32 sums can flip a near-tie: the path must be a valid monotone alignment whose score (recomputed in fp64) is the optimum
        for n in range(N):
            if not np.isfinite(want_s[n]):
                assert (got_p[n] == -1).all()
                continue
            p = got_p[n, :il[n]]
            assert p[0] == 0 and p[-1] == tl[n] - 1 and ((np.diff(p) == 0) | (np.diff(p) == 1)).all() and (got_p[n, il[n]:] == -1).all()
            ys = y[n][p]
            sc = x[np.arange(il[n]), n, ys].sum() + sum(A[ys[t], ys[t - 1]] for t in range(1, il[n]))
            assert abs(sc - want_s[n]) <= 1e-4 * max(1.0, abs(want_s[n]))
    lab = labels.cpu().numpy()
    ok = got_p >= 0
    assert (lab[ok] == np.take_along_axis(y, np.maximum(got_p, 0), 1)[ok]).all() and (lab[~ok] == -1).all()
