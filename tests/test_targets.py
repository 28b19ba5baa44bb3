"""replabel packing / unpacking and length scaling.  CPU tier: the oracle against the technical report's worked
example and round-trip properties; GPU tier: the device kernels against the oracle, ScaledASGLoss against the README's recipe."""
import numpy as np
import pytest
import torch

from oracle import asg_oracle as O


def test_oracle_replabels_worked_example_and_cap():
    i, j, k, l, R = 0, 1, 2, 3, 10
    # doc/tech_report.tex:126:  i j j k l l l  ->  i j r_2 k l r_3   (here: extra repeats, so R+0 and R+1)
    assert O.pack_replabels([i, j, j, k, l, l, l], R, 3) == [i, j, R + 0, k, l, R + 1]
    assert O.pack_replabels([5, 5, 5, 5], R, 2) == [5, R + 1, 5]                 # a run longer than max_reps + 1 starts over
    assert O.pack_replabels([5, 5, 5, 5, 5, 5], R, 2) == [5, R + 1, 5, R + 1]
    assert O.pack_replabels([5, 5], R, 0) == [5, 5] and O.pack_replabels([], R, 2) == []
    assert O.unpack_replabels([R + 1, 4, R + 0], R, 2) == [4, 4]               # a leading replabel has nothing to repeat


def test_oracle_replabels_round_trip_and_no_repeats():
    rng = np.random.default_rng(0)
    for _ in range(300):
        L, m = int(rng.integers(1, 5)), int(rng.integers(1, 4))
        toks = rng.integers(0, L, int(rng.integers(0, 40))).tolist()
        packed = O.pack_replabels(toks, L, m)
        assert O.unpack_replabels(packed, L, m) == toks
        assert len(packed) <= len(toks)
        # the point of the exercise: equal neighbours only remain where a capped run starts over right after its replabel
        for a, b in zip(packed, packed[1:]):
            assert a != b or a >= L


def test_oracle_length_scale():
    assert O.length_scale("none", 100, 10) == 1.0
    assert O.length_scale("input_sz", 100, 10) == 0.01 and O.length_scale("target_sz", 100, 10) == 0.1
    assert abs(O.length_scale("input_sz_sqrt", 100, 10) - 0.1) < 1e-15 and O.length_scale("target_sz_sqrt", 100, 0) == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("max_reps", [0, 1, 2, 3])
def test_device_replabels_match_the_oracle(max_reps):
    import torch_asg
    rng = np.random.default_rng(max_reps)
    N, S, L = 37, 61, 3                                    # few labels: long runs
    y = rng.integers(0, L, (N, S))
    y[0, :] = 1                                            # one run over the whole row
    tl = rng.integers(0, S + 1, N); tl[0] = S; tl[1] = 0; tl[2] = 1
    yt = torch.tensor(y, device="cuda")
    packed, plen = torch_asg.pack_replabels(yt, torch.tensor(tl, device="cuda"), L, max_reps, fill=-1)
    assert packed.shape == (N, S) and packed.dtype == torch.long
    packed_c, plen_c = packed.cpu().numpy(), plen.cpu().numpy()
    for n in range(N):
        want = O.pack_replabels(y[n, :tl[n]].tolist(), L, max_reps)
        assert plen_c[n] == len(want) and packed_c[n, :len(want)].tolist() == want and (packed_c[n, len(want):] == -1).all(), n
    # round trip on the device; a transposed (strided) view as input
    un, ulen = torch_asg.unpack_replabels(packed, plen, L, max_reps, S, fill=-1)
    un_c, ulen_c = un.cpu().numpy(), ulen.cpu().numpy()
    for n in range(N):
        assert ulen_c[n] == tl[n] and un_c[n, :tl[n]].tolist() == y[n, :tl[n]].tolist() and (un_c[n, tl[n]:] == -1).all(), n
    p2, l2 = torch_asg.pack_replabels(yt.t().contiguous().t(), torch.tensor(tl, device="cuda"), L, max_reps, fill=-1)
    assert torch.equal(p2, packed) and torch.equal(l2, plen)
    # truncation is reported through the length
    if max_reps:
        short, slen = torch_asg.unpack_replabels(packed[:1], plen[:1], L, max_reps, 5, fill=-1)
        assert int(slen[0]) == S and short.shape == (1, 5) and short[0].tolist() == [1] * 5


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["none", "input_sz", "input_sz_sqrt", "target_sz", "target_sz_sqrt"])
def test_scaled_loss_is_the_readme_recipe(mode):
    import torch_asg
    torch.manual_seed(0)
    T, N, L, S = 40, 5, 9, 7
    dev = "cuda"
    x = torch.randn(T, N, L, device=dev).log_softmax(-1).requires_grad_(True)
    y = torch.randint(0, L, (N, S), device=dev)
    il = torch.tensor([40, 33, 25, 40, 21], device=dev); tl = torch.tensor([7, 5, 7, 1, 3], device=dev)
    crit = torch_asg.ASGLoss(L, reduction="mean").to(dev)
    with torch.no_grad():
        crit.transition.copy_(0.2 * torch.randn(L, L, device=dev))
    scaled = torch_asg.ScaledASGLoss(crit, mode=mode, reduction="sum")
    loss = scaled(x, y, il, tl)
    loss.backward()
    assert crit.reduction == "mean"                        # the wrapped criterion is left as it was
    ref = O.asg_loss(x.detach().double().cpu().numpy(), y.cpu().numpy(), crit.transition.detach().double().cpu().numpy(), il.cpu().numpy(), tl.cpu().numpy(),
                     g=np.array([O.length_scale(mode, int(il[n]), int(tl[n])) for n in range(N)]))
    sc = np.array([O.length_scale(mode, int(il[n]), int(tl[n])) for n in range(N)])
    assert abs(float(loss) - float((ref["loss"] * sc).sum())) <= 1e-4 * abs(float((ref["loss"] * sc).sum()))
    assert np.abs(x.grad.double().cpu().numpy() - ref["gx"]).max() <= 1e-4 * np.abs(ref["gx"]).max()
    assert np.abs(crit.transition.grad.double().cpu().numpy() - ref["gA"]).max() <= 1e-4 * np.abs(ref["gA"]).max()
