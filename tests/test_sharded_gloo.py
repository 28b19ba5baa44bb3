"""CPU tier: the multi-GPU host logic (torch_asg/sharded.py) under world_size-2 ``gloo``.

The per-rank arithmetic is a stand-in here -- a tiny differentiable log-domain ASG written with torch ops
(tests only; the product has no CPU path) -- what is under test is the sharding, the fused
``[L*L + 2]`` all-reduce, and the global handling of ``reduction='mean'``: two ranks holding half the
batch each must reproduce the single-process loss, transition gradient and input gradients."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _sharded():
    spec = importlib.util.spec_from_file_location("asg_sharded", os.path.join(ROOT, "torch-asg_b200", "torch_asg", "sharded.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class TorchASG(torch.nn.Module):
    """Stand-in criterion (same constructor/forward surface as ASGLoss), autograd through torch ops."""

    def __init__(self, num_labels, reduction="mean"):
        super().__init__()
        self.reduction = reduction
        self.transition = torch.nn.Parameter(torch.zeros(num_labels, num_labels, dtype=torch.float64))

    def forward(self, x, y, il=None, tl=None):
        T, N, L = x.shape
        out = []
        for n in range(N):
            Tn = T if il is None else int(il[n])
            Sn = y.shape[1] if tl is None else int(tl[n])
            a = x[0, n]
            for t in range(1, Tn):
                a = x[t, n] + torch.logsumexp(self.transition + a[None, :], 1)
            full = torch.logsumexp(a, 0)
            ys = y[n, :Sn]
            neg = torch.full((Sn,), -1e30, dtype=x.dtype)     # finite "zero": keeps autograd NaN-free
            b = torch.cat([x[0, n, ys[:1]], neg[1:]])
            H = self.transition[ys, ys]
            D = self.transition[ys[1:], ys[:-1]]
            for t in range(1, Tn):
                stay = b + H
                move = torch.cat([neg[:1], b[:-1] + D])
                b = x[t, n, ys] + torch.logsumexp(torch.stack([stay, move]), 0)
            out.append(full - b[Sn - 1])
        per = torch.stack(out)
        return per.sum() if self.reduction == "sum" else per.mean() if self.reduction == "mean" else per


def _data():
    g = torch.Generator().manual_seed(0)
    T, N, L, S = 9, 5, 4, 3       # N odd: uneven shards
    x = torch.randn(T, N, L, generator=g, dtype=torch.float64)
    y = torch.randint(0, L, (N, S), generator=g)
    il = torch.tensor([9, 7, 9, 5, 8])
    tl = torch.tensor([3, 2, 3, 1, 2])
    A = 0.3 * torch.randn(L, L, generator=g, dtype=torch.float64)
    return x, y, il, tl, A


def _worker(rank, world, port, reduction, q, explicit_batch=True):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = _sharded()
        x, y, il, tl, A = _data()
        xs, ys, ils, tls = sh.shard_batch(x, y, il, tl, world, rank)
        xs = xs.clone().requires_grad_(True)
        crit = TorchASG(A.shape[0], reduction)
        with torch.no_grad():
            crit.transition.copy_(A)
        # explicit_batch False: the global utterance count must come from an all-reduce (uneven shards: 3 + 2 utterances)
        loss = sh.sharded_asg_step(crit, xs, ys, ils, tls, global_batch=x.shape[1] if explicit_batch else None)
        q.put((rank, float(loss), crit.transition.grad.numpy().copy(), xs.grad.numpy().copy(), crit.reduction))
    finally:
        dist.destroy_process_group()


def test_reduction_none_is_rejected():
    sh = _sharded()
    x, y, il, tl, A = _data()
    with pytest.raises(ValueError):
        sh.sharded_asg_step(TorchASG(A.shape[0], "none"), x.clone().requires_grad_(True), y, il, tl)


@pytest.mark.parametrize("reduction,explicit_batch", [("mean", True), ("sum", True), ("mean", False)])
def test_two_rank_gloo_matches_single_process(reduction, explicit_batch):
    sh = _sharded()
    x, y, il, tl, A = _data()
    crit = TorchASG(A.shape[0], reduction)
    with torch.no_grad():
        crit.transition.copy_(A)
    xw = x.clone().requires_grad_(True)
    ref_loss = crit(xw, y, il, tl)
    ref_loss.backward()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, reduction, q, explicit_batch)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, loss, gA, gx, red in res:
        assert red == reduction                       # the criterion's setting is restored
        assert abs(loss - float(ref_loss)) < 1e-12
        assert np.abs(gA - crit.transition.grad.numpy()).max() < 1e-12
        b, e = sh.shard_bounds(x.shape[1], 2, rank)
        assert np.abs(gx - xw.grad.numpy()[:, b:e]).max() < 1e-12


def test_shard_bounds_cover_batch():
    sh = _sharded()
    for N in (1, 5, 64, 513):
        for W in (1, 2, 3, 8):
            spans = [sh.shard_bounds(N, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_bounds(4, 2, 2)


def test_pack_unpack_roundtrip():
    sh = _sharded()
    gA = torch.arange(9, dtype=torch.float32).view(3, 3)
    buf = sh.pack_for_allreduce(gA, torch.tensor(2.5), 4.0)
    assert buf.shape == (11,)
    g2, s, c = sh.unpack_after_allreduce(buf, 3)
    assert torch.equal(g2, gA) and float(s) == 2.5 and float(c) == 4.0
