"""GPU tier: the host-independent step (workspace + CUDA graphs: torch_asg/graphed.py, C ABI asg_forward /
asg_backward / asg_step) against the numpy oracle and against the eager reference-shaped route.

What is pinned here:
* ``GraphedASGLoss(crit)(x, ...)`` + ``loss.backward()`` and ``GraphedASGLoss.step`` give the oracle's loss, input
  gradient and transition gradient (fp32 1e-4 / fp64 1e-9, variable lengths, every reduction);
* a graph REPLAY sees new data written into the same tensors (the graph holds addresses, not values), and tensors at
  new addresses get their own graph;
* graph replays are bit-identical to the eager zero-allocation calls and to the plain ``ASGLoss`` route;
* the step allocates nothing once captured.
"""
import numpy as np
import pytest
import torch

from oracle import asg_oracle as O

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-4, torch.float64: 1e-9}


def rel(a, ref, floor=1e-30):
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), floor))


def make(T, N, L, S, seed, dtype, varlen=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((T, N, L))
    x = x - np.log(np.exp(x).sum(-1, keepdims=True))
    A = 0.3 * rng.standard_normal((L, L))
    y = rng.integers(0, L, (N, S))
    il = rng.integers(max(S, T // 2), T + 1, N) if varlen else np.full(N, T)
    tl = rng.integers(1, S + 1, N) if varlen else np.full(N, S)
    return x, A, y, il, tl


def oracle(x, A, y, il, tl, reduction):
    N = x.shape[1]
    g = np.full(N, 1.0 / N if reduction == "mean" else 1.0)
    r = O.asg_loss(x, y, A, il, tl, g=g)
    loss = r["loss"].mean() if reduction == "mean" else r["loss"].sum() if reduction == "sum" else r["loss"]
    return loss, r["gx"], r["gA"]


def device_case(x, A, y, il, tl, dtype, reduction):
    import torch_asg
    dev = "cuda"
    L = x.shape[2]
    crit = torch_asg.ASGLoss(L, reduction=reduction).to(dev).to(dtype)
    with torch.no_grad():
        crit.transition.copy_(torch.tensor(A, dtype=dtype))
    xt = torch.tensor(x, dtype=dtype, device=dev, requires_grad=True)
    return crit, xt, torch.tensor(y, device=dev), torch.tensor(il, device=dev), torch.tensor(tl, device=dev)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
@pytest.mark.parametrize("use_graph", [True, False])
def test_graphed_autograd_matches_oracle(dtype, reduction, use_graph):
    import torch_asg
    x, A, y, il, tl = make(60, 5, 12, 9, 1, dtype)
    crit, xt, yt, ilt, tlt = device_case(x, A, y, il, tl, dtype, reduction)
    g = torch_asg.GraphedASGLoss(crit, use_graph=use_graph)
    want_loss, want_gx, want_gA = oracle(x, A, y, il, tl, reduction)
    for rep in range(3):                     # capture, then two replays
        xt.grad = None
        crit.transition.grad = None
        loss = g(xt, yt, ilt, tlt)
        (loss.sum() if reduction == "none" else loss).backward()
        assert rel(loss.detach().cpu().numpy(), want_loss, 1.0) <= TOL[dtype], (rep, "loss")
        assert rel(xt.grad.cpu().numpy(), want_gx) <= TOL[dtype], (rep, "gx")
        assert rel(crit.transition.grad.cpu().numpy(), want_gA) <= TOL[dtype], (rep, "gA")
    if use_graph and reduction != "none":
        assert g.graph_replays >= 4          # forward + backward of the two later calls at least


def test_replay_sees_new_data_and_new_addresses():
    import torch_asg
    dtype = torch.float32
    T, N, L, S = 80, 6, 20, 10
    x, A, y, il, tl = make(T, N, L, S, 2, dtype)
    crit, xt, yt, ilt, tlt = device_case(x, A, y, il, tl, dtype, "mean")
    g = torch_asg.GraphedASGLoss(crit)
    g(xt, yt, ilt, tlt).backward()
    # (a) same tensors, new contents: every input of the graph is rewritten in place, including the transition matrix
    x2, A2, y2, il2, tl2 = make(T, N, L, S, 3, dtype)
    with torch.no_grad():
        xt.copy_(torch.tensor(x2, dtype=dtype))
        yt.copy_(torch.tensor(y2))
        ilt.copy_(torch.tensor(il2))
        tlt.copy_(torch.tensor(tl2))
        crit.transition.copy_(torch.tensor(A2, dtype=dtype))
    replays = g.graph_replays
    xt.grad = None
    crit.transition.grad = None
    loss = g(xt, yt, ilt, tlt)
    loss.backward()
    assert g.graph_replays == replays + 2
    want_loss, want_gx, want_gA = oracle(x2, A2, y2, il2, tl2, "mean")
    assert rel(float(loss), want_loss, 1.0) <= 1e-4
    assert rel(xt.grad.cpu().numpy(), want_gx) <= 1e-4
    assert rel(crit.transition.grad.cpu().numpy(), want_gA) <= 1e-4
    # (b) tensors at new addresses: a new graph, same answer
    xt3 = xt.detach().clone().requires_grad_(True)
    crit.transition.grad = None
    loss3 = g(xt3, yt.clone(), ilt.clone(), tlt.clone())
    loss3.backward()
    assert torch.equal(loss3.detach(), loss.detach())
    assert torch.equal(xt3.grad, xt.grad)
    # (c) default lengths (None) behave like full lengths (asg.py:113-117)
    la = g(xt3, yt)
    lb = crit(xt3, yt)
    assert rel(float(la), float(lb), 1.0) <= 1e-6


@pytest.mark.parametrize("reduction", ["mean", "sum"])
def test_one_graph_step_bit_identical_to_other_routes(reduction):
    """step() (one graph), graphed autograd (two graphs), eager workspace calls and the plain ASGLoss route run the same
    kernels on the same data: bit-identical gradients, and the loss agrees to rounding of the final reduction."""
    import torch_asg
    dtype = torch.float32
    x, A, y, il, tl = make(120, 7, 40, 25, 4, dtype)
    crit, xt, yt, ilt, tlt = device_case(x, A, y, il, tl, dtype, reduction)
    # plain eager route
    loss0 = crit(xt, yt, ilt, tlt)
    loss0.backward()
    gx0, gA0 = xt.grad.clone(), crit.transition.grad.clone()
    res = {}
    for name, use_graph in (("graph", True), ("eager-ws", False)):
        g = torch_asg.GraphedASGLoss(crit, use_graph=use_graph)
        for _ in range(2):
            xt.grad = None
            crit.transition.grad = None
            loss = g(xt, yt, ilt, tlt)
            loss.backward()
        res[name] = (loss.detach().clone(), xt.grad.clone(), crit.transition.grad.clone())
        for _ in range(2):
            xt.grad = None
            lstep = g.step(xt, yt, ilt, tlt)
        res[name + "-step"] = (lstep.clone(), xt.grad.clone(), crit.transition.grad.clone())
    for name, (l, gx, gA) in res.items():
        assert torch.equal(gx, gx0), name
        assert torch.equal(gA, gA0), name
        assert abs(float(l) - float(loss0)) <= 2e-6 * max(1.0, abs(float(loss0))), name
    want_loss, want_gx, want_gA = oracle(x, A, y, il, tl, reduction)
    assert rel(float(loss0), want_loss, 1.0) <= 1e-4 and rel(gx0.cpu().numpy(), want_gx) <= 1e-4 and rel(gA0.cpu().numpy(), want_gA) <= 1e-4


def test_forward_only_graph_and_eval():
    import torch_asg
    dtype = torch.float32
    x, A, y, il, tl = make(90, 4, 32, 20, 5, dtype)
    crit, xt, yt, ilt, tlt = device_case(x, A, y, il, tl, dtype, "none")
    g = torch_asg.GraphedASGLoss(crit)
    want = O.asg_loss(x, y, A, il, tl, want_grad=False)["loss"]
    with torch.no_grad():
        a = g(xt, yt, ilt, tlt)
        b = g(xt, yt, ilt, tlt)
    assert torch.equal(a, b)
    assert rel(a.cpu().numpy(), want, 1.0) <= 1e-4
    crit.eval()
    c = g(xt, yt, ilt, tlt)
    assert torch.equal(a, c) and not c.requires_grad


def test_step_allocates_nothing_once_captured():
    import torch_asg
    dtype = torch.float32
    x, A, y, il, tl = make(200, 8, 40, 30, 6, dtype, varlen=False)
    crit, xt, yt, ilt, tlt = device_case(x, A, y, il, tl, dtype, "mean")
    g = torch_asg.GraphedASGLoss(crit)
    for _ in range(3):
        g.step(xt, yt, ilt, tlt)
    torch.cuda.synchronize()
    before = torch.cuda.memory_stats()["allocation.all.allocated"]
    for _ in range(10):
        g.step(xt, yt, ilt, tlt)
    torch.cuda.synchronize()
    assert torch.cuda.memory_stats()["allocation.all.allocated"] == before
    assert g.kernel_launches >= 13 * 4        # 4 kernels per step (2 scans, grad, finalize)


def test_cfg2_graphed_step_against_eager():
    """BASELINE.json configs[1] (T=1000, N=64, L=40, S=200) through the graphs: bit-identical to the eager route."""
    import torch_asg
    torch.manual_seed(0)
    T, N, L, S = 1000, 64, 40, 200
    dev = "cuda"
    x = torch.randn(T, N, L, device=dev).log_softmax(-1).requires_grad_(True)
    y = torch.randint(0, L, (N, S), device=dev)
    crit = torch_asg.ASGLoss(L).to(dev)
    with torch.no_grad():
        crit.transition.copy_(0.1 * torch.randn(L, L, device=dev))
    loss0 = crit(x, y)
    loss0.backward()
    gx0, gA0 = x.grad.clone(), crit.transition.grad.clone()
    g = torch_asg.GraphedASGLoss(crit)
    for _ in range(3):
        x.grad = None
        crit.transition.grad = None
        loss = g(x, y)
        loss.backward()
    assert torch.equal(x.grad, gx0) and torch.equal(crit.transition.grad, gA0)
    assert abs(float(loss) - float(loss0)) <= 2e-6 * abs(float(loss0))
    l2 = g.step(x, y)
    assert torch.equal(x.grad, gx0) and torch.equal(crit.transition.grad, gA0)
    assert abs(float(l2) - float(loss0)) <= 2e-6 * abs(float(loss0))
