"""Batch-sharded ASG over the GPUs of one box (one process per GPU, ``torch.distributed``).

Utterances are independent: rank ``r`` owns a contiguous slice of the batch, runs the whole
path (four scans + gradient assembly) on it, and the only cross-rank quantities are sums over utterances --
the transition gradient ``gA[L,L]`` (reference: the ``.sum({0,1})`` of ``fully_connected_lattice.cpp:60`` and
the scatter of ``force_aligned_lattice.cpp:220-230`` both run over the batch) and the reduced loss
(``asg.py:137-140``).  They travel in ONE all-reduce of a fused ``[L*L + 2]`` buffer
(``gA``, ``sum of the losses``, ``count of utterances``); ``grad_inputs`` never leaves its rank (it feeds
that rank's acoustic model).  There is no other data-path collective.  Like the reference's own ``.sum()`` /
``.mean()``, the loss sum is NOT filtered: one infeasible utterance (target longer than its input, loss ``+inf``)
makes the reduced loss ``+inf`` on every rank (its gradients stay finite).

This file is the EAGER sharded step (per-call allocations, a Python ``all_reduce``); the host-independent route --
persistent workspace, CUDA graphs, the all-reduce captured in the backward graph -- is ``torch_asg.GraphedASGLoss``
(``graphed.py``), which ``bench.py`` times.

The reference itself has no distributed code; this is the B200-side answer to BASELINE.json's
"Batch-sharded" config.  The arithmetic lives in the CUDA extension; this file is host plumbing only and
is exercised on CPU by ``tests/test_sharded_gloo.py`` (world_size 2, ``gloo``) with a stand-in compute
function.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world_size, rank):
    """Contiguous, balanced split of ``batch`` utterances: the first ``batch % world_size`` ranks get one
    extra.  Returns ``(begin, end)``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(batch, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_batch(inputs, targets, input_lengths, target_lengths, world_size, rank):
    """Views (no copies) of this rank's utterances: ``inputs[:, b:e, :]`` is a strided view -- the
    kernels take strides -- and ``targets[b:e]`` / lengths are row slices."""
    b, e = shard_bounds(inputs.shape[1], world_size, rank)
    il = None if input_lengths is None else input_lengths[b:e]
    tl = None if target_lengths is None else target_lengths[b:e]
    return inputs[:, b:e, :], targets[b:e], il, tl


def pack_for_allreduce(grad_transition, loss_sum, count):
    """Fused buffer ``[L*L + 2]`` in the gradient's dtype: gA (flattened), local loss sum, local count."""
    flat = grad_transition.reshape(-1)
    buf = flat.new_empty(flat.numel() + 2)
    buf[:-2].copy_(flat)
    buf[-2] = loss_sum
    buf[-1] = count
    return buf


def unpack_after_allreduce(buf, num_labels):
    return buf[:-2].view(num_labels, num_labels), buf[-2], buf[-1]


class FusedAllReduce:
    """Persistent fused buffer ``[L*L + 2]`` = (transition gradient, loss sum, utterance count).

    ``criterion.transition.grad`` is made a VIEW of the buffer's head, so autograd accumulates the local
    transition gradient straight into it and the single ``all_reduce`` needs no packing or unpacking copies.
    Everything the step touches is a persistent device tensor or view: no Python scalar is ever written into a
    CUDA tensor by indexing (``buf[-1] = 64.0`` is a synchronous pageable host-to-device copy: 160 us per step on
    the B200 boxes, more than the all-reduce itself).
    """

    def __init__(self, criterion, group=None):
        p = criterion.transition
        self.num_labels = p.shape[0]
        self.buf = torch.zeros(p.numel() + 2, dtype=p.dtype, device=p.device)
        self.grad_view = self.buf[:-2].view_as(p)
        self.loss_slot = self.buf[-2:-1]
        self.count_slot = self.buf[-1:]
        self.loss_out = self.buf[-2]
        self.count_out = self.buf[-1]
        self.group = group
        self._seeds = {}

    def prepare(self, criterion):
        self.buf.zero_()
        criterion.transition.grad = self.grad_view

    def seed(self, scale):
        """0-dim device tensor holding ``scale`` (the upstream gradient of the local loss sum), cached."""
        t = self._seeds.get(scale)
        if t is None:
            t = self._seeds[scale] = torch.full((), scale, dtype=self.buf.dtype, device=self.buf.device)
        return t

    def seed_vectors(self, n, scale):
        """Per-utterance upstream gradients ``(+scale, -scale)`` of (FCC score, FAC score), cached."""
        key = (n, scale)
        t = self._seeds.get(key)
        if t is None:
            up = torch.full((n,), scale, dtype=self.buf.dtype, device=self.buf.device)
            t = self._seeds[key] = (up, -up)
        return t

    def reduce(self, loss_sum, count):
        self.loss_slot.copy_(loss_sum.reshape(1))
        self.count_slot.fill_(count)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        return self.loss_out, self.count_out


def global_utterance_count(n_local, device, group=None):
    """Sum of the ranks' local batch sizes.  Shards may be uneven (``shard_bounds`` gives the first ranks one extra), so
    the count is all-reduced, never assumed to be ``n_local * world``: with 'mean' every rank must seed its backward
    with the same ``1/global_batch``.  One small collective and a host read -- pass ``global_batch`` to skip it."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return n_local
    t = torch.tensor([float(n_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(round(float(t.item())))


def _direct_step(criterion, native, inputs, targets, input_lengths, target_lengths, scale, fused):
    """Forward + backward of this rank's shard straight through the native boundary (no autograd graph): the two
    native calls, one loss sum and one copy of the transition gradient into the fused buffer.  The autograd engine
    costs ~0.1 ms of host time per step (a thread hand-off and four graph nodes); with several ranks per box the step
    was bound by host issue time, not by the GPU, so the sharded path skips it.  ``inputs`` gets its gradient the way
    autograd would deliver it: accumulated into ``.grad`` for a leaf, propagated with ``backward`` otherwise."""
    targets, input_lengths, target_lengths = criterion._lengths(inputs, targets, input_lengths, target_lengths)
    T, N, L = inputs.shape
    S = targets.shape[1]
    x = inputs.detach()
    A = criterion.transition.detach()
    full, aligned, g_full, g_aligned, c_full, c_aligned = native.fast_asg_gpu_forward(
        x, targets, A, input_lengths, target_lengths, T, N, L, S)
    local_sum = (full - aligned).sum()
    up, down = fused.seed_vectors(N, scale)
    grad_A, grad_x = native.fast_asg_gpu_backward(up, down, g_full, g_aligned, c_full, c_aligned, targets,
                                                  input_lengths, target_lengths, T, N, L, S)
    fused.grad_view.copy_(grad_A)
    criterion.transition.grad = fused.grad_view
    if inputs.requires_grad:
        if inputs.grad_fn is None:
            inputs.grad = grad_x if inputs.grad is None else inputs.grad + grad_x
        else:
            inputs.backward(grad_x)
    return local_sum


def sharded_asg_step(criterion, inputs, targets, input_lengths=None, target_lengths=None, group=None,
                     global_batch=None, fused=None):
    """One forward+backward of ``criterion`` (an ``ASGLoss`` whose ``reduction`` is honoured GLOBALLY) on
    this rank's utterances, followed by the single fused all-reduce.

    ``inputs`` must be a leaf that requires grad (or feed from a graph that does).  On return:
    ``criterion.transition.grad`` holds the global transition gradient on every rank, ``inputs.grad`` the
    local input gradient (already scaled for the global reduction), and the returned scalar is the global
    reduced loss.  With ``reduction='mean'`` the upstream gradient is ``1/global_batch`` (not ``1/local``).
    Pass a :class:`FusedAllReduce` as ``fused`` to reuse its buffer across steps.
    """
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    n_local = inputs.shape[1]
    reduction = criterion.reduction
    if reduction not in ('mean', 'sum'):
        raise ValueError("sharded_asg_step reduces the loss over ALL ranks' utterances: reduction must be 'mean' or 'sum' "
                         "(per-utterance losses never leave their rank -- call the criterion directly for reduction='none')")
    if global_batch is None:
        global_batch = global_utterance_count(n_local, inputs.device, group)
    if fused is None:
        fused = FusedAllReduce(criterion, group)
    scale = 1.0 / global_batch if reduction == 'mean' else 1.0
    native = getattr(type(criterion), "_native", None)
    direct = (native is not None and hasattr(native, "fast_asg_gpu_backward") and inputs.is_cuda and criterion.training
              and not getattr(criterion, "forward_only", False) and not getattr(criterion, "gpu_no_stream_impl", False)
              and reduction in ('mean', 'sum'))
    if direct:
        local_sum = _direct_step(criterion, native, inputs, targets, input_lengths, target_lengths, scale, fused)
    else:
        fused.prepare(criterion)
        criterion.reduction = 'sum'
        try:
            local_sum = criterion(inputs, targets, input_lengths, target_lengths)
        finally:
            criterion.reduction = reduction
        local_sum.backward(fused.seed(scale))        # seeds the backward with the global scale: no extra graph node
        local_sum = local_sum.detach()
    loss_sum, count = fused.reduce(local_sum, float(n_local))
    if reduction == 'mean':
        return loss_sum / count
    return loss_sum
