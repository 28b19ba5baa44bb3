"""``ASGLossFromLogits`` -- ``log_softmax`` fused into the ASG criterion.

The reference's users normalise the network's activations before the loss (``test/test_asg.py:111-113``:
``asg_loss(inputs.log_softmax(-1), ...)``).  For ASG that pass is pure overhead:

* ``loss(log_softmax(z)) == loss(z)``: the fully-connected score is the normaliser of the criterion, so both scores move by the
  same ``sum_t logsumexp(z[t])`` and their difference does not change.  The kernels subtract every frame's maximum themselves.
* ``d loss / d z = gx - softmax(z) * sum_l gx`` and ``sum_l gx[t,n,:]`` is exactly zero for every utterance with a valid
  alignment (the two posteriors cancel).  Only an infeasible utterance (target longer than its input: loss ``+inf``, FAC
  gradients exactly 0) needs the softmax term; ``asg_logits_backward_fix`` adds it in one tiny launch.

So this module calls the same native entry points as ``ASGLoss`` on the RAW activations and saves, per step, the forward read +
write and the backward two reads + write of ``[T,N,L]`` that ``log_softmax`` costs (51 MB of HBM traffic at T=1000, N=64, L=40).
Results are those of ``ASGLoss(...)(z.log_softmax(-1), ...)`` (asserted against the oracle in tests/test_gpu_logits_viterbi.py).
"""
import torch
from torch import nn
from torch.autograd import Function

from .asg import ASGLoss


class _FromLogits(Function):
    @staticmethod
    def forward(ctx, z, transition, native, targets, input_lengths, target_lengths):
        T, N, L = z.shape
        S = targets.shape[1]
        full, aligned, g_full, g_aligned, c_full, c_aligned = native.fast_asg_gpu_forward(
            z, targets, transition, input_lengths, target_lengths, T, N, L, S)
        ctx.save_for_backward(g_full, g_aligned, c_full, c_aligned, targets, input_lengths, target_lengths, z, aligned)
        ctx.native = native
        return full - aligned

    @staticmethod
    def backward(ctx, grad_out):
        g_full, g_aligned, c_full, c_aligned, targets, il, tl, z, aligned = ctx.saved_tensors
        T, N, L = g_full.shape
        S = targets.shape[1]
        g = grad_out.contiguous()
        grad_A, grad_z = ctx.native.fast_asg_gpu_backward(g, -g, g_full, g_aligned, c_full, c_aligned, targets, il, tl, T, N, L, S)
        ctx.native.logits_backward_fix(z, aligned, g, 1.0, il, grad_z)
        return grad_z, grad_A, None, None, None, None


class ASGLossFromLogits(nn.Module):
    """``ASGLoss`` for raw activations: same constructor, same call, the results of ``ASGLoss(...)(z.log_softmax(-1), ...)``."""

    def __init__(self, num_labels, reduction='mean'):
        super().__init__()
        self.criterion = ASGLoss(num_labels, reduction=reduction)

    @property
    def transition(self):
        return self.criterion.transition

    def forward(self, inputs, targets, input_lengths=None, target_lengths=None):
        crit = self.criterion
        targets, input_lengths, target_lengths = crit._lengths(inputs, targets, input_lengths, target_lengths)
        if not inputs.is_cuda:
            raise RuntimeError("ASGLossFromLogits: CUDA tensors only (no CPU path)")
        per_utt = _FromLogits.apply(inputs, crit.transition, type(crit)._native, targets, input_lengths, target_lengths)
        if crit.reduction == 'sum':
            return per_utt.sum()
        if crit.reduction == 'mean':
            return per_utt.mean()
        return per_utt
