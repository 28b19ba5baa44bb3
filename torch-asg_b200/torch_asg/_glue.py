"""Host-side mirror of the reference's operator surface (``torch_asg/asg.py`` of zh217/torch-asg).

Everything in this file is glue: it owns no arithmetic.  It reproduces, name for name and
argument for argument, what a user of the reference sees --

* ``FAC`` / ``FCC`` -- ``autograd.Function``s over ``(transition, inputs, targets, input_lengths,
  target_lengths)``  (reference ``asg.py:7-34`` and ``:37-55``),
* ``ASGGPUFastForwardOnly`` over ``(inputs, outputs, transition, input_lengths, output_lengths)``
  (``asg.py:58-68``),
* ``ASGGPUFast`` over ``(inputs, transition, outputs, input_lengths, output_lengths)`` returning
  ``(full_scores, aligned_scores)`` (``asg.py:71-97``),
* ``ASGLoss(num_labels, reduction='mean', forward_only=False, gpu_no_stream_impl=False)`` with a
  zero-initialised ``transition`` parameter (``asg.py:100-142``)

-- and forwards, positionally, to the seven functions of the native boundary
(``native/extension.cpp:15-29``).  The classes are produced by :func:`bind_native` so the same
glue can be pointed at the product's CUDA module (``torch_asg_native``, done in ``asg.py``) or,
by the test-suite and the CPU-baseline leg only, at the compiled reference (``oracle/_ref``).
"""
import torch
from torch import nn
from torch.autograd import Function


def _dims(inputs, targets):
    frames, batch, labels = inputs.shape
    return frames, batch, labels, targets.shape[1]


def bind_native(native):
    """Create the autograd surface bound to ``native`` (a module exporting the reference's
    ``fully_connected_*``, ``force_aligned_*`` and -- for CUDA builds -- ``fast_asg_gpu_*``)."""

    class FAC(Function):
        """Force-aligned (numerator) score.  Reference: ``asg.py:7-34``."""

        @staticmethod
        def forward(ctx, transition, inputs, targets, input_lengths, target_lengths):
            T, N, L, S = _dims(inputs, targets)
            scores, alpha, beta, path_contrib = native.force_aligned_forward(
                inputs, targets, transition, input_lengths, target_lengths, T, N, L, S)
            ctx.save_for_backward(alpha, beta, path_contrib, targets, input_lengths,
                                  target_lengths, transition)
            return scores

        @staticmethod
        def backward(ctx, grad_out):
            (alpha, beta, path_contrib, targets, input_lengths, target_lengths,
             transition) = ctx.saved_tensors
            T, N, S = alpha.shape
            L = transition.shape[0]
            grad_transition, grad_inputs = native.force_aligned_backward(
                grad_out, alpha, beta, path_contrib, targets, input_lengths, target_lengths,
                T, N, L, S)
            return grad_transition, grad_inputs, None, None, None, None, None

    class FCC(Function):
        """Fully-connected (denominator) score.  Reference: ``asg.py:37-55``."""

        @staticmethod
        def forward(ctx, transition, inputs, targets, input_lengths, target_lengths):
            T, N, L = inputs.shape
            scores, alpha, beta, path_contrib = native.fully_connected_forward(
                inputs, transition, input_lengths, T, N, L)
            ctx.save_for_backward(alpha, beta, path_contrib)
            return scores

        @staticmethod
        def backward(ctx, grad_out):
            alpha, beta, path_contrib = ctx.saved_tensors
            T, N, L = alpha.shape
            grad_transition, grad_inputs = native.fully_connected_backward(
                grad_out, alpha, beta, path_contrib, T, N, L)
            return grad_transition, grad_inputs, None, None, None, None, None

    class ASGGPUFastForwardOnly(Function):
        """Loss without a backward: only the two beta scans run.  Reference: ``asg.py:58-68``."""

        @staticmethod
        def forward(ctx, inputs, outputs, transition, input_lengths, output_lengths):
            T, N, L, S = _dims(inputs, outputs)
            return native.fast_asg_gpu_forward_only(
                inputs, outputs, transition, input_lengths, output_lengths, T, N, L, S)

        @staticmethod
        def backward(ctx, *grad_outputs):
            return None

    class ASGGPUFast(Function):
        """All four scans in one forward, one fused gradient assembly in the backward.
        Reference: ``asg.py:71-97``."""

        @staticmethod
        def forward(ctx, inputs, transition, outputs, input_lengths, output_lengths):
            T, N, L, S = _dims(inputs, outputs)
            (full_scores, aligned_scores, full_gamma, aligned_gamma, full_path_contrib,
             aligned_path_contrib) = native.fast_asg_gpu_forward(
                inputs, outputs, transition, input_lengths, output_lengths, T, N, L, S)
            ctx.save_for_backward(full_gamma, aligned_gamma, full_path_contrib,
                                  aligned_path_contrib, outputs, input_lengths, output_lengths)
            return full_scores, aligned_scores

        @staticmethod
        def backward(ctx, grad_full, grad_aligned):
            (full_gamma, aligned_gamma, full_path_contrib, aligned_path_contrib, outputs,
             input_lengths, output_lengths) = ctx.saved_tensors
            T, N, L = full_gamma.shape
            S = outputs.shape[1]
            grad_transition, grad_inputs = native.fast_asg_gpu_backward(
                grad_full, grad_aligned, full_gamma, aligned_gamma, full_path_contrib,
                aligned_path_contrib, outputs, input_lengths, output_lengths, T, N, L, S)
            return grad_inputs, grad_transition, None, None, None

    class ASGLoss(nn.Module):
        """Auto Segmentation Criterion.  ``loss[n] = FCC[n] - FAC[n]``, then ``reduction``
        ('mean' over the batch -- the default --, 'sum', anything else: none).
        Reference: ``asg.py:100-142``; option semantics ``README.md:84-94``."""

        _native = native          # the bound native module (torch_asg.sharded calls it directly)

        def __init__(self, num_labels, reduction='mean', forward_only=False,
                     gpu_no_stream_impl=False, validate_targets=False):
            super().__init__()
            # not in the reference: with validate_targets the labels inside target_lengths are checked against
            # [0, num_labels) before the call (one device synchronisation) and an IndexError names the utterance; the
            # reference gathers out of bounds for such labels.  Without it the kernels clamp the label and raise bit 8
            # of the utterance's range flags (asg_range_flags / GraphedASGLoss.check()).
            self.validate_targets = validate_targets
            self.num_labels = num_labels
            self.reduction = reduction
            self.forward_only = forward_only
            self.gpu_no_stream_impl = gpu_no_stream_impl
            self.transition = nn.Parameter(torch.zeros(num_labels, num_labels))

        def _lengths(self, inputs, targets, input_lengths, target_lengths):
            T, N, _, S = _dims(inputs, targets)
            if target_lengths is None:
                target_lengths = targets.new_full((N,), S)
            if input_lengths is None:
                input_lengths = target_lengths.new_full((N,), T)
            if S > T:  # a target cannot be longer than the utterance: keep its first T labels
                targets = targets[:, :T]
                target_lengths = torch.clamp(target_lengths, max=T)
            return targets, input_lengths, target_lengths

        def forward(self, inputs, targets, input_lengths=None, target_lengths=None):
            targets, input_lengths, target_lengths = self._lengths(
                inputs, targets, input_lengths, target_lengths)
            if self.validate_targets:
                S = targets.shape[1]
                live = torch.arange(S, device=targets.device)[None, :] < target_lengths.to(targets.device)[:, None]
                bad = live & ((targets < 0) | (targets >= self.num_labels))
                if bool(bad.any()):
                    n = int(bad.any(dim=1).nonzero()[0])
                    raise IndexError("ASGLoss: utterance %d has a target label outside [0, %d)" % (n, self.num_labels))
            args = (targets, input_lengths, target_lengths)
            if self.gpu_no_stream_impl or not inputs.is_cuda:
                per_utt = (FCC.apply(self.transition, inputs, *args)
                           - FAC.apply(self.transition, inputs, *args))
            elif self.forward_only or not self.training:
                per_utt = ASGGPUFastForwardOnly.apply(inputs, targets, self.transition,
                                                      input_lengths, target_lengths)
            else:
                full, aligned = ASGGPUFast.apply(inputs, self.transition, *args)
                per_utt = full - aligned
            if self.reduction == 'sum':
                return per_utt.sum()
            if self.reduction == 'mean':
                return per_utt.mean()
            return per_utt

    return {"FAC": FAC, "FCC": FCC, "ASGGPUFastForwardOnly": ASGGPUFastForwardOnly,
            "ASGGPUFast": ASGGPUFast, "ASGLoss": ASGLoss}
