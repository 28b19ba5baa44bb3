"""Target preprocessing and length scaling -- what the reference says it leaves to the user
(``README.md:93-94``: "we have also omitted scaling based on input/output lengths"; replabels: ``doc/tech_report.tex:126``).

* ``pack_replabels`` / ``unpack_replabels``: ASG transcripts cannot repeat a label, so runs are written with "repeat" tokens;
  wav2letter++ convention (flashlight ``packReplabels``): token ``first_rep_label + k - 1`` (1 <= k <= max_reps) = "the previous
  label k MORE times", a run longer than ``max_reps + 1`` starts over.  Runs on the device (C ABI ``asg_pack_replabels`` /
  ``asg_unpack_replabels``), int64 in, int64 out.
* ``length_scale`` / ``ScaledASGLoss``: wav2letter's criterion scale modes (none, input_sz, input_sz_sqrt, target_sz,
  target_sz_sqrt) applied per utterance before the reduction -- exactly the recipe the reference's README gives ("use the None
  reduction and scale the individual scores").
"""
import torch
from torch import nn

from . import asg as _asg

SCALE_MODES = ("none", "input_sz", "input_sz_sqrt", "target_sz", "target_sz_sqrt")


def pack_replabels(targets, target_lengths, first_rep_label, max_reps, fill=0):
    """(packed [N,S] int64, lengths [N] int64).  ``targets`` must be a CUDA int64 tensor; padding beyond the new length = fill."""
    if target_lengths is None:
        target_lengths = targets.new_full((targets.shape[0],), targets.shape[1])
    return _asg.torch_asg_native.replabels(targets, target_lengths, int(first_rep_label), int(max_reps), int(fill), True, 0)


def unpack_replabels(targets, target_lengths, first_rep_label, max_reps, max_length, fill=0):
    """The inverse: (unpacked [N,max_length], lengths [N]); a length above ``max_length`` means the row was truncated."""
    if target_lengths is None:
        target_lengths = targets.new_full((targets.shape[0],), targets.shape[1])
    return _asg.torch_asg_native.replabels(targets, target_lengths, int(first_rep_label), int(max_reps), int(fill), False, int(max_length))


def length_scale(mode, input_lengths, target_lengths, dtype=torch.float32):
    """Per-utterance factor of wav2letter's CriterionScaleMode: 1, 1/T_n, sqrt(1/T_n), 1/S_n, sqrt(1/S_n) (1 where the size is 0)."""
    if mode not in SCALE_MODES:
        raise ValueError("mode must be one of %s" % (SCALE_MODES,))
    if mode == "none":
        return torch.ones(input_lengths.shape, dtype=dtype, device=input_lengths.device)
    size = (input_lengths if mode.startswith("input") else target_lengths).to(dtype)
    inv = torch.where(size > 0, 1.0 / size.clamp(min=1), torch.ones_like(size))
    return inv.sqrt() if mode.endswith("sqrt") else inv


class ScaledASGLoss(nn.Module):
    """``ASGLoss`` with wav2letter's length scaling: ``reduce(scale(T_n, S_n) * loss[n])``.  ``criterion`` is an ``ASGLoss`` (its own
    reduction is ignored: the per-utterance losses are taken, scaled, then reduced with ``reduction``)."""

    def __init__(self, criterion, mode="target_sz_sqrt", reduction="mean"):
        super().__init__()
        if mode not in SCALE_MODES:
            raise ValueError("mode must be one of %s" % (SCALE_MODES,))
        self.criterion, self.mode, self.reduction = criterion, mode, reduction

    def forward(self, inputs, targets, input_lengths=None, target_lengths=None):
        T, N = inputs.shape[0], inputs.shape[1]
        if target_lengths is None:
            target_lengths = targets.new_full((N,), targets.shape[1])
        if input_lengths is None:
            input_lengths = target_lengths.new_full((N,), T)
        keep = self.criterion.reduction
        self.criterion.reduction = "none"
        try:
            per_utt = self.criterion(inputs, targets, input_lengths, target_lengths)
        finally:
            self.criterion.reduction = keep
        out = per_utt * length_scale(self.mode, input_lengths.to(per_utt.device), target_lengths.to(per_utt.device), per_utt.dtype)
        if self.reduction == "mean":
            return out.mean()
        if self.reduction == "sum":
            return out.sum()
        return out
