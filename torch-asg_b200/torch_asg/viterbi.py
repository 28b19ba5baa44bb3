"""Viterbi forced alignment on the ASG lattice (a TODO of the reference, ``README.md:33``).

``forced_align(inputs, targets, transition, input_lengths=None, target_lengths=None)`` returns, per utterance, the score of the
best monotone alignment of the target to the frames (the force-aligned lattice of ``force_aligned_lattice.cpp:84-111`` in the
tropical semiring of ``doc/tech_report.tex:84-87``) and that alignment itself: ``positions[n,t]`` = index into the target,
``labels[n,t]`` = the label emitted at frame ``t`` (both ``-1`` on padding frames and for utterances whose target does not fit).
"""
import torch

from . import asg as _asg


def forced_align(inputs, targets, transition, input_lengths=None, target_lengths=None):
    native = _asg.torch_asg_native
    T, N, _ = inputs.shape
    S = targets.shape[1]
    if target_lengths is None:
        target_lengths = targets.new_full((N,), S)
    if input_lengths is None:
        input_lengths = target_lengths.new_full((N,), T)
    if S > T:                                   # same truncation as ASGLoss (asg.py:119-122)
        targets = targets[:, :T]
        target_lengths = torch.clamp(target_lengths, max=T)
    scores, positions = native.viterbi_align(inputs.detach(), targets, transition.detach(), input_lengths, target_lengths)
    labels = torch.where(positions >= 0, targets.to(positions.device).gather(1, positions.clamp_min(0)), positions)
    return scores, positions, labels
