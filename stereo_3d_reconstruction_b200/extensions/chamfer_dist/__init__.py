# -*- coding: utf-8 -*-
"""Chamfer distance, forward only -- the B200 replacement of the reference's
`extensions/chamfer_dist` CUDA extension (/root/reference/README.md:62-65: path and
`python setup.py install --user`; the extension's source is on the Stereo2Point branch and not on
disk, so the class names below follow the same author's public GRNet extension from memory
[RECALL] and are a convenience, not a verified API).

No `setup.py install` step: the kernel ships inside libs3d_b200.so (s3d_chamfer_forward).
"""
import torch

from ... import ops


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = ops.chamfer_forward(xyz1.contiguous().float(), xyz2.contiguous().float())
        ctx.mark_non_differentiable(dist1, dist2, idx1, idx2)
        return dist1, dist2, idx1, idx2

    @staticmethod
    def backward(ctx, *grads):  # pragma: no cover
        raise NotImplementedError('inference-only build: chamfer_dist backward is out of scope (SURVEY.md 2, row 12)')


def _apply(xyz1, xyz2):
    if torch.is_grad_enabled() and (xyz1.requires_grad or xyz2.requires_grad):
        # never hand silently gradient-less distances back to a training loop
        raise NotImplementedError('inference-only build: chamfer_dist has no backward; call it under torch.no_grad() or on '
                                  'detached tensors (SURVEY.md 2, row 12)')
    return ChamferFunction.apply(xyz1, xyz2)


class ChamferDistance(torch.nn.Module):
    """forward(xyz1 [B,N,3], xyz2 [B,M,3]) -> scalar mean(dist1) + mean(dist2) (squared L2)."""

    def forward(self, xyz1, xyz2, return_all=False):
        dist1, dist2, idx1, idx2 = _apply(xyz1, xyz2)
        if return_all:
            return dist1, dist2, idx1, idx2
        return torch.mean(dist1) + torch.mean(dist2)


def chamfer_per_sample(xyz1, xyz2):
    """[B] fp64 tensor of mean(dist1[b]) + mean(dist2[b]) -- the per-shard statistic the test driver reduces."""
    dist1, dist2, _, _ = _apply(xyz1, xyz2)
    return dist1.double().mean(1) + dist2.double().mean(1)
