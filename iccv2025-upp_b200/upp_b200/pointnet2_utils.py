"""Mirror of the two pointnet2_ops.pointnet2_utils symbols the UPP hot path calls
(reference utils/misc.py:10,18-19; tools/runner_module.py:151-153,450-455).

Same names, argument meaning and autograd conventions as upstream pointnet2_ops 3.0.0:
  furthest_point_sample(xyz, npoint) -> (B, npoint) int32, non-differentiable
  gather_operation(features, idx)    -> (B, C, npoint), differentiable w.r.t. features
"""
import torch
from torch.autograd import Function

from . import ops


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        out = ops.fps(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n = features.size(2)
        return ops.gather(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return ops.gather_grad(grad_out.contiguous(), idx, ctx.n), None


gather_operation = GatherOperation.apply
