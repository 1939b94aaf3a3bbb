"""Host-side mirror of the reference's operator-level Python for the hot path: same names,
signatures, return conventions and error behaviour, over the sm_100a kernels.

  fps                         utils/misc.py:13-20
  Group                       models/Point_MAE_unify.py:51-92
  ChamferFunction             extensions/chamfer_dist/__init__.py:13-25
  ChamferDistanceL2/_split/L1 extensions/chamfer_dist/__init__.py:28-84
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops
from .knn import KNN


class _FpsGather(Function):
    """FPS + coordinate gather in one kernel; differentiable w.r.t. data through the gather."""

    @staticmethod
    def forward(ctx, data, number):
        idx, centers = ops.fps(data, number, want_centers=True)
        ctx.save_for_backward(idx)
        ctx.n = data.size(1)
        ctx.mark_non_differentiable(idx)
        return centers, idx

    @staticmethod
    def backward(ctx, grad_centers, _grad_idx):
        (idx,) = ctx.saved_tensors
        # (B,M,3) -> channel-first scatter-add -> (B,N,3)
        g = ops.gather_grad(grad_centers.transpose(1, 2).contiguous(), idx, ctx.n)
        return g.transpose(1, 2).contiguous(), None


def fps(data, number):
    """data (B,N,3) -> (fps_data (B,number,3) contiguous, fps_idx (B,number) int32)."""
    return _FpsGather.apply(data.contiguous(), number)


class _FusedGroup(Function):
    @staticmethod
    def forward(ctx, xyz, num_group, group_size):
        nb, center, idx, cidx = ops.group(xyz, num_group, group_size)
        ctx.save_for_backward(idx, cidx)
        ctx.n = xyz.size(1)
        ctx.mark_non_differentiable(idx, cidx)
        return nb, center, idx, cidx

    @staticmethod
    def backward(ctx, g_nb, g_center, _gi, _gc):
        idx, cidx = ctx.saved_tensors
        g_center = g_center.contiguous() if g_center is not None else None
        return ops.group_backward(g_nb.contiguous(), g_center, idx, cidx, ctx.n), None, None


class Group(nn.Module):
    """FPS + kNN patch divider.

    forward(xyz (B,N,3), require_index=False, gather_idx=False) ->
        neighborhood (B,G,k,3) centre-subtracted, center (B,G,3)
        [+ idx, center_idx when require_index: flat (B*G*k,) / (B*G,) with a b*N base when
           gather_idx=False, (B,G,k) int64 / (B,G) int64 when gather_idx=True]
    fused=True (default) runs the single C-ABI call upp_group_f32; fused=False composes
    fps -> KNN -> gather exactly as the reference's Python does.
    """

    def __init__(self, num_group, group_size, fused=True):
        super().__init__()
        self.num_group = num_group
        self.group_size = group_size
        self.fused = fused
        self.knn = KNN(k=self.group_size, transpose_mode=True)

    def forward(self, xyz, require_index=False, gather_idx=False):
        batch_size, num_points, _ = xyz.shape
        if self.fused:
            neighborhood, center, idx, center_idx = _FusedGroup.apply(
                xyz.contiguous(), self.num_group, self.group_size)
        else:
            center, center_idx = fps(xyz, self.num_group)
            _, idx = self.knn(xyz, center)
            assert idx.size(1) == self.num_group
            assert idx.size(2) == self.group_size
            picked = torch.gather(xyz, 1, idx.reshape(batch_size, -1, 1).expand(-1, -1, 3))
            neighborhood = picked.view(batch_size, self.num_group, self.group_size, 3) - center.unsqueeze(2)
        if not require_index:
            return neighborhood, center
        if gather_idx:
            return neighborhood, center, idx, center_idx.long()
        base = torch.arange(0, batch_size, device=xyz.device) * num_points
        return (neighborhood, center, (idx + base.view(-1, 1, 1)).view(-1),
                (center_idx + base.view(-1, 1)).view(-1))


class ChamferFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = ops.chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        grad_xyz1, grad_xyz2 = ops.chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
        return grad_xyz1, grad_xyz2


def _drop_zero_points(xyz1, xyz2):
    """ignore_zeros rule (only applied at batch size 1): drop points whose coordinate SUM is 0."""
    keep1 = torch.sum(xyz1, dim=2).ne(0)
    keep2 = torch.sum(xyz2, dim=2).ne(0)
    return xyz1[keep1].unsqueeze(dim=0), xyz2[keep2].unsqueeze(dim=0)


class _ChamferBase(nn.Module):
    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def _dists(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1, xyz2 = _drop_zero_points(xyz1, xyz2)
        return ChamferFunction.apply(xyz1.contiguous(), xyz2.contiguous())


class ChamferDistanceL2(_ChamferBase):
    def forward(self, xyz1, xyz2):
        dist1, dist2 = self._dists(xyz1, xyz2)
        return torch.mean(dist1) + torch.mean(dist2)


class ChamferDistanceL2_split(_ChamferBase):
    def forward(self, xyz1, xyz2):
        dist1, dist2 = self._dists(xyz1, xyz2)
        return torch.mean(dist1), torch.mean(dist2)


class ChamferDistanceL1(_ChamferBase):
    def forward(self, xyz1, xyz2):
        dist1, dist2 = self._dists(xyz1, xyz2)
        return (torch.mean(torch.sqrt(dist1)) + torch.mean(torch.sqrt(dist2))) / 2
