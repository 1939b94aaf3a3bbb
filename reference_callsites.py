"""The reference's CALL SITES for the hot path, restated over the reference's own imports.

TEST / BENCH INFRASTRUCTURE (bench.py's `e2e_dropin` arm, tests/): this is what a user of the reference runs when
nothing but the three third-party modules is swapped -- `from pointnet2_ops import pointnet2_utils`,
`from knn_cuda import KNN`, `import chamfer` resolve to iccv2025-upp_b200/dropin/ -- and every line of the
reference's Python around them stays as it is:

  fps                     utils/misc.py:13-20            furthest_point_sample -> transpose -> gather_operation -> transpose
  Group                   models/Point_MAE_unify.py:51-92  misc.fps -> KNN -> flat / batched index gather -> centre subtraction
  ChamferFunction         extensions/chamfer_dist/__init__.py:13-25
  ChamferDistanceL1 / L2  extensions/chamfer_dist/__init__.py:28-44, 64-84

A restatement, not a copy: tests/test_oracle.py::test_callsite_restatement_equals_live_reference runs the reference's
real source (lifted by AST from /root/reference, build container only) and these classes over the same recording
stand-ins and requires the same sequence of third-party calls with the same arguments and the same results.
"""
import torch
import torch.nn as nn

import chamfer                                   # extensions/chamfer_dist/__init__.py:10
from knn_cuda import KNN                         # models/Point_MAE_unify.py:16
from pointnet2_ops import pointnet2_utils        # utils/misc.py:10


def fps(data, number):
    """data (B,N,3) -> (sampled points (B,number,3) contiguous, indices (B,number) int32); utils/misc.py:13-20."""
    picked = pointnet2_utils.furthest_point_sample(data, number)
    channel_first = data.transpose(1, 2).contiguous()
    return pointnet2_utils.gather_operation(channel_first, picked).transpose(1, 2).contiguous(), picked


class Group(nn.Module):
    """models/Point_MAE_unify.py:51-92 -- same constructor, same forward signature and return conventions."""

    def __init__(self, num_group, group_size):
        super().__init__()
        self.num_group, self.group_size = num_group, group_size
        self.knn = KNN(k=group_size, transpose_mode=True)

    def forward(self, xyz, require_index=False, gather_idx=False):
        B, N, _ = xyz.shape
        G, k = self.num_group, self.group_size
        center, center_idx = fps(xyz, G)
        _, idx = self.knn(xyz, center)
        assert idx.size(1) == G
        assert idx.size(2) == k
        if gather_idx:
            nb = torch.gather(xyz, 1, idx.reshape(B, -1, 1).expand(-1, -1, 3))
            center_idx = center_idx.long()
        else:
            first = torch.arange(0, B, device=xyz.device) * N          # flat row of each cloud's point 0
            idx = (idx + first.view(-1, 1, 1)).view(-1)
            center_idx = (center_idx + first.view(-1, 1)).view(-1)
            nb = xyz.view(B * N, -1)[idx, :]
        nb = nb.view(B, G, k, 3).contiguous() - center.unsqueeze(2)
        return (nb, center, idx, center_idx) if require_index else (nb, center)


class ChamferFunction(torch.autograd.Function):
    """extensions/chamfer_dist/__init__.py:13-25 over the top-level `chamfer` module."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, dist2, idx1, idx2 = chamfer.forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        g1, g2 = chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
        return g1, g2


def _maybe_drop_zero_rows(xyz1, xyz2, ignore_zeros):
    if xyz1.size(0) == 1 and ignore_zeros:  # only at batch size 1, on the coordinate SUM (__init__.py:36-41)
        xyz1 = xyz1[torch.sum(xyz1, dim=2).ne(0)].unsqueeze(dim=0)
        xyz2 = xyz2[torch.sum(xyz2, dim=2).ne(0)].unsqueeze(dim=0)
    return xyz1, xyz2


class ChamferDistanceL2(nn.Module):
    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        d1, d2 = ChamferFunction.apply(*_maybe_drop_zero_rows(xyz1, xyz2, self.ignore_zeros))
        return torch.mean(d1) + torch.mean(d2)


class ChamferDistanceL1(nn.Module):
    def __init__(self, ignore_zeros=False):
        super().__init__()
        self.ignore_zeros = ignore_zeros

    def forward(self, xyz1, xyz2):
        d1, d2 = ChamferFunction.apply(*_maybe_drop_zero_rows(xyz1, xyz2, self.ignore_zeros))
        return (torch.mean(torch.sqrt(d1)) + torch.mean(torch.sqrt(d2))) / 2
