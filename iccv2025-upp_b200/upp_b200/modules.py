"""Host-side mirror of the reference's operator-level Python for the hot path: same names,
signatures, return conventions and error behaviour, over the sm_100a kernels.

  fps                         utils/misc.py:13-20
  Group                       models/Point_MAE_unify.py:51-92
  ChamferFunction             extensions/chamfer_dist/__init__.py:13-25
  ChamferDistanceL2/_split/L1 extensions/chamfer_dist/__init__.py:28-84
  knn_points                  pytorch3d.ops.knn_points as called at models/Point_MAE_pretask_dev.py:680
  propagate                   models/Point_MAE_unify.py:22-48
  interpolate_features        the interpolation inside PointNetFeaturePropagation.forward
                              (models/Point_MAE_unify_segment.py:289-313, models/Point_MAE_pretask_dev.py:437-461)
"""
import collections

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
        ctx.set_materialize_grads(False)  # no zero-fill launches for the index output's (undefined) gradient
        return centers, idx

    @staticmethod
    def backward(ctx, grad_centers, _grad_idx):
        if grad_centers is None:
            return None, None
        (idx,) = ctx.saved_tensors
        # (B,M,3) -> scatter-add -> (B,N,3), row-major: no transposes (the reference pays two, utils/misc.py:19)
        return ops.rows_scatter_add(grad_centers.contiguous(), idx, ctx.n), None


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
        # undefined gradients arrive as None instead of zero tensors: autograd otherwise launches one fill kernel per
        # output (neighbourhoods, centres and both index tensors) in front of every backward
        ctx.set_materialize_grads(False)
        return nb, center, idx, cidx

    @staticmethod
    def backward(ctx, g_nb, g_center, _gi, _gc):
        if g_nb is None and g_center is None:
            return None, None, None
        idx, cidx = ctx.saved_tensors
        if g_nb is None:  # only the centres were used downstream
            g_nb = g_center.new_zeros(idx.shape + (3,))
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
        ctx.set_materialize_grads(False)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        if grad_dist1 is None and grad_dist2 is None:
            return None, None
        if grad_dist1 is None:  # only one direction was used downstream (ChamferDistanceL2_split callers)
            grad_dist1 = xyz1.new_zeros(xyz1.shape[:2])
        if grad_dist2 is None:
            grad_dist2 = xyz2.new_zeros(xyz2.shape[:2])
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


class Selection:
    """A kept neighbour selection: (idx, weight, dist) of `k` nearest sources per target for one (xyz1, xyz2) pair.
    Pass it to propagate / interpolate_features (`selection=`) to pay for the selection once when several calls share
    their geometry -- the six SA-unit propagate calls of one forward (models/Point_MAE_pretask_dev.py:298), the feature
    propagation levels of the segmentation head."""
    __slots__ = ("idx", "weight", "dist", "k", "eps")

    def __init__(self, idx, weight, dist, k, eps):
        self.idx, self.weight, self.dist, self.k, self.eps = idx, weight, dist, int(k), float(eps)


def select_neighbors(xyz1, xyz2, k, eps):
    """Selection + inverse-distance weights of propagate / the feature-propagation interpolation alone
    (upp_interp_select_f32) -> Selection.  k is clipped to the number of sources, as the reference's slice does."""
    k = min(int(k), xyz2.shape[1])
    idx, w, d = ops.interp_select(xyz1.detach().contiguous(), xyz2.detach().contiguous(), k, eps)
    return Selection(idx, w, d, k, eps)


class _Interpolate(Function):
    """out = (base or 0) + alpha * sum_j w_j points2[idx_j]; differentiable w.r.t. points2, base and -- through the
    weights, as autograd is through the reference's square_distance -- xyz1 / xyz2.  `sel`: a Selection made for the
    same (xyz1, xyz2, k, eps): the forward is then the blend alone."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, points2, base, k, eps, alpha, sel=None):
        if sel is None:
            out, idx, w, d = ops.interp_forward(xyz1, xyz2, points2, k, eps, base=base, alpha=alpha)
        else:
            if sel.k != k or sel.eps != float(eps) or tuple(sel.idx.shape[:2]) != tuple(xyz1.shape[:2]):
                raise ValueError("selection was made for a different (xyz1, k, eps)")
            idx, w, d = sel.idx, sel.weight, sel.dist
            out = ops.interp_blend(points2, idx, w, base=base, alpha=alpha)
        ctx.save_for_backward(xyz1, xyz2, points2, idx, w, d)
        ctx.eps, ctx.alpha, ctx.has_base = float(eps), float(alpha), base is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        xyz1, xyz2, points2, idx, w, d = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        need_xyz = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        terms = (d, points2, xyz1, xyz2, ctx.eps) if need_xyz else None
        gp2, g1, g2 = ops.interp_backward(grad_out, idx, w, points2.size(1), alpha=ctx.alpha, xyz_terms=terms)
        return (g1 if ctx.needs_input_grad[0] else None, g2 if ctx.needs_input_grad[1] else None,
                gp2 if ctx.needs_input_grad[2] else None,
                grad_out if (ctx.has_base and ctx.needs_input_grad[3]) else None, None, None, None, None)


def interpolate_features(xyz1, xyz2, points2, k, eps=1e-4, selection=None):
    """The interpolation of PointNetFeaturePropagation.forward: xyz1 (B,N,3), xyz2 (B,S,3), points2 (B,S,D)
    -> (B,N,D).  S == 1 repeats the single source row, as the reference does; k is clipped to S like the
    reference's slice `[:, :, :k]`.  selection: a Selection from select_neighbors(xyz1, xyz2, k, eps) to reuse."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if S == 1:
        return points2.repeat(1, N, 1)
    return _Interpolate.apply(xyz1.contiguous(), xyz2.contiguous(), points2.contiguous(), None,
                              min(int(k), S), float(eps), 1.0, selection)


def _propagate_wide(xyz1, xyz2, points1, points2, k, eps):
    """More than 32 neighbours (only the reference's DEFAULT de_neighbors=64 on clouds of more than 32 sources gets here;
    its call sites pass 6 or 8): the reference's formulation on torch's CUDA kernels -- square_distance, full sort,
    first k, inverse-distance weights, gather -- so that the mirror runs whatever its signature advertises."""
    B, N, _ = xyz1.shape
    d = -2 * torch.matmul(xyz1, xyz2.permute(0, 2, 1))
    d = d + torch.sum(xyz1 ** 2, -1).view(B, N, 1) + torch.sum(xyz2 ** 2, -1).view(B, 1, -1)
    d, idx = d.sort(dim=-1)
    d, idx = d[:, :, :k], idx[:, :, :k]
    r = 1.0 / (d + eps)
    w = r / torch.sum(r, dim=2, keepdim=True)
    picked = torch.gather(points2.unsqueeze(1).expand(-1, N, -1, -1), 2, idx.unsqueeze(-1).expand(-1, -1, -1, points2.shape[-1]))
    return points1 + 0.3 * torch.sum(picked * w.view(B, N, k, 1), dim=2)


def propagate(xyz1, xyz2, points1, points2, de_neighbors=64, dist_e=1e-8, selection=None):
    """points1 + 0.3 * (inverse-distance interpolation of points2 over the de_neighbors nearest of xyz2);
    same signature and defaults as the reference function (models/Point_MAE_unify.py:22-48).  de_neighbors is clipped to
    S like the reference's slice.  Up to 32 neighbours (the UPP configs use 3..16) run on the interpolation kernels; more
    -- the bare default on a cloud of more than 32 sources -- on the reference's own torch formulation (GPU).
    selection: a Selection from select_neighbors(xyz1, xyz2, de_neighbors, dist_e) to reuse across calls."""
    S = xyz2.shape[1]
    k = min(int(de_neighbors), S)
    if k > 32:
        if selection is not None:
            raise ValueError("a kept Selection covers at most 32 neighbours")
        return _propagate_wide(xyz1, xyz2, points1, points2, k, float(dist_e))
    return _Interpolate.apply(xyz1.contiguous(), xyz2.contiguous(), points2.contiguous(), points1.contiguous(),
                              k, float(dist_e), 0.3, selection)


_KNN = collections.namedtuple("KNN", "dists idx knn")  # pytorch3d's return type


class _KnnPoints(Function):
    """Squared distances + indices from the kernel; gradients of the distances w.r.t. both clouds
    (d/dp1 = 2 (p1 - p2[idx]), d/dp2[idx] = -2 (p1 - p2[idx])), as pytorch3d provides."""

    @staticmethod
    def forward(ctx, p1, p2, K):
        d, i, _ = ops.knn_points(p1, p2, K)
        ctx.save_for_backward(p1, p2, i)
        ctx.mark_non_differentiable(i)
        ctx.set_materialize_grads(False)
        return d, i

    @staticmethod
    def backward(ctx, grad_d, _grad_i):
        if grad_d is None:
            return None, None, None
        p1, p2, i = ctx.saved_tensors
        B, N1, K = i.shape
        flat = i.reshape(B, N1 * K, 1).expand(-1, -1, 3)
        diff = p1.unsqueeze(2) - torch.gather(p2, 1, flat).view(B, N1, K, 3)
        v = 2.0 * grad_d.unsqueeze(-1) * diff
        g2 = torch.zeros_like(p2).scatter_add_(1, flat, -v.reshape(B, N1 * K, 3))
        return v.sum(2), g2, None


def knn_points(p1, p2, lengths1=None, lengths2=None, norm=2, K=1, version=-1, return_nn=False, return_sorted=True):
    """Drop-in for pytorch3d.ops.knn_points on equal-length 3-D clouds (the UPP call site:
    `knn_points(noise, partial, K=4, return_nn=True)`): -> KNN(dists (B,N1,K) squared ascending,
    idx (B,N1,K) int64, knn (B,N1,K,3) or None).  Heterogeneous lengths / L1 norm are outside the path."""
    if lengths1 is not None or lengths2 is not None:
        raise NotImplementedError("upp_b200.knn_points covers equal-length clouds (lengths1/lengths2 = None)")
    if norm != 2:
        raise NotImplementedError("upp_b200.knn_points covers norm=2 (the UPP call site)")
    p1, p2 = p1.contiguous(), p2.contiguous()
    if not (p1.requires_grad or p2.requires_grad):
        d, i, nn_ = ops.knn_points(p1, p2, K, want_nn=return_nn)
        return _KNN(dists=d, idx=i, knn=nn_)
    d, i = _KnnPoints.apply(p1, p2, int(K))
    nn_ = None
    if return_nn:  # knn_gather: differentiable w.r.t. p2
        B, N1, k = i.shape
        nn_ = torch.gather(p2, 1, i.reshape(B, N1 * k, 1).expand(-1, -1, 3)).view(B, N1, k, 3)
    return _KNN(dists=d, idx=i, knn=nn_)
