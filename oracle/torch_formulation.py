"""The reference's PURE-TORCH formulation of the hot path, run on host cores.

TEST / BASELINE INFRASTRUCTURE ONLY (bench.py cpu_baseline + --impl reference, tests).
This is the CPU comparison line BASELINE.json's north_star names: the reference ships
no CPU implementation of its CUDA ops, but it does carry torch formulations of the
same geometry which run anywhere:

  square_distance   models/modules.py:13-32 / models/dgcnn_group.py:21-40
                    (expanded -2ab + a^2 + b^2 form; less accurate than the CUDA ops)
  knn_point         models/dgcnn_group.py:8-19  (topk, largest=False, sorted=False)
  index_points      models/modules.py:35-51
  farthest sampling datasets/ModelNetDataset.py:29-50 (numpy, random start; restated
                    batched in torch with start index 0 to match the CUDA op)
  Group.forward     models/Point_MAE_unify.py:58-92
  interpolate       models/Point_MAE_unify_segment.py:289-313 / models/Point_MAE_unify.py:22-48
                    (square_distance -> full sort -> first k -> inverse-distance weights -> index_points)
  Chamfer L1/L2     extensions/chamfer_dist/__init__.py:28-84 reductions over a
                    cdist-based nearest-neighbour distance

It is used for TIMING and loose checks only; bit-level parity uses oracle/upp_oracle.c.
"""
import torch


def square_distance(src, dst):
    """(B,N,C),(B,M,C) -> (B,N,M) squared distances, expanded form."""
    inner = torch.matmul(src, dst.transpose(1, 2))
    d = inner.mul_(-2.0)
    d += (src * src).sum(-1)[:, :, None]
    d += (dst * dst).sum(-1)[:, None, :]
    return d


def knn_point(nsample, xyz, new_xyz):
    """indices (B,S,nsample) of the nsample nearest xyz points per new_xyz point (unsorted)."""
    d = square_distance(new_xyz, xyz)
    return torch.topk(d, nsample, dim=-1, largest=False, sorted=False)[1]


def index_points(points, idx):
    """points (B,N,C), idx (B,...) -> (B,...,C)."""
    B = points.shape[0]
    shape = [B] + [1] * (idx.dim() - 1)
    batch = torch.arange(B, device=points.device).view(shape).expand_as(idx)
    return points[batch, idx, :]


def farthest_point_sample(xyz, npoint):
    """Batched torch FPS, start index 0, direct-difference distances. (B,N,3) -> (B,npoint) int64."""
    B, N, _ = xyz.shape
    out = torch.zeros(B, npoint, dtype=torch.long, device=xyz.device)
    mind = torch.full((B, N), 1e10, dtype=xyz.dtype, device=xyz.device)
    far = torch.zeros(B, dtype=torch.long, device=xyz.device)
    rows = torch.arange(B, device=xyz.device)
    for i in range(npoint):
        out[:, i] = far
        c = xyz[rows, far, :].unsqueeze(1)
        d = ((xyz - c) ** 2).sum(-1)
        mind = torch.minimum(mind, d)
        far = mind.argmax(-1)
    return out


def fps(data, number):
    """utils/misc.py:13-20 semantics on the torch formulation."""
    idx = farthest_point_sample(data, number)
    return index_points(data, idx), idx


def group(xyz, num_group, group_size):
    """Group.forward semantics: (neighborhood (B,G,k,3), center (B,G,3))."""
    center, _ = fps(xyz, num_group)
    idx = knn_point(group_size, xyz, center)
    nb = index_points(xyz, idx)
    return nb - center.unsqueeze(2), center


def chamfer_sq(xyz1, xyz2):
    """cdist-based nearest-neighbour squared distances, differentiable."""
    d = torch.cdist(xyz1, xyz2) ** 2
    return d.min(dim=2)[0], d.min(dim=1)[0]


def chamfer_l2(xyz1, xyz2):
    d1, d2 = chamfer_sq(xyz1, xyz2)
    return d1.mean() + d2.mean()


def chamfer_l1(xyz1, xyz2):
    d1, d2 = chamfer_sq(xyz1, xyz2)
    return (d1.sqrt().mean() + d2.sqrt().mean()) / 2


def interpolate(xyz1, xyz2, points2, k, eps):
    """PointNetFeaturePropagation / propagate interpolation: (B,N,3),(B,S,3),(B,S,C) -> (B,N,C)."""
    B, N, _ = xyz1.shape
    d, idx = square_distance(xyz1, xyz2).sort(dim=-1)
    d, idx = d[:, :, :k], idx[:, :, :k]
    r = 1.0 / (d + eps)
    w = r / r.sum(dim=2, keepdim=True)
    return (index_points(points2, idx) * w.view(B, N, k, 1)).sum(dim=2)
