"""Host-side mirror of the geometry metrics of the reference's utils/metrics.py, on the
Chamfer kernels:

  f_score              Metrics._get_f_score            utils/metrics.py:70-92  (open3d nearest-neighbour distances on CPU,
                                                       one cloud at a time) -> one batched Chamfer forward
  chamfer_distance_l1  Metrics._get_chamfer_distancel1 utils/metrics.py:103-106 (ChamferDistanceL1(ignore_zeros=True) x 1000)
  chamfer_distance_l2  Metrics._get_chamfer_distancel2 utils/metrics.py:108-111 (ChamferDistanceL2(ignore_zeros=True) x 1000)
"""
import torch

from . import ops
from .modules import ChamferDistanceL1, ChamferDistanceL2

_CD_L1 = ChamferDistanceL1(ignore_zeros=True)
_CD_L2 = ChamferDistanceL2(ignore_zeros=True)


def f_score(pred, gt, th=0.01):
    """F-Score at distance threshold `th`: per cloud, precision = share of pred points whose nearest gt point is
    closer than th, recall = the same from gt to pred, F = 2PR/(P+R) (0 when both are 0); mean over the batch --
    exactly the reference's recursion over single clouds.  pred (B,N,3), gt (B,M,3) CUDA f32 -> 0-dim tensor."""
    assert pred.size(0) == gt.size(0)
    d1, d2, _, _ = ops.chamfer_forward(pred.contiguous(), gt.contiguous())
    precision = (torch.sqrt(d1) < th).float().mean(dim=1)
    recall = (torch.sqrt(d2) < th).float().mean(dim=1)
    denom = recall + precision
    f = torch.where(denom > 0, 2 * recall * precision / denom.clamp_min(1e-30), torch.zeros_like(denom))
    return f.mean()


def chamfer_distance_l1(pred, gt):
    return _CD_L1(pred, gt) * 1000


def chamfer_distance_l2(pred, gt):
    return _CD_L2(pred, gt) * 1000
