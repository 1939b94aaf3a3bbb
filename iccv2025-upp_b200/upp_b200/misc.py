"""Host-side mirror of the reference's utils/misc.py helpers that sit on the FPS path.

  fps                   utils/misc.py:13-20      (re-exported from modules)
  seprate_point_cloud   utils/misc.py:205-256    
  random_dropping       utils/misc.py:308-315    (KITTI fine-tuning: FPS to a random size, zero-padded to 2048)

The reference's seprate_point_cloud walks the batch in a Python loop and, per cloud, sorts the points by
distance to a random viewpoint, crops, and calls fps() on (1, n - num_crop, 3) and (1, num_crop, 3): per training step
B x (norm + argsort + gathers) and 2*B one-CTA FPS launches (tools/runner_module.py:131, tools/runner_pretask.py:179,
tools/runner_unify_seg.py:212).  num_crop is drawn ONCE per call, so every cloud of the batch has the same two
lengths: here the whole batch is ONE crop launch (upp_crop_split_f32: distances, in-shared-memory sort, split, both
gathers) plus two batched FPS launches.  The random viewpoints are drawn exactly as the reference draws them (one
torch.randn(1,1,3) per cloud from the CPU generator, random.sample for a list of fixed points), so a seeded run selects
the same crops; they reach the device in one copy.
"""
import random

import torch
import torch.nn.functional as F

from . import ops
from .modules import fps

__all__ = ["fps", "seprate_point_cloud", "random_dropping"]


def _draw_viewpoints(batch, fixed_points=None):
    """The reference's per-cloud viewpoint draws (utils/misc.py:224-231), in its order, consuming the CPU torch generator
    and Python's `random` exactly as its loop does -> (batch, 3) float32 on the CPU.  Random viewpoints are normalised
    in ONE F.normalize over the stacked draws (bit-identical to normalising each draw on its own: the norm is a
    three-term sum per row either way; tests/test_host.py checks it)."""
    draws = []
    for _ in range(batch):
        if fixed_points is None:
            draws.append(torch.randn(1, 1, 3))
        else:
            if isinstance(fixed_points, list):
                fixed_point = random.sample(fixed_points, 1)[0]
            else:
                fixed_point = fixed_points
            draws.append(fixed_point.reshape(1, 1, 3).to("cpu", torch.float32))
    centers = torch.cat(draws, 0)
    if fixed_points is None:
        centers = F.normalize(centers, p=2, dim=-1)
    return centers.reshape(batch, 3)


def seprate_point_cloud(xyz, num_points, crop, fixed_points=None, padding_zeros=False, sample_points=1024,
                        incomplete_shape=True):
    """Same signature, return convention and RNG consumption as the reference function:
    -> (input_data (B, n_in, 3), crop_data (B, n_crop, 3)), both contiguous; (xyz, None) when crop == num_points.
    xyz must be a CUDA tensor (the reference moves its viewpoints with .cuda(); there is no CPU path here)."""
    _, n, c = xyz.shape
    assert n == num_points
    assert c == 3
    if crop == num_points:
        return xyz, None
    if isinstance(crop, list):
        num_crop = random.randint(crop[0], crop[1])
    else:
        num_crop = crop
    B = xyz.shape[0]
    centers = _draw_viewpoints(B, fixed_points).to(xyz.device, non_blocking=True)  # ONE host-to-device copy
    pts = xyz.contiguous().float()
    if n <= 8192:
        input_data, crop_data = ops.crop_split(pts, centers, num_crop, padding_zeros=padding_zeros)
    else:  # beyond the crop kernel's shared-memory sort: the same computation on torch's CUDA kernels
        dist = torch.norm(centers.view(B, 1, 1, 3) - pts.unsqueeze(1), p=2, dim=-1)
        idx = torch.argsort(dist, dim=-1, descending=False, stable=True)[:, 0]
        take = lambda ix: torch.gather(pts, 1, ix.unsqueeze(-1).expand(-1, -1, 3))  # noqa: E731
        crop_data = take(idx[:, :num_crop])
        if padding_zeros:
            input_data = pts.clone()
            input_data.scatter_(1, idx[:, :num_crop].unsqueeze(-1).expand(-1, -1, 3), crop_data * 0)
        else:
            input_data = take(idx[:, num_crop:])
    if isinstance(crop, list):
        input_data = fps(input_data, sample_points)[0]
        crop_data = fps(crop_data, sample_points)[0]
    else:
        if incomplete_shape and input_data.shape[1] > sample_points:
            input_data = fps(input_data, sample_points)[0]
        if incomplete_shape and crop_data.shape[1] > sample_points:
            crop_data = fps(crop_data, sample_points)[0]
    return input_data.contiguous(), crop_data.contiguous()


def random_dropping(pc, e):
    """utils/misc.py:308-315, same RNG draw (one torch.randint from the CPU generator): FPS down to a random number of
    points in [1, max(64, 768 // (e // 50 + 1))), zero rows appended up to 2048 points (the rows ChamferDistance's
    ignore_zeros drops again).  As written the reference passes fps()'s (data, idx) TUPLE on to .size() and raises
    (its fps, utils/misc.py:13-20, returns both; the function predates that); the evident intent -- the sampled
    points -- is what this mirror implements."""
    up_num = max(64, 768 // (e // 50 + 1))
    random_num = int(torch.randint(1, up_num, (1, 1))[0, 0])
    pc = fps(pc, random_num)[0]
    padding = torch.zeros(pc.size(0), 2048 - pc.size(1), 3, device=pc.device, dtype=pc.dtype)
    return torch.cat([pc, padding], dim=1)
