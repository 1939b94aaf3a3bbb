"""Multi-GPU plumbing: one process per GPU, the batch of clouds sharded across ranks.

Clouds are independent, so FPS / kNN / Group / Chamfer run with NO data-path
collective.  The only exchange is the scalar Chamfer loss: each rank's forward kernel leaves
{sum d1, sum d2, sum sqrt d1, sum sqrt d2} in a 4-float buffer which is all-reduced (SUM) once,
on the compute stream, and divided by the GLOBAL element counts -- the batch-sharded equivalent
of torch.mean over the whole batch followed by dist_utils.reduce_tensor
(reference utils/dist_utils.py:41-48, tools/runner_pretask.py:241).
"""
import torch
import torch.distributed as dist
from torch.autograd import Function

from . import ops


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(batch, rank, world_size):
    """Contiguous shard [lo, hi) of `batch` clouds for `rank`; remainders go to the low ranks
    (the reference's DistributedSampler split, tools/builder.py:17-23, without padding)."""
    base, rem = divmod(int(batch), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, rank=None, world_size=None):
    """Slice dim 0 of a (B, ...) tensor to this rank's clouds."""
    if rank is None or world_size is None:
        rank, world_size = world()
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def reduce_sums(sums, group=None):
    """SUM all-reduce of the 4-float partial-sum buffer (NCCL on GPU tensors, gloo on CPU)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def chamfer_loss_from_sums(sums, n1_global, n2_global, kind):
    """Loss value from globally reduced sums; kind in {'l2', 'l1', 'l2_split'}."""
    if kind == "l2":
        return sums[0] / n1_global + sums[1] / n2_global
    if kind == "l2_split":
        return sums[0] / n1_global, sums[1] / n2_global
    if kind == "l1":
        return (sums[2] / n1_global + sums[3] / n2_global) / 2
    raise ValueError(kind)


class _ShardedChamfer(Function):
    """Chamfer loss over a batch sharded across ranks.  Forward: local kernel + one all-reduce of
    4 floats.  Backward: no collective -- grad_dist is the constant 1/(B_global*N) (times the
    sqrt chain for L1), so each rank's coordinate gradients are exactly the rows the unsharded
    computation would produce for its clouds."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, kind, n_global_clouds, group):
        d1, d2, i1, i2, sums = ops.chamfer_forward(xyz1, xyz2, want_sums=True)
        reduce_sums(sums, group)
        n1 = float(n_global_clouds * xyz1.size(1))
        n2 = float(n_global_clouds * xyz2.size(1))
        ctx.save_for_backward(xyz1, xyz2, i1, i2, d1, d2)
        ctx.kind, ctx.n1, ctx.n2 = kind, n1, n2
        return chamfer_loss_from_sums(sums, n1, n2, kind)

    @staticmethod
    def backward(ctx, grad_loss):
        xyz1, xyz2, i1, i2, d1, d2 = ctx.saved_tensors
        if ctx.kind == "l2":
            g1 = (grad_loss / ctx.n1).expand_as(d1)
            g2 = (grad_loss / ctx.n2).expand_as(d2)
        else:  # l1: d/dd mean(sqrt(d))/2 = 1/(4 n sqrt(d)); inf at d == 0, as in the reference
            g1 = grad_loss / (4.0 * ctx.n1) / torch.sqrt(d1)
            g2 = grad_loss / (4.0 * ctx.n2) / torch.sqrt(d2)
        gx1, gx2 = ops.chamfer_backward(xyz1, xyz2, i1, i2, g1, g2)
        return gx1, gx2, None, None, None


def sharded_chamfer(xyz1_local, xyz2_local, kind="l1", n_global_clouds=None, group=None):
    """Chamfer-L1 / L2 loss of the GLOBAL batch from this rank's shard of clouds.
    Every rank returns the same scalar; gradients flow to the local clouds only."""
    if kind not in ("l1", "l2"):
        raise ValueError("kind must be 'l1' or 'l2'")
    if n_global_clouds is None:
        n = torch.tensor([xyz1_local.size(0)], dtype=torch.int64, device=xyz1_local.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(n, group=group)
        n_global_clouds = int(n.item())
    return _ShardedChamfer.apply(xyz1_local.contiguous(), xyz2_local.contiguous(), kind,
                                 int(n_global_clouds), group)
