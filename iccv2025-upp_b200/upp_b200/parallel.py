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


class PeerExchange:
    """Exchange buffers for the fused Chamfer-sums all-reduce (upp_chamfer_fwd_sharded_f32): every rank's buffer
    mapped into every process, so that the kernel that finishes a rank's sums can store them straight into its
    peers' memory over NVLink.  Mapping: torch symmetric memory (CUDA VMM handles exchanged through the process
    group's store) first, classic CUDA IPC handles second.  Construction is a collective; raises if neither
    mapping works on this system (callers then keep the NCCL path)."""

    def __init__(self, group=None, device=None, timeout_s=None):
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerExchange needs an initialised process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.UPP_MAX_PEERS:
            raise RuntimeError(f"world size {self.world} > UPP_MAX_PEERS")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n = 2 * self.world * 8  # floats: slots[2 parities][world][8]
        self._keep = []
        try:
            ptrs, self.how = self._map_symmetric(group), "torch symmetric memory"
        except Exception as first:  # noqa: BLE001 -- any failure of the preferred mapping: try the classic one
            try:
                ptrs, self.how = self._map_ipc(group), "CUDA IPC handles"
            except Exception as second:  # noqa: BLE001
                raise RuntimeError(f"peer mapping unavailable (symmetric memory: {first!r}; IPC: {second!r})") from second
        self.seq = torch.zeros(4, dtype=torch.int32, device=self.device)
        # status word in pinned host memory (device-visible through UVA): a wait for a peer that runs out leaves the failed
        # call's sequence number here, and check() raises -- the loss is never NaN-poisoned silently
        self.status = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.struct = _lib.PeerExchangeStruct()
        for r in range(self.world):
            self.struct.slots[r] = ptrs[r]
        self.struct.rank, self.struct.world, self.struct.seq = self.rank, self.world, self.seq.data_ptr()
        self.struct.status = self.status.data_ptr()
        self.struct.timeout_cycles = int(timeout_s * 1.9e9) if timeout_s else 0  # 0: the library default (~2.3 minutes)
        torch.cuda.synchronize(self.device)
        dist.barrier(group)  # every buffer is zeroed and mapped before anyone's first call

    def check(self):
        """Raise if a fused exchange timed out waiting for a peer (reads one pinned host word: no CUDA synchronisation of
        its own -- call it where the step's loss has been read back, e.g. next to loss.item())."""
        failed = int(self.status[0])
        if failed:
            raise RuntimeError(f"upp_b200 peer exchange: rank {self.rank} gave up waiting for a peer in exchange #{failed} "
                               "(a rank stalled, died, or skipped a collective call); the sums of that call are NaN")

    def _map_symmetric(self, group):
        import torch.distributed._symmetric_memory as symm_mem
        buf = symm_mem.empty(self.n, dtype=torch.float32, device=self.device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        if len(ptrs) != self.world or ptrs[self.rank] != buf.data_ptr():
            raise RuntimeError("unexpected symmetric-memory handle layout")
        self._keep += [buf, hdl]
        return ptrs

    def _map_ipc(self, group):
        buf = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        handle = buf.untyped_storage()._share_cuda_()
        got = [None] * self.world
        dist.all_gather_object(got, (handle, buf.storage_offset()), group)
        ptrs = []
        for r, (h, off) in enumerate(got):
            if r == self.rank:
                ptrs.append(buf.data_ptr())
                continue
            st = torch.UntypedStorage._new_shared_cuda(*h)
            peer = torch.empty(0, dtype=torch.float32, device=st.device).set_(st, off, (self.n,))
            probe = torch.empty(1, dtype=torch.float32, device=self.device)
            probe.copy_(peer[:1])  # makes torch enable peer access between the two devices
            self._keep += [st, peer]
            ptrs.append(peer.data_ptr())
        self._keep.append(buf)
        return ptrs


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


class GradStats:
    """Receives the gradient statistics of a sharded_chamfer backward: `sq_norm` = [sum ||dL/dxyz1||^2, sum ||dL/dxyz2||^2]
    over the GLOBAL batch (all ranks' clouds), computed and all-reduced inside the backward kernel, identical on every
    rank.  total_norm() is what torch.nn.utils.clip_grad_norm_ (tools/runner_module.py:204) would report for the
    unsharded coordinate gradients."""

    def __init__(self):
        self.sq_norm = None

    def total_norm(self, which=(0, 1)):
        if self.sq_norm is None:
            raise RuntimeError("GradStats: no backward has run yet")
        return torch.sqrt(sum(self.sq_norm[i] for i in which))


class _ShardedChamfer(Function):
    """Chamfer loss over a batch sharded across ranks.  Forward: local kernel + one all-reduce of
    4 floats.  Backward: no collective for the gradients -- grad_dist is the constant 1/(B_global*N) (times the
    sqrt chain for L1), so each rank's coordinate gradients are exactly the rows the unsharded
    computation would produce for its clouds (times `scale`, see sharded_chamfer)."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, kind, n_global_clouds, group, peers, scale, stats):
        # the path is chosen from RANK-INVARIANT information only (the per-cloud sizes, equal on every rank): a rank
        # whose shard is empty still takes part in the fused exchange (it contributes zeros) -- it must never end up in
        # NCCL while its peers spin on the exchange buffers
        fused = peers is not None and max(xyz1.size(1), xyz2.size(1)) >= 128 and min(xyz1.size(1), xyz2.size(1)) >= 1
        if fused:
            # one kernel chain: local Chamfer + sums + all-reduce over NVLink peer memory, no NCCL launch
            d1, d2, i1, i2, sums = ops.chamfer_forward_sharded(xyz1, xyz2, peers)
        else:
            d1, d2, i1, i2, sums = ops.chamfer_forward(xyz1, xyz2, want_sums=True)
            reduce_sums(sums, group)
        n1 = float(n_global_clouds * xyz1.size(1))
        n2 = float(n_global_clouds * xyz2.size(1))
        ctx.save_for_backward(xyz1, xyz2, i1, i2, d1, d2)
        ctx.kind, ctx.n1, ctx.n2, ctx.scale = kind, n1, n2, float(scale)
        ctx.stats, ctx.peers, ctx.group, ctx.fused = stats, peers, group, fused
        return chamfer_loss_from_sums(sums, n1, n2, kind)

    @staticmethod
    def backward(ctx, grad_loss):
        xyz1, xyz2, i1, i2, d1, d2 = ctx.saved_tensors
        grad_loss = grad_loss * ctx.scale
        if ctx.kind == "l2":
            g1 = (grad_loss / ctx.n1).expand_as(d1)
            g2 = (grad_loss / ctx.n2).expand_as(d2)
        else:  # l1: d/dd mean(sqrt(d))/2 = 1/(4 n sqrt(d)); inf at d == 0, as in the reference
            g1 = grad_loss / (4.0 * ctx.n1) / torch.sqrt(d1)
            g2 = grad_loss / (4.0 * ctx.n2) / torch.sqrt(d2)
        if ctx.stats is None:
            gx1, gx2 = ops.chamfer_backward(xyz1, xyz2, i1, i2, g1, g2)
        else:
            gx1, gx2, sq = ops.chamfer_backward(xyz1, xyz2, i1, i2, g1, g2, want_sqnorm=True,
                                                peers=ctx.peers if ctx.fused else None)
            if not ctx.fused:
                reduce_sums(sq, ctx.group)
            ctx.stats.sq_norm = sq[:2]
        return gx1, gx2, None, None, None, None, None, None


def sharded_chamfer(xyz1_local, xyz2_local, kind="l1", n_global_clouds=None, group=None, peers=None,
                    grad_scale="global", stats=None):
    """Chamfer-L1 / L2 loss of the GLOBAL batch from this rank's shard of clouds (an empty shard is fine).
    Every rank returns the same scalar; gradients flow to the local clouds only.
    peers: a PeerExchange -> the sums are all-reduced inside the Chamfer kernels over NVLink peer memory
    (bit-identical on every rank); None -> one NCCL all-reduce of 16 bytes.
    grad_scale: how the backward is normalised.
      "global" (default): d(global mean loss)/d(local clouds) -- the rows of the unsharded gradient.  Summing parameter
                gradients over ranks then gives the unsharded gradient.
      "ddp":    times world_size.  The reference trains under DistributedDataParallel, which AVERAGES parameter gradients
                over ranks, with every rank back-propagating its LOCAL mean loss (tools/runner_pretask.py:226-241,
                main.py:46-53); the global-mean loss back-propagated per rank and then averaged would be world_size times
                smaller -- a different effective learning rate and grad_norm_clip threshold.  With "ddp" the averaged
                gradient equals the reference's (exactly for equal shards; for uneven shards it is the gradient of the
                global mean, which is what the reference's per-rank means approximate).
    stats: a GradStats -> its sq_norm receives [sum ||dL/dxyz1||^2, sum ||dL/dxyz2||^2] over all ranks' clouds after
           backward (exchanged inside the backward kernel with `peers`, else one more 16-byte NCCL all-reduce)."""
    if kind not in ("l1", "l2"):
        raise ValueError("kind must be 'l1' or 'l2'")
    if n_global_clouds is None:
        n = torch.tensor([xyz1_local.size(0)], dtype=torch.int64, device=xyz1_local.device)
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(n, group=group)
        n_global_clouds = int(n.item())
    if grad_scale not in ("global", "ddp"):
        raise ValueError("grad_scale must be 'global' or 'ddp'")
    ws = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    return _ShardedChamfer.apply(xyz1_local.contiguous(), xyz2_local.contiguous(), kind,
                                 int(n_global_clouds), group, peers, float(ws) if grad_scale == "ddp" else 1.0, stats)
