"""Tensor-level entry points over the C ABI: checks, allocation, stream/device plumbing.

torch is used for device memory, streams and autograd only; every computation below is a
libupp_geom.so kernel.  Argument rules follow the ops these replace (file:line in each
docstring): CUDA-only, fp32, contiguous; violations raise instead of printing.
"""
import functools
import os

import torch

from . import _lib

# UPP_NVTX=1 wraps every op in an NVTX range named after its C-ABI entry point, so that the ops
# line up with the kernels in an nsys / ncu timeline.  Off by default: a range push / pop costs ~1 us of host time per op.
_NVTX = os.environ.get("UPP_NVTX") == "1"


def _traced(fn):
    if not _NVTX:
        return fn

    @functools.wraps(fn)
    def wrapper(*a, **k):
        torch.cuda.nvtx.range_push("upp_b200." + fn.__name__)
        try:
            return fn(*a, **k)
        finally:
            torch.cuda.nvtx.range_pop()
    return wrapper


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(name, t):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (CPU not supported; there is no CPU fallback)")


def _need(name, t, dtype, ndim):
    _need_cuda(name, t)
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _xyz(name, t):
    _need(name, t, torch.float32, 3)
    if t.shape[2] != 3:
        raise ValueError(f"{name} must be (B, N, 3), got {tuple(t.shape)}")


class _on:
    """Make t's device current for the duration of the call (only when it is not already)."""
    __slots__ = ("prev", "dev")

    def __init__(self, t):
        self.dev = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def _ptr(t):
    return t.data_ptr() if t is not None else None


@_traced
def fps(xyz, npoint, want_centers=False):
    """Farthest point sampling; replaces pointnet2_utils.furthest_point_sample (utils/misc.py:18)
    and, with want_centers, also the gather of utils/misc.py:19.
    xyz (B,N,3) f32 CUDA -> idx (B,npoint) int32 [, centers (B,npoint,3) f32]."""
    _xyz("xyz", xyz)
    npoint = int(npoint)
    if npoint < 0:
        raise ValueError("npoint must be >= 0")
    B, N, _ = xyz.shape
    if N == 0 and npoint > 0 and B > 0:
        raise ValueError("cannot sample from an empty cloud")
    lib = _lib.load()
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    centers = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device) if want_centers else None
    wbytes = int(lib.upp_fps_workspace_bytes(B, N, npoint))
    ws = torch.empty(wbytes, dtype=torch.uint8, device=xyz.device) if wbytes else None
    with _on(xyz):
        rc = lib.upp_fps_f32(_ptr(xyz), B, N, npoint, _ptr(idx), _ptr(centers), _ptr(ws), wbytes, _stream(xyz))
    _lib.check(rc, "upp_fps_f32")
    return (idx, centers) if want_centers else idx


@_traced
def gather(features, idx):
    """features (B,C,N) f32, idx (B,M) int32 -> (B,C,M); replaces gather_operation fwd (utils/misc.py:19)."""
    _need("features", features, torch.float32, 3)
    _need("idx", idx, torch.int32, 2)
    B, C, N = features.shape
    if idx.shape[0] != B:
        raise ValueError(f"idx batch {idx.shape[0]} != features batch {B}")
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=features.device)
    with _on(features):
        rc = _lib.load().upp_gather_f32(_ptr(features), _ptr(idx), B, C, N, M, _ptr(out), _stream(features))
    _lib.check(rc, "upp_gather_f32")
    return out


@_traced
def gather_grad(grad_out, idx, N):
    """grad_out (B,C,M), idx (B,M) int32 -> grad_features (B,C,N) (scatter-add)."""
    _need("grad_out", grad_out, torch.float32, 3)
    _need("idx", idx, torch.int32, 2)
    B, C, M = grad_out.shape
    g = torch.empty((B, C, int(N)), dtype=torch.float32, device=grad_out.device)
    with _on(grad_out):
        rc = _lib.load().upp_gather_grad_f32(_ptr(grad_out), _ptr(idx), B, C, int(N), M, _ptr(g), _stream(grad_out))
    _lib.check(rc, "upp_gather_grad_f32")
    return g


@_traced
def rows_scatter_add(grad_rows, idx, N):
    """grad_rows (B,M,C), idx (B,M) int32 -> (B,N,C): gradient of the row-major gather rows = x[b, idx[b,j], :]
    (what misc.fps returns as fps_data)."""
    _need("grad_rows", grad_rows, torch.float32, 3)
    _need("idx", idx, torch.int32, 2)
    B, M, C = grad_rows.shape
    g = torch.empty((B, int(N), C), dtype=torch.float32, device=grad_rows.device)
    with _on(grad_rows):
        rc = _lib.load().upp_rows_scatter_add_f32(_ptr(grad_rows), _ptr(idx), B, int(N), M, C, _ptr(g), _stream(grad_rows))
    _lib.check(rc, "upp_rows_scatter_add_f32")
    return g


@_traced
def knn(ref, query, k, want_dist=True):
    """ref (B,N,3), query (B,Q,3) -> D (B,Q,k) f32 Euclidean ascending, I (B,Q,k) int64;
    replaces KNN(k, transpose_mode=True).forward (models/Point_MAE_unify.py:69)."""
    _xyz("ref", ref)
    _xyz("query", query)
    if ref.shape[0] != query.shape[0]:
        raise ValueError(f"ref.shape={tuple(ref.shape)} != query.shape={tuple(query.shape)}")
    if ref.device != query.device:
        raise RuntimeError("ref and query must be on the same device")
    B, N, _ = ref.shape
    Q = query.shape[1]
    k = int(k)
    if not 1 <= k <= N:
        raise ValueError(f"k={k} must be in [1, N={N}]")
    D = torch.empty((B, Q, k), dtype=torch.float32, device=ref.device) if want_dist else None
    I = torch.empty((B, Q, k), dtype=torch.int64, device=ref.device)
    with _on(ref):
        rc = _lib.load().upp_knn_f32(_ptr(ref), _ptr(query), B, N, Q, k, _ptr(D), _ptr(I), _stream(ref))
    _lib.check(rc, "upp_knn_f32")
    return D, I


@_traced
def chamfer_forward(xyz1, xyz2, want_sums=False):
    """chamfer.forward (extensions/chamfer_dist/chamfer.cu:147-171):
    -> [dist1 (B,N), dist2 (B,M) f32 squared, idx1, idx2 int32] (+ 4-float partial sums)."""
    _xyz("xyz1", xyz1)
    _xyz("xyz2", xyz2)
    if xyz1.shape[0] != xyz2.shape[0]:
        raise ValueError("xyz1 and xyz2 must have the same batch size")
    if xyz1.device != xyz2.device:
        raise RuntimeError("xyz1 and xyz2 must be on the same device")
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    dev = xyz1.device
    d1 = torch.empty((B, N), dtype=torch.float32, device=dev)
    d2 = torch.empty((B, M), dtype=torch.float32, device=dev)
    i1 = torch.empty((B, N), dtype=torch.int32, device=dev)
    i2 = torch.empty((B, M), dtype=torch.int32, device=dev)
    sums = torch.empty(4, dtype=torch.float32, device=dev) if want_sums else None
    lib = _lib.load()
    wbytes = int(lib.upp_chamfer_fwd_workspace_bytes(B, N, M))
    ws = torch.empty(wbytes, dtype=torch.uint8, device=dev) if wbytes else None
    with _on(xyz1):
        rc = lib.upp_chamfer_fwd_f32(_ptr(xyz1), _ptr(xyz2), B, N, M, _ptr(d1), _ptr(d2), _ptr(i1), _ptr(i2),
                                     _ptr(sums), _ptr(ws), wbytes, _stream(xyz1))
    _lib.check(rc, "upp_chamfer_fwd_f32")
    return [d1, d2, i1, i2] + ([sums] if want_sums else [])


@_traced
def chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2, want_sqnorm=False, peers=None):
    """chamfer.backward (extensions/chamfer_dist/chamfer.cu:203-229) -> [grad_xyz1, grad_xyz2].
    grad_dist* may arrive non-contiguous (expanded scalars from mean/sqrt backward); they are
    densified here -- the reference silently assumes dense (chamfer.cu:217).  Deterministic (no atomics).
    want_sqnorm: also return a 4-float tensor whose first two entries are sum ||grad_xyz1||^2, sum ||grad_xyz2||^2
    (upp_chamfer_bwd_stats_f32) -- over the GLOBAL sharded batch when `peers` (a parallel.PeerExchange) is given: the
    gradient statistics a clip_grad_norm_ over the coordinate gradients needs, exchanged inside the kernel."""
    _xyz("xyz1", xyz1)
    _xyz("xyz2", xyz2)
    _need("idx1", idx1, torch.int32, 2)
    _need("idx2", idx2, torch.int32, 2)
    _need_cuda("grad_dist1", grad_dist1)
    _need_cuda("grad_dist2", grad_dist2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    if tuple(grad_dist1.shape) != (B, N) or tuple(grad_dist2.shape) != (B, M):
        raise ValueError("grad_dist shapes must be (B,N) and (B,M)")
    g1 = grad_dist1.to(torch.float32).contiguous()
    g2 = grad_dist2.to(torch.float32).contiguous()
    gx1 = torch.empty_like(xyz1)
    gx2 = torch.empty_like(xyz2)
    lib = _lib.load()
    if want_sqnorm:
        import ctypes
        sq = torch.zeros(4, dtype=torch.float32, device=xyz1.device)
        wbytes = int(lib.upp_chamfer_bwd_stats_workspace_bytes(B, N, M))
        ws = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=xyz1.device)
        pp = None
        if peers is not None:
            peers.struct.defer = 0
            pp = ctypes.addressof(peers.struct)
        with _on(xyz1):
            rc = lib.upp_chamfer_bwd_stats_f32(_ptr(xyz1), _ptr(xyz2), _ptr(idx1), _ptr(idx2), _ptr(g1), _ptr(g2), B, N, M,
                                               _ptr(gx1), _ptr(gx2), _ptr(sq), _ptr(ws), wbytes, pp, _stream(xyz1))
        _lib.check(rc, "upp_chamfer_bwd_stats_f32")
        return [gx1, gx2, sq]
    with _on(xyz1):
        rc = lib.upp_chamfer_bwd_f32(_ptr(xyz1), _ptr(xyz2), _ptr(idx1), _ptr(idx2), _ptr(g1),
                                     _ptr(g2), B, N, M, _ptr(gx1), _ptr(gx2), _stream(xyz1))
    _lib.check(rc, "upp_chamfer_bwd_f32")
    return [gx1, gx2]


@_traced
def group(xyz, num_group, group_size):
    """Fused Group divider (models/Point_MAE_unify.py:58-92):
    -> neighborhood (B,G,k,3), center (B,G,3), idx (B,G,k) int64 local, center_idx (B,G) int32."""
    _xyz("xyz", xyz)
    B, N, _ = xyz.shape
    G, k = int(num_group), int(group_size)
    if not 1 <= k <= N:
        raise ValueError(f"group_size={k} must be in [1, N={N}]")
    dev = xyz.device
    lib = _lib.load()
    nb = torch.empty((B, G, k, 3), dtype=torch.float32, device=dev)
    ce = torch.empty((B, G, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((B, G, k), dtype=torch.int64, device=dev)
    cidx = torch.empty((B, G), dtype=torch.int32, device=dev)
    wbytes = int(lib.upp_fps_workspace_bytes(B, N, G))
    ws = torch.empty(wbytes, dtype=torch.uint8, device=dev) if wbytes else None
    with _on(xyz):
        rc = lib.upp_group_f32(_ptr(xyz), B, N, G, k, _ptr(nb), _ptr(ce), _ptr(idx), _ptr(cidx),
                               _ptr(ws), wbytes, _stream(xyz))
    _lib.check(rc, "upp_group_f32")
    return nb, ce, idx, cidx


@_traced
def group_backward(grad_nb, grad_center, idx, center_idx, N):
    """Gradient of the fused Group w.r.t. xyz -> (B,N,3)."""
    _need("grad_nb", grad_nb, torch.float32, 4)
    _need("idx", idx, torch.int64, 3)
    _need("center_idx", center_idx, torch.int32, 2)
    B, G, k, _ = grad_nb.shape
    if grad_center is not None:
        _need("grad_center", grad_center, torch.float32, 3)
    gx = torch.empty((B, int(N), 3), dtype=torch.float32, device=grad_nb.device)
    with _on(grad_nb):
        rc = _lib.load().upp_group_bwd_f32(_ptr(grad_nb), _ptr(grad_center), _ptr(idx), _ptr(center_idx),
                                           B, int(N), G, k, _ptr(gx), _stream(grad_nb))
    _lib.check(rc, "upp_group_bwd_f32")
    return gx


@_traced
def interp_forward(xyz1, xyz2, points2, k, eps, base=None, alpha=1.0, want_dist=True):
    """k-nearest inverse-distance interpolation of points2 (B,S,C) from xyz2 (B,S,3) onto xyz1 (B,N,3):
    the body of propagate (models/Point_MAE_unify.py:22-48) and of PointNetFeaturePropagation's
    interpolation (models/Point_MAE_unify_segment.py:289-313) in one launch.
    -> out (B,N,C) = (base or 0) + alpha * interpolated, idx (B,N,k) int32, weight (B,N,k), dist (B,N,k)."""
    _xyz("xyz1", xyz1)
    _xyz("xyz2", xyz2)
    _need("points2", points2, torch.float32, 3)
    B, N, _ = xyz1.shape
    S, C = points2.shape[1], points2.shape[2]
    if xyz2.shape[0] != B or points2.shape[0] != B or xyz2.shape[1] != S:
        raise ValueError(f"shape mismatch: xyz1 {tuple(xyz1.shape)}, xyz2 {tuple(xyz2.shape)}, points2 {tuple(points2.shape)}")
    k = int(k)
    if not 1 <= k <= min(S, 32):
        raise ValueError(f"k={k} must be in [1, min(S={S}, 32)]")
    if base is not None:
        _need("base", base, torch.float32, 3)
        if tuple(base.shape) != (B, N, C):
            raise ValueError(f"base must be {(B, N, C)}, got {tuple(base.shape)}")
    dev = xyz1.device
    out = torch.empty((B, N, C), dtype=torch.float32, device=dev)
    idx = torch.empty((B, N, k), dtype=torch.int32, device=dev)
    w = torch.empty((B, N, k), dtype=torch.float32, device=dev)
    d = torch.empty((B, N, k), dtype=torch.float32, device=dev) if want_dist else None
    with _on(xyz1):
        rc = _lib.load().upp_interp_fwd_f32(_ptr(xyz1), _ptr(xyz2), _ptr(points2), _ptr(base), float(alpha), float(eps),
                                            B, N, S, C, k, _ptr(out), _ptr(idx), _ptr(w), _ptr(d), _stream(xyz1))
    _lib.check(rc, "upp_interp_fwd_f32")
    return out, idx, w, d


@_traced
def interp_select(xyz1, xyz2, k, eps, want_dist=True):
    """Selection half of interp_forward alone (upp_interp_select_f32): -> idx (B,N,k) int32, weight (B,N,k), dist (B,N,k)."""
    _xyz("xyz1", xyz1)
    _xyz("xyz2", xyz2)
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    if xyz2.shape[0] != B:
        raise ValueError("xyz1 and xyz2 must have the same batch size")
    k = int(k)
    if not 1 <= k <= min(S, 32):
        raise ValueError(f"k={k} must be in [1, min(S={S}, 32)]")
    dev = xyz1.device
    idx = torch.empty((B, N, k), dtype=torch.int32, device=dev)
    w = torch.empty((B, N, k), dtype=torch.float32, device=dev)
    d = torch.empty((B, N, k), dtype=torch.float32, device=dev) if want_dist else None
    with _on(xyz1):
        rc = _lib.load().upp_interp_select_f32(_ptr(xyz1), _ptr(xyz2), float(eps), B, N, S, k, _ptr(idx), _ptr(w), _ptr(d), _stream(xyz1))
    _lib.check(rc, "upp_interp_select_f32")
    return idx, w, d


@_traced
def interp_blend(points2, idx, weight, base=None, alpha=1.0):
    """Blend half of interp_forward from a saved selection (upp_interp_blend_f32): -> out (B,N,C), bit-identical to what
    interp_forward returns for the same inputs."""
    _need("points2", points2, torch.float32, 3)
    _need("idx", idx, torch.int32, 3)
    _need("weight", weight, torch.float32, 3)
    B, S, C = points2.shape
    N, k = idx.shape[1], idx.shape[2]
    if idx.shape[0] != B or tuple(weight.shape) != tuple(idx.shape):
        raise ValueError("selection does not match points2")
    if base is not None:
        _need("base", base, torch.float32, 3)
        if tuple(base.shape) != (B, N, C):
            raise ValueError(f"base must be {(B, N, C)}, got {tuple(base.shape)}")
    out = torch.empty((B, N, C), dtype=torch.float32, device=points2.device)
    with _on(points2):
        rc = _lib.load().upp_interp_blend_f32(_ptr(points2), _ptr(base), float(alpha), _ptr(idx), _ptr(weight), B, N, S, C, k,
                                              _ptr(out), _stream(points2))
    _lib.check(rc, "upp_interp_blend_f32")
    return out


@_traced
def interp_backward(grad_out, idx, weight, S, alpha=1.0, xyz_terms=None):
    """Gradients of interp_forward: grad_points2 (B,S,C) always; with xyz_terms = (dist, points2, xyz1, xyz2, eps)
    also grad_xyz1 (B,N,3), grad_xyz2 (B,S,3) (the path through the weights).  Deterministic, no atomics."""
    _need("grad_out", grad_out, torch.float32, 3)
    _need("idx", idx, torch.int32, 3)
    _need("weight", weight, torch.float32, 3)
    B, N, C = grad_out.shape
    k = idx.shape[2]
    dev = grad_out.device
    gp2 = torch.empty((B, int(S), C), dtype=torch.float32, device=dev)
    g1 = g2 = gd = None
    dist = points2 = xyz1 = xyz2 = None
    eps = 0.0
    if xyz_terms is not None:
        dist, points2, xyz1, xyz2, eps = xyz_terms
        _need("dist", dist, torch.float32, 3)
        _need("points2", points2, torch.float32, 3)
        _xyz("xyz1", xyz1)
        _xyz("xyz2", xyz2)
        g1 = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        g2 = torch.empty((B, int(S), 3), dtype=torch.float32, device=dev)
        gd = torch.empty((B, N, k), dtype=torch.float32, device=dev)
    lib = _lib.load()
    ws_bytes = int(lib.upp_interp_bwd_workspace_bytes(B, N, int(S), C, k))  # 0: no streamed path for this shape
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    with _on(grad_out):
        rc = lib.upp_interp_bwd_f32(_ptr(grad_out), _ptr(idx), _ptr(weight), _ptr(dist), _ptr(points2),
                                    _ptr(xyz1), _ptr(xyz2), float(alpha), float(eps), B, N, int(S), C, k,
                                    _ptr(gp2), _ptr(g1), _ptr(g2), _ptr(gd), _ptr(ws), ws_bytes, _stream(grad_out))
    _lib.check(rc, "upp_interp_bwd_f32")
    return gp2, g1, g2


@_traced
def knn_points(p1, p2, K, want_nn=False):
    """pytorch3d.ops.knn_points convention (models/Point_MAE_pretask_dev.py:680): p1 (B,N1,3) queries,
    p2 (B,N2,3) references -> dists (B,N1,K) f32 SQUARED ascending, idx (B,N1,K) int64 [, nn (B,N1,K,3)]."""
    _xyz("p1", p1)
    _xyz("p2", p2)
    if p1.shape[0] != p2.shape[0]:
        raise ValueError("pts1 and pts2 must have the same batch dimension.")
    if p1.device != p2.device:
        raise RuntimeError("p1 and p2 must be on the same device")
    B, N1, _ = p1.shape
    N2 = p2.shape[1]
    K = int(K)
    if not 1 <= K <= min(N2, 32):
        raise ValueError(f"K={K} must be in [1, min(N2={N2}, 32)]")
    dev = p1.device
    d = torch.empty((B, N1, K), dtype=torch.float32, device=dev)
    i = torch.empty((B, N1, K), dtype=torch.int64, device=dev)
    nn = torch.empty((B, N1, K, 3), dtype=torch.float32, device=dev) if want_nn else None
    with _on(p1):
        rc = _lib.load().upp_knn_points_f32(_ptr(p1), _ptr(p2), B, N1, N2, K, _ptr(d), _ptr(i), _ptr(nn), _stream(p1))
    _lib.check(rc, "upp_knn_points_f32")
    return d, i, nn


@_traced
def crop_split(xyz, viewpoints, num_crop, padding_zeros=False, want_order=False):
    """The crop of misc.seprate_point_cloud (utils/misc.py:232-239) for the whole batch in one launch:
    xyz (B,n,3), viewpoints (B,3) -> (input (B,n-num_crop,3) or (B,n,3) when padding_zeros, crop (B,num_crop,3)[, order (B,n) int32])."""
    _xyz("xyz", xyz)
    _need("viewpoints", viewpoints, torch.float32, 2)
    B, n, _ = xyz.shape
    if tuple(viewpoints.shape) != (B, 3):
        raise ValueError(f"viewpoints must be ({B}, 3), got {tuple(viewpoints.shape)}")
    num_crop = int(num_crop)
    if not 0 <= num_crop <= n:
        raise ValueError(f"num_crop={num_crop} must be in [0, n={n}]")
    dev = xyz.device
    crop = torch.empty((B, num_crop, 3), dtype=torch.float32, device=dev)
    inp = torch.empty((B, n if padding_zeros else n - num_crop, 3), dtype=torch.float32, device=dev)
    order = torch.empty((B, n), dtype=torch.int32, device=dev) if want_order else None
    with _on(xyz):
        rc = _lib.load().upp_crop_split_f32(_ptr(xyz), _ptr(viewpoints), B, n, num_crop, 1 if padding_zeros else 0,
                                            _ptr(crop), _ptr(inp), _ptr(order), _stream(xyz))
    _lib.check(rc, "upp_crop_split_f32")
    return (inp, crop, order) if want_order else (inp, crop)


@_traced
def peer_allreduce_finish(peers, device):
    """upp_peer_allreduce_finish_f32: second half of a deferred exchange -> global sums (4 floats)."""
    import ctypes
    sums = torch.empty(4, dtype=torch.float32, device=device)
    with _on(sums):
        rc = _lib.load().upp_peer_allreduce_finish_f32(ctypes.addressof(peers.struct), _ptr(sums), _stream(sums))
    _lib.check(rc, "upp_peer_allreduce_finish_f32")
    return sums


@_traced
def peer_allreduce(peers, local4=None, defer=False):
    """upp_peer_allreduce_f32: SUM all-reduce of 4 floats over the mapped peer buffers (local4 None: zeros -- what a rank
    with an EMPTY shard contributes).  -> global sums (4 floats), or the local ones when defer=True."""
    import ctypes
    peers.struct.defer = 1 if defer else 0
    out = torch.empty(4, dtype=torch.float32, device=peers.device)
    if local4 is not None:
        _need("local4", local4, torch.float32, 1)
    with _on(out):
        rc = _lib.load().upp_peer_allreduce_f32(ctypes.addressof(peers.struct), _ptr(local4), _ptr(out), _stream(out))
    _lib.check(rc, "upp_peer_allreduce_f32")
    return out


@_traced
def chamfer_forward_sharded(xyz1, xyz2, peers, defer=False):
    """upp_chamfer_fwd_sharded_f32: chamfer.forward of this rank's clouds, fused with the all-reduce of its sums
    over NVLink peer memory.  `peers` is a parallel.PeerExchange.  -> [dist1, dist2, idx1, idx2, global_sums].
    defer=True: the kernels only SEND (the 5th output holds the LOCAL sums); call peer_allreduce_finish(peers, device)
    later on the same stream -- where the loss value is consumed -- for the global sums."""
    if hasattr(peers.struct, "defer"):
        peers.struct.defer = 1 if defer else 0
    _xyz("xyz1", xyz1)
    _xyz("xyz2", xyz2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    dev = xyz1.device
    d1 = torch.empty((B, N), dtype=torch.float32, device=dev)
    d2 = torch.empty((B, M), dtype=torch.float32, device=dev)
    i1 = torch.empty((B, N), dtype=torch.int32, device=dev)
    i2 = torch.empty((B, M), dtype=torch.int32, device=dev)
    sums = torch.empty(4, dtype=torch.float32, device=dev)
    lib = _lib.load()
    wbytes = int(lib.upp_chamfer_fwd_workspace_bytes(B, N, M))
    ws = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
    import ctypes
    with _on(xyz1):
        rc = lib.upp_chamfer_fwd_sharded_f32(_ptr(xyz1), _ptr(xyz2), B, N, M, _ptr(d1), _ptr(d2), _ptr(i1), _ptr(i2),
                                             _ptr(sums), _ptr(ws), wbytes, ctypes.addressof(peers.struct), _stream(xyz1))
    _lib.check(rc, "upp_chamfer_fwd_sharded_f32")
    return [d1, d2, i1, i2, sums]
