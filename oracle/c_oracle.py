"""numpy/ctypes front-end to oracle/upp_oracle.c (TEST INFRASTRUCTURE ONLY).

Each wrapper names the reference file:line its C function follows; see the C
file for the restatement itself.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libupp_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile upp_oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "upp_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.upp_oracle_knn.restype = ctypes.c_int
        _lib.upp_oracle_group.restype = ctypes.c_int
        _lib.upp_oracle_fps_upstream_block.restype = ctypes.c_int
        _lib.upp_oracle_get_threads.restype = ctypes.c_int
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


def set_threads(n):
    lib().upp_oracle_set_threads(ctypes.c_int(int(n)))


def get_threads():
    return int(lib().upp_oracle_get_threads())


def upstream_block(N):
    """upstream opt_n_threads(N) (pointnet2_ops sampling_gpu.cu)."""
    return int(lib().upp_oracle_fps_upstream_block(ctypes.c_int(int(N))))


def fps(xyz, M, block_size=0):
    """pointnet2_utils.furthest_point_sample (reference utils/misc.py:18).
    xyz (B,N,3) f32 -> (B,M) int32.  block_size=0: lowest-index ties."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.zeros((B, M), dtype=np.int32)
    lib().upp_oracle_fps(_p(xyz, _f32p), B, N, int(M), _p(out, _i32p), int(block_size))
    return out


def gather(feat, idx):
    """pointnet2_utils.gather_operation fwd (reference utils/misc.py:19).
    feat (B,C,N) f32, idx (B,M) int32 -> (B,C,M)."""
    feat = _f32(feat)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, N = feat.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), dtype=np.float32)
    lib().upp_oracle_gather(_p(feat, _f32p), _p(idx, _i32p), B, C, N, M, _p(out, _f32p))
    return out


def gather_grad(gout, idx, N):
    """gather_operation backward: scatter-add (B,C,M) -> (B,C,N)."""
    gout = _f32(gout)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, C, M = gout.shape
    out = np.empty((B, C, N), dtype=np.float32)
    lib().upp_oracle_gather_grad(_p(gout, _f32p), _p(idx, _i32p), B, C, int(N), M, _p(out, _f32p))
    return out


def knn(ref, query, k):
    """knn_cuda.KNN(k, transpose_mode=True)(ref, query)
    (reference models/Point_MAE_unify.py:56,69).
    ref (B,N,3), query (B,Q,3) -> D (B,Q,k) f32 Euclidean ascending, I (B,Q,k) int64."""
    ref = _f32(ref)
    query = _f32(query)
    B, N, _ = ref.shape
    Q = query.shape[1]
    D = np.empty((B, Q, k), dtype=np.float32)
    I = np.empty((B, Q, k), dtype=np.int64)
    rc = lib().upp_oracle_knn(_p(ref, _f32p), _p(query, _f32p), B, N, Q, int(k),
                              _p(D, _f32p), _p(I, _i64p))
    if rc != 0:
        raise ValueError(f"knn: k={k} must be in [1, N={N}]")
    return D, I


def chamfer_fwd(xyz1, xyz2):
    """chamfer.forward (reference extensions/chamfer_dist/chamfer.cu:147-171).
    -> dist1 (B,N), dist2 (B,M) squared f32; idx1, idx2 int32."""
    xyz1 = _f32(xyz1)
    xyz2 = _f32(xyz2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    d1 = np.empty((B, N), dtype=np.float32)
    d2 = np.empty((B, M), dtype=np.float32)
    i1 = np.empty((B, N), dtype=np.int32)
    i2 = np.empty((B, M), dtype=np.int32)
    lib().upp_oracle_chamfer_fwd(_p(xyz1, _f32p), _p(xyz2, _f32p), B, N, M,
                                 _p(d1, _f32p), _p(d2, _f32p), _p(i1, _i32p), _p(i2, _i32p))
    return d1, d2, i1, i2


def chamfer_bwd(xyz1, xyz2, idx1, idx2, g1, g2):
    """chamfer.backward (reference extensions/chamfer_dist/chamfer.cu:203-229)."""
    xyz1 = _f32(xyz1)
    xyz2 = _f32(xyz2)
    idx1 = np.ascontiguousarray(idx1, dtype=np.int32)
    idx2 = np.ascontiguousarray(idx2, dtype=np.int32)
    g1 = _f32(g1)
    g2 = _f32(g2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    gx1 = np.empty((B, N, 3), dtype=np.float32)
    gx2 = np.empty((B, M, 3), dtype=np.float32)
    lib().upp_oracle_chamfer_bwd(_p(xyz1, _f32p), _p(xyz2, _f32p), _p(idx1, _i32p),
                                 _p(idx2, _i32p), _p(g1, _f32p), _p(g2, _f32p), B, N, M,
                                 _p(gx1, _f32p), _p(gx2, _f32p))
    return gx1, gx2


def group(xyz, G, k):
    """Group.forward (reference models/Point_MAE_unify.py:58-92), gather_idx=True
    index convention: idx (B,G,k) int64 local, center_idx (B,G) int32."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    nb = np.empty((B, G, k, 3), dtype=np.float32)
    ce = np.empty((B, G, 3), dtype=np.float32)
    idx = np.empty((B, G, k), dtype=np.int64)
    cidx = np.empty((B, G), dtype=np.int32)
    rc = lib().upp_oracle_group(_p(xyz, _f32p), B, N, int(G), int(k), _p(nb, _f32p),
                                _p(ce, _f32p), _p(idx, _i64p), _p(cidx, _i32p))
    if rc != 0:
        raise ValueError(f"group: k={k} must be in [1, N={N}]")
    return nb, ce, idx, cidx


def interp_fwd(xyz1, xyz2, feat2, k, eps, base=None, alpha=1.0):
    """propagate / PointNetFeaturePropagation interpolation (reference
    models/Point_MAE_unify.py:22-48, models/Point_MAE_unify_segment.py:289-313).
    -> out (B,N,C), idx (B,N,k) int32, weight (B,N,k), dist (B,N,k)."""
    xyz1, xyz2, feat2 = _f32(xyz1), _f32(xyz2), _f32(feat2)
    B, N, _ = xyz1.shape
    S, C = feat2.shape[1], feat2.shape[2]
    out = np.empty((B, N, C), dtype=np.float32)
    idx = np.empty((B, N, k), dtype=np.int32)
    w = np.empty((B, N, k), dtype=np.float32)
    d = np.empty((B, N, k), dtype=np.float32)
    bp = _p(_f32(base), _f32p) if base is not None else None
    lib().upp_oracle_interp_fwd.restype = ctypes.c_int
    rc = lib().upp_oracle_interp_fwd(_p(xyz1, _f32p), _p(xyz2, _f32p), _p(feat2, _f32p), bp,
                                     ctypes.c_float(alpha), ctypes.c_float(eps), B, N, S, C, int(k),
                                     _p(out, _f32p), _p(idx, _i32p), _p(w, _f32p), _p(d, _f32p))
    if rc != 0:
        raise ValueError(f"interp: k={k} must be in [1, S={S}]")
    return out, idx, w, d


def interp_bwd(gout, feat2, xyz1, xyz2, idx, weight, dist, eps, alpha=1.0):
    """Analytic gradients of interp_fwd -> grad_feat2 (B,S,C), grad_xyz1 (B,N,3), grad_xyz2 (B,S,3)."""
    gout, feat2, xyz1, xyz2 = _f32(gout), _f32(feat2), _f32(xyz1), _f32(xyz2)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    weight, dist = _f32(weight), _f32(dist)
    B, N, C = gout.shape
    S = feat2.shape[1]
    k = idx.shape[2]
    gf = np.empty((B, S, C), dtype=np.float32)
    g1 = np.empty((B, N, 3), dtype=np.float32)
    g2 = np.empty((B, S, 3), dtype=np.float32)
    lib().upp_oracle_interp_bwd(_p(gout, _f32p), _p(feat2, _f32p), _p(xyz1, _f32p), _p(xyz2, _f32p),
                                _p(idx, _i32p), _p(weight, _f32p), _p(dist, _f32p), ctypes.c_float(alpha),
                                ctypes.c_float(eps), B, N, S, C, k, _p(gf, _f32p), _p(g1, _f32p), _p(g2, _f32p))
    return gf, g1, g2


def knn_points(p1, p2, K):
    """pytorch3d.ops.knn_points(p1, p2, K=K) convention (reference models/Point_MAE_pretask_dev.py:680):
    -> dists (B,N1,K) squared ascending, idx (B,N1,K) int64."""
    p1, p2 = _f32(p1), _f32(p2)
    B, N1, _ = p1.shape
    N2 = p2.shape[1]
    D = np.empty((B, N1, K), dtype=np.float32)
    I = np.empty((B, N1, K), dtype=np.int64)
    lib().upp_oracle_knn_points.restype = ctypes.c_int
    rc = lib().upp_oracle_knn_points(_p(p1, _f32p), _p(p2, _f32p), B, N1, N2, int(K), _p(D, _f32p), _p(I, _i64p))
    if rc != 0:
        raise ValueError(f"knn_points: K={K} must be in [1, N2={N2}]")
    return D, I


def crop_order(xyz, centers):
    """Ascending order of every cloud's points by distance to its viewpoint (reference utils/misc.py:232-233:
    torch.norm(center - points) -> torch.argsort), stable.  xyz (B,n,3), centers (B,3) -> (B,n) int32."""
    xyz, centers = _f32(xyz), _f32(centers)
    B, n, _ = xyz.shape
    order = np.empty((B, n), np.int32)
    lib().upp_oracle_crop_order(_p(xyz, _f32p), _p(centers, _f32p), ctypes.c_int(B), ctypes.c_int(n), _p(order, _i32p))
    return order


def crop_split(xyz, centers, num_crop, padding_zeros=False):
    """misc.seprate_point_cloud's crop (utils/misc.py:232-239) -> (input_data, crop_data)."""
    xyz = _f32(xyz)
    order = crop_order(xyz, centers).astype(np.int64)
    crop = np.take_along_axis(xyz, order[:, :num_crop, None], 1)
    if padding_zeros:
        inp = xyz.copy()
        np.put_along_axis(inp, order[:, :num_crop, None].repeat(3, -1), crop * np.float32(0), 1)
    else:
        inp = np.take_along_axis(xyz, order[:, num_crop:, None], 1)
    return inp, crop
