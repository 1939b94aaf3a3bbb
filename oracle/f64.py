"""float64 brute-force statements used for tolerance checks (TEST INFRASTRUCTURE ONLY)."""
import numpy as np


def pair_sq(a, b):
    """(B,N,3),(B,M,3) -> (B,N,M) float64 squared distances by direct differences."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    diff = a[:, :, None, :] - b[:, None, :, :]
    return (diff * diff).sum(-1)


def chamfer_fwd(xyz1, xyz2):
    d = pair_sq(xyz1, xyz2)
    return d.min(2), d.min(1), d.argmin(2), d.argmin(1)


def knn(ref, query, k):
    d = pair_sq(query, ref)
    order = np.argsort(d, axis=-1, kind="stable")[:, :, :k]
    return np.sqrt(np.take_along_axis(d, order, -1)), order


def fps(xyz, M):
    """start 0, skip |p|^2 <= 1e-3, lowest-index ties; float64 distances."""
    xyz = np.asarray(xyz, dtype=np.float64)
    B, N, _ = xyz.shape
    out = np.zeros((B, M), dtype=np.int64)
    for b in range(B):
        p = xyz[b]
        valid = (p * p).sum(-1) > 1e-3
        md = np.full(N, 1e10)
        old = 0
        for j in range(1, M):
            d = ((p - p[old]) ** 2).sum(-1)
            md = np.where(valid, np.minimum(md, d), md)
            cand = np.where(valid, md, -1.0)
            old = int(cand.argmax()) if valid.any() else 0
            out[b, j] = old
    return out
