"""CPU (`-m "not gpu"`) tests that pin the oracle: against the golden fixtures generated from the
reference's own Python (tests/golden/make_golden.py), against float64 brute force, against hand
known-answer cases, and -- where /root/reference exists -- against the reference's functions live."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as O
from oracle import f64, ref_lift

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def cube(B, N, seed):
    return (np.random.default_rng(seed).random((B, N, 3), dtype=np.float32) * 2 - 1)


# ---------------------------------------------------------------- golden fixtures ------------

def test_golden_knn_point_reference_torch():
    g = gold("golden_knn_point.npz")
    D, I = O.knn(g["ref"], g["query"], int(g["k"]))
    assert np.array_equal(np.sort(I, -1), g["idx_sorted_by_index"])
    assert (np.diff(D, axis=-1) >= 0).all()  # ascending, unlike the unsorted torch topk


def test_golden_fps_reference_numpy():
    g = gold("golden_fps_numpy.npz")
    idx = O.fps(g["xyz"], int(g["npoint"]))
    picked = np.take_along_axis(g["xyz"], idx[:, :, None].astype(np.int64), 1)
    assert np.array_equal(picked, g["picked"])
    assert np.array_equal(idx, O.fps(g["xyz"], int(g["npoint"]), block_size=O.upstream_block(g["xyz"].shape[1])))


def test_golden_chamfer_modules_reference_python():
    g = gold("golden_chamfer_modules.npz")
    d1, d2, i1, i2 = O.chamfer_fwd(g["xyz1"], g["xyz2"])
    l2 = d1.astype(np.float64).mean() + d2.astype(np.float64).mean()
    l1 = (np.sqrt(d1.astype(np.float64)).mean() + np.sqrt(d2.astype(np.float64)).mean()) / 2
    assert abs(l2 - g["l2_loss"]) <= 1e-5 * abs(g["l2_loss"])
    assert abs(l1 - g["l1_loss"]) <= 1e-5 * abs(g["l1_loss"])
    np.testing.assert_allclose([d1.mean(), d2.mean()], g["l2_split"], rtol=1e-5)
    # gradients: L2 -> grad_dist = 1/(B*N); L1 -> 1/(4*B*N*sqrt(d))
    gx1, gx2 = O.chamfer_bwd(g["xyz1"], g["xyz2"], i1, i2, np.full_like(d1, 1 / d1.size), np.full_like(d2, 1 / d2.size))
    np.testing.assert_allclose(gx1, g["l2_g1"], rtol=2e-4, atol=1e-8)
    np.testing.assert_allclose(gx2, g["l2_g2"], rtol=2e-4, atol=1e-8)
    gx1, gx2 = O.chamfer_bwd(g["xyz1"], g["xyz2"], i1, i2, 0.25 / d1.size / np.sqrt(d1), 0.25 / d2.size / np.sqrt(d2))
    np.testing.assert_allclose(gx1, g["l1_g1"], rtol=2e-4, atol=1e-8)
    np.testing.assert_allclose(gx2, g["l1_g2"], rtol=2e-4, atol=1e-8)
    # ignore_zeros at B=1: points whose coordinate SUM is zero are dropped first
    z1, z2 = g["z1"], g["z2"]
    k1, k2 = z1[:, z1[0].sum(-1) != 0], z2[:, z2[0].sum(-1) != 0]
    e1, e2, _, _ = O.chamfer_fwd(k1, k2)
    assert abs(e1.mean() + e2.mean() - g["l2_ignore_zeros"]) <= 1e-5 * g["l2_ignore_zeros"]
    assert abs((np.sqrt(e1).mean() + np.sqrt(e2).mean()) / 2 - g["l1_ignore_zeros"]) <= 1e-5 * g["l1_ignore_zeros"]
    f1, f2, _, _ = O.chamfer_fwd(z1, z2)
    assert abs(f1.mean() + f2.mean() - g["l2_keep_zeros"]) <= 1e-5 * max(g["l2_keep_zeros"], 1e-12)


def test_golden_group_reference_python():
    g = gold("golden_group.npz")
    nb, ce, idx, cidx = O.group(g["xyz"], int(g["G"]), int(g["k"]))
    assert np.array_equal(cidx.astype(np.int64), g["center_idx"])
    assert np.array_equal(idx, g["idx"])
    assert np.array_equal(ce, g["center"])
    assert np.array_equal(nb, g["neighborhood"])
    B, N = g["xyz"].shape[:2]
    base = np.arange(B) * N
    assert np.array_equal((idx + base[:, None, None]).reshape(-1), g["flat_idx"])
    assert np.array_equal((cidx.astype(np.int64) + base[:, None]).reshape(-1), g["flat_center_idx"])


# ---------------------------------------------------------------- known-answer cases ---------

def test_fps_unit_square_known_answer():
    sq = np.array([[[1, 1, 1], [2, 1, 1], [2, 2, 1], [1, 2, 1]]], np.float32)
    # start 0; farthest from 0 is the diagonal (2); then 1 and 3 tie at distance 1 -> lowest index
    assert O.fps(sq, 4).tolist() == [[0, 2, 1, 3]]


def test_fps_collinear_and_skip_rule():
    line = np.array([[[1 + i, 0, 0] for i in range(6)]], np.float32)
    assert O.fps(line, 3).tolist() == [[0, 5, 2]]  # midpoint tie 2|3 -> lowest index
    pts = np.array([[[1, 0, 0], [0.01, 0.01, 0.0], [0, 3, 0], [0, 0, 0.02]]], np.float32)
    assert O.fps(pts, 4).tolist() == [[0, 2, 0, 0]]  # points 1 and 3 sit inside |p|^2 <= 1e-3: never picked
    # the skip test is a double compare against 1e-3: float(1e-3) itself is NOT skipped
    edge = np.array([[[5, 0, 0], [np.sqrt(np.float32(1e-3)), 0, 0]]], np.float32)
    mag = np.float32(edge[0, 1, 0]) * np.float32(edge[0, 1, 0])
    assert O.fps(edge, 2).tolist() == [[0, 1 if float(mag) > 1e-3 else 0]]


def test_fps_upstream_thread_major_tie_break_differs_only_on_ties():
    ax = np.arange(5, dtype=np.float32)
    grid = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3) + 1
    low = O.fps(grid, 30, block_size=0)
    up = O.fps(grid, 30, block_size=O.upstream_block(grid.shape[1]))
    assert low.shape == up.shape  # both valid FPS orders; they may differ (exact ties)
    rnd = cube(3, 700, 5)
    assert np.array_equal(O.fps(rnd, 64), O.fps(rnd, 64, block_size=O.upstream_block(700)))
    assert O.upstream_block(700) == 512 and O.upstream_block(1024) == 512 and O.upstream_block(33) == 32


def test_knn_known_answer_and_ties():
    ref = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [3, 4, 0]]], np.float32)
    q = np.array([[[0, 0, 0]]], np.float32)
    D, I = O.knn(ref, q, 5)
    assert I.tolist() == [[[0, 1, 2, 3, 4]]]  # three exact ties at distance 1 keep index order
    assert D.tolist() == [[[0.0, 1.0, 1.0, 1.0, 5.0]]]
    with pytest.raises(ValueError):
        O.knn(ref, q, 6)


def test_chamfer_known_answer():
    a = np.array([[[0, 0, 0], [1, 0, 0]]], np.float32)
    b = np.array([[[0, 0, 1], [1, 0, 0], [1, 0, 0]]], np.float32)
    d1, d2, i1, i2 = O.chamfer_fwd(a, b)
    assert d1.tolist() == [[1.0, 0.0]] and i1.tolist() == [[0, 1]]  # duplicate refs: lowest index
    assert d2.tolist() == [[1.0, 0.0, 0.0]] and i2.tolist() == [[0, 1, 1]]
    gx1, gx2 = O.chamfer_bwd(a, b, i1, i2, np.ones_like(d1), np.ones_like(d2))
    assert gx1[0, 0].tolist() == [0.0, 0.0, -4.0] and gx2[0, 0].tolist() == [0.0, 0.0, 4.0]


# ---------------------------------------------------------------- float64 brute force --------

@pytest.mark.parametrize("N", [1, 3, 5, 31, 33, 511, 513, 1023, 1025])
def test_oracle_vs_float64(N):
    xyz = cube(2, N, N)
    M = min(N, 24)
    fo, ff = O.fps(xyz, M), f64.fps(xyz, M)
    assert (fo == ff).mean() > 0.99  # fp32 near-ties may legitimately reorder a sample
    k = min(N, 8)
    q = cube(2, 6, N + 1)
    D, I = O.knn(xyz, q, k)
    Df, If = f64.knn(xyz, q, k)
    np.testing.assert_allclose(D, Df, rtol=1e-5, atol=1e-7)
    assert (I == If).mean() > 0.99
    b = cube(2, max(N // 2, 1), N + 2)
    d1, d2, i1, i2 = O.chamfer_fwd(xyz, b)
    e1, e2, j1, j2 = f64.chamfer_fwd(xyz, b)
    np.testing.assert_allclose(d1, e1, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(d2, e2, rtol=1e-5, atol=1e-7)
    assert (i1 == j1).mean() > 0.99 and (i2 == j2).mean() > 0.99


def test_chamfer_l1_exact_zero_gives_nan_like_reference():
    a = cube(1, 8, 1)
    b = np.concatenate([a[:, :4], cube(1, 4, 2)], 1)
    d1, d2, i1, i2 = O.chamfer_fwd(b, a)
    assert (d1[0, :4] == 0).all()
    with np.errstate(divide="ignore"):
        g1 = 0.25 / d1.size / np.sqrt(d1)
    g2 = np.zeros_like(d2)
    with np.errstate(invalid="ignore"):
        gx1, _ = O.chamfer_bwd(b, a, i1, i2, g1, g2)
    assert np.isnan(gx1[0, :4]).all() and not np.isnan(gx1[0, 4:]).any()


def test_gather_and_grad():
    rng = np.random.default_rng(0)
    feat = rng.standard_normal((2, 3, 10)).astype(np.float32)
    idx = rng.integers(0, 10, (2, 6)).astype(np.int32)
    out = O.gather(feat, idx)
    assert np.array_equal(out, np.take_along_axis(feat, idx[:, None, :].astype(np.int64).repeat(3, 1), 2))
    go = rng.standard_normal((2, 3, 6)).astype(np.float32)
    gg = O.gather_grad(go, idx, 10)
    want = np.zeros((2, 3, 10), np.float32)
    for b in range(2):
        for j in range(6):
            want[b, :, idx[b, j]] += go[b, :, j]
    np.testing.assert_allclose(gg, want, rtol=1e-6)


# ---------------------------------------------------------------- live reference (container only)

@pytest.mark.skipif(not ref_lift.available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference_functions():
    H = ref_lift.torch_helpers()
    ref, q = cube(2, 300, 1), cube(2, 20, 2)
    sd = H.square_distance(torch.from_numpy(q), torch.from_numpy(ref)).numpy()
    np.testing.assert_allclose(sd, f64.pair_sq(q, ref), rtol=1e-4, atol=1e-5)  # expanded form: looser
    D, I = O.knn(ref, q, 16)
    idx = H.knn_point(16, torch.from_numpy(ref), torch.from_numpy(q)).numpy()
    assert (np.sort(idx, -1) == np.sort(I, -1)).mean() > 0.99
    got = H.index_points(torch.from_numpy(ref), torch.from_numpy(I)).numpy()
    assert np.array_equal(got, np.take_along_axis(ref[:, None], I[..., None], 2))


@pytest.mark.skipif(not ref_lift.available(), reason="/root/reference not present (GPU box)")
def test_reference_chamfer_modules_run_unmodified_over_oracle():
    class Impl:
        @staticmethod
        def forward(a, b):
            return [torch.from_numpy(x) for x in O.chamfer_fwd(a.detach().numpy(), b.detach().numpy())]

        @staticmethod
        def backward(a, b, i1, i2, g1, g2):
            return [torch.from_numpy(x) for x in O.chamfer_bwd(a.detach().numpy(), b.detach().numpy(), i1.numpy(),
                                                                 i2.numpy(), g1.contiguous().numpy(), g2.contiguous().numpy())]
    M = ref_lift.chamfer_modules(Impl)
    g = gold("golden_chamfer_modules.npz")
    a = torch.from_numpy(g["xyz1"]).requires_grad_(True)
    b = torch.from_numpy(g["xyz2"]).requires_grad_(True)
    loss = M.ChamferDistanceL1()(a, b)
    loss.backward()
    assert abs(loss.item() - g["l1_loss"]) <= 1e-5 * g["l1_loss"]
    np.testing.assert_allclose(a.grad.numpy(), g["l1_g1"], rtol=2e-4, atol=1e-8)
