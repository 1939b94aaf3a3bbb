"""GPU parity tests proper: the CUDA path (through the C ABI, via upp_b200.ops) against the CPU
oracle on the same seeded inputs.  Bar: indices bit-exact, values within 1e-5 relative (fp32),
Chamfer distances bit-exact (same fma spelling as the reference kernel)."""
import numpy as np
import pytest
import torch

from conftest import unit_sphere

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # BASELINE.json north_star: "within 1e-5 relative in fp32"


@pytest.fixture(scope="module")
def U():
    import upp_b200
    return upp_b200


@pytest.fixture(scope="module")
def O():
    from oracle import c_oracle
    return c_oracle


def cube(B, N, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, N, 3, generator=g) * 2 - 1


# ------------------------------------------------------------------ FPS ----------------------

FPS_SHAPES = [  # (B, N, M)  -- SURVEY.md Appendix A census + edges
    (3, 1, 1), (2, 5, 5), (2, 31, 7), (4, 32, 32), (4, 33, 16), (4, 64, 32), (3, 511, 64), (3, 513, 64),
    (3, 972, 32), (4, 1024, 64), (2, 1023, 256), (2, 1025, 64), (3, 1096, 32), (2, 1228, 1024),
    (2, 1459, 32), (2, 1536, 128), (2, 1624, 32), (2, 1843, 1536), (2, 2048, 128), (1, 6144, 1024),
    (2, 8192, 1024), (2, 2500, 300), (2, 3000, 64), (2, 4096, 200), (2, 4100, 64), (1, 5000, 300), (1, 7000, 64),
]


FPS_KERNELS = {  # name -> tuning environment (read per launch by fps_launch)
    "default": {},                                                        # heuristic: v2 warps x pairs; large clouds: clusters / deferred search
    "pruned": {"UPP_FPS_PRUNED": "1"},                                    # Morton buckets + exact pruning (experiment, N > 2048)
    "pruned_from_256": {"UPP_FPS_PRUNED": "1", "UPP_FPS_PRUNED_MIN": "255"},  # ... forced onto small clouds too
    "pruned_nw4": {"UPP_FPS_PRUNED": "1", "UPP_FPS_PRUNED_NW": "4", "UPP_FPS_PRUNED_MIN": "255"},
    "pruned_nw16": {"UPP_FPS_PRUNED": "1", "UPP_FPS_PRUNED_NW": "16"},
    "rows": {"UPP_FPS_PRUNED": "2"},                                      # register-resident rows + exact pruning (N > 2048)
    "rows_from_64": {"UPP_FPS_PRUNED": "2", "UPP_FPS_PRUNED_MIN": "63"},  # ... forced onto small clouds too
    "v2_nw1": {"UPP_FPS_NW": "1", "UPP_FPS_P2": "8"},                     # single warp, no barrier (N <= 512)
    "v2_nw2": {"UPP_FPS_NW": "2", "UPP_FPS_P2": "8"},                     # (N <= 1024)
    "v2_nw4_redux": {"UPP_FPS_NW": "4", "UPP_FPS_P2": "8", "UPP_FPS_S2": "1"},
    "v2_nw8_redux": {"UPP_FPS_NW": "8", "UPP_FPS_P2": "8", "UPP_FPS_S2": "1"},
    "v2_nw8_scan": {"UPP_FPS_NW": "8", "UPP_FPS_P2": "7", "UPP_FPS_S2": "0"},
    "v2_nw16_scan": {"UPP_FPS_NW": "16", "UPP_FPS_P2": "8", "UPP_FPS_S2": "0"},
    "v2_nw16_redux": {"UPP_FPS_NW": "16", "UPP_FPS_P2": "8", "UPP_FPS_S2": "1"},
    "v2_nw32": {"UPP_FPS_NW": "32", "UPP_FPS_P2": "4", "UPP_FPS_S2": "1"},
    # deferred decision-tree slot search (large clouds)
    "v2_tree_nw8_p16": {"UPP_FPS_NW": "8", "UPP_FPS_P2": "16", "UPP_FPS_S2": "0", "UPP_FPS_SEARCH": "2"},
    "v2_tree_nw8_p12_redux": {"UPP_FPS_NW": "8", "UPP_FPS_P2": "12", "UPP_FPS_S2": "1", "UPP_FPS_SEARCH": "2"},
    "v2_tree_nw8_p5": {"UPP_FPS_NW": "8", "UPP_FPS_P2": "5", "UPP_FPS_S2": "0", "UPP_FPS_SEARCH": "2"},
    "v2_tree_nw16_p8": {"UPP_FPS_NW": "16", "UPP_FPS_P2": "8", "UPP_FPS_S2": "1", "UPP_FPS_SEARCH": "2"},
    "v2_tree_nw32_p4": {"UPP_FPS_NW": "32", "UPP_FPS_P2": "4", "UPP_FPS_S2": "1", "UPP_FPS_SEARCH": "2"},
    # one cloud per cluster of 2 / 4 / 8 CTAs (DSMEM exchange; slices of >= 256 points, else the heuristic's choice)
    "cluster2": {"UPP_FPS_CLUSTER": "2"},
    "cluster4": {"UPP_FPS_CLUSTER": "4"},
    "cluster8": {"UPP_FPS_CLUSTER": "8"},
    "cluster2_nw8": {"UPP_FPS_CLUSTER": "2", "UPP_FPS_CLUSTER_NW": "8"},  # 4 / 8 warps per CTA (8: clusters of <= 4 CTAs)
    "cluster4_nw8": {"UPP_FPS_CLUSTER": "4", "UPP_FPS_CLUSTER_NW": "8"},
    "cluster4_nw4": {"UPP_FPS_CLUSTER": "4", "UPP_FPS_CLUSTER_NW": "4"},
    # up to two selections per exchange round (one look-ahead; measured experiment, off by default)
    "cluster4_ahead": {"UPP_FPS_CLUSTER": "4", "UPP_FPS_CLUSTER_AHEAD": "1"},
    "cluster8_ahead": {"UPP_FPS_CLUSTER": "8", "UPP_FPS_CLUSTER_AHEAD": "1"},
    "v1": {"UPP_FPS_IMPL": "1", "UPP_FPS_W4": "0"},                       # round-1a strided kernels
    "v1_w4": {"UPP_FPS_IMPL": "1", "UPP_FPS_W4": "1"},
}


@pytest.fixture(params=sorted(FPS_KERNELS))
def fps_kernel(request, monkeypatch):
    """Every register-resident FPS kernel family must agree with the oracle.  A forced v2 shape that
    cannot hold the cloud (warps x 64 x pairs < N) falls back to the heuristic's choice."""
    for k, v in FPS_KERNELS[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


@pytest.mark.parametrize("B,N,M", FPS_SHAPES)
def test_fps_matches_oracle_cube(U, O, dev, fps_kernel, B, N, M):
    xyz = cube(B, N, 100 + N + M)
    got = U.ops.fps(xyz.to(dev), M).cpu().numpy()
    assert got.dtype == np.int32 and got.shape == (B, M)
    assert np.array_equal(got, O.fps(xyz.numpy(), M))


@pytest.mark.parametrize("B,N,M", [(4, 1024, 64), (2, 2048, 128), (2, 8192, 1024), (3, 1096, 32)])
def test_fps_unit_sphere_skip_quirk(U, O, dev, fps_kernel, B, N, M):
    """Dataset-normalised clouds have points with |p|^2 <= 1e-3: never selected (upstream quirk)."""
    xyz = unit_sphere(torch.randn(B, N, 3, generator=torch.Generator().manual_seed(7)) * 0.3)
    near = (xyz.pow(2).sum(-1) <= 1e-3)
    assert near.any(), "test input must exercise the skip rule"
    got = U.ops.fps(xyz.to(dev), M).cpu().numpy()
    assert np.array_equal(got, O.fps(xyz.numpy(), M))
    sel = torch.from_numpy(got).long()
    picked_near = torch.gather(near, 1, sel)[:, 1:]
    assert not picked_near.any()


def test_fps_large_batch_heuristic_path(U, O, dev):
    xyz = cube(300, 200, 77)  # B >= 2*148 switches to the 4-warp kernel
    assert np.array_equal(U.ops.fps(xyz.to(dev), 24).cpu().numpy(), O.fps(xyz.numpy(), 24))


def test_fps_exact_ties_lowest_index(U, O, dev, fps_kernel):
    """Integer lattice => many exactly equal distances; tie-break must be the lowest index."""
    ax = torch.arange(8, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3) + 1.0
    xyz = torch.cat([grid, grid.flip(1)], 0).contiguous()
    got = U.ops.fps(xyz.to(dev), 100).cpu().numpy()
    assert np.array_equal(got, O.fps(xyz.numpy(), 100))


def test_fps_pruned_large_lattice_ties_and_degenerate_clouds(U, O, dev, fps_kernel):
    """Large clouds (the Morton-bucket kernel by default): a 16^3 lattice (thousands of exactly equal distances: the
    arg-max must resolve to the lowest ORIGINAL index although the kernel works on a sorted copy), a cloud of duplicates,
    a flat cloud (one axis constant), every point inside the skip radius, and M > N."""
    ax = torch.arange(16, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3) * 0.25 + 0.5
    perm = torch.randperm(4096, generator=torch.Generator().manual_seed(3))
    xyz = torch.cat([grid, grid[:, perm]], 0).contiguous()
    assert np.array_equal(U.ops.fps(xyz.to(dev), 300).cpu().numpy(), O.fps(xyz.numpy(), 300))
    dup = cube(2, 3000, 8)
    dup[:, 1500:] = dup[:, :1500]
    assert np.array_equal(U.ops.fps(dup.to(dev), 200).cpu().numpy(), O.fps(dup.numpy(), 200))
    flat = cube(2, 2600, 9)
    flat[..., 2] = 0.25
    assert np.array_equal(U.ops.fps(flat.to(dev), 100).cpu().numpy(), O.fps(flat.numpy(), 100))
    tiny = cube(1, 2300, 10) * 0.01
    assert np.array_equal(U.ops.fps(tiny.to(dev), 40).cpu().numpy(), O.fps(tiny.numpy(), 40))
    small = cube(1, 2100, 11)
    assert np.array_equal(U.ops.fps(small.to(dev), 2200).cpu().numpy(), O.fps(small.numpy(), 2200))
    idx, cen = U.ops.fps(xyz.to(dev), 64, True)
    assert torch.equal(cen.cpu(), torch.gather(xyz, 1, idx.cpu().long()[..., None].expand(-1, -1, 3)))


def test_fps_duplicates_all_skipped_and_m_gt_n(U, O, dev):
    xyz = cube(2, 40, 3)
    xyz[:, 20:] = xyz[:, :20]  # duplicated points
    assert np.array_equal(U.ops.fps(xyz.to(dev), 40).cpu().numpy(), O.fps(xyz.numpy(), 40))
    tiny = cube(2, 64, 4) * 0.01  # every point inside the skip radius -> index 0 forever
    got = U.ops.fps(tiny.to(dev), 8).cpu().numpy()
    assert np.array_equal(got, np.zeros((2, 8), np.int32))
    assert np.array_equal(got, O.fps(tiny.numpy(), 8))
    small = cube(2, 10, 5)
    assert np.array_equal(U.ops.fps(small.to(dev), 25).cpu().numpy(), O.fps(small.numpy(), 25))


@pytest.mark.parametrize("B,N,M", [(16, 8192, 300), (32, 6144, 200), (37, 4096, 64), (18, 5121, 64), (74, 3073, 16), (75, 3100, 16)])
def test_fps_cluster_heuristic_batches(U, O, dev, B, N, M):
    """Batches the heuristic itself sends to the cluster kernel (N > 3072; 8 CTAs while B * 8 <= 148, else 4 while B * 4 <= 296) and its edges, with the
    fused centre gather, against the oracle."""
    xyz = unit_sphere(torch.randn(B, N, 3, generator=torch.Generator().manual_seed(B + N)) * 0.3)
    idx, centers = U.ops.fps(xyz.to(dev), M, True)
    want = O.fps(xyz.numpy(), M)
    assert np.array_equal(idx.cpu().numpy(), want)
    assert torch.equal(centers.cpu(), torch.gather(xyz, 1, torch.from_numpy(want).long()[..., None].expand(-1, -1, 3)))


def test_fps_large_n_workspace_path(U, O, dev):
    xyz = cube(2, 10000, 11)
    assert np.array_equal(U.ops.fps(xyz.to(dev), 48).cpu().numpy(), O.fps(xyz.numpy(), 48))


def test_fps_centers_and_misc_fps_grad(U, O, dev):
    xyz = cube(3, 500, 12)
    x = xyz.to(dev).requires_grad_(True)
    centers, idx = U.fps(x, 20)
    o_idx = O.fps(xyz.numpy(), 20)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    want = np.take_along_axis(xyz.numpy(), o_idx[:, :, None].astype(np.int64), 1)
    assert np.array_equal(centers.detach().cpu().numpy(), want)
    w = torch.randn(3, 20, 3, generator=torch.Generator().manual_seed(1)).to(dev)
    (centers * w).sum().backward()
    g = np.zeros((3, 500, 3), np.float32)
    for b in range(3):
        for j in range(20):
            g[b, o_idx[b, j]] += w[b, j].cpu().numpy()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g, rtol=RTOL, atol=0)


def test_rows_scatter_add_equals_transposed_gather_grad(U, O, dev):
    """The row-major gradient of misc.fps's gather == gather_operation's backward on the transposed tensors."""
    g = torch.Generator().manual_seed(21)
    gr = torch.randn(3, 50, 3, generator=g)
    idx = torch.randint(0, 70, (3, 50), generator=g, dtype=torch.int32)  # duplicates accumulate
    got = U.ops.rows_scatter_add(gr.to(dev), idx.to(dev), 70)
    want = O.gather_grad(gr.transpose(1, 2).contiguous().numpy(), idx.numpy(), 70).transpose(0, 2, 1)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-6, atol=1e-6)
    wide = torch.randn(2, 9, 17, generator=g)
    i2 = torch.randint(0, 12, (2, 9), generator=g, dtype=torch.int32)
    want2 = torch.zeros(2, 12, 17).scatter_add_(1, i2.long().unsqueeze(-1).expand(-1, -1, 17), wide)
    np.testing.assert_allclose(U.ops.rows_scatter_add(wide.to(dev), i2.to(dev), 12).cpu().numpy(), want2.numpy(), rtol=1e-6, atol=1e-6)


def test_fps_empty(U, dev):
    assert U.ops.fps(torch.zeros(0, 16, 3, device=dev), 4).shape == (0, 4)
    assert U.ops.fps(torch.zeros(2, 16, 3, device=dev), 0).shape == (2, 0)


# ------------------------------------------------------------------ gather -------------------


def test_gather_fwd_bwd(U, O, dev):
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(3, 7, 50, generator=g)
    idx = torch.randint(0, 50, (3, 20), generator=g, dtype=torch.int32)
    idx[0, :5] = 3  # duplicates accumulate in the gradient
    f = feat.to(dev).requires_grad_(True)
    out = U.pointnet2_utils.gather_operation(f, idx.to(dev))
    assert np.array_equal(out.detach().cpu().numpy(), O.gather(feat.numpy(), idx.numpy()))
    go = torch.randn(3, 7, 20, generator=g)
    out.backward(go.to(dev))
    np.testing.assert_allclose(f.grad.cpu().numpy(), O.gather_grad(go.numpy(), idx.numpy(), 50),
                               rtol=RTOL, atol=1e-7)


# ------------------------------------------------------------------ kNN ----------------------

KNN_SHAPES = [  # (B, N, Q, k)
    (2, 32, 32, 16), (2, 32, 32, 32), (3, 64, 32, 8), (2, 972, 32, 16), (4, 1024, 64, 32),
    (2, 1096, 32, 16), (2, 1536, 128, 32), (2, 2048, 128, 32), (2, 1624, 32, 16), (1, 5000, 70, 32),
    (2, 33, 5, 1), (2, 100, 9, 31), (2, 3, 4, 3), (1, 300, 10, 48), (1, 200, 3, 200),
]


@pytest.mark.parametrize("B,N,Q,k", KNN_SHAPES)
def test_knn_matches_oracle(U, O, dev, B, N, Q, k):
    ref = cube(B, N, 200 + N)
    qry = torch.cat([ref[:, : Q // 2], cube(B, Q - Q // 2, 300 + Q)], 1).contiguous()  # half are ref points
    D, I = U.ops.knn(ref.to(dev), qry.to(dev), k)
    oD, oI = O.knn(ref.numpy(), qry.numpy(), k)
    assert I.dtype == torch.int64 and tuple(I.shape) == (B, Q, k)
    assert np.array_equal(I.cpu().numpy(), oI)
    assert np.array_equal(D.cpu().numpy(), oD)  # same fma order + IEEE sqrt -> bit-exact


def test_knn_ties_keep_lower_index(U, O, dev):
    ax = torch.arange(6, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3).contiguous()
    qry = grid[:, ::7].contiguous()
    D, I = U.ops.knn(grid.to(dev), qry.to(dev), 32)
    oD, oI = O.knn(grid.numpy(), qry.numpy(), 32)
    assert np.array_equal(I.cpu().numpy(), oI)
    assert np.array_equal(D.cpu().numpy(), oD)


def test_knn_module_modes_and_errors(U, O, dev):
    ref, qry = cube(2, 128, 1), cube(2, 16, 2)
    oD, oI = O.knn(ref.numpy(), qry.numpy(), 8)
    D, I = U.KNN(k=8, transpose_mode=True)(ref.to(dev), qry.to(dev))
    assert np.array_equal(I.cpu().numpy(), oI)
    D2, I2 = U.KNN(k=8, transpose_mode=False)(ref.transpose(1, 2).to(dev), qry.transpose(1, 2).to(dev))
    assert tuple(I2.shape) == (2, 8, 16)
    assert np.array_equal(I2.transpose(1, 2).cpu().numpy(), oI)
    assert np.array_equal(D2.transpose(1, 2).cpu().numpy(), oD)
    with pytest.raises(ValueError):
        U.ops.knn(ref.to(dev), qry.to(dev), 129)  # k > N: upstream UB, rejected here
    with pytest.raises(RuntimeError):
        U.ops.knn(ref, qry, 4)  # CPU tensors: no fallback


# ------------------------------------------------------------------ Chamfer ------------------

CH_SHAPES = [  # (B, N, M)
    (4, 64, 128), (2, 1, 1), (3, 5, 3), (2, 32, 1024), (2, 1024, 1024), (2, 2048, 2048), (1, 2048, 8192),
    (2, 513, 2049), (2, 1000, 4099), (3, 300, 200), (2, 128, 128), (2, 129, 255), (3, 257, 131), (70, 256, 300),
]


CH_KERNELS = {  # name -> tuning environment (read per launch by chamfer_fwd_launch)
    "slots": {},                                                    # default: plain-store partial slots (no atomics, no memset) + one finalize launch
    "slots_1chunk": {"UPP_CH_CHUNKS": "1"},                         # rows resolved directly by the CTA that scanned them
    "slots_3chunks": {"UPP_CH_CHUNKS": "3"},                        # row side through per-chunk partials too
    "slots_many_chunks": {"UPP_CH_CHUNKS": "32"},
    "slots_cg16": {"UPP_CH_FUSED": "5", "UPP_CH_CHUNKS": "3"},      # row-side groups of 16 columns
    "slots_cg8_5persm": {"UPP_CH_FUSED": "6"},
    "folded": {"UPP_CH_FUSED": "1"},                                # finalize + sums folded into the distance kernel (tickets)
    "folded_1chunk": {"UPP_CH_FUSED": "4", "UPP_CH_CHUNKS": "1"},
    "folded_3chunks": {"UPP_CH_FUSED": "2", "UPP_CH_CHUNKS": "3"},
    "folded_many_chunks": {"UPP_CH_FUSED": "3", "UPP_CH_CHUNKS": "32"},
    "packed": {"UPP_CH_VARIANT": "30"},                             # round-1 path: memset + packed kernel (RED.MIN keys) + finalize launch
    "packed_1chunk": {"UPP_CH_VARIANT": "30", "UPP_CH_CHUNKS": "1"},
    "packed_3chunks": {"UPP_CH_VARIANT": "30", "UPP_CH_CHUNKS": "3"},
    "packed_many_chunks": {"UPP_CH_VARIANT": "30", "UPP_CH_CHUNKS": "32"},
    "packed_r4": {"UPP_CH_VARIANT": "31", "UPP_CH_CHUNKS": "2"},
    "packed_r6": {"UPP_CH_VARIANT": "33"},
    "packed_w8": {"UPP_CH_VARIANT": "32"},
    "packed_r12": {"UPP_CH_VARIANT": "34"},
    "packed_r16": {"UPP_CH_VARIANT": "35", "UPP_CH_CHUNKS": "2"},
    "packed_r16w2": {"UPP_CH_VARIANT": "36"},
    "scalar_single_pass": {"UPP_CH_VARIANT": "20"},                 # round-1b kernel
    "two_pass": {"UPP_CH_VARIANT": "0"},                            # directed kernel (no workspace needed)
}


@pytest.fixture(params=sorted(CH_KERNELS))
def chamfer_path(request, monkeypatch):
    """Every forward kernel family must give identical (bit-exact) results."""
    for k, v in CH_KERNELS[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


@pytest.mark.parametrize("B,N,M", CH_SHAPES)
def test_chamfer_forward_bit_exact(U, O, dev, chamfer_path, B, N, M):
    g = torch.Generator().manual_seed(N * 7 + M)
    a, b = torch.rand(B, N, 3, generator=g), torch.rand(B, M, 3, generator=g)
    d1, d2, i1, i2 = [t.cpu().numpy() for t in U.chamfer.forward(a.to(dev), b.to(dev))]
    o1, o2, j1, j2 = O.chamfer_fwd(a.numpy(), b.numpy())
    assert i1.dtype == np.int32 and i2.dtype == np.int32
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2)
    assert np.array_equal(d1, o1) and np.array_equal(d2, o2)


def test_chamfer_ties_and_duplicates_both_kernels(U, O, dev, chamfer_path):
    """Lattice points + duplicated points: many exactly equal distances in both directions."""
    ax = torch.arange(7, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3)
    a = torch.cat([grid, grid[:, :57]], 1).repeat(3, 1, 1).contiguous()          # 400 points, 57 duplicates
    b = (grid[:, torch.randperm(343, generator=torch.Generator().manual_seed(1))] + 0.5).repeat(3, 1, 1).contiguous()
    d1, d2, i1, i2 = [t.cpu().numpy() for t in U.chamfer.forward(a.to(dev), b.to(dev))]
    o1, o2, j1, j2 = O.chamfer_fwd(a.numpy(), b.numpy())
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2)
    assert np.array_equal(d1, o1) and np.array_equal(d2, o2)
    d1, d2, i1, i2 = [t.cpu().numpy() for t in U.chamfer.forward(a.to(dev), a.flip(1).contiguous().to(dev))]
    o1, o2, j1, j2 = O.chamfer_fwd(a.numpy(), a.flip(1).contiguous().numpy())
    assert np.array_equal(i1, j1) and np.array_equal(i2, j2) and np.array_equal(d1, o1) and np.array_equal(d2, o2)


def test_chamfer_against_reference_cuda_extension(U, dev, chamfer_path):
    """The reference's own chamfer.cu, compiled unmodified into oracle/_ref (when it travelled)."""
    from oracle import ref_gpu
    ref = ref_gpu.load()
    if ref is None:
        pytest.skip("oracle/_ref/chamfer_ref*.so not built")
    torch.cuda.set_device(0)  # the reference allocates on the current device, default stream
    for (B, N, M, seed) in [(4, 64, 128, 0), (8, 1024, 1024, 1), (4, 2048, 2048, 2), (2, 2048, 8192, 3),
                            (3, 777, 1301, 4)]:
        g = torch.Generator().manual_seed(seed)
        a, b = torch.rand(B, N, 3, generator=g).to(dev), torch.rand(B, M, 3, generator=g).to(dev)
        mine = U.chamfer.forward(a, b)
        torch.cuda.synchronize()
        theirs = ref.forward(a, b)
        torch.cuda.synchronize()
        for x, y in zip(mine, theirs):
            assert torch.equal(x, y)
        g1, g2 = torch.rand(B, N, generator=g).to(dev), torch.rand(B, M, generator=g).to(dev)
        mg = U.chamfer.backward(a, b, mine[2], mine[3], g1, g2)
        torch.cuda.synchronize()
        tg = ref.backward(a, b, theirs[2], theirs[3], g1, g2)
        torch.cuda.synchronize()
        for x, y in zip(mg, tg):  # both sides sum with float atomics in unspecified order
            torch.testing.assert_close(x, y, rtol=1e-4, atol=1e-6)


def test_chamfer_backward_matches_oracle(U, O, dev):
    g = torch.Generator().manual_seed(9)
    a, b = torch.rand(3, 400, 3, generator=g), torch.rand(3, 250, 3, generator=g)
    g1, g2 = torch.randn(3, 400, generator=g), torch.randn(3, 250, generator=g)
    _, _, i1, i2 = U.chamfer.forward(a.to(dev), b.to(dev))
    gx1, gx2 = U.chamfer.backward(a.to(dev), b.to(dev), i1, i2, g1.to(dev), g2.to(dev))
    o1, o2 = O.chamfer_bwd(a.numpy(), b.numpy(), i1.cpu().numpy(), i2.cpu().numpy(), g1.numpy(), g2.numpy())
    np.testing.assert_allclose(gx1.cpu().numpy(), o1, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gx2.cpu().numpy(), o2, rtol=1e-4, atol=1e-6)


def test_chamfer_modules_l1_l2_and_autograd(U, O, dev):
    g = torch.Generator().manual_seed(10)
    a, b = torch.rand(4, 256, 3, generator=g), torch.rand(4, 300, 3, generator=g)
    o1, o2, _, _ = O.chamfer_fwd(a.numpy(), b.numpy())
    l2 = U.ChamferDistanceL2()(a.to(dev), b.to(dev)).item()
    l1 = U.ChamferDistanceL1()(a.to(dev), b.to(dev)).item()
    s1, s2 = U.ChamferDistanceL2_split()(a.to(dev), b.to(dev))
    assert abs(l2 - (o1.mean() + o2.mean())) <= RTOL * abs(l2)
    assert abs(l1 - (np.sqrt(o1).mean() + np.sqrt(o2).mean()) / 2) <= RTOL * abs(l1)
    assert abs(s1.item() - o1.mean()) <= RTOL * abs(o1.mean()) and abs(s2.item() - o2.mean()) <= RTOL * abs(o2.mean())
    # gradient through the non-contiguous expanded grad of mean(): compare with float64 autograd
    ad = a.to(dev).requires_grad_(True)
    bd = b.to(dev).requires_grad_(True)
    U.ChamferDistanceL2()(ad, bd).backward()
    a64 = a.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    dm = (a64[:, :, None, :] - b64[:, None, :, :]).pow(2).sum(-1)
    (dm.min(2)[0].mean() + dm.min(1)[0].mean()).backward()
    torch.testing.assert_close(ad.grad.cpu().double(), a64.grad, rtol=1e-4, atol=1e-9)
    torch.testing.assert_close(bd.grad.cpu().double(), b64.grad, rtol=1e-4, atol=1e-9)


def test_chamfer_l1_exact_zero_nan_rows(U, dev):
    """partial subset of gt (runner_pretask.py:223): d == 0 -> sqrt' = inf -> inf*0 = NaN in exactly
    the rows the reference poisons, nowhere else."""
    g = torch.Generator().manual_seed(11)
    gt = torch.rand(2, 128, 3, generator=g)
    pred = torch.cat([gt[:, :32], torch.rand(2, 32, 3, generator=g)], 1).contiguous()
    p = pred.to(dev).requires_grad_(True)
    U.ChamferDistanceL1()(p, gt.to(dev)).backward()
    nan_rows = torch.isnan(p.grad).any(-1).cpu()
    assert nan_rows[:, :32].all()
    assert not nan_rows[:, 32:].any()


def test_chamfer_ignore_zeros_b1(U, O, dev):
    g = torch.Generator().manual_seed(12)
    a, b = torch.rand(1, 100, 3, generator=g), torch.rand(1, 120, 3, generator=g)
    a[0, 60:] = 0  # zero-padded rows as produced by misc.random_dropping (utils/misc.py:313-314)
    got = U.ChamferDistanceL2(ignore_zeros=True)(a.to(dev), b.to(dev)).item()
    o1, o2, _, _ = O.chamfer_fwd(a[:, :60].numpy(), b.numpy())
    assert abs(got - (o1.mean() + o2.mean())) <= RTOL * abs(got)


def test_chamfer_empty_and_sums(U, O, dev):
    out = U.chamfer.forward(torch.zeros(2, 0, 3, device=dev), torch.rand(2, 5, 3, device=dev))
    assert out[0].shape == (2, 0) and torch.count_nonzero(out[1]) == 0 and torch.count_nonzero(out[3]) == 0
    g = torch.Generator().manual_seed(13)
    a, b = torch.rand(5, 700, 3, generator=g), torch.rand(5, 900, 3, generator=g)
    d1, d2, _, _, sums = U.ops.chamfer_forward(a.to(dev), b.to(dev), want_sums=True)
    want = np.array([d1.double().sum().item(), d2.double().sum().item(),
                     d1.double().sqrt().sum().item(), d2.double().sqrt().sum().item()])
    np.testing.assert_allclose(sums.cpu().numpy(), want, rtol=1e-5)


@pytest.mark.parametrize("B,N,M", [(5, 700, 900), (5, 900, 700), (64, 32, 1024), (3, 1024, 32), (2, 2048, 2048),
                                   (2, 40, 50), (1, 1, 300), (7, 257, 4099)])
def test_chamfer_partial_sums_every_path(U, dev, chamfer_path, B, N, M):
    """{sum d1, sum d2, sum sqrt d1, sum sqrt d2} from the fused finalize (packed paths: ticketed two-level
    reduction) or the cluster kernel (other paths): right slot whichever cloud became the row side,
    identical bits run to run, and the dist/idx outputs unchanged by asking for the sums."""
    g = torch.Generator().manual_seed(B * 1000 + N + M)
    a, b = torch.rand(B, N, 3, generator=g).to(dev), torch.rand(B, M, 3, generator=g).to(dev)
    d1, d2, i1, i2, sums = U.ops.chamfer_forward(a, b, want_sums=True)
    e1, e2, j1, j2 = U.ops.chamfer_forward(a, b)
    assert torch.equal(d1, e1) and torch.equal(d2, e2) and torch.equal(i1, j1) and torch.equal(i2, j2)
    want = np.array([d1.double().sum().item(), d2.double().sum().item(),
                     d1.double().sqrt().sum().item(), d2.double().sqrt().sum().item()])
    np.testing.assert_allclose(sums.cpu().numpy(), want, rtol=2e-5)
    for _ in range(3):
        again = U.ops.chamfer_forward(a, b, want_sums=True)[4]
        assert torch.equal(again, sums), "partial sums must be run-to-run deterministic"


# ------------------------------------------------------------------ Group --------------------


@pytest.mark.parametrize("B,N,G,k", [(4, 1024, 64, 32), (3, 1096, 32, 16), (3, 32, 32, 16), (3, 64, 32, 8),
                                     (2, 2048, 128, 32), (2, 972, 32, 16)])
@pytest.mark.parametrize("fused", [True, False])
def test_group_matches_oracle(U, O, dev, B, N, G, k, fused):
    xyz = unit_sphere(cube(B, N, 400 + N))
    o_nb, o_ce, o_idx, o_cidx = O.group(xyz.numpy(), G, k)
    grp = U.Group(G, k, fused=fused)
    nb, ce, idx, cidx = grp(xyz.to(dev), require_index=True, gather_idx=True)
    assert np.array_equal(cidx.cpu().numpy(), o_cidx.astype(np.int64))
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(ce.cpu().numpy(), o_ce)
    assert np.array_equal(nb.cpu().numpy(), o_nb)
    # flat index convention (gather_idx=False): + b*N base
    nb2, ce2, fidx, fcidx = grp(xyz.to(dev), require_index=True, gather_idx=False)
    base = (np.arange(B) * N)
    assert np.array_equal(fidx.cpu().numpy(), (o_idx + base[:, None, None]).reshape(-1))
    assert np.array_equal(fcidx.cpu().numpy(), (o_cidx.astype(np.int64) + base[:, None]).reshape(-1))
    assert torch.equal(nb2, nb) and torch.equal(ce2, ce)


@pytest.mark.parametrize("fused", [True, False])
def test_group_backward(U, O, dev, fused):
    xyz = cube(2, 200, 21)
    _, _, o_idx, o_cidx = O.group(xyz.numpy(), 8, 4)
    x = xyz.to(dev).requires_grad_(True)
    nb, ce = U.Group(8, 4, fused=fused)(x)
    g = torch.Generator().manual_seed(3)
    wn, wc = torch.randn(2, 8, 4, 3, generator=g), torch.randn(2, 8, 3, generator=g)
    ((nb * wn.to(dev)).sum() + (ce * wc.to(dev)).sum()).backward()
    want = np.zeros((2, 200, 3), np.float64)
    for b in range(2):
        for gg in range(8):
            want[b, o_cidx[b, gg]] += wc[b, gg].numpy() - wn[b, gg].numpy().sum(0)
            for j in range(4):
                want[b, o_idx[b, gg, j]] += wn[b, gg, j].numpy()
    np.testing.assert_allclose(x.grad.cpu().numpy(), want, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ drop-in imports ----------


def test_dropin_modules_resolve_reference_imports(U, O, dev):
    import chamfer  # reference extensions/chamfer_dist/__init__.py:10
    from knn_cuda import KNN  # reference models/Point_MAE_unify.py:16
    from pointnet2_ops import pointnet2_utils  # reference utils/misc.py:10
    xyz = cube(2, 256, 31)
    x = xyz.to(dev)
    idx = pointnet2_utils.furthest_point_sample(x, 16)
    data = pointnet2_utils.gather_operation(x.transpose(1, 2).contiguous(), idx).transpose(1, 2).contiguous()
    o_idx = O.fps(xyz.numpy(), 16)
    assert np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(data.cpu().numpy(), np.take_along_axis(xyz.numpy(), o_idx[:, :, None].astype(np.int64), 1))
    _, I = KNN(k=8, transpose_mode=True)(x, data)
    assert np.array_equal(I.cpu().numpy(), O.knn(xyz.numpy(), data.cpu().numpy(), 8)[1])
    d1, d2, i1, i2 = chamfer.forward(x, data)
    assert torch.count_nonzero(d2) == 0  # sampled points are cloud points: exact zeros
    assert U.launch_count() > 0


# ------------------------------------------------------------------ golden fixtures ----------
# generated from the reference's own Python in the build container (tests/golden/make_golden.py)

def _gold(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


def test_golden_reference_knn_point_and_numpy_fps(U, dev):
    g = _gold("golden_knn_point.npz")
    _, I = U.ops.knn(torch.from_numpy(g["ref"]).to(dev), torch.from_numpy(g["query"]).to(dev), int(g["k"]))
    assert np.array_equal(np.sort(I.cpu().numpy(), -1), g["idx_sorted_by_index"])
    g = _gold("golden_fps_numpy.npz")
    centers, _ = U.fps(torch.from_numpy(g["xyz"]).to(dev), int(g["npoint"]))
    assert np.array_equal(centers.cpu().numpy(), g["picked"])


def test_golden_reference_chamfer_modules(U, dev):
    g = _gold("golden_chamfer_modules.npz")
    for name, mod in (("l1", U.ChamferDistanceL1()), ("l2", U.ChamferDistanceL2())):
        a = torch.from_numpy(g["xyz1"]).to(dev).requires_grad_(True)
        b = torch.from_numpy(g["xyz2"]).to(dev).requires_grad_(True)
        loss = mod(a, b)
        loss.backward()
        assert abs(loss.item() - g[f"{name}_loss"]) <= RTOL * abs(g[f"{name}_loss"])
        np.testing.assert_allclose(a.grad.cpu().numpy(), g[f"{name}_g1"], rtol=2e-4, atol=1e-8)
        np.testing.assert_allclose(b.grad.cpu().numpy(), g[f"{name}_g2"], rtol=2e-4, atol=1e-8)
    s1, s2 = U.ChamferDistanceL2_split()(torch.from_numpy(g["xyz1"]).to(dev), torch.from_numpy(g["xyz2"]).to(dev))
    np.testing.assert_allclose([s1.item(), s2.item()], g["l2_split"], rtol=RTOL)
    z1, z2 = torch.from_numpy(g["z1"]).to(dev), torch.from_numpy(g["z2"]).to(dev)
    assert abs(U.ChamferDistanceL2(ignore_zeros=True)(z1, z2).item() - g["l2_ignore_zeros"]) <= RTOL * g["l2_ignore_zeros"]
    assert abs(U.ChamferDistanceL1(ignore_zeros=True)(z1, z2).item() - g["l1_ignore_zeros"]) <= RTOL * g["l1_ignore_zeros"]
    assert abs(U.ChamferDistanceL2(ignore_zeros=False)(z1, z2).item() - g["l2_keep_zeros"]) <= RTOL * g["l2_keep_zeros"]


@pytest.mark.parametrize("fused", [True, False])
def test_golden_reference_group_forward(U, dev, fused):
    g = _gold("golden_group.npz")
    grp = U.Group(int(g["G"]), int(g["k"]), fused=fused)
    x = torch.from_numpy(g["xyz"]).to(dev)
    nb, ce, idx, cidx = grp(x, require_index=True, gather_idx=True)
    assert np.array_equal(idx.cpu().numpy(), g["idx"]) and np.array_equal(cidx.cpu().numpy(), g["center_idx"])
    assert np.array_equal(nb.cpu().numpy(), g["neighborhood"]) and np.array_equal(ce.cpu().numpy(), g["center"])
    _, _, fidx, fcidx = grp(x, require_index=True, gather_idx=False)
    assert np.array_equal(fidx.cpu().numpy(), g["flat_idx"]) and np.array_equal(fcidx.cpu().numpy(), g["flat_center_idx"])


# ------------------------------------------------------------------ interpolation ---

INTERP_SHAPES = [  # (B, N targets, S sources, C channels, k)
    (3, 64, 32, 384, 8),      # Block propagate: level-1 <- level-2 centres (models/Point_MAE_pretask_dev.py:298)
    (3, 32, 32, 384, 6),      # completion propagate de_neighbors=6 (models/Point_MAE_unify.py:598)
    (2, 2048, 128, 1152, 3),  # seg feature propagation (models/Point_MAE_unify_segment.py:420,605-607)
    (2, 1096, 32, 96, 16),    # rectify prompter propagation (models/Point_MAE_pretask_dev.py:454-461)
    (2, 100, 2500, 10, 5),    # more sources than one staged tile; C not a multiple of 4
    (1, 7, 3, 1, 3), (2, 33, 40, 4, 1), (2, 5, 64, 2304, 32),
]


@pytest.mark.parametrize("B,N,S,C,k", INTERP_SHAPES)
@pytest.mark.parametrize("with_base", [False, True])
def test_interp_forward_matches_oracle(U, O, dev, B, N, S, C, k, with_base):
    g = torch.Generator().manual_seed(N * 31 + S)
    x1, x2 = torch.rand(B, N, 3, generator=g) * 2 - 1, torch.rand(B, S, 3, generator=g) * 2 - 1
    p2 = torch.randn(B, S, C, generator=g)
    base = torch.randn(B, N, C, generator=g) if with_base else None
    alpha, eps = (0.3, 1e-3) if with_base else (1.0, 1e-4)
    out, idx, w, d = U.ops.interp_forward(x1.to(dev), x2.to(dev), p2.to(dev), k, eps,
                                          base=base.to(dev) if with_base else None, alpha=alpha)
    o_out, o_idx, o_w, o_d = O.interp_fwd(x1.numpy(), x2.numpy(), p2.numpy(), k, eps,
                                          base=base.numpy() if with_base else None, alpha=alpha)
    assert idx.dtype == torch.int32 and np.array_equal(idx.cpu().numpy(), o_idx)   # bit-exact selection
    assert np.array_equal(d.cpu().numpy(), o_d) and np.array_equal(w.cpu().numpy(), o_w)
    np.testing.assert_allclose(out.cpu().numpy(), o_out, rtol=RTOL, atol=1e-6)


def test_interp_ties_and_coincident_points(U, O, dev):
    """Targets that ARE sources (level-2 centres are a subset of level-1 centres) and lattice ties."""
    ax = torch.arange(4, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3).repeat(2, 1, 1)
    x2 = grid[:, ::2].contiguous()
    p2 = torch.randn(2, x2.shape[1], 8, generator=torch.Generator().manual_seed(2))
    for eps in (1e-8, 1e-4):
        out, idx, w, d = U.ops.interp_forward(grid.to(dev), x2.to(dev), p2.to(dev), 6, eps)
        o_out, o_idx, o_w, o_d = O.interp_fwd(grid.numpy(), x2.numpy(), p2.numpy(), 6, eps)
        assert np.array_equal(idx.cpu().numpy(), o_idx) and np.array_equal(w.cpu().numpy(), o_w)
        np.testing.assert_allclose(out.cpu().numpy(), o_out, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("B,N,S,C,k", [(3, 64, 32, 384, 8), (2, 700, 48, 1152, 3), (2, 300, 2100, 24, 16), (1, 9, 4, 2500, 4)])
def test_interp_backward_matches_oracle_and_is_deterministic(U, O, dev, B, N, S, C, k):
    g = torch.Generator().manual_seed(S + C)
    x1, x2 = torch.rand(B, N, 3, generator=g) * 2 - 1, torch.rand(B, S, 3, generator=g) * 2 - 1
    p2, go = torch.randn(B, S, C, generator=g), torch.randn(B, N, C, generator=g)
    eps, alpha = 1e-3, 0.3
    out, idx, w, d = U.ops.interp_forward(x1.to(dev), x2.to(dev), p2.to(dev), k, eps, alpha=alpha)
    terms = (d, p2.to(dev), x1.to(dev), x2.to(dev), eps)
    gp2, g1, g2 = U.ops.interp_backward(go.to(dev), idx, w, S, alpha=alpha, xyz_terms=terms)
    o_gp2, o_g1, o_g2 = O.interp_bwd(go.numpy(), p2.numpy(), x1.numpy(), x2.numpy(), idx.cpu().numpy(),
                                     w.cpu().numpy(), d.cpu().numpy(), eps, alpha=alpha)
    scale = max(1.0, float(np.abs(o_gp2).max()))
    np.testing.assert_allclose(gp2.cpu().numpy(), o_gp2, rtol=1e-4, atol=1e-5 * scale)
    for got, want in ((g1, o_g1), (g2, o_g2)):
        s = max(1e-6, float(np.abs(want).max()))
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-3, atol=2e-4 * s)
    only_feat, n1, n2 = U.ops.interp_backward(go.to(dev), idx, w, S, alpha=alpha)
    assert n1 is None and n2 is None and torch.equal(only_feat, gp2)
    for _ in range(2):  # atomic-free: bit-identical run to run
        a, b1, b2 = U.ops.interp_backward(go.to(dev), idx, w, S, alpha=alpha, xyz_terms=terms)
        assert torch.equal(a, gp2) and torch.equal(b1, g1) and torch.equal(b2, g2)


WIDE_SHAPES = [  # shapes the wide-feature kernels accept: C a multiple of 128, few sources
    (2, 2048, 128, 1152, 3),  # seg feature propagation (BASELINE configs[4])
    (3, 64, 32, 384, 8),      # Block propagate
    (2, 333, 150, 256, 5),    # ragged target count, S > 128 (blend only; the streamed backward declines)
    (1, 130, 128, 128, 8),    # one channel chunk, S and k at the streamed backward's limits, N % 64 != 0
    (2, 17, 5, 128, 1),
    (2, 1096, 32, 96, 16),    # rectify-prompter propagation: narrow rows (one 96-channel chunk), thread-per-target k = 16
    (2, 300, 40, 64, 6), (3, 70, 20, 4, 2), (2, 260, 128, 100, 8),
]


@pytest.mark.parametrize("B,N,S,C,k", WIDE_SHAPES)
@pytest.mark.parametrize("with_base", [False, True])
def test_interp_forward_two_phase_is_bit_identical(U, O, dev, monkeypatch, B, N, S, C, k, with_base):
    """Selection + shared-memory blend (interp_blend_kernel, forced by UPP_INTERP_PATH=1) against the one-launch
    kernel (UPP_INTERP_PATH=0): same arithmetic in the same order, so bit-equal; and against the oracle."""
    g = torch.Generator().manual_seed(N * 7 + C)
    x1, x2 = torch.rand(B, N, 3, generator=g) * 2 - 1, torch.rand(B, S, 3, generator=g) * 2 - 1
    p2 = torch.randn(B, S, C, generator=g)
    base = torch.randn(B, N, C, generator=g).to(dev) if with_base else None
    alpha, eps = (0.3, 1e-3) if with_base else (1.0, 1e-4)
    got = {}
    # "1s": the selection by the thread-per-target kernel (k <= 4), "1": by the warp-per-target kernel
    # (1s1 / 1s2 / 1s4: with 1, 2 or 4 threads per target)
    for tag, path, select, tpt in (("0", "0", "0", "1"), ("1", "1", "0", "1"), ("1s1", "1", "1", "1"), ("1s2", "1", "1", "2"),
                                   ("1s4", "1", "1", "4")):
        monkeypatch.setenv("UPP_INTERP_PATH", path)
        monkeypatch.setenv("UPP_INTERP_SELECT", select)
        monkeypatch.setenv("UPP_INTERP_TPT", tpt)
        n0 = U.launch_count()
        got[tag] = U.ops.interp_forward(x1.to(dev), x2.to(dev), p2.to(dev), k, eps, base=base, alpha=alpha)
        assert U.launch_count() - n0 == (1 if path == "0" else 2)
    for tag in ("1", "1s1", "1s2", "1s4"):
        for a, b in zip(got["0"], got[tag]):
            assert torch.equal(a, b), tag
    o_out, o_idx, o_w, o_d = O.interp_fwd(x1.numpy(), x2.numpy(), p2.numpy(), k, eps,
                                          base=base.cpu().numpy() if with_base else None, alpha=alpha)
    assert np.array_equal(got["1s4"][1].cpu().numpy(), o_idx)
    np.testing.assert_allclose(got["1s4"][0].cpu().numpy(), o_out, rtol=RTOL, atol=1e-6)


@pytest.mark.parametrize("B,N,S,C,k", WIDE_SHAPES)
def test_interp_backward_streamed_matches_source_side_kernel(U, O, dev, monkeypatch, B, N, S, C, k):
    """CSR + streamed feature gradient (UPP_INTERP_PATH=1) against the source-side kernel (=0): per source the same
    fmaf sequence in (n, j) order -- bit-equal where the source-side kernel runs one thread group (C > 1024),
    1e-5 otherwise -- deterministic, with and without the coordinate terms, and against the oracle."""
    g = torch.Generator().manual_seed(S * 3 + C)
    x1, x2 = torch.rand(B, N, 3, generator=g) * 2 - 1, torch.rand(B, S, 3, generator=g) * 2 - 1
    p2, go = torch.randn(B, S, C, generator=g), torch.randn(B, N, C, generator=g)
    eps, alpha = 1e-3, 0.3
    out, idx, w, d = U.ops.interp_forward(x1.to(dev), x2.to(dev), p2.to(dev), k, eps, alpha=alpha)
    terms = (d, p2.to(dev), x1.to(dev), x2.to(dev), eps)
    res = {}
    for path in ("0", "1"):
        monkeypatch.setenv("UPP_INTERP_PATH", path)
        n0 = U.launch_count()
        res[path, "feat"] = U.ops.interp_backward(go.to(dev), idx, w, S, alpha=alpha)
        n1 = U.launch_count()
        res[path, "xyz"] = U.ops.interp_backward(go.to(dev), idx, w, S, alpha=alpha, xyz_terms=terms)
        # CSR paths: streamed kernel (128-channel chunks, k <= 8) or the CSR gather kernel (one chunk of <= 128 channels, k <= 16)
        streamed = path == "1" and S <= 128 and ((C % 128 == 0 and k <= 8) or (C <= 128 and C % 4 == 0 and k <= 16))
        acc = streamed and C <= 128 and S * C <= 3072 and not (C % 128 == 0 and k <= 8)  # small source block: accumulate + combine
        # features: csr + stream / gather, or accumulate + combine; with coordinate terms: + target, (csr,) xyz2
        assert n1 - n0 == (2 if streamed else 1) and U.launch_count() - n1 == ((5 if acc else 4) if streamed else 2)
    ref, new = res["0", "feat"][0], res["1", "feat"][0]
    if C > 1024:
        assert torch.equal(ref, new)
    else:
        scale = max(1.0, float(ref.abs().max()))
        np.testing.assert_allclose(new.cpu().numpy(), ref.cpu().numpy(), rtol=1e-5, atol=1e-6 * scale)
    assert torch.equal(res["1", "xyz"][0], new)                       # same feature gradient with the coordinate terms
    assert torch.equal(res["1", "xyz"][1], res["0", "xyz"][1])         # grad_xyz1: the same target-side kernel either way
    want2 = res["0", "xyz"][2]                                         # grad_xyz2: read off the CSR, another summation order
    s2 = max(1e-6, float(want2.abs().max()))
    np.testing.assert_allclose(res["1", "xyz"][2].cpu().numpy(), want2.cpu().numpy(), rtol=1e-4, atol=1e-5 * s2)
    monkeypatch.setenv("UPP_INTERP_PATH", "1")
    for _ in range(2):
        assert torch.equal(U.ops.interp_backward(go.to(dev), idx, w, S, alpha=alpha)[0], new)
        again = U.ops.interp_backward(go.to(dev), idx, w, S, alpha=alpha, xyz_terms=terms)
        assert all(torch.equal(a, b) for a, b in zip(again, res["1", "xyz"]))   # atomic-free: bit-identical run to run
    o_gp2, _, _ = O.interp_bwd(go.numpy(), p2.numpy(), x1.numpy(), x2.numpy(), idx.cpu().numpy(),
                               w.cpu().numpy(), d.cpu().numpy(), eps, alpha=alpha)
    scale = max(1.0, float(np.abs(o_gp2).max()))
    np.testing.assert_allclose(new.cpu().numpy(), o_gp2, rtol=1e-4, atol=1e-5 * scale)


@pytest.mark.parametrize("tpt", ["1", "2", "4"])
def test_interp_thread_select_duplicate_sources_keep_index_order(U, O, dev, monkeypatch, tpt):
    """Sources that coincide (exactly equal distances inside one partner list of the multi-thread merge): the lower source
    index must come first whatever the number of threads per target."""
    g = torch.Generator().manual_seed(9)
    x2 = torch.rand(2, 64, 3, generator=g) * 2 - 1
    x2[:, 1::2] = x2[:, 0::2]               # every source duplicated: pairs of equal distances everywhere
    x2[:, 40:44] = x2[:, 8:9]               # and a run of five coincident sources across the parts
    x1 = torch.rand(2, 300, 3, generator=g) * 2 - 1
    p2 = torch.randn(2, 64, 128, generator=g)
    monkeypatch.setenv("UPP_INTERP_PATH", "1")
    monkeypatch.setenv("UPP_INTERP_SELECT", "1")
    monkeypatch.setenv("UPP_INTERP_TPT", tpt)
    for k in (2, 3, 4):
        out, idx, w, d = U.ops.interp_forward(x1.to(dev), x2.to(dev), p2.to(dev), k, 1e-4)
        o_out, o_idx, o_w, o_d = O.interp_fwd(x1.numpy(), x2.numpy(), p2.numpy(), k, 1e-4)
        assert np.array_equal(idx.cpu().numpy(), o_idx) and np.array_equal(w.cpu().numpy(), o_w)


def test_interp_selection_reuse_is_bit_identical(U, O, dev):
    """upp_interp_select_f32 + upp_interp_blend_f32 (a kept Selection, SURVEY 8f row 1: six SA-unit propagate calls on the
    same geometry) against the one-shot forward: bit-identical outputs, one launch per reuse, gradients unchanged."""
    g = torch.Generator().manual_seed(14)
    for (B, N, S, C, k, eps, with_base) in [(3, 64, 32, 384, 8, 1e-3, True), (2, 2048, 128, 1152, 3, 1e-4, False), (2, 1096, 32, 96, 16, 1e-3, True),
                                            (2, 100, 40, 10, 5, 1e-8, True), (4, 2100, 64, 256, 6, 1e-3, False)]:
        x1, x2 = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev), (torch.rand(B, S, 3, generator=g) * 2 - 1).to(dev)
        feats = [torch.randn(B, S, C, generator=g).to(dev) for _ in range(3)]
        base = torch.randn(B, N, C, generator=g).to(dev) if with_base else None
        sel = U.select_neighbors(x1, x2, k, eps)
        for f in feats:
            want, idx, w, d = U.ops.interp_forward(x1, x2, f, k, eps, base=base, alpha=0.3 if with_base else 1.0)
            n0 = U.launch_count()
            got = U.ops.interp_blend(f, sel.idx, sel.weight, base=base, alpha=0.3 if with_base else 1.0)
            assert U.launch_count() - n0 == 1
            assert torch.equal(sel.idx, idx) and torch.equal(sel.weight, w) and torch.equal(sel.dist, d)
            assert torch.equal(got, want)
    # through the mirrors, with autograd
    x1, x2 = (torch.rand(2, 200, 3, generator=g) * 2 - 1).to(dev), (torch.rand(2, 48, 3, generator=g) * 2 - 1).to(dev)
    p1 = torch.randn(2, 200, 64, generator=g).to(dev)
    sel = U.select_neighbors(x1, x2, 8, 1e-3)
    for _ in range(2):
        p2a = torch.randn(2, 48, 64, generator=g).to(dev).requires_grad_(True)
        p2b = p2a.detach().clone().requires_grad_(True)
        a = U.propagate(x1, x2, p1, p2a, de_neighbors=8, dist_e=1e-3, selection=sel)
        b = U.propagate(x1, x2, p1, p2b, de_neighbors=8, dist_e=1e-3)
        assert torch.equal(a, b)
        go = torch.randn_like(a)
        a.backward(go)
        b.backward(go)
        assert torch.equal(p2a.grad, p2b.grad)
    with pytest.raises(ValueError):
        U.propagate(x1, x2, p1, p2a, de_neighbors=6, dist_e=1e-3, selection=sel)
    # the reference's default de_neighbors=64 on more than 32 sources: served (torch formulation), not refused
    from oracle import torch_formulation as T
    wide = U.propagate(x1, x2, p1, p2a.detach())
    want = p1.cpu() + 0.3 * T.interpolate(x1.cpu(), x2.cpu(), p2a.detach().cpu(), 48, 1e-8)
    torch.testing.assert_close(wide.cpu(), want, rtol=1e-3, atol=1e-3)


def test_golden_reference_propagate_and_feature_propagation(U, dev):
    """The reference's own propagate / PointNetFeaturePropagation outputs and float64 autograd gradients
    (tests/golden/golden_interp.npz) against the drop-in functions, autograd included."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_interp.npz"))
    T = lambda k: torch.from_numpy(g[k]).to(dev)  # noqa: E731
    t = [T(k).requires_grad_(True) for k in ("a_xyz1", "a_xyz2", "a_p1", "a_p2")]
    out = U.propagate(*t, de_neighbors=8, dist_e=1e-3)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["a_out"], rtol=1e-5, atol=2e-6)
    (out * T("a_w")).sum().backward()
    for v, name, tol in zip(t, ("a_gx1", "a_gx2", "a_gp1", "a_gp2"), (1e-3, 1e-3, 1e-6, 1e-4)):
        np.testing.assert_allclose(v.grad.cpu().numpy(), g[name], rtol=tol, atol=tol * max(1.0, float(np.abs(g[name]).max())) * 0.1)
    np.testing.assert_allclose(U.propagate(T("a_xyz1"), T("a_xyz2"), T("a_p1"), T("a_p2"), de_neighbors=6).cpu().numpy(),
                               g["b_out"], rtol=1e-5, atol=2e-6)
    t = [T(k).requires_grad_(True) for k in ("c_xyz1", "c_xyz2", "c_p2")]
    out = U.interpolate_features(t[0], t[1], t[2], 3, eps=1e-4)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["c_out"], rtol=1e-5, atol=2e-6)
    (out * T("c_w")).sum().backward()
    for v, name, tol in zip(t, ("c_gx1", "c_gx2", "c_gp2"), (1e-3, 1e-3, 1e-4)):
        np.testing.assert_allclose(v.grad.cpu().numpy(), g[name], rtol=tol, atol=tol * max(1.0, float(np.abs(g[name]).max())) * 0.1)
    one = U.interpolate_features(T("c_xyz1"), T("c_xyz2")[:, :1].contiguous(), T("c_p2")[:, :1].contiguous(), 3)
    np.testing.assert_allclose(one.cpu().numpy(), g["d_out"], rtol=0, atol=0)  # S == 1: repeated
    with pytest.raises(ValueError):
        U.ops.interp_forward(T("a_xyz1"), T("a_xyz2"), T("a_p2"), 33, 1e-3)
    with pytest.raises(RuntimeError):
        U.ops.interp_forward(T("a_xyz1").cpu(), T("a_xyz2"), T("a_p2"), 3, 1e-3)


# ------------------------------------------------------------------ knn_points -----

@pytest.mark.parametrize("B,N1,N2,K", [(4, 20, 1024, 4), (2, 52, 1076, 4), (3, 100, 64, 8), (2, 7, 5, 5), (2, 300, 2500, 32), (1, 1, 1, 1)])
def test_knn_points_matches_oracle(U, O, dev, B, N1, N2, K):
    """pytorch3d convention (models/Point_MAE_pretask_dev.py:680): squared distances ascending, int64 idx, nn gather."""
    g = torch.Generator().manual_seed(N1 + N2)
    p1, p2 = torch.rand(B, N1, 3, generator=g) * 2 - 1, torch.rand(B, N2, 3, generator=g) * 2 - 1
    out = U.knn_points(p1.to(dev), p2.to(dev), K=K, return_nn=True)
    D, I = O.knn_points(p1.numpy(), p2.numpy(), K)
    assert out.idx.dtype == torch.int64 and np.array_equal(out.idx.cpu().numpy(), I)
    assert np.array_equal(out.dists.cpu().numpy(), D)                       # same fma spelling: bit-exact
    want_nn = np.take_along_axis(p2.numpy()[:, None], I[..., None].repeat(3, -1), 2)
    assert np.array_equal(out.knn.cpu().numpy(), want_nn)
    assert U.knn_points(p1.to(dev), p2.to(dev), K=K).knn is None
    # float64 brute force: same neighbours wherever the gap to the next candidate exceeds fp32 rounding
    d64 = ((p1.double()[:, :, None] - p2.double()[:, None]) ** 2).sum(-1)
    np.testing.assert_allclose(out.dists.cpu().numpy(), np.sort(d64.numpy(), -1)[..., :K], rtol=1e-5, atol=1e-7)


def test_knn_points_autograd_and_errors(U, dev):
    g = torch.Generator().manual_seed(5)
    p1 = (torch.rand(2, 30, 3, generator=g)).to(dev).requires_grad_(True)
    p2 = (torch.rand(2, 90, 3, generator=g)).to(dev).requires_grad_(True)
    out = U.knn_points(p1, p2, K=4, return_nn=True)
    wd, wn = torch.rand_like(out.dists), torch.rand_like(out.knn)
    ((out.dists * wd).sum() + (out.knn * wn).sum()).backward()
    q1, q2 = p1.detach().double().requires_grad_(True), p2.detach().double().requires_grad_(True)
    nn64 = torch.gather(q2, 1, out.idx.reshape(2, -1, 1).expand(-1, -1, 3)).view(2, 30, 4, 3)
    d64 = ((q1.unsqueeze(2) - nn64) ** 2).sum(-1)
    ((d64 * wd.double()).sum() + (nn64 * wn.double()).sum()).backward()
    torch.testing.assert_close(p1.grad.double(), q1.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(p2.grad.double(), q2.grad, rtol=1e-4, atol=1e-6)
    with pytest.raises(ValueError):
        U.knn_points(p1.detach(), p2.detach(), K=91)
    with pytest.raises(NotImplementedError):
        U.knn_points(p1.detach(), p2.detach(), K=4, lengths1=torch.tensor([30, 30]))


# ------------------------------------------------------------------ seprate_point_cloud ----

SEPRATE_CASES = {
    "fixed_crop": dict(crop=128, sample_points=256),
    "range_crop": dict(crop=[100, 200], sample_points=64),
    "fixed_view": dict(crop=128, fixed_points=torch.Tensor([1, 1, 1]), sample_points=256),
    "view_list": dict(crop=150, fixed_points=[torch.Tensor([1, 1, 1]), torch.Tensor([-1, 1, 0]), torch.Tensor([0, -1, 1])]),
    "padding": dict(crop=128, padding_zeros=True, sample_points=1024),
    "no_fps": dict(crop=128, incomplete_shape=False),
}


@pytest.mark.parametrize("case", sorted(SEPRATE_CASES))
def test_golden_reference_seprate_point_cloud(U, dev, case):
    """The batched mirror against the reference's own per-cloud loop (tests/golden/golden_seprate.npz), same seeds:
    identical crops and identical FPS selections (2 batched FPS launches instead of 2*B batch-1 launches)."""
    import os
    import random
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_seprate.npz"))
    xyz = torch.from_numpy(g["xyz"]).to(dev)
    random.seed(11)
    torch.manual_seed(11)
    c0 = U.launch_count()
    a, b = U.misc.seprate_point_cloud(xyz, 512, **SEPRATE_CASES[case])
    assert U.launch_count() - c0 <= 3  # one crop launch + at most two batched FPS launches (reference: 2*B FPS launches + B sorts)
    assert a.is_contiguous() and b.is_contiguous()
    assert np.array_equal(a.cpu().numpy(), g[case + "_input"])
    assert np.array_equal(b.cpu().numpy(), g[case + "_crop"])
    same, none = U.misc.seprate_point_cloud(xyz, 512, 512)
    assert same is xyz and none is None


@pytest.mark.parametrize("B,n,num_crop,pad", [(32, 8192, 2048, False), (32, 8192, 4096, False), (5, 2048, 512, True), (3, 777, 100, False),
                                              (2, 33, 33, False), (2, 100, 0, False), (4, 5000, 1228, True)])
def test_crop_split_matches_oracle(U, O, dev, B, n, num_crop, pad):
    """upp_crop_split_f32 (distance to the viewpoint, in-shared-memory stable sort, split, gathers: one launch for the batch)
    against the oracle restatement of utils/misc.py:232-239, at the runners' real shape (B32, 8192 points) and ragged ones;
    duplicated points exercise the stable (distance, index) order."""
    g = torch.Generator().manual_seed(n + num_crop)
    xyz = unit_sphere(torch.randn(B, n, 3, generator=g) * 0.35).contiguous()
    if n >= 100:
        xyz[:, 50:60] = xyz[:, 10:20]  # exact ties
    centers = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).contiguous()
    n0 = U.launch_count()
    inp, crp, order = U.ops.crop_split(xyz.to(dev), centers.to(dev), num_crop, padding_zeros=pad, want_order=True)
    assert U.launch_count() - n0 == 1
    assert np.array_equal(order.cpu().numpy(), O.crop_order(xyz.numpy(), centers.numpy()))
    o_in, o_crop = O.crop_split(xyz.numpy(), centers.numpy(), num_crop, padding_zeros=pad)
    assert np.array_equal(inp.cpu().numpy(), o_in) and np.array_equal(crp.cpu().numpy(), o_crop)
    # the torch formulation of the reference (norm + argsort) selects the same points wherever distances are distinct
    d = torch.norm(centers[:, None] - xyz, p=2, dim=-1)
    same = (torch.argsort(d, dim=-1, stable=True).int() == order.cpu()).float().mean().item()
    assert same > 0.999


def test_random_dropping_mirror(U, dev):
    """utils/misc.py:308-315: FPS to a random size (same CPU-generator draw), zero rows up to 2048 points."""
    xyz = cube(3, 2048, 4).to(dev)
    for e in (0, 120, 400):
        torch.manual_seed(5 + e)
        out = U.misc.random_dropping(xyz, e)
        torch.manual_seed(5 + e)
        n = int(torch.randint(1, max(64, 768 // (e // 50 + 1)), (1, 1))[0, 0])
        assert tuple(out.shape) == (3, 2048, 3)
        assert torch.equal(out[:, :n], U.fps(xyz, n)[0]) and not out[:, n:].any()


def test_chamfer_sharded_entry_world1_equals_plain(U, dev):
    """upp_chamfer_fwd_sharded_f32 with a one-rank exchange (no peers to wait for) == upp_chamfer_fwd_f32 + sums;
    the N-rank exchange itself is exercised by scripts/check_multigpu.py under torchrun."""
    from upp_b200 import _lib

    class _One:
        struct = _lib.PeerExchangeStruct()
    _One.struct.rank, _One.struct.world, _One.struct.seq = 0, 1, None
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(3, 500, 3, generator=g).to(dev), torch.rand(3, 700, 3, generator=g).to(dev)
    got = U.ops.chamfer_forward_sharded(a, b, _One)
    want = U.ops.chamfer_forward(a, b, want_sums=True)
    for x, y in zip(got, want):
        assert torch.equal(x, y)


def test_metrics_f_score_and_cd_x1000(U, dev):
    """utils/metrics.py:70-111 semantics against float64 brute force (what open3d computes), per cloud then averaged."""
    g = torch.Generator().manual_seed(17)
    gt = torch.rand(3, 400, 3, generator=g) * 0.2
    pred = gt[:, torch.randperm(400, generator=g)[:300]] + 0.004 * torch.randn(3, 300, 3, generator=g)
    d = torch.cdist(pred.double(), gt.double())
    fs = []
    for b in range(3):
        p, r = (d[b].min(1)[0] < 0.01).double().mean(), (d[b].min(0)[0] < 0.01).double().mean()
        fs.append(2 * r * p / (r + p) if r + p else 0.0)
    got = U.metrics.f_score(pred.to(dev), gt.to(dev))
    assert got.dim() == 0 and abs(got.item() - float(sum(fs) / 3)) < 2e-3  # a point within fp32 rounding of th may flip
    far = U.metrics.f_score(pred.to(dev), (gt + 5.0).to(dev))
    assert far.item() == 0.0
    l1 = ((d.min(2)[0].mean() + d.min(1)[0].mean()) / 2 * 1000).item()
    l2 = (((d ** 2).min(2)[0].mean() + (d ** 2).min(1)[0].mean()) * 1000).item()
    assert abs(U.metrics.chamfer_distance_l1(pred.to(dev), gt.to(dev)).item() - l1) <= 1e-5 * l1
    assert abs(U.metrics.chamfer_distance_l2(pred.to(dev), gt.to(dev)).item() - l2) <= 1e-5 * l2


# ------------------------------------------------------------------ single-launch Group ----

GROUP_KERNELS = {  # name -> tuning environment (read per launch by group_fused_launch, honoured under UPP_TUNING=1)
    "fused_default": {},                                          # cluster size / warps from the heuristic
    "fused_cluster1": {"UPP_GROUP_CLUSTER": "1"},                 # producer and consumers in one CTA
    "fused_cluster1_w8": {"UPP_GROUP_CLUSTER": "1", "UPP_GROUP_WARPS": "8"},
    "fused_cluster2": {"UPP_GROUP_CLUSTER": "2"},                 # consumers on their own SMs, centres published over DSMEM
    "fused_cluster3": {"UPP_GROUP_CLUSTER": "3"},
    "fused_cluster4": {"UPP_GROUP_CLUSTER": "4"},
    "fused_cluster8": {"UPP_GROUP_CLUSTER": "8", "UPP_GROUP_WARPS": "8"},
    "two_launch": {"UPP_GROUP_FUSED": "0"},                       # fps_launch + knn_launch (round-1 path)
}


@pytest.fixture(params=sorted(GROUP_KERNELS))
def group_kernel(request, monkeypatch):
    for k, v in GROUP_KERNELS[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


@pytest.mark.parametrize("B,N,G,k", [(4, 1024, 64, 32), (3, 1096, 32, 16), (3, 32, 32, 16), (3, 64, 32, 8), (2, 2048, 128, 32),
                                     (5, 972, 32, 16), (2, 1536, 128, 32), (3, 200, 7, 5), (2, 130, 64, 32), (2, 40, 50, 3),
                                     (1, 513, 1, 1), (2, 256, 256, 8)])
def test_group_single_launch_variants(U, O, dev, group_kernel, B, N, G, k):
    """The single-launch Group (kNN pipelined behind the FPS chain, one cluster per cloud) in every cluster shape, and
    the two-launch path, against the oracle: bit-equal indices, centres and neighbourhoods; launch count as advertised."""
    xyz = unit_sphere(cube(B, N, 900 + N + G))
    o_nb, o_ce, o_idx, o_cidx = O.group(xyz.numpy(), G, k)
    n0 = U.launch_count()
    nb, ce, idx, cidx = U.ops.group(xyz.to(dev), G, k)
    assert U.launch_count() - n0 == (2 if group_kernel == "two_launch" else 1)
    assert np.array_equal(cidx.cpu().numpy(), o_cidx) and np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(ce.cpu().numpy(), o_ce) and np.array_equal(nb.cpu().numpy(), o_nb)


def test_group_single_launch_lattice_ties_and_many_clouds(U, O, dev, group_kernel):
    ax = torch.arange(6, dtype=torch.float32)
    grid = (torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).reshape(1, -1, 3) + 1.0)
    xyz = torch.cat([grid, grid.flip(1)], 0).contiguous()
    for got, want in zip(U.ops.group(xyz.to(dev), 40, 16), O.group(xyz.numpy(), 40, 16)):
        assert np.array_equal(got.cpu().numpy(), want)
    many = cube(300, 96, 5)  # more clouds than SMs: several clusters per SM
    for got, want in zip(U.ops.group(many.to(dev), 16, 8), O.group(many.numpy(), 16, 8)):
        assert np.array_equal(got.cpu().numpy(), want)


# ------------------------------------------------------------------ deterministic backward -----------

def test_backward_kernels_are_deterministic_and_collision_safe(U, O, dev):
    """Every scatter-add on the path is an ordered gather (scatter.cuh): bit-identical run to run, also when every
    entry lands on the same destination; values against the oracle / a float64 accumulation."""
    g = torch.Generator().manual_seed(41)
    a, b = torch.rand(3, 700, 3, generator=g).to(dev), (torch.rand(3, 900, 3, generator=g) * 0.05 + 0.5).to(dev)  # b is a small blob:
    g1, g2 = torch.randn(3, 700, generator=g).to(dev), torch.randn(3, 900, generator=g).to(dev)                    # few points of a attract all of b
    _, _, i1, i2 = U.chamfer.forward(a, b)
    first = U.chamfer.backward(a, b, i1, i2, g1, g2)
    for _ in range(4):
        again = U.chamfer.backward(a, b, i1, i2, g1, g2)
        assert torch.equal(first[0], again[0]) and torch.equal(first[1], again[1])
    o1, o2 = O.chamfer_bwd(a.cpu().numpy(), b.cpu().numpy(), i1.cpu().numpy(), i2.cpu().numpy(), g1.cpu().numpy(), g2.cpu().numpy())
    np.testing.assert_allclose(first[0].cpu().numpy(), o1, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(first[1].cpu().numpy(), o2, rtol=1e-4, atol=1e-5)
    # all entries on ONE destination, list longer than one staged tile, duplicates elsewhere
    rows = torch.randn(2, 9000, 3, generator=g).to(dev)
    idx = torch.zeros(2, 9000, dtype=torch.int32)
    idx[1] = torch.randint(0, 50, (9000,), generator=g, dtype=torch.int32)
    idx = idx.to(dev)
    got = U.ops.rows_scatter_add(rows, idx, 300)
    want = torch.zeros(2, 300, 3, dtype=torch.float64).scatter_add_(1, idx.cpu().long().unsqueeze(-1).expand(-1, -1, 3), rows.cpu().double())
    np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-3)
    assert all(torch.equal(got, U.ops.rows_scatter_add(rows, idx, 300)) for _ in range(3))
    gg = U.ops.gather_grad(rows.transpose(1, 2).contiguous(), idx, 300)
    assert torch.equal(gg, got.transpose(1, 2))  # same additions in the same order, channel-first layout
    # Group backward: every neighbour slot of every group on the same few points
    B, N, G, k = 2, 64, 32, 16
    gnb, gce = torch.randn(B, G, k, 3, generator=g).to(dev), torch.randn(B, G, 3, generator=g).to(dev)
    ii = torch.randint(0, 3, (B, G, k), generator=g).to(dev)
    ci = torch.randint(0, 3, (B, G), generator=g, dtype=torch.int32).to(dev)
    gx = U.ops.group_backward(gnb, gce, ii, ci, N)
    want = torch.zeros(B, N, 3, dtype=torch.float64)
    want.scatter_add_(1, ii.cpu().reshape(B, -1, 1).expand(-1, -1, 3), gnb.cpu().double().reshape(B, -1, 3))
    want.scatter_add_(1, ci.cpu().long().unsqueeze(-1).expand(-1, -1, 3), (gce.cpu().double() - gnb.cpu().double().sum(2)))
    np.testing.assert_allclose(gx.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    assert all(torch.equal(gx, U.ops.group_backward(gnb, gce, ii, ci, N)) for _ in range(3))
    assert not gx[:, 3:].any()  # untouched destinations are written as zeros (no memset in front of the kernel)


# ------------------------------------------------------------------ BASELINE.json configs at their real batch sizes ----

def _cmp_group(U, O, dev, xyz, G, k):
    nb, ce, idx, cidx = U.ops.group(xyz.to(dev), G, k)
    o_nb, o_ce, o_idx, o_cidx = O.group(xyz.numpy(), G, k)
    assert np.array_equal(cidx.cpu().numpy(), o_cidx) and np.array_equal(idx.cpu().numpy(), o_idx)
    assert np.array_equal(ce.cpu().numpy(), o_ce) and np.array_equal(nb.cpu().numpy(), o_nb)
    return ce


def test_baseline_config_c1_real_batch(U, O, dev):
    """BASELINE.json configs[0]: Group(64, 32) on B=32 x 1024 points (uniform cube and unit-sphere normalised), default dispatch."""
    for xyz in (cube(32, 1024, 0), unit_sphere(torch.randn(32, 1024, 3, generator=torch.Generator().manual_seed(1)) * 0.35)):
        _cmp_group(U, O, dev, xyz.contiguous(), 64, 32)


def test_baseline_config_c3_real_batch(U, O, dev):
    """BASELINE.json configs[2]: Chamfer L1 / L2 fwd + bwd at B=64, 2048 vs 2048 (default dispatch: wave-aware chunks)."""
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(64, 2048, 3, generator=g), torch.rand(64, 2048, 3, generator=g)
    d1, d2, i1, i2, sums = U.ops.chamfer_forward(a.to(dev), b.to(dev), want_sums=True)
    o1, o2, j1, j2 = O.chamfer_fwd(a.numpy(), b.numpy())
    assert np.array_equal(i1.cpu().numpy(), j1) and np.array_equal(i2.cpu().numpy(), j2)
    assert np.array_equal(d1.cpu().numpy(), o1) and np.array_equal(d2.cpu().numpy(), o2)
    want = np.array([o1.astype(np.float64).sum(), o2.astype(np.float64).sum(), np.sqrt(o1.astype(np.float64)).sum(), np.sqrt(o2.astype(np.float64)).sum()])
    np.testing.assert_allclose(sums.cpu().numpy(), want, rtol=1e-5)
    for mod, ref in ((U.ChamferDistanceL1(), (np.sqrt(o1).mean() + np.sqrt(o2).mean()) / 2), (U.ChamferDistanceL2(), o1.mean() + o2.mean())):
        x, y = a.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)
        loss = mod(x, y)
        loss.backward()
        assert abs(loss.item() - ref) <= RTOL * abs(ref)
    n = float(o1.size)
    gx1, gx2 = O.chamfer_bwd(a.numpy(), b.numpy(), j1, j2, np.full_like(o1, 1.0 / n), np.full_like(o2, 1.0 / n))  # L2: d mean / d dist
    np.testing.assert_allclose(x.grad.cpu().numpy(), gx1, rtol=1e-4, atol=1e-10)
    np.testing.assert_allclose(y.grad.cpu().numpy(), gx2, rtol=1e-4, atol=1e-10)


def test_baseline_config_c4_real_batch(U, O, dev):
    """BASELINE.json configs[3]: B=128 clouds of 8192 points -> misc.fps(1024) -> Group(64, 32); default dispatch (one CTA per
    cloud with the deferred tree search at this batch), and the batch sharded 8 ways (16 clouds: the cluster kernel)."""
    xyz = unit_sphere(torch.randn(128, 8192, 3, generator=torch.Generator().manual_seed(3)) * 0.35).contiguous()
    idx, centers = U.ops.fps(xyz.to(dev), 1024, True)
    want = O.fps(xyz.numpy(), 1024)
    assert np.array_equal(idx.cpu().numpy(), want)
    assert torch.equal(centers.cpu(), torch.gather(xyz, 1, torch.from_numpy(want).long()[..., None].expand(-1, -1, 3)))
    _cmp_group(U, O, dev, centers.cpu().contiguous(), 64, 32)
    shard = U.ops.fps(xyz[:16].contiguous().to(dev), 1024)
    assert np.array_equal(shard.cpu().numpy(), want[:16])


def test_baseline_config_c5_real_batch(U, O, dev):
    """BASELINE.json configs[4], geometry part: Group(128, 32) on B=32 x 2048 points + 3-NN propagation of 1152-d features
    2048 <- 128, forward and feature gradient (oracle on a sample of clouds; the full batch is launched)."""
    g = torch.Generator().manual_seed(5)
    xyz = unit_sphere(torch.randn(32, 2048, 3, generator=g) * 0.35).contiguous()
    ce = _cmp_group(U, O, dev, xyz, 128, 32)
    feat, go = torch.randn(32, 128, 1152, generator=g), torch.randn(32, 2048, 1152, generator=g)
    out, idx, w, d = U.ops.interp_forward(xyz.to(dev), ce, feat.to(dev), 3, 1e-4)
    gp2 = U.ops.interp_backward(go.to(dev), idx, w, 128)[0]
    for b in (0, 13, 31):
        s = slice(b, b + 1)
        o_out, o_idx, o_w, o_d = O.interp_fwd(xyz[s].numpy(), ce[s].cpu().numpy(), feat[s].numpy(), 3, 1e-4)
        assert np.array_equal(idx[s].cpu().numpy(), o_idx) and np.array_equal(w[s].cpu().numpy(), o_w)
        np.testing.assert_allclose(out[s].cpu().numpy(), o_out, rtol=RTOL, atol=1e-6)
        o_g = O.interp_bwd(go[s].numpy(), feat[s].numpy(), xyz[s].numpy(), ce[s].cpu().numpy(), o_idx, o_w, o_d, 1e-4)[0]
        np.testing.assert_allclose(gp2[s].cpu().numpy(), o_g, rtol=1e-4, atol=1e-5 * max(1.0, float(np.abs(o_g).max())))


def test_baseline_config_c2_census_real_batch(U, O, dev):
    """BASELINE.json configs[1] geometry census at B=32: every FPS / Group shape of one UPP
    ModelNet40 classification step, default dispatch, against the oracle."""
    g = torch.Generator().manual_seed(2)
    pts = torch.cat([unit_sphere(torch.randn(32, 1024, 3, generator=g) * 0.35), torch.randn(32, 72, 3, generator=g) * 0.6], 1).contiguous()
    ce1 = _cmp_group(U, O, dev, pts, 32, 16)                       # 1096 -> Group(32,16)
    _cmp_group(U, O, dev, ce1.cpu().contiguous(), 32, 16)          # 32 -> Group(32,16)
    keep = pts[:, :972].contiguous()
    _cmp_group(U, O, dev, keep, 32, 16)                            # 972 -> Group(32,16)
    reb = pts[:, :1024].contiguous()
    i1, c1 = U.ops.fps(reb.to(dev), 256, True)                     # 1024 -> 256
    assert np.array_equal(i1.cpu().numpy(), O.fps(reb.numpy(), 256))
    cat = torch.cat([keep, c1.cpu()], 1).contiguous()
    i2, c2 = U.ops.fps(cat.to(dev), 1024, True)                    # 1228 -> 1024
    assert np.array_equal(i2.cpu().numpy(), O.fps(cat.numpy(), 1024))
    ce4 = _cmp_group(U, O, dev, c2.cpu().contiguous(), 64, 32)     # 1024 -> Group(64,32)
    _cmp_group(U, O, dev, ce4.cpu().contiguous(), 32, 8)           # 64 -> Group(32,8)


# ------------------------------------------------------------------ the reference's call sites over the drop-ins ----

def test_reference_callsites_over_dropins(U, O, dev):
    """reference_callsites.py (the reference's misc.fps / Group.forward / ChamferDistanceL1,L2 call sequences, pinned to
    the real source by tests/test_oracle.py::test_callsite_restatement_equals_live_reference) importing pointnet2_ops,
    knn_cuda and chamfer exactly as the reference does -- resolved by the drop-in packages -- against the oracle, both
    index conventions, autograd included."""
    import reference_callsites as R
    xyz = unit_sphere(cube(4, 1024, 77)).contiguous()
    o_nb, o_ce, o_idx, o_cidx = O.group(xyz.numpy(), 64, 32)
    x = xyz.to(dev).requires_grad_(True)
    grp = R.Group(64, 32)
    nb, ce, idx, cidx = grp(x, require_index=True, gather_idx=True)
    assert np.array_equal(idx.cpu().numpy(), o_idx) and np.array_equal(cidx.cpu().numpy(), o_cidx.astype(np.int64))
    assert np.array_equal(nb.detach().cpu().numpy(), o_nb) and np.array_equal(ce.detach().cpu().numpy(), o_ce)
    nb2, ce2, fidx, fcidx = grp(x, require_index=True, gather_idx=False)
    base = np.arange(4) * 1024
    assert np.array_equal(fidx.cpu().numpy(), (o_idx + base[:, None, None]).reshape(-1))
    assert np.array_equal(fcidx.cpu().numpy(), (o_cidx.astype(np.int64) + base[:, None]).reshape(-1))
    assert torch.equal(nb2, nb) and torch.equal(ce2, ce)
    # same numbers as the fused module, gradients included
    gw, cw = torch.randn(4, 64, 32, 3, generator=torch.Generator().manual_seed(1)).to(dev), torch.randn(4, 64, 3, generator=torch.Generator().manual_seed(2)).to(dev)
    ((nb2 * gw).sum() + (ce2 * cw).sum()).backward()
    y = xyz.to(dev).requires_grad_(True)
    fnb, fce = U.Group(64, 32)(y)
    ((fnb * gw).sum() + (fce * cw).sum()).backward()
    assert torch.equal(fnb, nb2) and torch.equal(fce, ce2)
    torch.testing.assert_close(x.grad, y.grad, rtol=1e-5, atol=1e-6)
    # Chamfer modules over `import chamfer`
    g = _gold("golden_chamfer_modules.npz")
    for name, mod in (("l1", R.ChamferDistanceL1()), ("l2", R.ChamferDistanceL2())):
        a = torch.from_numpy(g["xyz1"]).to(dev).requires_grad_(True)
        b = torch.from_numpy(g["xyz2"]).to(dev).requires_grad_(True)
        loss = mod(a, b)
        loss.backward()
        assert abs(loss.item() - g[f"{name}_loss"]) <= RTOL * abs(g[f"{name}_loss"])
        np.testing.assert_allclose(a.grad.cpu().numpy(), g[f"{name}_g1"], rtol=2e-4, atol=1e-8)
    z1, z2 = torch.from_numpy(g["z1"]).to(dev), torch.from_numpy(g["z2"]).to(dev)
    assert abs(R.ChamferDistanceL2(ignore_zeros=True)(z1, z2).item() - g["l2_ignore_zeros"]) <= RTOL * g["l2_ignore_zeros"]


# ------------------------------------------------------------------ N-rank exchange under pytest ----------------

def test_multigpu_sharded_path_under_torchrun():
    """scripts/check_multigpu.py under torchrun on min(2, device_count) GPUs: sharded Chamfer loss and gradients ==
    unsharded, fused NVLink peer all-reduce == NCCL and bit-identical across ranks, CUDA-graph replay, deferred finish,
    empty shard, DDP-compatible gradient scaling, gradient-statistics exchange, Group shard-invariant.  Skipped on one GPU."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (the driver's multi-GPU tier runs it)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", str(port), os.path.join(root, "scripts", "check_multigpu.py")],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multigpu ok: 2/2 ranks" in out.stdout
