"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN PYTHON.

Runs only where /root/reference exists (the build container): the reference's functions are
lifted by AST (oracle/ref_lift.py) -- never copied -- executed on seeded inputs, and inputs +
outputs are stored as small .npz files that travel to the GPU box.

  golden_knn_point.npz       models/dgcnn_group.py:8-19 knn_point (+ square_distance :21-40)
  golden_fps_numpy.npz       datasets/ModelNetDataset.py:29-50 farthest_point_sample (start forced to 0)
  golden_chamfer_modules.npz extensions/chamfer_dist/__init__.py:13-84 ChamferFunction +
                             ChamferDistanceL1 / L2 / L2_split (ignore_zeros at B=1 included), run
                             unmodified over a float64 brute-force `chamfer` stand-in, with autograd
  golden_group.npz           models/Point_MAE_unify.py:51-92 Group.forward run unmodified over
                             utils/misc.py:13-20 fps and float64/numpy stand-ins for the two
                             third-party ops

  golden_interp.npz          models/Point_MAE_unify.py:22-48 propagate and the interpolation of
                             models/Point_MAE_unify_segment.py:277-325 PointNetFeaturePropagation (empty MLP),
                             run unmodified in fp32 (outputs) and float64 (autograd gradients)

    python tests/golden/make_golden.py            # everything
    python tests/golden/make_golden.py interp     # only golden_interp.npz
    python tests/golden/make_golden.py crop       # only golden_seprate.npz

  golden_seprate.npz         utils/misc.py:205-256 seprate_point_cloud run unmodified (Tensor.cuda patched to the
                             identity, fps = utils/misc.py:13-20 over a float64 FPS stand-in), seeded
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import f64, ref_lift  # noqa: E402


def lattice_cloud(B, N, seed):
    """Points on a jittered lattice: pairwise squared distances are separated by far more than fp32
    rounding of either distance formula, so index results are formula-independent."""
    g = np.random.default_rng(seed)
    side = int(np.ceil(N ** (1 / 3))) + 1
    out = np.zeros((B, N, 3), np.float32)
    for b in range(B):
        cells = g.choice(side ** 3, size=N, replace=False)
        ijk = np.stack(np.unravel_index(cells, (side,) * 3), -1).astype(np.float64)
        # jitter on a 2^-7 grid: all coordinates (and their squares / products) exact in fp32
        jit = g.integers(-24, 25, size=(N, 3)) / 128.0
        out[b] = (ijk + jit + 1.0).astype(np.float32)
    return out


class _F64Chamfer:
    """`chamfer` stand-in: float64 brute force on CPU tensors, reference return convention."""

    @staticmethod
    def forward(xyz1, xyz2):
        d1, d2, i1, i2 = f64.chamfer_fwd(xyz1.detach().numpy(), xyz2.detach().numpy())
        return [torch.from_numpy(d1).to(xyz1.dtype), torch.from_numpy(d2).to(xyz1.dtype),
                torch.from_numpy(i1.astype(np.int32)), torch.from_numpy(i2.astype(np.int32))]

    @staticmethod
    def backward(xyz1, xyz2, idx1, idx2, g1, g2):
        a, b = xyz1.detach().double(), xyz2.detach().double()
        ga, gb = torch.zeros_like(a), torch.zeros_like(b)
        for (p, q, idx, g, gp, gq) in ((a, b, idx1, g1, ga, gb), (b, a, idx2, g2, gb, ga)):
            v = 2 * g.double().unsqueeze(-1) * (p - torch.gather(q, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)))
            gp += v
            gq.scatter_add_(1, idx.long().unsqueeze(-1).expand(-1, -1, 3), -v)
        return ga.to(xyz1.dtype), gb.to(xyz2.dtype)


def make_interp():
    """propagate + PointNetFeaturePropagation interpolation, from the reference's own code."""
    R = ref_lift.interpolation()
    g = torch.Generator().manual_seed(9)
    out = {}
    # case A: propagate(level-1 centres <- level-2 centres), de_neighbors=8, dist_e=1e-3 (models/Point_MAE_pretask_dev.py:298)
    x1 = torch.from_numpy(lattice_cloud(3, 64, 21))
    x2 = torch.from_numpy(lattice_cloud(3, 32, 22))
    p1 = torch.randn(3, 64, 48, generator=g)
    p2 = torch.randn(3, 32, 48, generator=g)
    wgt = torch.randn(3, 64, 48, generator=g).double()  # fp32-representable loss weights
    out.update(a_xyz1=x1.numpy(), a_xyz2=x2.numpy(), a_p1=p1.numpy(), a_p2=p2.numpy(), a_w=wgt.numpy())
    out["a_out"] = R.propagate(x1, x2, p1, p2, de_neighbors=8, dist_e=1e-3).numpy()
    t = [v.double().clone().requires_grad_(True) for v in (x1, x2, p1, p2)]
    (R.propagate(*t, de_neighbors=8, dist_e=1e-3) * wgt).sum().backward()
    for name, v in zip(("a_gx1", "a_gx2", "a_gp1", "a_gp2"), t):
        out[name] = v.grad.numpy()
    # case B: propagate with its defaults clipped by S (de_neighbors=6, dist_e=1e-8; models/Point_MAE_unify.py:598)
    out["b_out"] = R.propagate(x1, x2, p1, p2, de_neighbors=6).numpy()
    # case C: feature propagation, 3 neighbours, eps 1e-4, wide channels (models/Point_MAE_unify_segment.py:420,605-607)
    fp = R.PointNetFeaturePropagation(in_channel=0, mlp=[], interpolate_neighbors=3)
    y1 = torch.from_numpy(lattice_cloud(2, 200, 23))
    y2 = torch.from_numpy(lattice_cloud(2, 40, 24))
    q2 = torch.randn(2, 40, 36, generator=g)
    q1 = torch.randn(2, 200, 5, generator=g)
    out.update(c_xyz1=y1.numpy(), c_xyz2=y2.numpy(), c_p2=q2.numpy(), c_p1=q1.numpy())
    out["c_out"] = fp(y1, y2, None, q2).detach().numpy()             # interpolated only
    out["c_out_cat"] = fp(y1, y2, q1, q2).detach().numpy()           # cat([points1, interpolated])
    wc = torch.randn(2, 200, 36, generator=g).double()
    t = [v.double().clone().requires_grad_(True) for v in (y1, y2, q2)]
    (fp(t[0], t[1], None, t[2]) * wc).sum().backward()
    out.update(c_w=wc.numpy(), c_gx1=t[0].grad.numpy(), c_gx2=t[1].grad.numpy(), c_gp2=t[2].grad.numpy())
    # case D: a single source point (S == 1) is repeated
    out["d_out"] = fp(y1, y2[:, :1], None, q2[:, :1]).detach().numpy()
    out = {k: (v.astype(np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "golden_interp.npz"), **out)


def make_seprate():
    """seprate_point_cloud from the reference's own code, seeded; the cases of its call sites."""
    import random

    class _P2:
        @staticmethod
        def furthest_point_sample(xyz, npoint):
            return torch.from_numpy(f64.fps(xyz.numpy(), npoint).astype(np.int32))

        @staticmethod
        def gather_operation(features, idx):
            return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1))

    spc = ref_lift.seprate_point_cloud(ref_lift.misc_fps(_P2))
    xyz = torch.from_numpy(lattice_cloud(5, 512, 31))
    out = {"xyz": xyz.numpy()}
    cases = {  # name -> kwargs (tools/runner_module.py:131, tools/runner_pretask.py:179,369, tools/runner_finetune.py:267)
        "fixed_crop": dict(crop=128, sample_points=256),
        "range_crop": dict(crop=[100, 200], sample_points=64),
        "fixed_view": dict(crop=128, fixed_points=torch.Tensor([1, 1, 1]), sample_points=256),
        "view_list": dict(crop=150, fixed_points=[torch.Tensor([1, 1, 1]), torch.Tensor([-1, 1, 0]), torch.Tensor([0, -1, 1])]),
        "padding": dict(crop=128, padding_zeros=True, sample_points=1024),
        "no_fps": dict(crop=128, incomplete_shape=False),
    }
    with ref_lift.cpu_cuda():
        for name, kw in cases.items():
            random.seed(11)
            torch.manual_seed(11)
            a, b = spc(xyz, 512, **kw)
            out[name + "_input"], out[name + "_crop"] = a.numpy(), b.numpy()
    np.savez_compressed(os.path.join(HERE, "golden_seprate.npz"), **out)


def main():
    assert ref_lift.available(), "needs /root/reference"
    if sys.argv[1:] == ["interp"]:
        make_interp()
        return
    if sys.argv[1:] == ["crop"]:
        make_seprate()
        return
    make_interp()
    make_seprate()
    H = ref_lift.torch_helpers()

    # ---- knn_point ----
    ref = lattice_cloud(3, 200, 1)
    qry = np.concatenate([ref[:, :10], lattice_cloud(3, 14, 2)], 1)
    idx = H.knn_point(16, torch.from_numpy(ref), torch.from_numpy(qry)).numpy()  # unsorted top-k
    np.savez_compressed(os.path.join(HERE, "golden_knn_point.npz"), ref=ref, query=qry, k=16,
                        idx_sorted_by_index=np.sort(idx, -1))

    # ---- numpy FPS (random start patched to 0; no near-origin points: lattice is offset by +1) ----
    class _NP:
        random = types.SimpleNamespace(randint=lambda lo, hi: 0)

        def __getattr__(self, name):
            return getattr(np, name)
    fps_fn = ref_lift.lift("datasets/ModelNetDataset.py", ["farthest_point_sample"], {"np": _NP()}).farthest_point_sample
    cloud = lattice_cloud(4, 300, 3)
    picked = np.stack([fps_fn(cloud[b], 40) for b in range(4)])  # (4,40,3) coordinates
    np.savez_compressed(os.path.join(HERE, "golden_fps_numpy.npz"), xyz=cloud, npoint=40, picked=picked)

    # ---- Chamfer modules ----
    M = ref_lift.chamfer_modules(_F64Chamfer)
    g = torch.Generator().manual_seed(4)
    out = {}
    a = torch.rand(3, 96, 3, generator=g, dtype=torch.float64)
    b = torch.rand(3, 130, 3, generator=g, dtype=torch.float64)
    out["xyz1"], out["xyz2"] = a.numpy().astype(np.float32), b.numpy().astype(np.float32)
    a32 = torch.from_numpy(out["xyz1"]).double()
    b32 = torch.from_numpy(out["xyz2"]).double()
    for name, mod in (("l1", M.ChamferDistanceL1()), ("l2", M.ChamferDistanceL2())):
        x, y = a32.clone().requires_grad_(True), b32.clone().requires_grad_(True)
        loss = mod(x, y)
        loss.backward()
        out[f"{name}_loss"], out[f"{name}_g1"], out[f"{name}_g2"] = loss.item(), x.grad.numpy(), y.grad.numpy()
    s1, s2 = M.ChamferDistanceL2_split()(a32, b32)
    out["l2_split"] = np.array([s1.item(), s2.item()])
    z1, z2 = a32[:1].clone(), b32[:1].clone()
    z1[0, 50:] = 0  # zero-padded rows (utils/misc.py:313-314)
    z2[0, 100:] = 0
    out["z1"], out["z2"] = z1.numpy().astype(np.float32), z2.numpy().astype(np.float32)
    out["l2_ignore_zeros"] = M.ChamferDistanceL2(ignore_zeros=True)(z1, z2).item()
    out["l1_ignore_zeros"] = M.ChamferDistanceL1(ignore_zeros=True)(z1, z2).item()
    out["l2_keep_zeros"] = M.ChamferDistanceL2(ignore_zeros=False)(z1, z2).item()
    np.savez_compressed(os.path.join(HERE, "golden_chamfer_modules.npz"), **out)

    # ---- Group.forward, unmodified, over stand-ins for the third-party ops ----
    class _P2:
        @staticmethod
        def furthest_point_sample(xyz, npoint):
            return torch.from_numpy(f64.fps(xyz.numpy(), npoint).astype(np.int32))

        @staticmethod
        def gather_operation(features, idx):
            return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1))

    class _KNN(torch.nn.Module):
        def __init__(self, k, transpose_mode=False):
            super().__init__()
            self.k = k

        def forward(self, ref, query):
            d, i = f64.knn(ref.numpy(), query.numpy(), self.k)
            return torch.from_numpy(d.astype(np.float32)), torch.from_numpy(i)

    misc = types.SimpleNamespace(fps=ref_lift.misc_fps(_P2))
    Group = ref_lift.group_class(misc, _KNN)
    xyz = torch.from_numpy(lattice_cloud(2, 256, 5))
    nb, ce, idx, cidx = Group(16, 8)(xyz, require_index=True, gather_idx=True)
    nb2, ce2, fidx, fcidx = Group(16, 8)(xyz, require_index=True, gather_idx=False)
    np.savez_compressed(os.path.join(HERE, "golden_group.npz"), xyz=xyz.numpy(), G=16, k=8, neighborhood=nb.numpy(),
                        center=ce.numpy(), idx=idx.numpy(), center_idx=cidx.numpy(), flat_idx=fidx.numpy(),
                        flat_center_idx=fcidx.numpy())
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
