"""Lift pure functions/classes out of the read-only reference tree by AST, without
importing its (broken) packages.  Works ONLY where /root/reference exists (the build
container); used by tests/golden/make_golden.py and by `-m "not gpu"` tests that skip
when the tree is absent.  TEST INFRASTRUCTURE ONLY.
"""
import ast
import os
import types

REF_ROOT = os.environ.get("UPP_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(REF_ROOT)


def lift(rel_path, names, env=None):
    """Exec only the named top-level defs/classes of a reference file in a fresh namespace."""
    path = os.path.join(REF_ROOT, rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    keep = [n for n in tree.body
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    if missing:
        raise KeyError(f"{rel_path}: {sorted(missing)} not found")
    mod = ast.Module(body=keep, type_ignores=[])
    ns = dict(env or {})
    exec(compile(mod, path, "exec"), ns)
    return types.SimpleNamespace(**{n: ns[n] for n in names})


def torch_helpers():
    """square_distance / index_points (models/modules.py:13-51), knn_point
    (models/dgcnn_group.py:8-19), numpy farthest_point_sample
    (datasets/ModelNetDataset.py:29-50)."""
    import numpy as np
    import torch
    env = {"torch": torch, "np": np}
    m = lift("models/modules.py", ["square_distance", "index_points"], env)
    env2 = dict(env, square_distance=m.square_distance)
    k = lift("models/dgcnn_group.py", ["knn_point"], env2)
    f = lift("datasets/ModelNetDataset.py", ["farthest_point_sample"], env)
    return types.SimpleNamespace(square_distance=m.square_distance, index_points=m.index_points,
                                 knn_point=k.knn_point, farthest_point_sample=f.farthest_point_sample)


def chamfer_modules(chamfer_impl):
    """The reference's own ChamferFunction / ChamferDistanceL1 / L2 / L2_split
    (extensions/chamfer_dist/__init__.py:13-84) bound to `chamfer_impl`, an object with
    forward(xyz1, xyz2) and backward(...) -- i.e. whatever stands in for `import chamfer`."""
    import torch
    return lift("extensions/chamfer_dist/__init__.py",
                ["ChamferFunction", "ChamferDistanceL2", "ChamferDistanceL2_split", "ChamferDistanceL1"],
                {"torch": torch, "chamfer": chamfer_impl})


def group_class(misc_impl, knn_cls):
    """The reference's own Group (models/Point_MAE_unify.py:51-92) bound to stand-ins for
    `utils.misc` (needs .fps) and `knn_cuda.KNN`."""
    import torch
    import torch.nn as nn
    return lift("models/Point_MAE_unify.py", ["Group"],
                {"torch": torch, "nn": nn, "misc": misc_impl, "KNN": knn_cls}).Group


def misc_fps(pointnet2_utils_impl):
    """The reference's own utils.misc.fps (utils/misc.py:13-20) bound to a stand-in for
    `pointnet2_ops.pointnet2_utils`."""
    return lift("utils/misc.py", ["fps"], {"pointnet2_utils": pointnet2_utils_impl}).fps


def interpolation():
    """The reference's own propagate (models/Point_MAE_unify.py:22-48) and PointNetFeaturePropagation
    (models/Point_MAE_unify_segment.py:277-325), bound to its square_distance / index_points
    (models/modules.py:13-51)."""
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    h = lift("models/modules.py", ["square_distance", "index_points"], {"torch": torch})
    env = {"torch": torch, "nn": nn, "F": F, "square_distance": h.square_distance, "index_points": h.index_points}
    p = lift("models/Point_MAE_unify.py", ["propagate"], env)
    fp = lift("models/Point_MAE_unify_segment.py", ["PointNetFeaturePropagation"], env)
    return types.SimpleNamespace(propagate=p.propagate, PointNetFeaturePropagation=fp.PointNetFeaturePropagation)


def seprate_point_cloud(fps_impl):
    """The reference's own seprate_point_cloud (utils/misc.py:205-256) bound to a stand-in for fps
    (utils/misc.py:13-20).  The function calls .cuda() on its viewpoints: run it under `cpu_cuda()`."""
    import random
    import torch
    import torch.nn.functional as F
    return lift("utils/misc.py", ["seprate_point_cloud"],
                {"torch": torch, "F": F, "random": random, "fps": fps_impl}).seprate_point_cloud


class cpu_cuda:
    """Context manager: Tensor.cuda() is the identity (lets the reference's host code run in a GPU-less container)."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._orig
