"""Loader for oracle/_ref/chamfer_ref*.so: the reference's extensions/chamfer_dist
(chamfer.cu + chamfer_cuda.cpp) compiled UNMODIFIED for sm_100a by oracle/Makefile from the
sources where they lie under /root/reference.  It exposes forward/backward exactly as the
reference's `chamfer` module does and is the GPU-side oracle for the Chamfer rows
(TEST INFRASTRUCTURE ONLY; needs a GPU to run).
"""
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def path():
    hits = sorted(glob.glob(os.path.join(_HERE, "_ref", "chamfer_ref*.so")))
    return hits[0] if hits else None


def load():
    p = path()
    if p is None:
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("chamfer_ref", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
