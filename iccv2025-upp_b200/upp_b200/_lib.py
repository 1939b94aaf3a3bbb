"""ctypes binding of libupp_geom.so (include/upp_geom.h).

There is no CPU path and no fallback: if the CUDA library is missing or a call returns a
non-zero code this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.abspath(os.path.join(_HERE, os.pardir, "lib", "libupp_geom.so"))

_vp = ctypes.c_void_p
_i = ctypes.c_int
_sz = ctypes.c_size_t
_f = ctypes.c_float

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/upp_geom.h
_SIGNATURES = {
    "upp_version": [],
    "upp_error_string": [_i],
    "upp_launch_count": [],
    "upp_fps_workspace_bytes": [_i, _i, _i],
    "upp_fps_f32": [_vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp],
    "upp_gather_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "upp_gather_grad_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "upp_rows_scatter_add_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "upp_knn_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "upp_chamfer_fwd_workspace_bytes": [_i, _i, _i],
    "upp_chamfer_fwd_f32": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "upp_chamfer_fwd_sharded_f32": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp],
    "upp_peer_allreduce_finish_f32": [_vp, _vp, _vp],
    "upp_peer_allreduce_f32": [_vp, _vp, _vp, _vp],
    "upp_chamfer_bwd_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "upp_chamfer_bwd_stats_workspace_bytes": [_i, _i, _i],
    "upp_chamfer_bwd_stats_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp, _vp],
    "upp_group_f32": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp],
    "upp_group_bwd_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "upp_crop_split_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "upp_knn_points_f32": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "upp_interp_fwd_f32": [_vp, _vp, _vp, _vp, _f, _f, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "upp_interp_select_f32": [_vp, _vp, _f, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "upp_interp_blend_f32": [_vp, _vp, _f, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "upp_interp_bwd_workspace_bytes": [_i, _i, _i, _i, _i],
    "upp_interp_bwd_f32": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _vp],
}
_RESTYPES = {
    "upp_error_string": ctypes.c_char_p,
    "upp_launch_count": ctypes.c_ulonglong,
    "upp_fps_workspace_bytes": _sz,
    "upp_chamfer_fwd_workspace_bytes": _sz,
    "upp_chamfer_bwd_stats_workspace_bytes": _sz,
    "upp_interp_bwd_workspace_bytes": _sz,
}

EXPORTS = tuple(_SIGNATURES)

UPP_MAX_PEERS = 16


class PeerExchangeStruct(ctypes.Structure):
    """upp_peer_exchange of include/upp_geom.h."""
    _fields_ = [("slots", ctypes.c_void_p * UPP_MAX_PEERS), ("rank", ctypes.c_int), ("world", ctypes.c_int),
                ("seq", ctypes.c_void_p), ("defer", ctypes.c_int), ("status", ctypes.c_void_p),
                ("timeout_cycles", ctypes.c_longlong)]

_lib = None


def load():
    """Load libupp_geom.so once; raise ImportError (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"upp_b200: CUDA library not built: {LIB_PATH} is missing. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C "
            "iccv2025-upp_b200/csrc`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _i)
    _lib = lib
    return lib


def error_string(rc):
    return load().upp_error_string(int(rc)).decode()


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed: {error_string(rc)} (code {rc})")


def launch_count():
    return int(load().upp_launch_count())
