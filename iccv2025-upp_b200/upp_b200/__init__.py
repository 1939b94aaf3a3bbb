"""upp_b200 -- B200 (sm_100a) implementation of the UPP point-geometry hot path behind the
reference's own operator API.

    pointnet2_utils.furthest_point_sample / gather_operation   (pointnet2_ops)
    KNN                                                        (knn_cuda)
    chamfer.forward / chamfer.backward                         (extensions/chamfer_dist)
    fps, Group, ChamferFunction, ChamferDistanceL1/L2/L2_split (reference Python, mirrored)
    propagate, interpolate_features                            (kNN inverse-distance feature interpolation)
    knn_points                                                 (pytorch3d.ops.knn_points convention)
    misc.seprate_point_cloud                                   (batched crop + FPS, utils/misc.py:205-256)
    metrics.f_score / chamfer_distance_l1 / _l2                (utils/metrics.py:70-111 on the Chamfer kernels)
    parallel                                                   (batch sharding + NCCL loss all-reduce)

Everything computes in libupp_geom.so (hand-written CUDA); importing this package without the
built library raises -- there is no CPU or eager fallback.
"""
from . import _lib

_lib.load()  # fail loudly, at import, if the CUDA library is missing

from . import chamfer, metrics, misc, ops, parallel, pointnet2_utils  # noqa: E402,F401
from .knn import KNN  # noqa: E402,F401
from .modules import (ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split,  # noqa: E402,F401
                      ChamferFunction, Group, Selection, fps, interpolate_features, knn_points, propagate,
                      select_neighbors)

__version__ = "0.1.0"
launch_count = _lib.launch_count
