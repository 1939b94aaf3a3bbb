"""Drop-in `pointnet2_ops.pointnet2_utils` (the two symbols the UPP hot path uses)."""
from upp_b200.pointnet2_utils import (FurthestPointSampling, GatherOperation,  # noqa: F401
                                      furthest_point_sample, gather_operation)
