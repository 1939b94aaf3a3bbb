"""Drop-in `pointnet2_ops` package: resolves `from pointnet2_ops import pointnet2_utils`
(reference utils/misc.py:10) to the sm_100a implementation."""
from upp_b200 import pointnet2_utils  # noqa: F401
