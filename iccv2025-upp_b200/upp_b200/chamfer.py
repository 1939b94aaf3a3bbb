"""Stand-in for the reference's compiled `chamfer` extension module
(extensions/chamfer_dist/chamfer_cuda.cpp:36-39): forward / backward with the same argument
order and return lists, so extensions/chamfer_dist/__init__.py runs unmodified on top of it."""
from . import ops


def forward(xyz1, xyz2):
    """-> [dist1, dist2, idx1, idx2] (chamfer.cu:147-171)."""
    return ops.chamfer_forward(xyz1, xyz2)


def backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    """-> [grad_xyz1, grad_xyz2] (chamfer.cu:203-229)."""
    return ops.chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
