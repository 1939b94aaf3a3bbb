"""CPU oracle for the UPP point-geometry hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``iccv2025-upp_b200/``) never does; it fails loudly without its CUDA library.

``oracle.c_oracle``      numpy front-end to ``upp_oracle.c`` (bit-exact fmaf restatement).
``oracle.torch_formulation``  the reference's pure-torch formulation (CPU baseline of record).
``oracle.f64``           float64 brute force used for tolerance checks.
``oracle.ref_gpu``       loader for ``oracle/_ref/chamfer_ref*.so`` = the reference's own
                         chamfer.cu compiled unmodified (GPU box only).

Parity status: Chamfer PINNED (against the reference's CUDA file and Python modules);
FPS / gather / kNN PARITY UNPINNED (third-party sources absent, see upp_oracle.c header).
"""
from . import c_oracle  # noqa: F401
