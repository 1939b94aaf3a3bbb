"""Experiment: is the blend kernel bound by the 512-byte-strided output pattern?  Same output bytes (302 MB), same
sources, but C = 128 (one channel chunk: every CTA writes whole rows, consecutive targets contiguous) against C = 1152."""
import os
os.environ.setdefault("UPP_TUNING", "1")  # the UPP_* variant switches below are honoured only with this set
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
from upp_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
os.environ["UPP_INTERP_PATH"] = "1"
for (B, N, S, C) in [(32, 2048, 128, 1152), (32, 2048 * 9, 128, 128)]:
    x1 = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    x2 = (torch.rand(B, S, 3, generator=g) * 2 - 1).to(dev)
    p2 = torch.randn(B, S, C, generator=g).to(dev)
    for _ in range(3):
        ops.interp_forward(x1, x2, p2, 3, 1e-4)
    torch.cuda.synchronize()
