"""Tuning aid: the seg-shape interpolation forward under the selection kernel's threads-per-target variants
(UPP_INTERP_TPT) and the one-launch kernel (UPP_INTERP_PATH=0).  CUDA-event timed, L2 flushed between runs."""
import json
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
B, N, S, C, k = 32, 2048, 128, 1152, 3
x1 = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
x2 = (torch.rand(B, S, 3, generator=g) * 2 - 1).to(dev)
p2 = torch.randn(B, S, C, generator=g).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=20):
    ts = []
    for _ in range(3):
        fn()
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for env in ({"UPP_INTERP_PATH": "0"}, {"UPP_INTERP_SELECT": "0"}, {"UPP_INTERP_TPT": "1"}, {"UPP_INTERP_TPT": "2"},
            {"UPP_INTERP_TPT": "4"}):
    for kk in ("UPP_INTERP_PATH", "UPP_INTERP_SELECT", "UPP_INTERP_TPT"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    us = timeit(lambda: ops.interp_forward(x1, x2, p2, k, 1e-4))
    print(json.dumps({"op": f"interp_fwd B{B} N{N} S{S} C{C} k{k}", "env": env, "us": round(us, 2)}))
