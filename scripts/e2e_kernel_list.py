"""Device activities (kernels, copies) of ONE headline step through the module API and through the unchanged call sites,
in start order (torch profiler): what sits on the end-to-end critical path besides our kernels."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), os.path.join(ROOT, "iccv2025-upp_b200", "dropin"), ROOT):
    sys.path.insert(0, p)
import torch
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
W = bench.GpuWorkload(bench.HEADLINE, bench.make_inputs(bench.HEADLINE, 32, 0), dev, 1, None, "none")
_, keys = W.h2d_bytes()
W.h2d = {k: W.host[k] for k in keys}
dd = dict(W.d)
for api, tag in ((None, "modules"), (W.callsites(), "dropin")):
    for _ in range(3):
        W.run_modules(dd, api).item()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        W.run_modules(dd, api).item()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    print("==", tag, len(evs), "device activities")
    t0 = evs[0].time_range.start
    for e in evs:
        print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name[:90]}")
