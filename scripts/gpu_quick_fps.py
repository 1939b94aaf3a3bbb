"""Large-cloud FPS timings on the GPU box (CUDA-graph replay, CUDA events): default dispatch, one CTA per cloud, the
pruned kernels (UPP_FPS_PRUNED=1: shared-memory buckets, 2: register-resident rows)."""
import os
import sys
os.environ.setdefault("UPP_TUNING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), ROOT):
    sys.path.insert(0, p)
import json  # noqa: E402
import statistics  # noqa: E402
import torch  # noqa: E402
import upp_b200  # noqa: E402

o = upp_b200.ops
dev = torch.device("cuda:0")


def timeit(fn, reps=5, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for it in range(iters + 2):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(s.elapsed_time(e) / reps * 1e3)
    return round(statistics.median(ts), 2)


def env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


VARIANTS = (("default", {}), ("clusters + look-ahead", dict(UPP_FPS_CLUSTER_AHEAD=1)), ("one CTA per cloud", dict(UPP_FPS_CLUSTER=0)), ("buckets", dict(UPP_FPS_PRUNED=1)),
            ("rows", dict(UPP_FPS_PRUNED=2, UPP_FPS_PRUNED_MIN=63)))
SHAPES = ((128, 8192, 1024), (16, 8192, 1024), (32, 6144, 1024), (64, 6144, 1024), (32, 4096, 1024), (32, 2048, 1024),
          (32, 2500, 300), (64, 8192, 128), (32, 1228, 1024))
if len(sys.argv) > 1:
    SHAPES = tuple(tuple(int(v) for v in a.split("x")) for a in sys.argv[1:])
g = torch.Generator().manual_seed(0)
for kind in ("ball", "surface"):
    for B, N, M in SHAPES:
        x = torch.randn(B, N, 3, generator=g)
        if kind == "surface":
            x = x / x.norm(dim=2, keepdim=True)         # points on a sphere: what a scanned shape looks like
        else:
            x = x * 0.35
            x = x - x.mean(1, keepdim=True)
            x = x / x.norm(dim=2).max(dim=1)[0].view(-1, 1, 1)
        x = x.to(dev)
        want = None
        for tag, kw in VARIANTS:
            env(**kw)
            us = timeit(lambda: o.fps(x, M, True))
            got = o.fps(x, M, True)[0]
            env(**{kk: None for kk in kw})
            if want is None:
                want = got
            print(json.dumps(dict(op=f"fps {kind} B{B} N{N} M{M} [{tag}]", us=us, us_per_round=round(us / max(M - 1, 1), 4),
                                  same_as_default=bool(torch.equal(got, want)))), flush=True)
