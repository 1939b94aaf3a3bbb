"""Quick per-op timings on the GPU box (CUDA-graph replay, CUDA events): the ops changed in round 2, with their A/B switches."""
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


def timeit(fn, reps=20, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    ts = []
    for it in range(iters + 3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(s.elapsed_time(e) / reps * 1e3)
    return round(statistics.median(ts), 2)


def rec(name, us, **kw):
    print(json.dumps(dict(op=name, us=us, **kw)), flush=True)


def env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


g = torch.Generator().manual_seed(0)
# large-cloud FPS: Morton buckets + exact pruning vs the plain kernels (one CTA per cloud / clusters of CTAs)
for B, N, M in ((128, 8192, 1024), (16, 8192, 1024), (32, 6144, 1024), (32, 4096, 1024), (32, 2500, 300), (64, 8192, 128)):
    x = (torch.randn(B, N, 3, generator=g) * 0.35)
    x = x - x.mean(1, keepdim=True)
    x = (x / x.norm(dim=2).max(dim=1)[0].view(-1, 1, 1)).to(dev)
    for tag, kw in (("default", {}), ("pruned nw8", dict(UPP_FPS_PRUNED=1)), ("pruned nw4", dict(UPP_FPS_PRUNED=1, UPP_FPS_PRUNED_NW=4)),
                    ("pruned nw16", dict(UPP_FPS_PRUNED=1, UPP_FPS_PRUNED_NW=16)), ("one CTA per cloud", dict(UPP_FPS_CLUSTER=0))):
        env(**kw)
        rec(f"fps B{B} N{N} M{M} [{tag}]", timeit(lambda: o.fps(x, M, True), reps=5, iters=5))
        env(**{kk: None for kk in kw})
# Group: fused (cluster shapes) vs two launches
for B, N, G, k in ((32, 1024, 64, 32), (128, 1024, 64, 32), (32, 2048, 128, 32), (32, 1096, 32, 16), (32, 64, 32, 8), (64, 1024, 64, 32), (16, 1024, 64, 32)):
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    for tag, kw in (("default", {}), ("plain all-lanes sleep40", dict(UPP_GROUP_SYNC=2 | (40 << 4))), ("plain lane0 sleep40", dict(UPP_GROUP_SYNC=0 | (40 << 4))),
                    ("plain lane0 spin", dict(UPP_GROUP_SYNC=0)), ("atomic spin", dict(UPP_GROUP_SYNC=1)), ("atomic sleep100", dict(UPP_GROUP_SYNC=1 | (100 << 4))),
                    ("atomic sleep20", dict(UPP_GROUP_SYNC=1 | (20 << 4))), ("two_launch", dict(UPP_GROUP_FUSED=0)), ("cs1", dict(UPP_GROUP_CLUSTER=1)), ("cs2", dict(UPP_GROUP_CLUSTER=2)), ("cs3", dict(UPP_GROUP_CLUSTER=3)),
                    ("cs4", dict(UPP_GROUP_CLUSTER=4)), ("cs8_w8", dict(UPP_GROUP_CLUSTER=8, UPP_GROUP_WARPS=8)), ("cs4_w8", dict(UPP_GROUP_CLUSTER=4, UPP_GROUP_WARPS=8))):
        env(**kw)
        try:
            rec(f"group B{B} N{N} G{G} k{k} [{tag}]", timeit(lambda: o.group(x, G, k)))
        except Exception as ex:  # noqa: BLE001
            rec(f"group B{B} N{N} G{G} k{k} [{tag}]", None, error=str(ex)[:80])
        env(**{kk: None for kk in kw})
# Chamfer forward: fused vs keyed, chunk choices
for B, N, M in ((64, 2048, 2048), (32, 1024, 1024), (64, 2048, 8192), (64, 32, 1024)):
    a, b = torch.rand(B, N, 3, generator=g).to(dev), torch.rand(B, M, 3, generator=g).to(dev)
    for tag, kw in (("slots", {}), ("keyed", dict(UPP_CH_VARIANT=30)), ("folded cg8 6/SM", dict(UPP_CH_FUSED=1)), ("slots cg16", dict(UPP_CH_FUSED=5)),
                    ("slots cg8 5/SM", dict(UPP_CH_FUSED=6)), ("slots_c1", dict(UPP_CH_CHUNKS=1)), ("slots_c2", dict(UPP_CH_CHUNKS=2)),
                    ("slots_c3", dict(UPP_CH_CHUNKS=3)), ("slots_c4", dict(UPP_CH_CHUNKS=4)), ("slots_c5", dict(UPP_CH_CHUNKS=5)), ("slots_c6", dict(UPP_CH_CHUNKS=6)),
                    ("slots_c8", dict(UPP_CH_CHUNKS=8)), ("slots cg8 5/SM c4", dict(UPP_CH_FUSED=6, UPP_CH_CHUNKS=4)),
                    ("slots cg16 c4", dict(UPP_CH_FUSED=5, UPP_CH_CHUNKS=4))):
        env(**kw)
        rec(f"chamfer_fwd B{B} {N}x{M} +sums [{tag}]", timeit(lambda: o.chamfer_forward(a, b, want_sums=True)))
        env(**{kk: None for kk in kw})
    d1, d2, i1, i2 = o.chamfer_forward(a, b)
    g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
    rec(f"chamfer_bwd B{B} {N}x{M}", timeit(lambda: o.chamfer_backward(a, b, i1, i2, g1, g2)))
    rec(f"chamfer_bwd+stats B{B} {N}x{M}", timeit(lambda: o.chamfer_backward(a, b, i1, i2, g1, g2, want_sqnorm=True)))
# scatter-add backward kernels
for B, N, G, k in ((32, 1024, 64, 32), (32, 64, 32, 8), (32, 2048, 128, 32)):
    x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
    nb, ce, idx, cidx = o.group(x, G, k)
    gnb, gce = torch.randn_like(nb), torch.randn_like(ce)
    rec(f"group_bwd B{B} N{N} G{G} k{k}", timeit(lambda: o.group_backward(gnb, gce, idx, cidx, N)))
for B, N, M in ((32, 1228, 1024), (32, 1024, 256), (128, 8192, 1024)):
    rows = torch.randn(B, M, 3, generator=g).to(dev)
    idx = torch.stack([torch.randperm(N, generator=g)[:M] for _ in range(B)]).to(torch.int32).to(dev)
    rec(f"rows_scatter_add B{B} N{N} M{M}", timeit(lambda: o.rows_scatter_add(rows, idx, N)))

# seprate_point_cloud as the runners call it (B32, 8192 points, crop in [2048, 4096] -> both sides resampled to 1024)
import random  # noqa: E402
xyz = (torch.rand(32, 8192, 3, generator=g) * 2 - 1).to(dev)
for crop in (2048, [2048, 4096]):
    def run():
        random.seed(1)
        torch.manual_seed(1)
        return upp_b200.misc.seprate_point_cloud(xyz, 8192, crop, sample_points=1024)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    import time  # noqa: E402
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e6)
    rec(f"seprate_point_cloud B32 N8192 crop{crop} -> 1024 + 1024 (eager, wall clock incl. host RNG)", round(statistics.median(ts), 1))
c = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1).to(dev)
rec("crop_split B32 N8192 crop2048", timeit(lambda: o.crop_split(xyz, c, 2048)))
