"""Per-op device timings at the BASELINE shapes (CUDA events, L2 flushed between iterations).
Tuning aid; the numbers of record come from bench.py."""
import argparse
import json
import os
os.environ.setdefault("UPP_TUNING", "1")  # the UPP_* variant switches below are honoured only with this set
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
import upp_b200  # noqa: E402
from upp_b200 import ops  # noqa: E402


def timeit(fn, iters=20, warm=5, flush=None, reps=4, graph=True):
    """Kernel-only time: the op is captured into a CUDA graph and replayed (no Python / launch gaps); graph=False times
    eager calls (functions with host-side random draws / pageable copies cannot be captured)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        fn = g.replay
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):  # CUDA-event resolution on this part is ~2 us: average a few replays
            fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3 / reps)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--sweep-fps", action="store_true")
    ap.add_argument("--sweep-chamfer", action="store_true")
    ap.add_argument("--sweep-chamfer2", action="store_true")
    ap.add_argument("--sweep-fps-cluster", action="store_true")
    ap.add_argument("--sweep-fps2", action="store_true", help="v2 FPS kernel: warps x point-pairs x stage-2 variant")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    rows = []

    def rec(name, fn, work=None, graph=True):
        if args.only and args.only not in name:
            return
        med, best = timeit(fn, flush=flush, graph=graph, reps=4 if graph else 1)
        r = {"op": name, "us_median": round(med, 2), "us_min": round(best, 2)}
        if work:
            r.update(work(med))
        rows.append(r)
        print(json.dumps(r), flush=True)

    if args.sweep_fps2:
        shapes = [(32, 1228, 1024), (32, 1024, 256), (32, 64, 512), (32, 32, 512), (128, 1024, 512), (32, 2048, 512), (32, 1843, 1536),
                  (32, 1536, 512), (1, 2048, 1024), (512, 1024, 256), (32, 256, 256), (32, 128, 256), (32, 512, 256), (32, 384, 256)]
        if args.only == "big":
            shapes = [(128, 8192, 1024), (1, 6144, 1024), (32, 4096, 512), (32, 3000, 512), (32, 2500, 512), (128, 8192, 128)]
            args.only = ""
        for (B, N, M) in shapes:
            x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
            os.environ["UPP_FPS_IMPL"] = "1"
            want = ops.fps(x, M)
            rec(f"fps2-sweep B{B} N{N} M{M} v1", lambda: ops.fps(x, M), lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4)})
            os.environ.pop("UPP_FPS_IMPL")
            rec(f"fps2-sweep B{B} N{N} M{M} v2-default", lambda: ops.fps(x, M), lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4)})
            for search in (0, 2):
                for nw in (1, 2, 4, 8, 16, 32):
                    p2 = (N + nw * 64 - 1) // (nw * 64)
                    if search == 2:  # deferred tree search: the instantiated large-cloud combinations
                        if nw == 8 and p2 > 8:
                            p2 += p2 % 2
                        if not ((nw == 4 and 3 <= p2 <= 8) or (nw == 8 and 3 <= p2 <= 16) or (nw == 16 and 3 <= p2 <= 8) or (nw == 32 and 2 <= p2 <= 4)):
                            continue
                    elif p2 > (4 if nw == 32 else 8):
                        continue
                    for s2 in ((0,) if nw <= 2 or (nw == 4 and search == 2) else (1,) if nw == 32 else (0, 1)):
                        os.environ["UPP_FPS_NW"], os.environ["UPP_FPS_P2"], os.environ["UPP_FPS_S2"] = str(nw), str(p2), str(s2)
                        os.environ["UPP_FPS_SEARCH"] = str(search)
                        ok = bool(torch.equal(ops.fps(x, M), want))
                        rec(f"fps2-sweep B{B} N{N} M{M} nw{nw} p2={p2} s2={s2} search={search}", lambda: ops.fps(x, M),
                            lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4), "matches_v1": ok})
            for k in ("UPP_FPS_NW", "UPP_FPS_P2", "UPP_FPS_S2", "UPP_FPS_SEARCH"):
                os.environ.pop(k, None)
        return
    if args.sweep_fps_cluster:  # one cloud per cluster of CTAs against one CTA per cloud (UPP_FPS_CLUSTER=0)
        shapes = [(32, 6144, 1024), (32, 4096, 1024), (16, 8192, 1024), (8, 8192, 1024), (1, 6144, 1024),
                  (32, 2048, 1024), (64, 8192, 1024), (16, 4096, 512)]
        if args.only == "mid":  # where between 2048 and 4096 points the cluster starts to pay
            shapes = [(B, N, 512) for N in (2560, 3072, 3584, 5120) for B in (8, 32, 64)] + [(64, 4096, 512), (72, 6144, 512)]
            args.only = ""
        for (B, N, M) in shapes:
            x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
            want = None
            for cs in ("0", "2", "4", "8", None):
                if cs is None:
                    os.environ.pop("UPP_FPS_CLUSTER", None)
                else:
                    os.environ["UPP_FPS_CLUSTER"] = cs
                got = ops.fps(x, M)
                want = got if want is None else want
                rec(f"fps-cluster B{B} N{N} M{M} cluster={cs or 'auto'}", lambda: ops.fps(x, M),
                    lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4), "same_as_single_cta": bool(torch.equal(got, want))})
            os.environ.pop("UPP_FPS_CLUSTER", None)
        return
    if args.sweep_chamfer2:  # residency variants of the packed kernel (register cap -> CTAs per SM)
        for (B, N, M) in [(64, 2048, 2048), (32, 1024, 1024), (64, 2048, 8192)]:
            a = torch.rand(B, N, 3, generator=g).to(dev)
            b = torch.rand(B, M, 3, generator=g).to(dev)
            for v, name in [(-1, "default R8W4 83 regs 5/SM"), (37, "R8W4 80 regs 6/SM"), (38, "R8W4 <=128 regs 4/SM"),
                            (39, "R8W4 <=168 regs 3/SM"), (40, "R8W8 <=80 regs 3/SM")]:
                for ch in (0, 1, 2, 4):
                    os.environ["UPP_CH_VARIANT"] = str(v)
                    os.environ["UPP_CH_CHUNKS"] = str(ch)
                    rec(f"chamfer-sweep2 {name} chunks={ch or 'auto'} B{B} N{N} M{M} +sums", lambda: ops.chamfer_forward(a, b, True),
                        lambda us: {"tflops_8NM": round(8.0 * N * M * B / us / 1e6, 2)})
            os.environ.pop("UPP_CH_CHUNKS", None)
            os.environ.pop("UPP_CH_VARIANT", None)
        return
    if args.sweep_chamfer:
        for (B, N, M) in [(64, 2048, 2048), (32, 1024, 1024), (64, 2048, 8192), (64, 1024, 1024), (64, 32, 1024), (8, 2048, 2048)]:
            a = torch.rand(B, N, 3, generator=g).to(dev)
            b = torch.rand(B, M, 3, generator=g).to(dev)
            for v, ch, name in [(0, 0, "2pass R2T128"), (30, 1, "packed R8W4 1 chunk"), (30, 0, "packed R8W4 auto"),
                                (30, 2, "packed R8W4 2 chunks"), (30, 4, "packed R8W4 4 chunks"),
                                (32, 0, "packed R8W8 auto"), (34, 0, "packed R12W4 auto"), (34, 1, "packed R12W4 1 chunk"),
                                (35, 0, "packed R16W4 auto"), (35, 1, "packed R16W4 1 chunk"), (35, 2, "packed R16W4 2 chunks"),
                                (36, 0, "packed R16W2 auto"), (36, 1, "packed R16W2 1 chunk"), (-1, 0, "default")]:
                os.environ["UPP_CH_VARIANT"] = str(v)
                os.environ["UPP_CH_CHUNKS"] = str(ch)
                rec(f"chamfer-sweep {name} B{B} N{N} M{M} +sums", lambda: ops.chamfer_forward(a, b, True),
                    lambda us: {"tflops_8NM": round(8.0 * N * M * B / us / 1e6, 2)})
            os.environ.pop("UPP_CH_CHUNKS", None)
            os.environ.pop("UPP_CH_VARIANT", None)
        return
    if args.sweep_fps:
        for (B, N, M) in [(32, 1228, 1024), (32, 1024, 512), (32, 2048, 512), (32, 1536, 512), (32, 512, 256), (32, 256, 256), (32, 128, 128),
                          (32, 64, 64), (512, 1024, 256)]:
            x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
            rec(f"fps-sweep B{B} N{N} M{M} default", lambda: ops.fps(x, M), lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4)})
            if N <= 2048:
                os.environ["UPP_FPS_W4"] = "1"
                rec(f"fps-sweep B{B} N{N} M{M} W4", lambda: ops.fps(x, M), lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4)})
                os.environ.pop("UPP_FPS_W4", None)
                continue
            for p in (1, 2, 3, 4, 6, 8, 12, 16):
                threads = ((N + p - 1) // p + 31) // 32 * 32
                if threads > (512 if p > 8 else 1024) or threads < 32:
                    continue
                os.environ["UPP_FPS_THREADS"], os.environ["UPP_FPS_P"] = str(threads), str(p)
                rec(f"fps-sweep B{B} N{N} M{M} T{threads} P{p} (runtime-T)", lambda: ops.fps(x, M),
                    lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4)})
            os.environ.pop("UPP_FPS_THREADS", None)
            os.environ.pop("UPP_FPS_P", None)
        return
    for (B, N, M) in [(32, 1024, 64), (32, 1096, 32), (32, 32, 32), (32, 972, 32), (32, 1024, 256), (32, 1228, 1024),
                      (32, 64, 32), (128, 8192, 1024), (128, 1024, 64), (32, 2048, 128), (32, 1843, 1536),
                      (1, 6144, 1024), (1, 2048, 1024), (16, 8192, 1024)]:
        x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
        rec(f"fps B{B} N{N} M{M}", lambda: ops.fps(x, M),
            lambda us: {"us_per_iter": round(us / max(M - 1, 1), 4), "gflops": round(8.0 * N * (M - 1) * B / us / 1e3, 1)})
    # seprate_point_cloud as the runners call it (tools/runner_module.py:131: 8192 points, crop in [1/4, 3/4] of them,
    # both sides resampled to 1024): the reference's 2*B batch-1 FPS launches are two batched ones here, the larger
    # (up to 6144 points, B <= 37) on clusters of CTAs.  UPP_FPS_CLUSTER=0 gives the one-CTA-per-cloud number.
    if not args.only or "seprate" in args.only:
        import random
        x = (torch.rand(32, 8192, 3, generator=g) * 2 - 1).to(dev)
        for crop in (2048, 4096, 6144):
            for cl in (None, "0"):
                if cl is None:
                    os.environ.pop("UPP_FPS_CLUSTER", None)
                else:
                    os.environ["UPP_FPS_CLUSTER"] = cl

                def run():
                    random.seed(1)
                    torch.manual_seed(1)
                    return upp_b200.misc.seprate_point_cloud(x, 8192, crop, sample_points=1024)
                rec(f"seprate_point_cloud B32 N8192 crop{crop} -> 1024 + 1024 (eager){' [UPP_FPS_CLUSTER=0]' if cl else ''}", run,
                    graph=False)
        os.environ.pop("UPP_FPS_CLUSTER", None)
    for (B, N, Q, k) in [(32, 1024, 64, 32), (32, 1096, 32, 16), (32, 32, 32, 16), (32, 64, 32, 8), (128, 1024, 64, 32),
                         (32, 2048, 128, 32), (32, 1536, 128, 32)]:
        r = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
        q = r[:, :Q].contiguous()
        rec(f"knn B{B} N{N} Q{Q} k{k}", lambda: ops.knn(r, q, k), lambda us: {"gflops": round(8.0 * N * Q * B / us / 1e3, 1)})
    for (B, N, G, k) in [(32, 1024, 64, 32), (128, 1024, 64, 32), (32, 2048, 128, 32)]:
        x = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
        rec(f"group B{B} N{N} G{G} k{k}", lambda: ops.group(x, G, k))
    for (B, N, S, C, k) in [(32, 2048, 128, 1152, 3), (32, 64, 32, 384, 8), (32, 1096, 32, 96, 16)]:
        if args.only and "interp" not in args.only:
            break
        x1 = (torch.rand(B, N, 3, generator=g) * 2 - 1).to(dev)
        x2 = (torch.rand(B, S, 3, generator=g) * 2 - 1).to(dev)
        p2, go = torch.randn(B, S, C, generator=g).to(dev), torch.randn(B, N, C, generator=g).to(dev)
        hbm = 4.0 * B * (N * C + S * C)
        out, idx, w, d = ops.interp_forward(x1, x2, p2, k, 1e-4)
        # default dispatch, then the one-launch / source-side kernels (0) and the wide-feature kernels (1) forced
        for path, tag in ((None, ""), ("0", " [UPP_INTERP_PATH=0]"), ("1", " [UPP_INTERP_PATH=1]")):
            if path is None:
                os.environ.pop("UPP_INTERP_PATH", None)
            else:
                os.environ["UPP_INTERP_PATH"] = path
            rec(f"interp_fwd B{B} N{N} S{S} C{C} k{k}{tag}", lambda: ops.interp_forward(x1, x2, p2, k, 1e-4),
                lambda us: {"gbs": round(hbm / us / 1e3, 1)})
            rec(f"interp_bwd_feat B{B} N{N} S{S} C{C} k{k}{tag}", lambda: ops.interp_backward(go, idx, w, S),
                lambda us: {"gbs": round(hbm / us / 1e3, 1)})
            rec(f"interp_bwd_feat+xyz B{B} N{N} S{S} C{C} k{k}{tag}",
                lambda: ops.interp_backward(go, idx, w, S, xyz_terms=(d, p2, x1, x2, 1e-4)), lambda us: {"gbs": round(hbm / us / 1e3, 1)})
        os.environ.pop("UPP_INTERP_PATH", None)
    if args.only and "interp" in args.only:
        return
    try:
        from oracle import ref_gpu
        ref = ref_gpu.load()   # the reference's own chamfer.cu, compiled unmodified (legacy default stream)
    except Exception:
        ref = None
    for (B, N, M) in [(64, 2048, 2048), (64, 1024, 1024), (64, 32, 1024), (64, 2048, 8192), (32, 1024, 1024)]:
        a = torch.rand(B, N, 3, generator=g).to(dev)
        b = torch.rand(B, M, 3, generator=g).to(dev)
        if ref is not None and not args.only:
            o = ref.forward(a, b)
            g1r, g2r = torch.rand_like(o[0]), torch.rand_like(o[1])
            for name, fn in (("REFERENCE chamfer_fwd", lambda: ref.forward(a, b)),
                             ("REFERENCE chamfer_bwd", lambda: ref.backward(a, b, o[2], o[3], g1r, g2r))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(10):
                    flush.zero_()
                    torch.cuda.synchronize()
                    s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0.record()
                    fn()
                    e0.record()
                    torch.cuda.synchronize()
                    ts.append(s0.elapsed_time(e0) * 1e3)
                ts.sort()
                r = {"op": f"{name} B{B} N{N} M{M} (eager, incl. torch::zeros)", "us_median": round(ts[len(ts) // 2], 2)}
                rows.append(r)
                print(json.dumps(r), flush=True)
        rec(f"chamfer_fwd B{B} N{N} M{M}", lambda: ops.chamfer_forward(a, b),
            lambda us: {"tflops_8NM": round(8.0 * N * M * B / us / 1e6, 2), "tflops_16NM": round(16.0 * N * M * B / us / 1e6, 2)})
        d1, d2, i1, i2 = ops.chamfer_forward(a, b)
        g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
        rec(f"chamfer_bwd B{B} N{N} M{M}", lambda: ops.chamfer_backward(a, b, i1, i2, g1, g2),
            lambda us: {"gbs": round(56.0 * (N + M) * B / us / 1e3, 1)})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "time_ops.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
