#!/usr/bin/env python
"""bench.py -- clouds/s for the UPP point-geometry hot path (FPS + kNN Group + Chamfer fwd/bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic clouds.  The headline workload
`upp_cls_geometry+chamfer` = every FPS / Group call of one UPP ModelNet40 classification forward+backward
(BASELINE.json configs[1]; shape census SURVEY.md 3.1 / Appendix A, B=32, 1024 points + 72 noise points)
followed by the Completion-Prompter Chamfer-L1 fwd+bwd on the rebuilt 1024 points (tools/runner_pretask.py:222).

Prints ONE JSON line (rank 0).  Top level = the headline workload:
  value        device-resident throughput (inputs in HBM, CUDA-graph replay, CUDA events, L2 flushed between steps,
               max over ranks).  K steps form one timed region; the region is repeated until >= 200 steps have been
               timed and the MEDIAN region is reported (`timed_regions`), so the GPU is busy for seconds, not 5 ms.
  e2e          the same metric through the public module API with pinned-host inputs copied H2D inside the step and
               the loss read back D2H every step;  e2e_dropin: through the reference's UNCHANGED call sites
               (reference_callsites.py: misc.fps -> pointnet2_utils + 2 transposes -> KNN -> index gather) over the
               drop-in pointnet2_ops / knn_cuda / chamfer modules.
  roofline     the dominant kernel measured live with CUDA events, against MEASURED_PEAKS.json
  cpu_baseline the reference's pure-torch formulation (oracle/torch_formulation.py) on this box's host cores
               (all cores, and the reference's own OMP_NUM_THREADS=5, main.py:2-3)
  configs      one sub-record per other BASELINE.json config -- c1 (configs[0]), c3 (configs[2]), c4 (configs[3]),
               c5 (configs[4]) -- each with value / ms_per_step / e2e / kernels / roofline (/ cpu_baseline at N=1);
               c3 also times the reference's own chamfer.cu (oracle/_ref, compiled unmodified) on the same GPU, and
               carries `training_shapes`: forward / backward of the three shapes the pre-training calls Chamfer on.
  step_ms      p10 / p50 / p90 of the per-step CUDA-event times (device arm and every e2e arm)
  e2e_dropin_launch_blocking   the headline step through the unchanged call sites, eager, with CUDA_LAUNCH_BLOCKING=1
               as the reference ships (main.py:5) -- a child process, host wall clock
  strong       (N > 1) C4 with its GLOBAL batch of 128 clouds split over the N ranks (BASELINE.json configs[3]).
`--workload c1|c3|c4|c5` runs that config alone as the top-level record (profiling); `--no-configs` skips the
sub-records.  N > 1 (torchrun): batch sharded, the Chamfer-loss all-reduce fused into the kernels over NVLink peer
memory (NCCL fallback), whole step still one CUDA graph per rank.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "iccv2025-upp_b200"), os.path.join(ROOT, "iccv2025-upp_b200", "dropin"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "point clouds/sec for FPS+kNN group+Chamfer fwd/bwd"
UNIT = "clouds/s"
N_SM, FP32_LANES = 148, 128
HEADLINE = "upp_cls_geometry+chamfer"
MIN_TIMED_STEPS = 200


# ----------------------------------------------------------------------------- inputs --------

def unit_sphere(x):
    x = x - x.mean(dim=1, keepdim=True)
    return x / x.norm(dim=2).max(dim=1)[0].view(-1, 1, 1)


def make_inputs(workload, B, seed):
    """Synthetic host tensors of the workload's shapes (float32, CPU)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)  # noqa: E731
    if workload == HEADLINE:
        clean = unit_sphere(torch.randn(B, 1024, 3, generator=g) * 0.35)
        lidar = clean[:, torch.randint(0, 1024, (48,), generator=g)] * (1.02 + 0.28 * r(1, 48, 1))
        gn = torch.randn(B, 24, 3, generator=g) * 0.2
        gn = gn + gn / gn.norm(dim=-1, keepdim=True) * 0.9
        return {"pts": torch.cat([clean, lidar, gn], 1).contiguous(),          # (B,1096,3) misc.py:28-46
                "rebuild": (clean[:, torch.randperm(1024, generator=g)] + 0.02 * torch.randn(B, 1024, 3, generator=g)).contiguous(),
                "target": clean.contiguous(),
                "w_nb": torch.randn(B, 64, 32, 3, generator=g), "w_c5": torch.randn(B, 32, 3, generator=g)}
    if workload == "c1":
        return {"pts": (r(B, 1024, 3) * 2 - 1).contiguous()}
    if workload == "c3":
        return {"xyz1": r(B, 2048, 3).contiguous(), "xyz2": r(B, 2048, 3).contiguous()}
    if workload == "c4":
        return {"pts": unit_sphere(torch.randn(B, 8192, 3, generator=g) * 0.35).contiguous()}
    if workload == "c5":
        return {"pts": unit_sphere(torch.randn(B, 2048, 3, generator=g) * 0.35).contiguous(),
                "feat": torch.randn(B, 128, 1152, generator=g),  # (B,128,1152) group features to propagate
                # what the segmentation head hands back for the propagated features (device-resident, like the features)
                "w_up": torch.randn(B, 2048, 1152, generator=g)}
    raise SystemExit(f"unknown workload {workload}")


DEFAULT_B = {HEADLINE: 32, "c1": 32, "c3": 64, "c4": 128, "c5": 32}
DESCR = {
    HEADLINE: "all FPS/Group calls of one UPP ModelNet40 cls fwd+bwd (BASELINE configs[1] census: "
              "Group(32,16)x3, Group(64,32), Group(32,8), fps 1024->256, fps 1228->1024) + "
              "Chamfer-L1 fwd+bwd 1024 vs 1024, B=32 per GPU",
    "c1": "Point-MAE Group divider FPS 64 + kNN k=32, B=32, N=1024 (BASELINE configs[0])",
    "c3": "Chamfer L1 fwd+bwd B=64, 2048 vs 2048 (BASELINE configs[2])",
    "c4": "ShapeNet55-scale grouping FPS 8192->1024 then Group(64,32), B=128 (BASELINE configs[3])",
    "c5": "ShapeNetPart Group(128,32) on 2048 points + kNN feature propagation 2048<-128, 3-NN, 1152-d, forward and "
          "feature-gradient backward, B=32 (BASELINE configs[4], geometry part)",
}
L2_NOTE = "flushed between timed steps (256 MiB memset outside the event pair); inputs are << L2"


def config_of(workload, B):
    """Identical in both arms (the driver compares the two lines' `config`); run details live in `run_info`."""
    return {"workload": workload, "description": DESCR[workload], "clouds_per_gpu": B, "l2": L2_NOTE}


# ----------------------------------------------------------------------------- GPU steps -----

class GpuWorkload:
    """The step expressed three times: `run_ops` on the C-ABI-level ops (static, CUDA-graph friendly), `run_modules`
    on the public module API with autograd (upp_b200.Group / fps / ChamferDistanceL1), and `run_callsites` on the
    reference's unchanged call sites over the drop-in third-party modules (reference_callsites.py)."""

    def __init__(self, name, host, dev, world, peers, collective):
        import upp_b200
        self.U, self.ops, self.par = upp_b200, upp_b200.ops, upp_b200.parallel
        self.name, self.dev, self.world = name, dev, world
        self.host = {k: v.pin_memory() for k, v in host.items()}
        self.d = {k: v.to(dev) for k, v in host.items()}
        self.B = next(iter(host.values())).shape[0]
        self.collect = None  # when set: list collecting (label, fn, args, kwargs) of every labelled op
        self.h2d = None      # e2e arms: key -> pinned host tensor copied inside the step
        self.side = [torch.cuda.Stream(device=dev) for _ in range(2)]  # independent branches of the step
        # N > 1: the loss all-reduce runs inside the Chamfer kernels over NVLink peer memory when the peers'
        # buffers can be mapped (parallel.PeerExchange); otherwise one NCCL all-reduce of 16 bytes
        self.peers, self.collective = peers, collective
        U = upp_b200
        self.g32_16, self.g64_32, self.g32_8 = U.Group(32, 16), U.Group(64, 32), U.Group(32, 8)
        self.g128_32 = U.Group(128, 32)
        self.cd_l1 = U.ChamferDistanceL1()
        self._cs = None

    def callsites(self):
        if self._cs is None:
            import reference_callsites as R
            self._cs = {"R": R, "g32_16": R.Group(32, 16), "g64_32": R.Group(64, 32), "g32_8": R.Group(32, 8),
                        "g128_32": R.Group(128, 32), "cd_l1": R.ChamferDistanceL1()}
        return self._cs

    # -- per-op bookkeeping (roofline pass times every labelled op alone, as its own CUDA graph) --
    def _t(self, label, fn, *a, **k):
        if self.collect is not None:
            self.collect.append((label, fn, a, k))
        return fn(*a, **k)

    def _fork(self):
        cur = torch.cuda.current_stream()
        for st in self.side:
            st.wait_stream(cur)
        return cur

    def _join(self, cur):
        for st in self.side:
            cur.wait_stream(st)

    def h2d_bytes(self):
        keys = {HEADLINE: ("pts", "rebuild", "target"), "c5": ("pts",)}.get(self.name, tuple(self.host))
        return sum((self.host[k].numel() + 3) // 4 * 16 for k in keys), keys

    # -- ops-level step --
    def run_ops(self, d):
        o, t = self.ops, self._t
        n = self.name
        if n == HEADLINE:
            # The step is a DAG with three independent branches: the serial FPS chain of the
            # completion stage + downstream grouping (critical path), the rectification-stage groupings, and
            # the Chamfer loss.  They run on three streams (fork/join; captured as one CUDA graph), so the
            # short branches fill the SMs the one-CTA-per-cloud FPS chain leaves idle.
            B = self.B
            nglob = float(B * self.world * 1024)
            keep = d["pts"][:, :972].contiguous()
            cur = self._fork()
            with torch.cuda.stream(self.side[0]):
                g1 = t("group N1096 G32 k16", o.group, d["pts"], 32, 16)
                t("group N32 G32 k16", o.group, g1[1], 32, 16)
                t("group N972 G32 k16", o.group, keep, 32, 16)
            with torch.cuda.stream(self.side[1]):
                if self.peers is not None:  # deferred exchange: the kernels send now, the wait + sum closes the step
                    d1, d2, j1, j2, sums = t("chamfer_fwd N1024 M1024", o.chamfer_forward_sharded, d["rebuild"], d["target"], self.peers, True)
                else:
                    d1, d2, j1, j2, sums = t("chamfer_fwd N1024 M1024", o.chamfer_forward, d["rebuild"], d["target"], True)
                    if self.world > 1:
                        self.par.reduce_sums(sums)
                loss = (sums[2] + sums[3]) / (2.0 * nglob)
                gd1 = (0.25 / nglob) / torch.sqrt(d1)
                gd2 = (0.25 / nglob) / torch.sqrt(d2)
                ga, _ = t("chamfer_bwd N1024 M1024", o.chamfer_backward, d["rebuild"], d["target"], j1, j2, gd1, gd2)
            i1, c1 = t("fps N1024 M256", o.fps, d["rebuild"], 256, True)
            cat = torch.cat([keep, c1], 1)
            i2, c2 = t("fps N1228 M1024", o.fps, cat, 1024, True)
            g4 = t("group N1024 G64 k32", o.group, c2, 64, 32)
            if self.peers is not None:
                # close the deferred exchange on the Chamfer side stream, but only once the critical path has come this
                # far: the peers have had ~0.23 ms to deliver, and the wait + loss arithmetic overlaps the chain's tail
                late = torch.cuda.Event()
                late.record(torch.cuda.current_stream())
                with torch.cuda.stream(self.side[1]):
                    self.side[1].wait_event(late)
                    sums = o.peer_allreduce_finish(self.peers, self.dev)
                    loss = (sums[2] + sums[3]) / (2.0 * nglob)
            g5 = t("group N64 G32 k8", o.group, g4[1], 32, 8)
            # backward chain: G5 -> G4 -> fps(1228->1024) gather -> fps(1024->256) gather (row-major scatter-adds);
            # four separate autograd nodes in the model (the encoder's backward sits between them), so four launches
            gc4 = t("group_bwd N64", o.group_backward, torch.zeros_like(g5[0]), d["w_c5"], g5[2], g5[3], 64)
            gx = t("group_bwd N1024", o.group_backward, d["w_nb"], gc4, g4[2], g4[3], 1024)
            gcat = t("fps_gather_bwd N1228", o.rows_scatter_add, gx, i2, 1228)
            gc1 = gcat[:, 972:].contiguous()
            greb = t("fps_gather_bwd N1024", o.rows_scatter_add, gc1, i1, 1024)
            self._join(cur)
            self.grad = ga + greb
            self.keepalive = (g1, g4, g5, d1, d2)
            return loss
        if n == "c1":
            nb, ce, _, _ = t("group N1024 G64 k32", o.group, d["pts"], 64, 32)
            return ce[0, 0, 0]
        if n == "c3":
            nglob = float(self.B * self.world * 2048)
            if self.peers is not None:
                d1, d2, j1, j2, sums = t("chamfer_fwd N2048 M2048", o.chamfer_forward_sharded, d["xyz1"], d["xyz2"], self.peers, True)
            else:
                d1, d2, j1, j2, sums = t("chamfer_fwd N2048 M2048", o.chamfer_forward, d["xyz1"], d["xyz2"], True)
                if self.world > 1:
                    self.par.reduce_sums(sums)
            gd1, gd2 = (0.25 / nglob) / torch.sqrt(d1), (0.25 / nglob) / torch.sqrt(d2)
            self.grad = t("chamfer_bwd N2048 M2048", o.chamfer_backward, d["xyz1"], d["xyz2"], j1, j2, gd1, gd2)
            if self.peers is not None:
                sums = o.peer_allreduce_finish(self.peers, self.dev)
            return (sums[2] + sums[3]) / (2.0 * nglob)
        if n == "c4":
            _, c = t("fps N8192 M1024", o.fps, d["pts"], 1024, True)
            nb, ce, _, _ = t("group N1024 G64 k32", o.group, c, 64, 32)
            return ce[0, 0, 0]
        if n == "c5":
            nb, ce, _, _ = t("group N2048 G128 k32", o.group, d["pts"], 128, 32)
            up, j, w, _ = t("interp N2048 S128 C1152 k3", o.interp_forward, d["pts"], ce, d["feat"], 3, 1e-4)
            self.grad = t("interp_bwd N2048 S128 C1152 k3", o.interp_backward, d["w_up"], j, w, 128)[0]
            self.keepalive = (nb, up)
            return up[0, 0, 0]
        raise SystemExit(n)

    # -- module-level step (public API + autograd), used for e2e --
    def run_modules(self, d, api=None):
        """h2d (set by the e2e arms): key -> pinned host tensor.  The step's inputs are then copied host->device INSIDE
        the step, each on the branch that consumes it -- the critical chain waits for its own 393 KB only, the other
        two copies overlap it.  api = None: upp_b200's modules (fused Group, FPS + gather in one kernel);
        api = self.callsites(): the reference's unchanged call sites over the drop-in modules."""
        U, n = self.U, self.name
        h = self.h2d
        if api is None:
            fps, g32_16, g64_32, g32_8, g128_32, cd_l1 = U.fps, self.g32_16, self.g64_32, self.g32_8, self.g128_32, self.cd_l1
        else:
            fps, g32_16, g64_32, g32_8, g128_32, cd_l1 = (api["R"].fps, api["g32_16"], api["g64_32"], api["g32_8"],
                                                           api["g128_32"], api["cd_l1"])
        if n == HEADLINE:
            if h:
                d["rebuild"].copy_(h["rebuild"], non_blocking=True)
            reb = d["rebuild"].detach().requires_grad_(True)
            cur = self._fork()
            keep_ready = torch.cuda.Event()
            with torch.cuda.stream(self.side[0]):
                if h:
                    d["pts"].copy_(h["pts"], non_blocking=True)
                keep = d["pts"][:, :972].contiguous()
                keep_ready.record(self.side[0])
                _, ce1 = g32_16(d["pts"])
                g32_16(ce1)
                g32_16(keep)
            with torch.cuda.stream(self.side[1]):
                if h:
                    d["target"].copy_(h["target"], non_blocking=True)
                if self.world > 1 and api is None:
                    cd = self.par.sharded_chamfer(reb, d["target"], "l1", n_global_clouds=self.B * self.world, peers=self.peers)
                else:
                    cd = cd_l1(reb, d["target"])  # call sites unchanged: the rank-local mean, as the reference computes it
            c1, _ = fps(reb, 256)
            cur.wait_event(keep_ready)
            c2, _ = fps(torch.cat([keep, c1], 1), 1024)
            nb4, ce4 = g64_32(c2)
            _, ce5 = g32_8(ce4)
            self._join(cur)
            # backward through the public autograd Functions, seeded with the downstream gradients directly (what the
            # encoder would hand back for the neighbourhoods / centres) instead of a synthetic scalar head
            torch.autograd.backward([cd, nb4, ce5], [None, d["w_nb"], d["w_c5"]])
            self.grad = reb.grad
            return cd.detach()
        if h:
            for k, v in h.items():
                d[k].copy_(v, non_blocking=True)
        if n == "c3":
            a = d["xyz1"].detach().requires_grad_(True)
            b = d["xyz2"].detach().requires_grad_(True)
            if self.world > 1 and api is None:
                loss = self.par.sharded_chamfer(a, b, "l1", n_global_clouds=self.B * self.world, peers=self.peers)
            else:
                loss = cd_l1(a, b)
            loss.backward()
            return loss.detach()
        if n == "c1":
            nb, ce = g64_32(d["pts"])
            return ce[0, 0, 0]
        if n == "c4":
            c, _ = fps(d["pts"], 1024)
            nb, ce = g64_32(c)
            return ce[0, 0, 0]
        if n == "c5":
            nb, ce = g128_32(d["pts"])
            feat = d["feat"].detach().requires_grad_(True)
            up = U.interpolate_features(d["pts"], ce, feat, 3, eps=1e-4)
            up.backward(d["w_up"])
            self.grad = feat.grad
            return up.detach()[0, 0, 0]
        raise SystemExit(n)


# algorithmic work per launch (SURVEY.md 8d / DESIGN.md "Kernels"): label -> (flops, bytes) per cloud
def op_work(label):
    f = label.split()
    kind = f[0]
    v = {x[0]: int(x[1:]) for x in f[1:] if x[1:].isdigit()}
    if kind == "fps":
        N, M = v["N"], v["M"]
        return 8.0 * N * (M - 1), 12.0 * N + 16.0 * M
    if kind == "group":
        N, G, k = v["N"], v["G"], v["k"]
        return 8.0 * N * (G - 1) + 8.0 * G * N, 12.0 * N + 12.0 * G * k + 12.0 * G + 8.0 * G * k + 4.0 * G
    if kind == "chamfer_fwd":
        N, M = v["N"], v["M"]
        return 8.0 * N * M, 20.0 * (N + M)
    if kind == "chamfer_bwd":
        N, M = v["N"], v["M"]
        return 0.0, 56.0 * (N + M)
    if kind == "interp":  # HBM: out written once, features + coordinates read once (the k re-reads hit L2)
        N, S, C, k = v["N"], v["S"], v["C"], v["k"]
        return 0.0, 4.0 * N * C + 4.0 * S * C + 12.0 * (N + S) + 12.0 * N * k
    if kind == "interp_bwd":  # grad_out read once, grad_feat2 written once, the selection read once
        N, S, C, k = v["N"], v["S"], v["C"], v["k"]
        return 0.0, 4.0 * N * C + 4.0 * S * C + 8.0 * N * k
    return 0.0, 0.0


def traffic_for(label):
    """dram__bytes_read.sum + dram__bytes_write.sum of the op's dominant kernel, per launch, from the committed
    `ncu --set full` captures (profiles/traffic.json: label -> {bytes, kernel, source}); None when not captured."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(label)
    except (OSError, ValueError):
        return None
    return t["bytes"] if t else None


def load_peaks():
    """MEASURED_PEAKS.json (driver-written, git-ignored) when it travelled to this box, else the committed snapshot of
    the same driver measurement (profiles/measured_peaks_snapshot.json), else the profiling recipe's fallback."""
    for path, how in ((os.path.join(ROOT, "MEASURED_PEAKS.json"), "MEASURED_PEAKS.json"),
                      (os.path.join(ROOT, "profiles", "measured_peaks_snapshot.json"),
                       "profiles/measured_peaks_snapshot.json (committed copy of the driver's MEASURED_PEAKS.json)")):
        try:
            return json.load(open(path)), how
        except (OSError, ValueError):
            continue
    return {"hbm_gbs": 6650.0}, "fallback 6650 GB/s (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks --------

class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.first = [], None, 0
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        """Samples from here on (the last one taken before included) are the ones reported."""
        self.first = max(0, len(self.rows) - 1)

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        """Clocks and throttle reasons of the samples since mark() (the sampler keeps running)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms ------

def cpu_step(workload, host):
    """The reference's pure-torch formulation of the same step, on host cores."""
    from oracle import torch_formulation as T
    if workload == HEADLINE:
        _, ce1 = T.group(host["pts"], 32, 16)
        T.group(ce1, 32, 16)
        keep = host["pts"][:, :972]
        T.group(keep, 32, 16)
        reb = host["rebuild"].clone().requires_grad_(True)
        c1, _ = T.fps(reb, 256)
        c2, _ = T.fps(torch.cat([keep, c1], 1), 1024)
        nb4, ce4 = T.group(c2, 64, 32)
        _, ce5 = T.group(ce4, 32, 8)
        loss = T.chamfer_l1(reb, host["target"]) + ((nb4 * host["w_nb"]).sum() + (ce5 * host["w_c5"]).sum()) * 1e-9
        loss.backward()
        return float(loss.detach())
    if workload == "c1":
        return float(T.group(host["pts"], 64, 32)[1][0, 0, 0])
    if workload == "c3":
        a = host["xyz1"].clone().requires_grad_(True)
        b = host["xyz2"].clone().requires_grad_(True)
        loss = T.chamfer_l1(a, b)
        loss.backward()
        return float(loss.detach())
    if workload == "c4":
        c, _ = T.fps(host["pts"], 1024)
        return float(T.group(c, 64, 32)[1][0, 0, 0])
    if workload == "c5":
        _, ce = T.group(host["pts"], 128, 32)
        feat = host["feat"].clone().requires_grad_(True)
        up = T.interpolate(host["pts"], ce, feat, 3, 1e-4)
        up.backward(host["w_up"])
        return float(up.detach()[0, 0, 0])
    raise SystemExit(workload)


def cpu_sample_batch(workload, B):
    """Bounded sample: the CPU arm processes this many clouds per step (same shapes per cloud)."""
    cap = {HEADLINE: 32, "c1": 32, "c3": 8, "c4": 4, "c5": 32}[workload]
    return min(B, cap)


def time_cpu(workload, B, seed, steps, warmup, budget_s=25.0, threads=None):
    all_threads = os.cpu_count() or 1
    threads = min(threads or all_threads, all_threads)
    torch.set_num_threads(threads)
    bs = cpu_sample_batch(workload, B)
    host = make_inputs(workload, bs, seed)
    for _ in range(warmup):
        cpu_step(workload, host)
    ts, t_all = [], time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(workload, host)
        ts.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    torch.set_num_threads(all_threads)
    sec = statistics.median(ts)
    return {"value": bs / sec, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{bs} clouds/step x {len(ts)} steps of the same per-cloud shapes, pure-torch formulation "
                      f"(oracle/torch_formulation.py), median step {sec * 1e3:.1f} ms"}, sec, bs


def cpu_baseline_record(workload, B, steps=5, warmup=1, budget_s=12.0):
    """All host cores, plus the reference's own thread setting (OMP_NUM_THREADS=5, main.py:2-3)."""
    cb, _, _ = time_cpu(workload, B, 0, steps, warmup, budget_s)
    five, _, _ = time_cpu(workload, B, 0, max(2, steps // 2), 1, budget_s / 2, threads=5)
    cb["omp5"] = {"value": five["value"], "cores": five["cores"], "sample": five["sample"],
                  "note": "torch.set_num_threads(5): the reference pins OMP_NUM_THREADS=5 (main.py:2-3)"}
    return cb


# ----------------------------------------------------------------------------- measurement ---

class Ctx:
    """Per-process measurement context: device, ranks, the shared L2-flush buffer and streams."""

    def __init__(self, args, rank, local_rank, world):
        self.args, self.rank, self.world = args, rank, world
        self.dev = torch.device("cuda", local_rank)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.hp = torch.cuda.Stream(device=self.dev, priority=-1)
        self.regions = max(1, math.ceil(MIN_TIMED_STEPS / max(1, args.steps)))
        import torch.distributed as dist
        self.dist = dist

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, step_fn, regions=None):
        """R regions of EXACTLY K steps each: per step flush L2, then one CUDA-event pair around the step; a barrier +
        synchronize on both sides of every region.  Returns the median region's ms (max over ranks per region)."""
        K = self.args.steps
        out = []
        self.last_steps_ms = []
        for _ in range(regions or self.regions):
            self.barrier()
            evs = []
            for _ in range(K):
                self.flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                step_fn()
                e.record()
                evs.append((s, e))
            self.barrier()
            per = [s.elapsed_time(e) for s, e in evs]
            self.last_steps_ms.extend(per)
            out.append(sum(per))
        t = torch.tensor(out, dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.median()), [float(x) for x in t]


def step_percentiles(ctx):
    """p10 / p50 / p90 of the per-step CUDA-event times of the last timed() call on this rank."""
    xs = sorted(ctx.last_steps_ms)
    q = lambda f: round(xs[min(len(xs) - 1, int(f * len(xs)))], 5)  # noqa: E731
    return {"p10": q(0.10), "p50": q(0.50), "p90": q(0.90), "n": len(xs)}


def capture(ctx, fn):
    """fn() -> result captured into a CUDA graph on the high-priority stream; (replay, result, mode, launches)."""
    import upp_b200
    if not ctx.args.no_graph:
        try:
            g = torch.cuda.CUDAGraph()
            c0 = upp_b200.launch_count()
            with torch.cuda.graph(g, stream=ctx.hp):
                res = fn()
            return g.replay, res, "cuda_graph_replay", upp_b200.launch_count() - c0, None
        except Exception as ex:  # capture unsupported: eager launches, and say so
            torch.cuda.synchronize()
            err = str(ex)[:160]
    else:
        err = None
    c0 = upp_b200.launch_count()
    res = fn()
    return None, res, "eager", upp_b200.launch_count() - c0, err


def time_gpu_reference_chamfer(W):
    """The reference's OWN chamfer.cu (compiled unmodified into oracle/_ref by oracle/Makefile), timed on this GPU on the
    c3 tensors: the GPU-vs-GPU anchor beside the GPU-vs-CPU ratio.  A baseline leg: the checker is timed, never shipped."""
    from oracle import ref_gpu
    ref = ref_gpu.load()
    if ref is None:
        return {"unavailable": "oracle/_ref/chamfer_ref*.so not built (needs /root/reference at build time)"}
    a, b = W.d["xyz1"], W.d["xyz2"]
    torch.cuda.synchronize()
    out = ref.forward(a, b)
    g1, g2 = torch.rand_like(out[0]), torch.rand_like(out[1])
    res = {}
    for name, fn in (("fwd", lambda: ref.forward(a, b)), ("bwd", lambda: ref.backward(a, b, out[2], out[3], g1, g2))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):  # the reference launches on the legacy default stream: events on it, synchronise around
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(torch.cuda.default_stream())
            fn()
            e.record(torch.cuda.default_stream())
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        res[name + "_ms"] = round(statistics.median(ts), 5)
    res["what"] = ("reference extensions/chamfer_dist/chamfer.cu compiled unmodified for sm_100a (oracle/_ref), B=%d 2048 vs 2048, "
                   "its own allocation (torch::zeros) included, CUDA events on the legacy default stream" % a.shape[0])
    return res


def time_chamfer_training_shapes(dev, B):
    """the three shapes the Completion-Prompter pre-training actually calls ChamferDistanceL1 on
    (/root/reference tools/runner_pretask.py:220-223 -- 32 predicted vs 1024, 1024 vs 1024, 2048 vs 8192, B = 64),
    forward and backward, each as its own CUDA graph; the reference's own chamfer.cu (oracle/_ref) beside it."""
    import upp_b200
    from oracle import ref_gpu
    o, ref = upp_b200.ops, ref_gpu.load()
    g = torch.Generator().manual_seed(5)
    out = []

    def graph_ms(fn, reps=20):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        ts = []
        for it in range(8):
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(reps):
                gr.replay()
            e0.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(s0.elapsed_time(e0) / reps)
        return round(statistics.median(ts), 5)

    def eager_ms(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(torch.cuda.default_stream())
            fn()
            e0.record(torch.cuda.default_stream())
            torch.cuda.synchronize()
            ts.append(s0.elapsed_time(e0))
        return round(statistics.median(ts), 5)

    for n, m in ((32, 1024), (1024, 1024), (2048, 8192)):
        a, b = torch.rand(B, n, 3, generator=g).to(dev), torch.rand(B, m, 3, generator=g).to(dev)
        d1, d2, i1, i2 = o.chamfer_forward(a, b)
        g1, g2 = torch.rand_like(d1), torch.rand_like(d2)
        fwd = graph_ms(lambda: o.chamfer_forward(a, b, True))
        bwd = graph_ms(lambda: o.chamfer_backward(a, b, i1, i2, g1, g2))
        row = {"shape": f"B{B} {n} vs {m}", "fwd_ms": fwd, "bwd_ms": bwd,
               "fwd_tflops_8NM": round(8.0 * n * m * B / (fwd * 1e-3) / 1e12, 3)}
        if ref is not None:
            r = ref.forward(a, b)
            row["reference_fwd_ms"] = eager_ms(lambda: ref.forward(a, b))
            row["reference_bwd_ms"] = eager_ms(lambda: ref.backward(a, b, r[2], r[3], g1, g2))
        out.append(row)
    return out


def as_shipped_leg():
    """The reference ships with CUDA_LAUNCH_BLOCKING=1 (main.py:5): every launch waits for the kernel.  One more line for
    that setting (SURVEY.md 8d "as shipped"): the headline step through the unchanged call sites, EAGER, the variable set
    before CUDA starts -- so it runs in a child process.  Returns the child's record or why it could not be had."""
    import subprocess
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--as-shipped-child"], env=env, capture_output=True,
                           text=True, timeout=240)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": (r.stderr or r.stdout)[-200:]}
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": str(ex)[:200]}


def as_shipped_child():
    assert os.environ.get("CUDA_LAUNCH_BLOCKING") == "1"
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    B = DEFAULT_B[HEADLINE]
    W = GpuWorkload(HEADLINE, make_inputs(HEADLINE, B, seed=0), dev, 1, None, "none (1 GPU)")
    _, keys = W.h2d_bytes()
    W.h2d = {k: W.host[k] for k in keys}
    api = W.callsites()
    dd = dict(W.d)
    for _ in range(5):
        W.run_modules(dd, api).item()
    torch.cuda.synchronize()
    ts = []
    for _ in range(50):
        t0 = time.perf_counter()
        W.run_modules(dd, api).item()
        ts.append((time.perf_counter() - t0) * 1e3)
    ms = statistics.median(ts)
    print(json.dumps({"value": B / (ms * 1e-3), "unit": UNIT, "ms_per_step": round(ms, 5),
                      "what": "headline step, reference call sites unchanged over the drop-in modules + autograd, eager launches "
                              "with CUDA_LAUNCH_BLOCKING=1 as the reference's main.py:5 sets it; H2D of the inputs and the "
                              "loss read-back inside the step; host wall clock, median of 50 steps"}), flush=True)


def measure(ctx, name, B, peers, collective, peaks, peak_src, want_cpu, scaling="weak", sampler=None):
    """One workload on this rank's GPU: device-resident arm, e2e arm(s), per-op roofline pass.  Collective calls inside:
    every rank must call it with the same arguments."""
    import upp_b200
    args, dev, world, rank = ctx.args, ctx.dev, ctx.world, ctx.rank
    W = GpuWorkload(name, make_inputs(name, B, seed=rank), dev, world, peers, collective)
    rec = {"config": config_of(name, B), "scaling": scaling}
    run_info = {"collective": W.collective}

    # ---- device-resident arm: CUDA-graph replay of the ops-level step.  The step's critical path (the serial FPS
    #      chain) is captured on a high-priority stream, the independent branches on default-priority side streams ----
    for _ in range(2):
        W.run_ops(W.d)
    torch.cuda.synchronize()
    replay, static_loss, mode, launches_per_step, err = capture(ctx, lambda: W.run_ops(W.d))
    run_info["launch_mode"] = mode
    if err:
        run_info["graph_error"] = err
    one_step = replay if replay is not None else (lambda: W.run_ops(W.d))
    for _ in range(args.warmup):
        one_step()
    if sampler is not None:
        sampler.mark()
    dev_ms, dev_regions = ctx.timed(one_step)
    dev_pct = step_percentiles(ctx)
    clocks = sampler.summary() if sampler is not None else None

    # ---- end-to-end arms: pinned host -> device every step (inside the captured step, on the branches that consume
    #      the inputs), loss read back every step ----
    h2d_bytes, h2d_keys = W.h2d_bytes()
    dd = dict(W.d)
    offs, total = {}, 0
    for k in h2d_keys:  # the step's inputs live in ONE pinned host arena and one device arena (16-byte aligned views)
        offs[k] = total
        total += (W.host[k].numel() + 3) // 4 * 4
    host_arena = torch.empty(total, dtype=torch.float32).pin_memory()
    dev_arena = torch.empty(total, dtype=torch.float32, device=dev)
    for k in h2d_keys:
        n = W.host[k].numel()
        host_arena[offs[k]:offs[k] + n].copy_(W.host[k].reshape(-1))
        dd[k] = dev_arena[offs[k]:offs[k] + n].view(W.host[k].shape)
    dev_arena.copy_(host_arena)
    W.h2d = {k: host_arena[offs[k]:offs[k] + W.host[k].numel()].view(W.host[k].shape) for k in h2d_keys}

    def e2e_arm(api, label):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                W.run_modules(dd, api)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        host_loss = torch.empty(1, dtype=torch.float32).pin_memory()

        def graph_fn():  # the step AND the read-back of its result: one pinned 4-byte D2H copy, a node of the same graph
            out = W.run_modules(dd, api)
            host_loss.copy_(out.reshape(1), non_blocking=True)
            return out
        rp, loss, emode, nlaunch, eerr = capture(ctx, graph_fn)

        def step():
            if rp is not None:
                rp()
                torch.cuda.current_stream().synchronize()   # the host needs the value: it waits for the copy
                return float(host_loss[0])                   # D2H of the step's result
            return W.run_modules(dd, api).item()
        for _ in range(args.warmup):
            step()
        ms, _ = ctx.timed(step)
        out = {"value": B * world * args.steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
               "d2h_bytes_per_step": 4, "ms_per_step": ms / args.steps, "step_ms": step_percentiles(ctx),
               "gpu_launches_per_step": int(nlaunch), "api": f"{label}, {emode}"}
        if eerr:
            out["graph_error"] = eerr
        return out

    rec["e2e"] = e2e_arm(None, "upp_b200 modules (fused Group / fps / ChamferDistanceL1" +
                         (" via parallel.sharded_chamfer" if world > 1 and name in (HEADLINE, "c3") else "") + ") + autograd")
    if name != "c5":
        rec["e2e_dropin"] = e2e_arm(W.callsites(), "reference call sites unchanged (reference_callsites.py: misc.fps -> "
                                    "pointnet2_utils.furthest_point_sample + gather_operation + 2 transposes, knn_cuda.KNN, "
                                    "index gather, chamfer.forward/backward) over the drop-in modules + autograd")
    W.h2d = None

    # ---- per-kernel pass (roofline): every labelled op of the step alone, as its own CUDA graph (no Python or
    #      launch gaps), CUDA events around REPS back-to-back replays.  Inputs are the step's own tensors: L2-warm,
    #      as they are inside the step (producer -> consumer). ----
    W.collect = []
    W.run_ops(W.d)
    torch.cuda.synchronize()
    ops_seen, W.collect = W.collect, None
    op_ms, REPS = {}, 20
    for label, fn, a, k in ops_seen:
        if label in op_ms:
            continue
        try:
            for _ in range(2):
                fn(*a, **k)
            torch.cuda.synchronize()
            g1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g1):
                fn(*a, **k)
            run = g1.replay
        except Exception:
            torch.cuda.synchronize()
            run = lambda fn=fn, a=a, k=k: fn(*a, **k)  # noqa: E731
        ts = []
        for it in range(3 + 7):
            s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(REPS):
                run()
            e0.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(s0.elapsed_time(e0) / REPS)
        op_ms[label] = statistics.median(ts)
    if world > 1 and peers is not None:  # the per-op pass replays sharded sends without their finish: realign the counters
        ctx.barrier()

    gpu_ref = time_gpu_reference_chamfer(W) if (name == "c3" and rank == 0) else None
    train_shapes = time_chamfer_training_shapes(dev, B) if (name == "c3" and rank == 0 and world == 1) else None
    if rank != 0:
        return None

    sm_max = peaks.get("sm_max_mhz") or 1965.0
    fp32_peak = N_SM * FP32_LANES * 2 * sm_max * 1e6 / 1e12  # TFLOP/s
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    kernels = []
    for label, ms in sorted(op_ms.items(), key=lambda kv: -kv[1]):
        fl, by = op_work(label)
        kernels.append({"op": label, "ms": round(ms, 5), "share": round(ms / sum(op_ms.values()), 4),
                        "tflops": round(fl * B / (ms * 1e-3) / 1e12, 4), "gbs": round(by * B / (ms * 1e-3) / 1e9, 2)})
    top = kernels[0]
    fl, by = op_work(top["op"])
    fp32_bound = fl > 0
    roofline = {"kernel": top["op"], "bound": "fp32" if fp32_bound else "hbm",
                "achieved": top["tflops"] if fp32_bound else top["gbs"],
                "peak": round(fp32_peak, 2) if fp32_bound else hbm_peak,
                "unit": "TFLOP/s" if fp32_bound else "GB/s",
                "frac": round((top["tflops"] / fp32_peak) if fp32_bound else (top["gbs"] / hbm_peak), 5),
                "traffic": traffic_for(top["op"]),
                "peak_source": (f"computed 148 SM x 128 FP32 lanes x 2 x {sm_max:.0f} MHz (the measured peaks carry no fp32-pipe figure; "
                                "K=3 distances are not a tensor-core contraction)") if fp32_bound else peak_src,
                "hbm_gbs": top["gbs"], "hbm_frac": round(top["gbs"] / hbm_peak, 5)}
    if top["op"].startswith("chamfer_fwd"):
        # a pair costs 3 FADD + FMUL + 2 FFMA = 6 FMA-pipe slots for 8 flop: 66.7 % of peak is the op-mix ceiling
        roofline["frac_of_opmix_ceiling"] = round(roofline["frac"] / (8.0 / 12.0), 5)
        roofline["note"] = "8NM flop (each pair once); the reference's two directed passes execute 16NM"
    if top["op"].startswith(("fps", "group")):
        roofline["note"] = "FPS is a serial chain of M-1 block arg-max rounds: latency-bound, see DESIGN.md"
        v = {x[0]: int(x[1:]) for x in top["op"].split()[1:] if x[1:].isdigit()}
        rounds = max(v.get("M", v.get("G", 2)) - 1, 1)
        lm = {"rounds": rounds, "us_per_round": round(top["ms"] * 1e3 / rounds, 4), "busy_sms": min(B, N_SM)}
        if top["op"].startswith("fps") and v["N"] <= 2048:
            # the bound that actually binds small-cloud FPS: (M-1) dependent rounds, each a chain of fixed instruction
            # latencies (LDS centre 29 + sub/mul/fma/fma 16 + FMNMX 4 + 3 x VIMNMX3 12 + REDUX 23 + vote/compare 12 +
            # STS/BAR 15 + LDS slots 29 + select 12 + address 4 = 156 cycles, + ~45 cycles of packed-FMA issue)
            floor_us = 201.0 / sm_max
            lm["chain_floor_us_per_round"] = round(floor_us, 4)
            lm["frac_of_chain_floor"] = round(floor_us / (top["ms"] * 1e3 / rounds), 3)
        roofline["latency_model"] = lm

    clouds = B * world * args.steps
    rec.update({"value": clouds / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / args.steps,
                "timed_regions": len(dev_regions), "region_ms": [round(x, 4) for x in dev_regions[:8]], "step_ms": dev_pct,
                "gpu_launches_per_step": int(launches_per_step), "roofline": roofline, "kernels": kernels,
                "run_info": run_info, "clocks": clocks})
    if gpu_ref is not None:
        rec["gpu_reference"] = gpu_ref
    if train_shapes is not None:
        rec["training_shapes"] = train_shapes
    if want_cpu:
        rec["cpu_baseline"] = cpu_baseline_record(name, B)
    return rec


# ----------------------------------------------------------------------------- main ----------

def reference_arm(args, scaling, B, world):
    """--impl reference: the reference's CPU formulation of the path (kind "port": oracle/torch_formulation.py) on all
    host cores, same config / metric / unit as the B200 arm; sub-records for the other BASELINE configs."""
    cb, sec, bs = time_cpu(args.workload, B, 0, args.steps, args.warmup, budget_s=100.0)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(args.workload, B), "run_info": {"clouds_per_step": bs}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.workload == HEADLINE and not args.no_configs:
        five, _, _ = time_cpu(args.workload, B, 0, 3, 1, 10.0, threads=5)
        line["cpu_baseline"]["omp5"] = {"value": five["value"], "cores": five["cores"], "sample": five["sample"]}
        line["configs"] = {}
        for name in ("c1", "c3", "c4", "c5"):
            Bc = DEFAULT_B[name]
            sub = cpu_baseline_record(name, Bc, steps=5, warmup=1, budget_s=10.0)
            line["configs"][name] = {"config": config_of(name, Bc), "value": sub["value"], "unit": UNIT, "cpu_baseline": sub}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(DEFAULT_B))
    ap.add_argument("--batch", type=int, default=0, help="clouds per GPU (default: the config's B)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: the config's batch is the GLOBAL batch, split evenly over the ranks "
                         "(SURVEY.md 8d C4); default is weak scaling (the config's batch per GPU)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1/c3/c4/c5 (and strong-scaling) sub-records")
    ap.add_argument("--as-shipped-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.as_shipped_child:
        as_shipped_child()
        return
    args.warmup = max(args.warmup, 3)
    args.steps = max(args.steps, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = args.batch or DEFAULT_B[args.workload]
    if args.strong:
        if B % world:
            raise SystemExit(f"--strong: global batch {B} is not divisible by {world} ranks")
        B //= world
    scaling = "strong" if args.strong else "weak"

    if args.impl == "reference":
        if rank == 0:
            reference_arm(args, scaling, B, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    ctx = Ctx(args, rank, local_rank, world)
    dist = ctx.dist
    if world > 1:
        dist.init_process_group("nccl", device_id=ctx.dev)
    import upp_b200

    peers, collective = None, "none (1 GPU)"
    if world > 1:
        try:
            peers = upp_b200.parallel.PeerExchange()
            collective = (f"fused in the Chamfer kernels over NVLink peer memory ({peers.how}); "
                          "device arm: deferred wait late on the Chamfer side stream")
        except RuntimeError as ex:
            collective = f"NCCL all_reduce of 4 floats (peer mapping unavailable: {str(ex)[:80]})"
    peaks, peak_src = load_peaks()

    # nvidia-smi is started BEFORE the warm-up and left to settle: launching it right in front of the timed region
    # perturbs the rank that owns it (measured at 2 GPUs: rank 0 lagged ~14 us per step and every peer waited for it)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        time.sleep(0.5)
    want_cpu = world == 1 and not args.no_cpu_baseline
    top = measure(ctx, args.workload, B, peers, collective, peaks, peak_src, want_cpu, scaling, sampler)

    subs, strong = {}, None
    if args.workload == HEADLINE and not args.no_configs and not args.strong and not args.batch:
        for name in ("c1", "c3", "c4", "c5"):
            torch.cuda.empty_cache()
            subs[name] = measure(ctx, name, DEFAULT_B[name], peers, collective, peaks, peak_src, want_cpu, "weak", sampler)
        if world > 1 and DEFAULT_B["c4"] % world == 0:
            torch.cuda.empty_cache()
            strong = measure(ctx, "c4", DEFAULT_B["c4"] // world, peers, collective, peaks, peak_src, False, "strong", sampler)
    if sampler is not None:
        sampler.stop()
    if rank != 0:
        _finish(world, dist)
        return

    line = {"metric": METRIC, "value": top["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": top["ms_per_step"], "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": top["config"], "run_info": top["run_info"],
            "clocks": top["clocks"], "timed_regions": top["timed_regions"], "region_ms": top["region_ms"],
            "step_ms": top["step_ms"],
            "e2e": top["e2e"], "gpu_launches": int(top["gpu_launches_per_step"] * args.steps * top["timed_regions"]),
            "gpu_launches_per_step": top["gpu_launches_per_step"], "roofline": top["roofline"], "kernels": top["kernels"]}
    for k in ("e2e_dropin", "cpu_baseline", "gpu_reference"):
        if k in top:
            line[k] = top[k]
    if world == 1 and args.workload == HEADLINE and not args.no_configs and not args.batch:
        line["e2e_dropin_launch_blocking"] = as_shipped_leg()
    if subs:
        line["configs"] = {k: v for k, v in subs.items() if v is not None}
    if strong is not None:
        strong["note"] = (f"BASELINE.json configs[3] as written: the GLOBAL batch of {DEFAULT_B['c4']} clouds split over {world} ranks "
                          f"({DEFAULT_B['c4'] // world} per GPU); compare `value` with configs.c4 of the 1-GPU run")
        line["strong"] = strong
    print(json.dumps(line), flush=True)
    _finish(world, dist)


def _finish(world, dist):
    """N > 1: leave without tearing the NCCL communicator down.  The CUDA graphs replayed above hold captured
    NCCL kernels, and destroy_process_group() behind them was seen to block until torchrun's timeout; every
    result has been printed and flushed by now, so the ranks rendezvous once and exit."""
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
