"""One small launch of every kernel family (and of the variants the dispatcher only picks for large problems, forced
through UPP_TUNING) for compute-sanitizer: memcheck / racecheck / synccheck / initcheck.  Results are checked against the
oracle so that a sanitizer-clean run is also a correct one.
    compute-sanitizer --tool racecheck python scripts/sanitize_target.py"""
import os
import sys

os.environ["UPP_TUNING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import upp_b200  # noqa: E402
from oracle import c_oracle as O  # noqa: E402

o = upp_b200.ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
done = []


def env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def check(name, ok):
    assert ok, name
    done.append(name)


# FPS: one CTA per cloud (small / deferred-search / v1), clusters of 2 / 4 / 8 CTAs (st.async + mbarrier exchange)
x = (torch.rand(3, 700, 3, generator=g) * 2 - 1)
want = O.fps(x.numpy(), 40)
for tag, kw in (("blk", {}), ("cluster2", dict(UPP_FPS_CLUSTER=2)), ("cluster2_nw8", dict(UPP_FPS_CLUSTER=2, UPP_FPS_CLUSTER_NW=8)),
                ("v1", dict(UPP_FPS_IMPL=1, UPP_FPS_W4=0)), ("nw8", dict(UPP_FPS_NW=8, UPP_FPS_P2=2, UPP_FPS_S2=1))):
    env(**kw)
    check("fps " + tag, np.array_equal(o.fps(x.to(dev), 40).cpu().numpy(), want))
    env(**{k: None for k in kw})
xb = (torch.rand(2, 2100, 3, generator=g) * 2 - 1)
wantb = O.fps(xb.numpy(), 24)
for tag, kw in (("cluster4", dict(UPP_FPS_CLUSTER=4)), ("cluster8", dict(UPP_FPS_CLUSTER=8)),
                ("tree", dict(UPP_FPS_NW=8, UPP_FPS_P2=5, UPP_FPS_S2=0, UPP_FPS_SEARCH=2))):
    env(**kw)
    check("fps " + tag, np.array_equal(o.fps(xb.to(dev), 24).cpu().numpy(), wantb))
    env(**{k: None for k in kw})
# the measured experiments (off by default): exact pruning on a Z-order-sorted copy (shared-memory buckets / register rows),
# clusters with one look-ahead selection per exchange round (M > 48: the look-ahead is not tried before)
wantc = O.fps(xb.numpy(), 80)
for tag, kw in (("pruned_buckets", dict(UPP_FPS_PRUNED=1, UPP_FPS_PRUNED_MIN=255)), ("pruned_rows", dict(UPP_FPS_PRUNED=2, UPP_FPS_PRUNED_MIN=63)),
                ("cluster8_ahead", dict(UPP_FPS_CLUSTER=8, UPP_FPS_CLUSTER_AHEAD=1))):
    env(**kw)
    check("fps " + tag, np.array_equal(o.fps(xb.to(dev), 80).cpu().numpy(), wantc))
    env(**{k: None for k in kw})
xc = (torch.rand(2, 4100, 3, generator=g) * 2 - 1)
env(UPP_FPS_CLUSTER=4, UPP_FPS_CLUSTER_AHEAD=1)
check("fps cluster4_ahead", np.array_equal(o.fps(xc.to(dev), 70).cpu().numpy(), O.fps(xc.numpy(), 70)))
env(UPP_FPS_CLUSTER=None, UPP_FPS_CLUSTER_AHEAD=None)
# Group: single launch in every cluster shape (producer / consumer warps, DSMEM centre table), two launches
xg = (torch.rand(3, 600, 3, generator=g) * 2 - 1)
wg = O.group(xg.numpy(), 20, 16)
for tag, kw in (("default", {}), ("cs1", dict(UPP_GROUP_CLUSTER=1)), ("cs2", dict(UPP_GROUP_CLUSTER=2)), ("cs4_w8", dict(UPP_GROUP_CLUSTER=4, UPP_GROUP_WARPS=8)),
                ("cs8_w8", dict(UPP_GROUP_CLUSTER=8, UPP_GROUP_WARPS=8)), ("two_launch", dict(UPP_GROUP_FUSED=0))):
    env(**kw)
    got = o.group(xg.to(dev), 20, 16)
    check("group " + tag, all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(got, wg)))
    env(**{k: None for k in kw})
nb, ce, idx, cidx = o.group(xg.to(dev), 20, 16)
gx = o.group_backward(torch.randn_like(nb), torch.randn_like(ce), idx, cidx, 600)
check("group_bwd", bool(torch.isfinite(gx).all()))
# Chamfer forward (slots / folded / keyed / directed) + backward (+ statistics)
a, b = torch.rand(3, 500, 3, generator=g), torch.rand(3, 700, 3, generator=g)
wc = O.chamfer_fwd(a.numpy(), b.numpy())
for tag, kw in (("slots", {}), ("slots_3chunks", dict(UPP_CH_CHUNKS=3)), ("folded", dict(UPP_CH_FUSED=1, UPP_CH_CHUNKS=2)), ("keyed", dict(UPP_CH_VARIANT=30)),
                ("scalar", dict(UPP_CH_VARIANT=20)), ("directed", dict(UPP_CH_VARIANT=0))):
    env(**kw)
    got = o.chamfer_forward(a.to(dev), b.to(dev), want_sums=True)
    check("chamfer " + tag, all(np.array_equal(x_.cpu().numpy(), y_) for x_, y_ in zip(got[:4], wc)))
    env(**{k: None for k in kw})
d1, d2, i1, i2 = o.chamfer_forward(a.to(dev), b.to(dev))
gg = o.chamfer_backward(a.to(dev), b.to(dev), i1, i2, torch.rand_like(d1), torch.rand_like(d2), want_sqnorm=True)
check("chamfer_bwd", bool(torch.isfinite(gg[0]).all()) and float(gg[2][0]) > 0)
# kNN, gather, rows_scatter_add, crop, interpolation (one-launch / two-phase / streamed backward), knn_points
D, I = o.knn(xg.to(dev), xg[:, :17].contiguous().to(dev), 9)
check("knn", np.array_equal(I.cpu().numpy(), O.knn(xg.numpy(), xg[:, :17].contiguous().numpy(), 9)[1]))
rows = torch.randn(3, 50, 3, generator=g)
ii = torch.randint(0, 70, (3, 50), generator=g, dtype=torch.int32)
check("rows_scatter_add", bool(torch.isfinite(o.rows_scatter_add(rows.to(dev), ii.to(dev), 70)).all()))
check("gather_grad", bool(torch.isfinite(o.gather_grad(rows.transpose(1, 2).contiguous().to(dev), ii.to(dev), 70)).all()))
xc = (torch.rand(3, 1000, 3, generator=g) * 2 - 1)
cc = torch.nn.functional.normalize(torch.randn(3, 3, generator=g), dim=-1)
inp, crp, order = o.crop_split(xc.to(dev), cc.to(dev), 300, want_order=True)
check("crop_split", np.array_equal(order.cpu().numpy(), O.crop_order(xc.numpy(), cc.numpy())))
x1, x2 = torch.rand(2, 200, 3, generator=g), torch.rand(2, 48, 3, generator=g)
p2, go = torch.randn(2, 48, 256, generator=g), torch.randn(2, 200, 256, generator=g)
for path in ("0", "1"):
    env(UPP_INTERP_PATH=path)
    out, j, w, d = o.interp_forward(x1.to(dev), x2.to(dev), p2.to(dev), 3, 1e-4)
    gp2 = o.interp_backward(go.to(dev), j, w, 48, xyz_terms=(d, p2.to(dev), x1.to(dev), x2.to(dev), 1e-4))
    check("interp path " + path, np.array_equal(j.cpu().numpy(), O.interp_fwd(x1.numpy(), x2.numpy(), p2.numpy(), 3, 1e-4)[1]))
env(UPP_INTERP_PATH=None)
sel = upp_b200.select_neighbors(x1.to(dev), x2.to(dev), 3, 1e-4)
check("interp blend", torch.equal(o.interp_blend(p2.to(dev), sel.idx, sel.weight), o.interp_forward(x1.to(dev), x2.to(dev), p2.to(dev), 3, 1e-4)[0]))
kp = upp_b200.knn_points(x1.to(dev), x2.to(dev), K=4, return_nn=True)
check("knn_points", np.array_equal(kp.idx.cpu().numpy(), O.knn_points(x1.numpy(), x2.numpy(), 4)[1]))
torch.cuda.synchronize()
print(f"sanitize target ok: {len(done)} checks: " + ", ".join(done))
