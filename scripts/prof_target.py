"""Runs one op a few times so that `ncu -k regex:<kernel>` can capture it (see profiles/README.md)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
from upp_b200 import ops  # noqa: E402

what = sys.argv[1]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for _ in range(4):
    if what == "chamfer":
        a, b = torch.rand(64, 2048, 3, generator=g).to(dev), torch.rand(64, 2048, 3, generator=g).to(dev)
        out = ops.chamfer_forward(a, b)
        ops.chamfer_backward(a, b, out[2], out[3], torch.rand_like(out[0]), torch.rand_like(out[1]))
    elif what == "chamfer_bwd_small":  # Completion-Prompter training shape 32 predicted vs 1024 (runner_pretask.py:220)
        a, b = torch.rand(64, 32, 3, generator=g).to(dev), torch.rand(64, 1024, 3, generator=g).to(dev)
        out = ops.chamfer_forward(a, b)
        ops.chamfer_backward(a, b, out[2], out[3], torch.rand_like(out[0]), torch.rand_like(out[1]))
    elif what == "group":  # C1: single-launch Group + its backward
        x = (torch.rand(32, 1024, 3, generator=g) * 2 - 1).to(dev)
        nb, ce, idx, cidx = ops.group(x, 64, 32)
        ops.group_backward(torch.randn_like(nb), torch.randn_like(ce), idx, cidx, 1024)
    elif what == "fps":
        ops.fps((torch.rand(32, 1228, 3, generator=g) * 2 - 1).to(dev), 1024)
    elif what == "fps8k":
        ops.fps((torch.rand(128, 8192, 3, generator=g) * 2 - 1).to(dev), 1024)
    elif what == "fps_cluster":  # seprate_point_cloud's large side: B x 6144 -> 1024 on clusters of 4 CTAs
        ops.fps((torch.rand(32, 6144, 3, generator=g) * 2 - 1).to(dev), 1024)
    elif what == "fps_cluster2":  # the 2-CTA cluster shape the dispatcher never picks (A/B: why it loses)
        os.environ["UPP_TUNING"], os.environ["UPP_FPS_CLUSTER"] = "1", "2"
        ops.fps((torch.rand(32, 6144, 3, generator=g) * 2 - 1).to(dev), 1024)
    elif what == "fps_cluster8":  # C4 sharded over 8 GPUs: 16 clouds of 8192 points per GPU
        ops.fps((torch.rand(16, 8192, 3, generator=g) * 2 - 1).to(dev), 1024)
    elif what in ("fps_rows", "fps_rows_small"):  # register-resident rows + exact pruning (fps_pruned.cu)
        os.environ["UPP_TUNING"], os.environ["UPP_FPS_PRUNED"], os.environ["UPP_FPS_PRUNED_MIN"] = "1", "2", "63"
        B, N = (128, 8192) if what == "fps_rows" else (32, 1228)
        x = torch.randn(B, N, 3, generator=g) * 0.35
        x = x - x.mean(1, keepdim=True)
        ops.fps((x / x.norm(dim=2).max(dim=1)[0].view(-1, 1, 1)).to(dev), 1024)
    elif what == "crop":
        x = (torch.rand(32, 8192, 3, generator=g) * 2 - 1).to(dev)
        c = torch.nn.functional.normalize(torch.randn(32, 3, generator=g), dim=-1).to(dev)
        ops.crop_split(x, c, 2048)
    elif what == "scatter":  # the two remaining gather-gradient shapes of the default step + the channel-first gather grad
        rows = torch.randn(32, 1024, 3, generator=g).to(dev)
        idx = torch.stack([torch.randperm(1228, generator=g)[:1024] for _ in range(32)]).to(torch.int32).to(dev)
        ops.rows_scatter_add(rows, idx, 1228)
        ops.gather_grad(rows.transpose(1, 2).contiguous(), idx, 1228)
    elif what == "knn":
        r = (torch.rand(32, 1024, 3, generator=g) * 2 - 1).to(dev)
        ops.knn(r, r[:, :64].contiguous(), 32)
    elif what == "interp":
        x1 = (torch.rand(32, 2048, 3, generator=g) * 2 - 1).to(dev)
        x2 = (torch.rand(32, 128, 3, generator=g) * 2 - 1).to(dev)
        p2, go = torch.randn(32, 128, 1152, generator=g).to(dev), torch.randn(32, 2048, 1152, generator=g).to(dev)
        out, idx, w, d = ops.interp_forward(x1, x2, p2, 3, 1e-4)
        ops.interp_backward(go, idx, w, 128, xyz_terms=(d, p2, x1, x2, 1e-4))
    torch.cuda.synchronize()
