"""Debug aid: one eager launch of the rows FPS kernel built with -DUPP_ROWS_PROFILE (prints per-phase cycles per round)."""
import os
import sys
os.environ["UPP_TUNING"] = "1"
os.environ["UPP_FPS_PRUNED"] = "2"
os.environ["UPP_FPS_PRUNED_MIN"] = "63"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "iccv2025-upp_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402
import upp_b200  # noqa: E402

g = torch.Generator().manual_seed(0)
for a in sys.argv[1:]:
    B, N, M = (int(v) for v in a.split("x"))
    x = torch.randn(B, N, 3, generator=g) * 0.35
    x = x - x.mean(1, keepdim=True)
    x = (x / x.norm(dim=2).max(dim=1)[0].view(-1, 1, 1)).cuda()
    print(f"== B{B} N{N} M{M}", flush=True)
    upp_b200.ops.fps(x, M, True)
    torch.cuda.synchronize()
