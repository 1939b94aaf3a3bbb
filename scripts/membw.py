"""Reference points for the HBM-bound kernels: device memset (write only), copy (read + write) and read-only sum at the
seg-propagation size (B32 x N2048 x C1152 fp32 = 302 MB), CUDA-event timed, L2 flushed between runs."""
import json
import torch

dev = torch.device("cuda:0")
n = 32 * 2048 * 1152
x = torch.empty(n, dtype=torch.float32, device=dev)
y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
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


for name, fn, byt in (("memset 302MB (write)", lambda: x.zero_(), 4 * n), ("copy 302MB (read+write)", lambda: y.copy_(x), 8 * n),
                      ("sum 302MB (read)", lambda: x.sum(), 4 * n)):
    us = timeit(fn)
    print(json.dumps({"op": name, "us": round(us, 2), "gbs": round(byt / us / 1e3, 1)}))
