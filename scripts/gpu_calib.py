"""Calibration: write-only and copy HBM bandwidth of this GPU with plain torch kernels (context for the roofline)."""
import torch, json
n = 1 << 30   # 8 GiB of float64
a = torch.empty(n, dtype=torch.float64, device="cuda")
b = torch.empty(n, dtype=torch.float64, device="cuda")
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
tf = t(lambda: a.fill_(1.0))
tc = t(lambda: b.copy_(a))
print(json.dumps({"fill_GBps": n * 8 / tf / 1e6, "copy_GBps": 2 * n * 8 / tc / 1e6, "fill_ms": tf, "copy_ms": tc}))
