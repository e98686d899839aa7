"""Raw pinned-memory D2H / H2D rate of the box (what bounds bench.py's e2e arm: 16.4 MB of rows per step)."""
import json
import torch

out = {}
for mb in (16, 64, 256):
    n = mb * (1 << 20) // 8
    dev = torch.empty(n, dtype=torch.float64, device="cuda")
    host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    for name, (dst, src) in (("d2h", (host, dev)), ("h2d", (dev, host))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        out["%s_%dMB_GBps" % (name, mb)] = round(20 * mb * (1 << 20) / (a.elapsed_time(b) * 1e-3) / 1e9, 2)
print(json.dumps(out))
