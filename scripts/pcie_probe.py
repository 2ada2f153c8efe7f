"""Host<->device copy bandwidth of this box with pinned memory (context for the e2e numbers)."""
import json
import torch

out = {}
for mb in (64, 512, 2048):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[f"{name}_{mb}MiB_GBps"] = round(3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
print(json.dumps(out))
