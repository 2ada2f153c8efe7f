"""Phase timings of a batch plan (n slices of 2048^2 CGPS-like fields), device-resident (diagnostic)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
ws = load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
S = 2048
imgs = np.stack([fieldgen.cgps_like(S, S, seed=i) for i in range(n)])
ctx = ws.default_context()
plan = ws.Plan(ctx, n, S, S)
d_img = torch.from_numpy(imgs).cuda()
off = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
ns = plan.find_local_minima(d_img.data_ptr(), 0, 0, off.data_ptr())
seeds = torch.empty((max(ns, 1), 2), dtype=torch.int32, device="cuda")
plan.find_local_minima(d_img.data_ptr(), seeds.data_ptr(), ns, off.data_ptr())
for kind in (0, 1, 1):
    plan.run(kind, 254, d_img.data_ptr(), seeds.data_ptr(), off.data_ptr(), ns)
    ph = plan.phase_ms()
    print("kind", kind, "slices", n, "seeds", ns, {k: round(v, 3) for k, v in ph.items()}, "total", round(sum(ph.values()), 3), plan.stats())
