"""Per-tile activation statistics of the flood on a field (diagnostic)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
ws = load()
S = int(sys.argv[1]); kind = sys.argv[2]
import os
sig = float(sys.argv[3]) if len(sys.argv) > 3 else 16.0
cache = f"/tmp/field_{kind}_{S}_{sig}.npy"
if os.path.exists(cache):
    img = np.load(cache)
else:
    img = fieldgen.uniform(S, S, 0) if kind == "uniform" else fieldgen.smooth(S, S, sig, 0)
    np.save(cache, img)
ctx = ws.default_context()
plan = ws.Plan(ctx, 1, S, S)
d_img = torch.from_numpy(img).cuda()
off = torch.zeros(2, dtype=torch.int32, device="cuda")
n = plan.find_local_minima(d_img.data_ptr(), 0, 0, off.data_ptr())
seeds = torch.empty((max(n, 1), 2), dtype=torch.int32, device="cuda")
plan.find_local_minima(d_img.data_ptr(), seeds.data_ptr(), n, off.data_ptr())
for _ in range(3):
    plan.run(0, 254, d_img.data_ptr(), seeds.data_ptr(), off.data_ptr(), n)
st = plan.stats(); ph = plan.phase_ms()
tiles = ((S + 63) // 64) * ((S + 31) // 32)
print(kind, S, "seeds", n, st, ph)
print("activations/tile %.2f  phases/activation %.2f  us per activation (444 CTAs) %.2f" % (
    st["tile_activations"] / tiles, st["flood_phases"] / st["tile_activations"], ph["flood"] * 1e3 * 444 / st["tile_activations"]))
print("consumers: busy %.0f Mcyc, waiting %.0f Mcyc (%.0f %% waiting); busy cycles per activation %.0f, per phase %.0f" % (
    st["flood_busy_kcycles"] / 1e3, st["flood_wait_kcycles"] / 1e3,
    100.0 * st["flood_wait_kcycles"] / max(1, st["flood_wait_kcycles"] + st["flood_busy_kcycles"]),
    st["flood_busy_kcycles"] * 1024.0 / st["tile_activations"], st["flood_busy_kcycles"] * 1024.0 / st["flood_phases"]))
if os.environ.get("WS_STATS_LEVELS"):
    T = ctx.d2h(plan.arrival_times_ptr, (S, S), np.uint32)
    lv = T >> 24
    fin = lv < 255
    print("levels used: distinct", np.unique(lv[fin]).size, "max hop", int((T[fin] & 0xFFFFFF).max()))
