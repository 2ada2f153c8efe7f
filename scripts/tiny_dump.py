import sys, numpy as np
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
from oracle import oracle as orc
ws = load()
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 64)
img = fieldgen.uniform(H, W, 1)
seeds = orc.find_local_minima(img)
ctx = ws.default_context()
plan = ws.Plan(ctx, 1, H, W)
d_img = ctx.dev_malloc(img.size); s32 = np.ascontiguousarray(seeds, dtype=np.uint32)
d_seeds = ctx.dev_malloc(max(1, s32.nbytes)); d_off = ctx.dev_malloc(8)
ctx.h2d(d_img, img); ctx.h2d(d_seeds, s32); ctx.h2d(d_off, np.array([0, len(s32)], np.uint32))
try:
    plan.run(0, 254, d_img, d_seeds, d_off, len(s32))
    print("run ok", plan.stats())
except Exception as e:
    print("run failed:", e)
T = ctx.d2h(plan.arrival_times_ptr, img.shape, np.uint32)
ref = orc.transform(orc.SEGMENTING, img, seeds)
exp = (ref.lvl.astype(np.uint32) << 24) | ref.hop
exp[ref.lvl == 255] = 0xFF000000
T = np.minimum(T, 0xFF000000)
bad = T != exp
print("bad pixels", int(bad.sum()), "of", T.size)
if bad.any():
    ys, xs = np.nonzero(bad)
    print("rows", np.unique(ys)[:40], "cols", np.unique(xs)[:40])
    for y, x in list(zip(ys, xs))[:12]:
        print(y, x, hex(T[y, x]), hex(exp[y, x]), "img", img[y, x])
    print("bad lower than exp:", int((T[bad] < exp[bad]).sum()), " higher:", int((T[bad] > exp[bad]).sum()))
