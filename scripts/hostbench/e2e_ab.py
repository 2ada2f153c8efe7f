"""A/B timing of the host pipeline on one box: WS_HOST_NO_AVX2=1 python scripts/hostbench/e2e_ab.py  vs  without."""
import sys, time, os, numpy as np
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
ws = load()
S = 16384
cache = "/tmp/field_uniform_16384_16.0.npy"
img = np.load(cache) if os.path.exists(cache) else fieldgen.uniform(S, S, 0)
seg = ws.TransformBuilder.default().build_segmenting()
mrg = ws.TransformBuilder.default().build_merging()
seeds = np.ascontiguousarray(seg.find_local_minima(img), dtype=np.uint64)
out = np.zeros((S, S), np.uint64)
def med(f, n=12):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    ts.sort()
    return 1e3 * ts[0], 1e3 * ts[len(ts) // 2]
a = med(lambda: seg.transform(img, seeds, out=out))
b = med(lambda: mrg.lake_counts(img, seeds))
print("segmenting transform min %.1f median %.1f ms, merging lake counts min %.1f median %.1f ms (pageable)" % (a + b), int(out[::97, ::89].sum()))
