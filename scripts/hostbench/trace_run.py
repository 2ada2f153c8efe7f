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
for _ in range(3):
    seg.transform(img, seeds, out=out); mrg.lake_counts(img, seeds)
print("---- traced step", file=sys.stderr, flush=True)
os.environ["WS_TRACE_NOW"] = "1"
t0 = time.perf_counter(); seg.transform(img, seeds, out=out); t1 = time.perf_counter(); mrg.lake_counts(img, seeds); t2 = time.perf_counter()
print("seg %.1f ms, merge %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1)))
