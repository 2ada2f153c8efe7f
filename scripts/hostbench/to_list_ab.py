"""transform_to_list of config 2 (512^2 uniform, merging): 256 rows of rows*cols+1 usize = 537 MB into a fresh array."""
import sys, time, numpy as np
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
ws = load()
img = fieldgen.uniform(512, 512, 0)
t = ws.TransformBuilder.default().build_merging()
seeds = t.find_local_minima(img)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); r = t.transform_to_list(img, seeds); ts.append(1e3 * (time.perf_counter() - t0)); del r
print("transform_to_list ms:", [round(x, 1) for x in ts])
