import sys, numpy as np
sys.path.insert(0, '.')
import fieldgen
from wsb200_loader import load
from oracle import oracle as orc
ws = load()
img = fieldgen.uniform(64, 64, 1)
t = ws.TransformBuilder.default().build_segmenting()
seeds = t.find_local_minima(img)
print('seeds', len(seeds), flush=True)
lab, lvl = t.transform_compact(img, seeds)
ref = orc.transform(orc.SEGMENTING, img, seeds)
print('lvl ok', np.array_equal(lvl, ref.lvl), 'lab ok', np.array_equal(lab.astype(np.uint64), ref.final), flush=True)
