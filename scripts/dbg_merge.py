import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import fieldgen
from wsb200_loader import load
from oracle import oracle as orc
ws = load()
img = fieldgen.obstacles(120, 90, 6)
s = orc.find_local_minima(img)
seeds = np.concatenate([s, np.array([[0, 3], [img.shape[0] - 1, 5], [0, 0]], np.uint64), s[:1]])
ref = orc.transform(orc.MERGING, img, seeds, want_history=True, fast_closure=True)
refs = orc.transform(orc.SEGMENTING, img, seeds)
t = ws.TransformBuilder.default().build_merging()
for it in range(4):
    lab, lvl = t.transform_compact(img, seeds)
    print("seg ok", np.array_equal(lab.astype(np.uint64), refs.final), np.array_equal(lvl, refs.lvl))
    hist = t.transform_history(img, seeds)
    lakes, unc = t.lake_counts(img, seeds)
    bad = [l for (l, snap), exp in zip(hist, ref.history) if not orc.same_partition(snap, exp)]
    badc = [l for l, exp in enumerate(ref.history) if lakes[l] != np.unique(exp[exp != 0]).size]
    print("run", it, "bad partition levels", bad[:10], len(bad), "bad counts", badc[:10], len(badc))
    if bad:
        l = bad[0]
        snap, exp = hist[l][1], ref.history[l]
        print(" level", l, "lakes gpu", np.unique(snap[snap != 0]).size, "ref", np.unique(exp[exp != 0]).size, "count api", int(lakes[l]))
        # find a pair of pixels merged in ref but not in gpu or vice versa
        pairs = np.unique(np.stack([snap.ravel(), exp.ravel()], 1), axis=0)
        from collections import Counter
        cg = Counter(pairs[:, 0].tolist()); ce = Counter(pairs[:, 1].tolist())
        print(" gpu labels mapping to >1 ref labels:", [k for k, v in cg.items() if v > 1][:5])
        print(" ref labels mapping to >1 gpu labels:", [k for k, v in ce.items() if v > 1][:5])
        k = [k for k, v in ce.items() if v > 1]
        if k:
            ys, xs = np.nonzero(exp == k[0])
            print("  ref lake", k[0], "bbox", ys.min(), ys.max(), xs.min(), xs.max(), "gpu labels", np.unique(snap[exp == k[0]]))
