"""Row-strip decomposition on the GPU: several strips (one ws_plan each) in one process, boundary rows
moved by device copies -- bit-exact with the oracle on the whole field (labels, levels, lakes per level)."""
import numpy as np
import pytest

import fieldgen
from wsb200_loader import load

pytestmark = pytest.mark.gpu

CASES = {
    "uniform": lambda: fieldgen.uniform(300, 260, 3),
    "smooth": lambda: fieldgen.smooth(257, 190, 5.0, 4),        # geodesics cross the cuts many times
    "obstacles": lambda: fieldgen.obstacles(190, 333, 5),
    "plateaus": lambda: fieldgen.plateaus(160, 200, 4, 3.0, 6),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("n", [2, 3, 7])
def test_cuda_strips_match_whole_field(oracle, name, n):
    import importlib
    import torch
    ws = load()
    st = importlib.import_module("rustronomy_watershed_b200.strips")
    img = CASES[name]()
    seeds = oracle.find_local_minima(img)
    seg = oracle.transform(oracle.SEGMENTING, img, seeds)
    lakes = []
    oracle.transform(oracle.MERGING, img, seeds, hook=lambda l, c: lakes.append(np.unique(c[c != 0]).size))
    ctx = ws.default_context()
    parts = st.partition_rows(img.shape[0], n)
    strips = []
    for sid in range(n):
        g = st.StripGeometry(sid, n, img.shape[0], parts[sid])
        lo, hi = g.local_rows
        strips.append(st.CudaStrip(ws, ctx, g, torch.from_numpy(img[lo:hi].copy()).cuda()))
    try:
        res = st.solve(strips, st.LocalComm(n), st.MERGING, 254)
        lab = np.concatenate([s.owned_labels() for s in strips])
        lvl = np.concatenate([s.owned_levels() for s in strips])
    finally:
        for s in strips:
            s.close()
    assert res.nseeds_total == len(seeds)
    assert np.array_equal(lvl, seg.lvl)
    assert np.array_equal(lab.astype(np.uint64), seg.final)
    assert np.array_equal(res.lake_counts, np.array(lakes, np.uint64))
