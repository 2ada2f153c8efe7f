"""Randomised parity sweep: many shapes / field kinds / seed sets, segmenting labels + levels
bit-exact, merging lake counts and uncoloured counts exact at every level."""
import numpy as np
import pytest

import fieldgen
from wsb200_loader import load

pytestmark = pytest.mark.gpu


def _case(i):
    rng = np.random.default_rng(1000 + i)
    rows, cols = int(rng.integers(3, 260)), int(rng.integers(3, 330))
    kind = i % 5
    if kind == 0:
        img = fieldgen.uniform(rows, cols, i)
    elif kind == 1:
        img = fieldgen.smooth(rows, cols, float(rng.uniform(1.0, 6.0)), i)
    elif kind == 2:
        img = fieldgen.obstacles(rows, cols, i)
    elif kind == 3:
        img = fieldgen.plateaus(rows, cols, int(rng.integers(2, 9)), float(rng.uniform(1.0, 4.0)), i)
    else:
        img = (fieldgen.uniform(rows, cols, i) // int(rng.integers(8, 64))).astype(np.uint8)   # few distinct values
    return img, rng


@pytest.mark.parametrize("i", range(30))
def test_random_field(oracle, i):
    ws = load()
    img, rng = _case(i)
    seeds = oracle.find_local_minima(img)
    mode = i % 3
    if mode == 1 or len(seeds) == 0:       # arbitrary seeds, incl. border pixels and duplicates
        n = int(rng.integers(1, 40))
        seeds = np.stack([rng.integers(0, img.shape[0], n), rng.integers(0, img.shape[1], n)], 1).astype(np.uint64)
        seeds = np.concatenate([seeds, seeds[:3]])
    elif mode == 2:                         # a sparse subset of the maxima
        seeds = seeds[:: int(rng.integers(2, 9))]
    lmax = int(rng.choice([254, 254, 200, 37]))
    seg = ws.TransformBuilder.default().set_max_water_lvl(lmax).build_segmenting()
    assert np.array_equal(seg.find_local_minima(img), oracle.find_local_minima(img))
    lab, lvl = seg.transform_compact(img, seeds)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds, lmax)
    assert np.array_equal(lvl, ref.lvl)
    assert np.array_equal(lab.astype(np.uint64), ref.final)
    lakes, unc = ws.TransformBuilder.default().set_max_water_lvl(lmax).build_merging().lake_counts(img, seeds)
    exp = []
    oracle.transform(oracle.MERGING, img, seeds, lmax,
                     hook=lambda l, c: exp.append((np.unique(c[c != 0]).size, int((c == 0).sum()))))
    assert [(int(a), int(b)) for a, b in zip(lakes, unc)] == exp


def test_merging_lake_counts_1024_uniform(oracle):
    """tests/core_bench.rs:29 shape: 1024^2 uniform, merging."""
    ws = load()
    img = fieldgen.uniform(1024, 1024, 0)
    t = ws.TransformBuilder.default().build_merging()
    seeds = t.find_local_minima(img)
    lakes, unc = t.lake_counts(img, seeds)
    exp = []
    oracle.transform(oracle.MERGING, img, seeds,
                     hook=lambda l, c: exp.append((np.unique(c[c != 0]).size, int((c == 0).sum()))))
    assert [(int(a), int(b)) for a, b in zip(lakes, unc)] == exp
