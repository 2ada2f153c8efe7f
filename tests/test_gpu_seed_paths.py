"""The two initialisation paths (csrc/flood.cu): seed lists in row-major order without repeats -- what
find_local_minima returns -- are placed by fill_rows and coloured from the list order; any other list takes
seed_init and the label plane.  Both must give the oracle's bytes, and a list must not change its result by
being permuted (colours follow the seeds)."""
import numpy as np
import pytest

import fieldgen
from wsb200_loader import load

pytestmark = pytest.mark.gpu


def _check(ws, oracle, img, seeds, lmax=254):
    seg = ws.TransformBuilder.default().set_max_water_lvl(lmax).build_segmenting()
    lab, lvl = seg.transform_compact(img, seeds)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds, lmax)
    assert np.array_equal(lvl, ref.lvl)
    assert np.array_equal(lab.astype(np.uint64), ref.final)
    lakes, unc = ws.TransformBuilder.default().set_max_water_lvl(lmax).build_merging().lake_counts(img, seeds)
    exp = []
    oracle.transform(oracle.MERGING, img, seeds, lmax,
                     hook=lambda l, c: exp.append((np.unique(c[c != 0]).size, int((c == 0).sum()))))
    assert [(int(a), int(b)) for a, b in zip(lakes, unc)] == exp


def test_no_seeds_and_one_seed(oracle):
    ws = load()
    img = fieldgen.uniform(70, 150, 51)
    _check(ws, oracle, img, np.zeros((0, 2), np.uint64))
    _check(ws, oracle, img, np.array([[35, 77]], np.uint64))
    _check(ws, oracle, img, np.array([[0, 0]], np.uint64))               # a border seed: coloured, floods nothing by itself


def test_sorted_unsorted_and_repeated_lists(oracle):
    ws = load()
    img = fieldgen.uniform(131, 197, 52)
    seeds = oracle.find_local_minima(img)
    _check(ws, oracle, img, seeds)                                         # sorted: fill_rows
    rng = np.random.default_rng(3)
    perm = rng.permutation(len(seeds))
    _check(ws, oracle, img, seeds[perm])                                   # permuted: seed_init (colours follow the list)
    rep = np.concatenate([seeds, seeds[::7], seeds[:5]])                   # repeated positions: the later entry wins
    _check(ws, oracle, img, rep)
    border = np.array([[0, 5], [0, 64], [63, 0], [64, 0], [130, 196], [31, 63], [32, 64]], np.uint64)
    both = np.concatenate([seeds, border])
    both = both[np.lexsort((both[:, 1], both[:, 0]))]
    both = both[np.r_[True, np.any(np.diff(both.astype(np.int64), axis=0) != 0, axis=1)]]
    _check(ws, oracle, img, both)                                          # sorted, with border and tile-corner seeds


def test_rows_wider_than_one_bitmap_pass(oracle):
    ws = load()
    img = fieldgen.uniform(5, 70001, 53)                                   # 70 001 columns: three passes of fill_rows
    seeds = oracle.find_local_minima(img)
    assert seeds[:, 1].max() > 66000
    _check(ws, oracle, img, seeds, lmax=200)


def test_batch_with_empty_and_full_slices(oracle):
    ws = load()
    imgs = np.stack([fieldgen.uniform(96, 130, 60), np.full((96, 130), 7, np.uint8), fieldgen.smooth(96, 130, 3.0, 61),
                     fieldgen.uniform(96, 130, 62)])
    t = ws.TransformBuilder.default().build_merging()
    seeds, off = t.find_local_minima_batch(imgs)
    assert off[2] == off[1]                                                # the flat slice has no strict maximum
    labels, counts = ws.TransformBuilder.default().build_segmenting().transform_batch(imgs, seeds, off), None
    _, counts = t.transform_batch(imgs, seeds, off, want_labels=False, want_lake_counts=True)
    for k in range(len(imgs)):
        s = seeds[int(off[k]):int(off[k + 1])]
        ref = oracle.transform(oracle.SEGMENTING, imgs[k], s)
        assert np.array_equal(labels[0][k], ref.final)
        exp = []
        oracle.transform(oracle.MERGING, imgs[k], s, hook=lambda l, c: exp.append(np.unique(c[c != 0]).size))
        assert [int(x) for x in counts[k]] == exp
