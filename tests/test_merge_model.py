"""The algorithm of csrc/merge.cu (per-tile Boruvka with FINAL / DEFERRED picks + global level-ordered union),
restated on the CPU in tests/merge_model.py, against the oracle's merging transform: lakes per level exact,
partitions at several levels equal up to a renumbering.  Needs no GPU: the rules the kernel follows are
checked here, the kernel itself in tests/test_gpu_*."""
import numpy as np
import pytest

import fieldgen
import merge_model as mm

CASES = {
    "uniform": lambda: fieldgen.uniform(150, 200, 3),
    "uniform_tiles": lambda: fieldgen.uniform(96, 192, 4),          # whole tiles, several in each direction
    "smooth": lambda: fieldgen.smooth(130, 170, 3.0, 5),
    "obstacles": lambda: fieldgen.obstacles(100, 140, 6),
    "plateaus": lambda: fieldgen.plateaus(90, 150, 5, 2.0, 7),
    "few_values": lambda: (fieldgen.uniform(120, 150, 8) // 32).astype(np.uint8),
}


@pytest.mark.parametrize("name", list(CASES))
def test_model_lake_counts_and_partitions(oracle, name):
    img = CASES[name]()
    seeds = oracle.find_local_minima(img)
    if name == "few_values" or len(seeds) == 0:
        rng = np.random.default_rng(1)
        seeds = np.stack([rng.integers(0, img.shape[0], 60), rng.integers(0, img.shape[1], 60)], 1).astype(np.uint64)
    seg = oracle.transform(oracle.SEGMENTING, img, seeds)
    exp, snaps = [], {}
    levels = (0, 30, 100, 180, 254)

    def hook(l, c):
        exp.append(np.unique(c[c != 0]).size)
        if l in levels:
            snaps[l] = c.copy()
    oracle.transform(oracle.MERGING, img, seeds, hook=hook)
    lab, lvl = seg.final.astype(np.int64), seg.lvl.astype(np.int64)
    got = mm.lake_counts(lab, lvl, len(seeds))
    assert np.array_equal(got, np.array(exp)), "lakes per level"
    assert np.array_equal(mm.lake_counts_forest(lab, lvl, len(seeds)), np.array(exp)), "lakes per level (forest rounds)"
    parts = mm.partitions(lab, lvl, len(seeds), levels)
    for L in levels:
        assert oracle.same_partition(parts[L], snaps[L]), f"partition at level {L}"
    # without contraction (every basin open) nothing is FINAL and the counts are the same
    assert np.array_equal(mm.lake_counts(lab, lvl, len(seeds), contract=False), np.array(exp))
    F, D, rounds = mm.reduce_image(lab, lvl, contract=False)
    assert len(F) == 0


def test_forest_rounds_with_open_nodes_compose(oracle):
    """Strip-style use of the rounds: cut the DEFERRED graph of an image into two halves by node id, mark the
    nodes that have edges in both halves open, run the rounds on each half, then on the union of the two
    DEFERRED lists: FINAL counts per level must add up to the forest of the whole graph."""
    img = fieldgen.uniform(128, 256, 11)
    seeds = oracle.find_local_minima(img)
    seg = oracle.transform(oracle.SEGMENTING, img, seeds)
    lab, lvl = seg.final.astype(np.int64), seg.lvl.astype(np.int64)
    _, D, _ = mm.reduce_image(lab, lvl, True)
    n = len(seeds) + 1
    whole, rest, _ = mm.forest_rounds(D, n)
    assert len(rest) == 0
    rng = np.random.default_rng(3)
    side = rng.integers(0, 2, len(D)).astype(bool)          # which half owns an edge
    touch = np.zeros((2, n), bool)
    for h in (0, 1):
        e = D[side == bool(h)]
        touch[h, e[:, 0]] = True
        touch[h, e[:, 1]] = True
    open_ = touch[0] & touch[1]
    hist = np.zeros(256, np.int64)
    gathered = []
    for h in (0, 1):
        F, Dh, _ = mm.forest_rounds(D[side == bool(h)], n, open_)
        hist += np.bincount(F[:, 2], minlength=256)
        gathered.append(Dh)
    F, Dg, _ = mm.forest_rounds(np.concatenate(gathered), n)
    assert len(Dg) == 0
    hist += np.bincount(F[:, 2], minlength=256)
    assert np.array_equal(hist, np.bincount(whole[:, 2], minlength=256))
