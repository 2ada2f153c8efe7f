"""The oracle against the reference's own unit-test vectors (the parity pin).

Vectors: tests/golden/reference_unit_vectors.json, extracted from
/root/reference/src/lib.rs by tests/golden/make_reference_vectors.py.
"""
import itertools
import json
import os

import numpy as np
import pytest

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_unit_vectors.json")))
ORD = {"Less": -1, "Equal": 0, "Greater": 1}


def test_find_px(oracle):
    """lib.rs:259-291"""
    v = G["test_find_px"]
    img, col = np.array(v["input"], np.uint8), np.array(v["colours"], np.uint64)
    idx, colours = oracle.find_flooded_px(img, col, v["lvl"])
    got = {(int(i) // img.shape[1], int(i) % img.shape[1]) for i in idx}
    for a in v["must_contain"]:
        assert tuple(a) in got
    # stronger than the reference's assertion: the exact set (by hand from lib.rs:224-231)
    assert got == {(1, 5), (2, 2), (4, 4), (4, 5), (5, 6)}
    assert set(int(c) for c in colours) == {1}


def test_merge_eq(oracle):
    """lib.rs:308-311"""
    a, b = G["test_merge_eq"]["equal"]
    assert oracle.merge_eq(a, b)
    assert not oracle.merge_eq([1, 2], [1, 3])


@pytest.mark.parametrize("name", ["test_merge_ord_small_big", "test_merge_ord_big_small"])
def test_merge_ord(oracle, name):
    """lib.rs:336-344, 369-377"""
    fn = oracle.sort_by_small_big if name.endswith("small_big") else oracle.sort_by_big_small
    for c in G[name]["cases"]:
        assert fn(c["a"], c["b"]) == ORD[c["ordering"]], c


def test_find_merge(oracle):
    """lib.rs:447-465"""
    v = G["test_find_merge"]
    res = oracle.find_merge(np.array(v["input"], np.uint64))
    assert len(res) == len(v["answer"])
    for p in res:
        assert any(oracle.merge_eq([int(p[0]), int(p[1])], a) for a in v["answer"])


def test_make_colour_map(oracle):
    """lib.rs:544-587; every ordering of every step instead of 10 random shuffles."""
    for sc in G["test_make_colour_map"]["scenarios"]:
        orders = [list(itertools.permutations(step)) for step in sc["steps"]]
        for combo in itertools.product(*orders):
            cmap = np.array(sc["start"], np.uint64)
            for step in combo:
                cmap = oracle.make_colour_map(cmap, np.array(step, np.uint64))
            assert cmap.tolist() == sc["expect"], (sc, combo)


def test_recolour(oracle):
    """lib.rs:594-626"""
    v = G["test_recolour"]
    out = oracle.recolour(np.array(v["input"], np.uint64), np.array(v["cmap"], np.uint64))
    assert out.tolist() == v["answer"]
    out2 = oracle.recolour(out, np.array(v["stale_cmap"], np.uint64))
    assert out2.tolist() == v["answer"]


def test_find_lake_sizes(oracle):
    """lib.rs:629-635: histogram of length H*W+1, index 0 = uncoloured."""
    col = np.array(G["test_recolour"]["answer"], np.uint64)
    s = oracle.find_lake_sizes(col)
    assert s.size == col.size + 1
    assert s[0] == (col == 0).sum() and s[1] == (col == 1).sum() and s[4] == 4 and s[5] == 5
    assert s.sum() == col.size
