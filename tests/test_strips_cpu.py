"""The multi-GPU row-strip protocol on CPU: strips.solve with (a) all strips in one process and
(b) one strip per rank of a world-size-2 / 3 gloo group, against the oracle on the whole field."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import fieldgen
from strip_numpy_backend import NumpyStrip
from wsb200_loader import load


def _strips_mod():
    load()
    import importlib
    return importlib.import_module("rustronomy_watershed_b200.strips")


def _reference(oracle, img):
    seeds = oracle.find_local_minima(img)
    seg = oracle.transform(oracle.SEGMENTING, img, seeds)
    lakes = []
    oracle.transform(oracle.MERGING, img, seeds, hook=lambda l, c: lakes.append(np.unique(c[c != 0]).size))
    return seeds, seg, np.array(lakes, np.uint64)


def _make(st, img, n, sid):
    parts = st.partition_rows(img.shape[0], n)
    g = st.StripGeometry(sid, n, img.shape[0], parts[sid])
    lo, hi = g.local_rows
    return NumpyStrip(g, img[lo:hi])


FIELDS = {
    "uniform": lambda: fieldgen.uniform(40, 36, 3),
    "smooth": lambda: fieldgen.smooth(45, 30, 3.0, 4),         # long geodesics cross the cuts several times
    "obstacles": lambda: fieldgen.obstacles(38, 33, 5),
}


def test_partition_rows():
    st = _strips_mod()
    assert st.partition_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert st.partition_rows(16384, 8)[-1] == (14336, 16384)
    with pytest.raises(ValueError):
        st.partition_rows(2, 3)
    g = st.StripGeometry(1, 3, 10, (4, 7))
    assert g.halo_top and g.halo_bottom and g.local_rows == (3, 8)
    assert st.StripGeometry(0, 3, 10, (0, 4)).local_rows == (0, 5)


@pytest.mark.parametrize("name", list(FIELDS))
@pytest.mark.parametrize("n", [1, 2, 3, 5])
def test_strips_one_process(oracle, name, n):
    st = _strips_mod()
    img = FIELDS[name]()
    seeds, seg, lakes = _reference(oracle, img)
    strips = [_make(st, img, n, s) for s in range(n)]
    res = st.solve(strips, st.LocalComm(n), st.MERGING, 254)
    assert res.nseeds_total == len(seeds)
    assert np.array_equal(np.concatenate([s.owned_labels() for s in strips]).astype(np.uint64), seg.final)
    assert np.array_equal(np.concatenate([s.owned_levels() for s in strips]), seg.lvl)
    assert np.array_equal(res.lake_counts, lakes)
    if n > 1 and name == "smooth":
        assert res.flood_rounds >= 2


def _worker(rank, world, init_file, name, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    st = _strips_mod()
    img = FIELDS[name]()
    strip = _make(st, img, world, rank)
    res = st.solve([strip], st.DistComm(), st.MERGING, 254)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), lab=strip.owned_labels(), lvl=strip.owned_levels(),
             lakes=res.lake_counts, rounds=np.array([res.flood_rounds, res.label_rounds]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name", [(2, "smooth"), (2, "uniform"), (3, "obstacles")])
def test_strips_gloo_ranks(oracle, world, name):
    img = FIELDS[name]()
    seeds, seg, lakes = _reference(oracle, img)
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "rendezvous")
        mp.spawn(_worker, args=(world, init_file, name, d), nprocs=world, join=True)
        parts = [np.load(os.path.join(d, f"r{r}.npz")) for r in range(world)]
    assert np.array_equal(np.concatenate([p["lab"] for p in parts]).astype(np.uint64), seg.final)
    assert np.array_equal(np.concatenate([p["lvl"] for p in parts]), seg.lvl)
    for p in parts:                                   # every rank ends with the same per-level lake counts
        assert np.array_equal(p["lakes"], lakes)
