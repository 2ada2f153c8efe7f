"""Parity at BASELINE.json sizes.

  * configs 5 and 3 against the OpenMP oracle itself: one 2048^2 CGPS-like slice and the 4096^2 sigma-8 field --
    levels and segmenting labels bit-exact, lakes per level exact, history snapshots at levels {0, 64, 127, 254}
    (segmenting: bit-exact; merging: equal up to a renumbering per level, lib.rs:539);
  * config 4 (16384^2 uniform and smoothed): lakes per level against an INDEPENDENT connected-components pass
    (scipy.ndimage.label over `level of colouring <= L`): the merging partition at level L is the set of
    4-connected components of the coloured pixels where at least one pixel of every adjacent pair is a window
    centre (find_merge looks from centres only, lib.rs:411-414) -- nothing of the engine's merge path (tile
    contraction, edge lists, union-find) takes part in the expected value;
  * determinism: the asynchronous flood runs 20 times on the 16384^2 smoothed field; arrival times and labels
    must be identical every time (the worklist protocol is lock-free, so a lost wake-up would show here);
  * the reference's own tie-break (random among the coloured neighbours, lib.rs:250-253): valid by the
    oracle's checker, reproducible for a given seed, different for another.
"""
import numpy as np
import pytest
import torch

import fieldgen
from conftest import big_field
from wsb200_loader import load

pytestmark = pytest.mark.gpu

SNAP_LEVELS = (0, 64, 127, 254)


class _Dev:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, shape, typestr):
    return torch.as_tensor(_Dev(ptr, shape, typestr), device="cuda")


def _plan_run(ws, kind, img, seeds_rc=None):
    """One device-level run; seeds from the engine's own find_local_minima unless given."""
    ctx = ws.default_context()
    R, C = img.shape
    plan = ws.Plan(ctx, 1, R, C)
    d_img = torch.from_numpy(img).cuda()
    off = torch.zeros(2, dtype=torch.int32, device="cuda")
    if seeds_rc is None:
        n = plan.find_local_minima(d_img.data_ptr(), 0, 0, off.data_ptr())
        seeds = torch.empty((max(n, 1), 2), dtype=torch.int32, device="cuda")
        plan.find_local_minima(d_img.data_ptr(), seeds.data_ptr(), n, off.data_ptr())
    else:
        n = len(seeds_rc)
        seeds = torch.from_numpy(np.ascontiguousarray(seeds_rc, dtype=np.int64).astype(np.int32)).cuda()
        off = torch.tensor([0, n], dtype=torch.int32, device="cuda")
    plan.run(kind, 254, d_img.data_ptr(), seeds.data_ptr(), off.data_ptr(), n)
    return ctx, plan, d_img, seeds, off, n


@pytest.mark.parametrize("name", ["cgps_2048", "smooth8_4096"])
def test_baseline_config_against_oracle(oracle, name):
    """BASELINE configs 5 (one slice of the batch) and 3, compared with the oracle pixel by pixel."""
    ws = load()
    img = fieldgen.cgps_like(2048, 2048, 0) if name == "cgps_2048" else fieldgen.smooth(4096, 4096, 8.0, 0)
    R, C = img.shape
    seeds = oracle.find_local_minima(img)
    ctx, plan, d_img, d_seeds, d_off, n = _plan_run(ws, 1, img)      # a merging run leaves the segmenting labels too
    assert n == len(seeds)
    assert np.array_equal(d_seeds[:n].cpu().numpy().astype(np.uint64), seeds)

    # --- segmenting: oracle's literal level loop, snapshots at the chosen levels through its hook
    seg_snap = {}
    ref = oracle.transform(oracle.SEGMENTING, img, seeds,
                           hook=lambda l, c: seg_snap.__setitem__(l, c.astype(np.uint32)) if l in SNAP_LEVELS else None)
    lvl = _view(plan.levels_ptr, (R, C), "|u1").cpu().numpy()
    lab = (_view(plan.labels_ptr, (R, C), "<i4").cpu().numpy().view(np.uint32) & 0x7FFFFFFF)
    assert np.array_equal(lvl, ref.lvl), "levels of colouring differ from the oracle"
    assert np.array_equal(lab, ref.final.astype(np.uint32)), "segmenting labels differ from the oracle"
    T = _view(plan.arrival_times_ptr, (R, C), "<i4").cpu().numpy().view(np.uint32)
    exp_T = (ref.lvl.astype(np.uint32) << 24) | ref.hop
    exp_T[ref.lvl == 255] = 0xFF000000
    assert np.array_equal(np.minimum(T, 0xFF000000), exp_T), "arrival times differ from the loop counters"
    out = torch.empty((R, C), dtype=torch.int64, device="cuda")
    for L in SNAP_LEVELS:
        plan.snapshot(0, 0, L, out.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy().astype(np.uint32), seg_snap[L]), f"segmenting snapshot at level {L}"
    del seg_snap, ref

    # --- merging: per-level lake counts + uncoloured counts, partitions at the chosen levels
    exp_counts, mrg_snap = [], {}

    def hook(l, c):
        exp_counts.append((np.unique(c[c != 0]).size, int((c == 0).sum())))
        if l in SNAP_LEVELS:
            mrg_snap[l] = c.astype(np.uint32)
    # (the literal make_colour_map is super-linear in the number of merges: used where it is affordable)
    oracle.transform(oracle.MERGING, img, seeds, fast_closure=(name == "cgps_2048"), hook=hook)
    counts = _view(plan.lake_counts_ptr, (256,), "<i4").cpu().numpy()[:255]
    assert [int(x) for x in counts] == [a for a, _ in exp_counts], "lakes per level differ from the oracle"
    lv = torch.from_numpy(lvl).cuda()
    unc = [int((lv > L).sum()) for L in range(255)]
    assert unc == [b for _, b in exp_counts], "uncoloured pixels per level differ from the oracle"
    for L in SNAP_LEVELS:
        plan.snapshot(1, 0, L, out.data_ptr())
        torch.cuda.synchronize()
        assert oracle.same_partition(out.cpu().numpy(), mrg_snap[L]), f"merging partition at level {L}"
    plan.close()


def _independent_lake_count(lvl_cpu: np.ndarray, L: int) -> int:
    import scipy.ndimage as ndi
    mask = lvl_cpu <= L
    # only window centres flood (lib.rs:220) and the seeds of find_local_minima are interior pixels, so no
    # border pixel is coloured and the centre-pixel rule of find_merge (lib.rs:411-414) never excludes a pair
    assert not mask[0].any() and not mask[-1].any() and not mask[:, 0].any() and not mask[:, -1].any()
    _, n = ndi.label(mask)          # default structure: 4-connectivity
    return int(n)


@pytest.mark.parametrize("field", ["uniform", "smooth"])
def test_lake_counts_against_independent_components_16384(field):
    """BASELINE config 4 on one GPU: the engine's lakes per level == number of connected components of the
    coloured set, counted by scipy on the host from the level-of-colouring image alone."""
    ws = load()
    S = 16384
    img = big_field(field, S)
    ctx, plan, d_img, d_seeds, d_off, n = _plan_run(ws, 1, img)
    counts = _view(plan.lake_counts_ptr, (256,), "<i4").cpu().numpy()
    lvl = _view(plan.levels_ptr, (S, S), "|u1").cpu().numpy()
    plan.close()
    levels = (3, 128, 254) if field == "uniform" else (40, 128, 254)
    for L in levels:
        assert _independent_lake_count(lvl, L) == int(counts[L]), f"lakes at level {L}"


def test_flood_is_deterministic_16384_smooth():
    """20 runs of the same flood (asynchronous worklist, ~3 activations per tile in a different order every
    time): arrival times, labels and levels identical to the first run, bit for bit."""
    ws = load()
    S = 16384
    img = big_field("smooth", S)
    ctx, plan, d_img, d_seeds, d_off, n = _plan_run(ws, 0, img)
    T0 = _view(plan.arrival_times_ptr, (S, S), "<i4").clone()
    L0 = _view(plan.labels_ptr, (S, S), "<i4").clone()
    V0 = _view(plan.levels_ptr, (S, S), "|u1").clone()
    activations = set()
    for it in range(20):
        plan.run(0, 254, d_img.data_ptr(), d_seeds.data_ptr(), d_off.data_ptr(), n)
        activations.add(plan.stats()["tile_activations"])
        assert torch.equal(_view(plan.arrival_times_ptr, (S, S), "<i4"), T0), f"arrival times differ in run {it}"
        assert torch.equal(_view(plan.labels_ptr, (S, S), "<i4"), L0), f"labels differ in run {it}"
        assert torch.equal(_view(plan.levels_ptr, (S, S), "|u1"), V0), f"levels differ in run {it}"
    plan.close()
    # (the schedule itself is NOT deterministic -- that is the point of the test)
    assert len(activations) >= 1


def test_random_tie_break_is_a_valid_reference_outcome(oracle):
    """WS_TIE_RANDOM (lib.rs:250-253): every pixel carries the colour of a neighbour coloured strictly earlier,
    THE colour where those agree (oracle.check_valid_segmentation); same seed -> same image; the levels do not
    depend on the tie-break; the merging statistics do not either."""
    ws = load()
    img = fieldgen.uniform(300, 420, 21)
    seeds = oracle.find_local_minima(img)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds)
    first = ws.TransformBuilder.default().build_segmenting().transform_compact(img, seeds)
    imgs = []
    for seed in (1, 1, 2):
        t = ws.TransformBuilder.default().set_tie_break("random", seed=seed).build_segmenting()
        lab, lvl = t.transform_compact(img, seeds)
        assert np.array_equal(lvl, ref.lvl)
        assert oracle.check_valid_segmentation(lab, ref.lvl, ref.hop, seeds) is None
        imgs.append(lab)
    assert np.array_equal(imgs[0], imgs[1]), "the same tie seed must reproduce the image"
    assert not np.array_equal(imgs[0], imgs[2]), "another tie seed should move contested pixels"
    assert not np.array_equal(imgs[0], first[0])
    # the oracle's own random mode agrees on everything that does not depend on the draw
    rnd = oracle.transform(oracle.SEGMENTING, img, seeds, tie=oracle.TIE_RANDOM, rng_seed=5)
    contested = imgs[0] != rnd.final.astype(np.uint32)
    assert contested.mean() < 0.5
    m_first = ws.TransformBuilder.default().build_merging().lake_counts(img, seeds)
    m_rand = ws.TransformBuilder.default().set_tie_break("random", seed=3).build_merging().lake_counts(img, seeds)
    assert np.array_equal(m_first[0], m_rand[0]) and np.array_equal(m_first[1], m_rand[1])


def test_lake_sizes_compact_equals_transform_to_list(oracle):
    ws = load()
    img = fieldgen.uniform(96, 130, 31)
    seeds = oracle.find_local_minima(img)
    for build in ("build_segmenting", "build_merging"):
        t = getattr(ws.TransformBuilder.default(), build)()
        lakes, sizes = t.lake_sizes_compact(img, seeds)
        full = t.transform_to_list(img, seeds)
        assert sizes.shape == (255, len(seeds) + 1)
        for L, (lv, row) in enumerate(full):
            assert lv == L
            assert np.array_equal(row[: len(seeds) + 1], sizes[L])
            assert not row[len(seeds) + 1:].any()
            assert int(lakes[L]) == int((row[1:] != 0).sum())
