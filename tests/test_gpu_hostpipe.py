"""The host side of the reference-facing calls: pageable caller memory through the page-locked ring and the
worker threads (csrc/hostpipe.h), page-locked memory directly, seeds found on the device (WS_SEEDS_AUTO) --
every route must give the bytes of the oracle."""
import numpy as np
import pytest
import torch

import fieldgen
from wsb200_loader import load

pytestmark = pytest.mark.gpu


def _pin(a: np.ndarray) -> np.ndarray:
    """A page-locked copy of `a` (kept alive by the returned array's base)."""
    if a.dtype == np.uint64:
        t = torch.from_numpy(a.view(np.int64).copy()).pin_memory()
        out = t.numpy().view(np.uint64)
    else:
        t = torch.from_numpy(a.copy()).pin_memory()
        out = t.numpy()
    out_holder.append(t)
    return out


out_holder = []


def test_pageable_pinned_and_host_widened_outputs_agree(oracle):
    ws = load()
    img = fieldgen.uniform(1536, 2048, 41)                      # 3 MB image, 5.5 MB of seeds: both staged
    seeds = oracle.find_local_minima(img)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds)
    seg = ws.TransformBuilder.default().build_segmenting()
    ctx = seg._ctx()
    a = seg.transform(img, seeds)                               # pageable in, pageable out
    assert np.array_equal(a, ref.final)
    pimg, pseeds = _pin(img), _pin(seeds)
    pout = _pin(np.zeros(img.shape, np.uint64))
    seg.transform(pimg, pseeds, out=pout)                       # page-locked in and out: widened on the device
    assert np.array_equal(pout, ref.final)
    ctx.set_option(ws._native.WS_OPT_PINNED_HOST_WIDEN, 1)
    try:
        pout[:] = 0
        seg.transform(pimg, pseeds, out=pout)                   # page-locked, widened by the host threads
        assert np.array_equal(pout, ref.final)
    finally:
        ctx.set_option(ws._native.WS_OPT_PINNED_HOST_WIDEN, 0)
    for n in (1, 3, 0):
        ctx.set_host_threads(n)
        assert np.array_equal(seg.transform(img, seeds), ref.final)
        lab, lvl = seg.transform_compact(img, seeds)
        assert np.array_equal(lab.astype(np.uint64), ref.final) and np.array_equal(lvl, ref.lvl)
    # unaligned pageable destination (the non-temporal stores want 16-byte alignment)
    raw = np.zeros(img.size + 1, np.uint64)
    out = raw[1:].reshape(img.shape)
    seg.transform(img, seeds, out=out)
    assert np.array_equal(out, ref.final)


def test_strided_views_through_the_ring(oracle):
    ws = load()
    base = fieldgen.uniform(2600, 3100, 42)
    seg = ws.TransformBuilder.default().build_segmenting()
    for view in (base[::2, ::3], base.T[100:1500, 7:2000], base[::-1, :][:1400, :1000], base[5:1300, 11:1711]):
        dense = np.ascontiguousarray(view)
        seeds = seg.find_local_minima(dense)
        assert np.array_equal(seg.find_local_minima(view), seeds)
        lab_v, lvl_v = seg.transform_compact(view, seeds)
        lab_d, lvl_d = seg.transform_compact(dense, seeds)
        assert np.array_equal(lab_v, lab_d) and np.array_equal(lvl_v, lvl_d)
    small = base[:300, :400]
    ref = oracle.transform(oracle.SEGMENTING, small, oracle.find_local_minima(small))
    assert np.array_equal(seg.transform(small, oracle.find_local_minima(small)), ref.final)


def test_seeds_found_on_the_device(oracle):
    ws = load()
    img = fieldgen.smooth(700, 900, 3.0, 43)
    seeds = oracle.find_local_minima(img)
    seg = ws.TransformBuilder.default().build_segmenting()
    mrg = ws.TransformBuilder.default().build_merging()
    assert np.array_equal(seg.transform(img, None), seg.transform(img, seeds))
    la, ua = mrg.lake_counts(img, None)
    lb, ub = mrg.lake_counts(img, seeds)
    assert np.array_equal(la, lb) and np.array_equal(ua, ub)
    seen = []
    hooked = ws.TransformBuilder.default().set_max_water_lvl(3).set_wlvl_hook(
        lambda c: seen.append([(col, rc) for col, rc in c.seeds])).build_segmenting()
    hooked.transform_with_hook(img, None)
    assert len(seen) == 4
    assert seen[0] == [(i + 1, (int(r), int(c))) for i, (r, c) in enumerate(seeds)]
    with pytest.raises(ws.WatershedError):
        ws.TransformBuilder.default().enable_edge_correction().build_segmenting().transform(img, None)
    with pytest.raises(ws.WatershedError):
        mrg.transform_to_list(img, None)


def test_out_of_bounds_seed_in_a_pageable_list(oracle):
    ws = load()
    img = fieldgen.uniform(600, 700, 44)
    seeds = np.tile(oracle.find_local_minima(img), (3, 1))      # > 1 MB of (usize, usize): staged + narrowed
    assert seeds.nbytes > (1 << 20)
    seg = ws.TransformBuilder.default().build_segmenting()
    seg.transform(img, seeds)
    bad = seeds.copy()
    bad[len(bad) // 2, 1] = 700                                  # the reference panics: lib.rs:1366 / 1676
    with pytest.raises(ws.WatershedError) as e:
        seg.transform(img, bad)
    assert e.value.status == ws._native.WS_ERR_SEED_OOB
    bad[len(bad) // 2, 1] = 2 ** 40 + 5                          # must not alias a valid column after narrowing
    with pytest.raises(ws.WatershedError):
        seg.transform(img, bad)
