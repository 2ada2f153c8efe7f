"""The oracle's drivers against an independent statement of the semantics (tests/spec_model.py):
arrival times, first-predecessor labels, connected-component merging -- on small images."""
import numpy as np
import pytest

import fieldgen
import spec_model as sm

CASES = {
    "uniform": lambda: fieldgen.uniform(40, 52, 1),
    "smooth": lambda: fieldgen.smooth(56, 48, 3.0, 2),
    "obstacles": lambda: fieldgen.obstacles(44, 40, 3),
    "plateaus": lambda: fieldgen.plateaus(36, 50, 5, 3.0, 4),
    "maze": lambda: fieldgen.maze(19, 17),
}


def _seeds(oracle, name, img):
    s = oracle.find_local_minima(img)
    if name == "maze":
        s = np.array([[1, 1]], np.uint64)
    return np.concatenate([s, np.array([[0, 3], [img.shape[0] - 1, 5]], np.uint64), s[:1]])


@pytest.mark.parametrize("lmax", [254, 90])
@pytest.mark.parametrize("name", list(CASES))
def test_drivers_match_spec(oracle, name, lmax):
    img = CASES[name]()
    seeds = _seeds(oracle, name, img)
    sl = [(int(a), int(b)) for a, b in seeds]
    r = oracle.transform(oracle.SEGMENTING, img, seeds, lmax, want_history=True)
    T = sm.arrival_times(img, sl, lmax)
    Torc = (r.lvl.astype(np.int64) << 24) | r.hop.astype(np.int64)
    Torc[r.lvl == 255] = sm.INF
    assert np.array_equal(T, Torc)
    lab = sm.segment_labels(T, sl)
    assert np.array_equal(lab, r.final.astype(np.int64))
    assert len(r.history) == lmax + 1
    for L, h in enumerate(r.history):
        assert np.array_equal(h.astype(np.int64), np.where(r.lvl <= L, lab, 0))
    # any random tie-break stays a valid segmentation with the same arrival times
    rr = oracle.transform(oracle.SEGMENTING, img, seeds, lmax, tie=oracle.TIE_RANDOM, rng_seed=5)
    assert np.array_equal(rr.lvl, r.lvl) and np.array_equal(rr.hop, r.hop)
    assert oracle.check_valid_segmentation(rr.final, rr.lvl, rr.hop, seeds) is None
    assert oracle.check_valid_segmentation(r.final, r.lvl, r.hop, seeds) is None
    # merging: literal closure == union-find closure == connected components
    m = oracle.transform(oracle.MERGING, img, seeds, lmax, want_history=True, fast_closure=False)
    m2 = oracle.transform(oracle.MERGING, img, seeds, lmax, want_history=True, fast_closure=True,
                          tie=oracle.TIE_RANDOM, rng_seed=3)
    for L in range(0, lmax + 1, 13):
        assert oracle.same_partition(sm.merging_partition(T, L), m.history[L]), L
    for a, b in zip(m.history, m2.history):
        assert oracle.same_partition(a, b)


def test_validity_checker_rejects_wrong_labels(oracle):
    img = fieldgen.smooth(40, 40, 3.0, 9)
    seeds = oracle.find_local_minima(img)
    r = oracle.transform(oracle.SEGMENTING, img, seeds)
    bad = r.final.copy()
    ys, xs = np.nonzero((r.lvl != 255) & (r.hop > 0))
    bad[ys[0], xs[0]] = bad.max() + 1
    assert oracle.check_valid_segmentation(bad, r.lvl, r.hop, seeds) is not None


def test_find_local_minima_is_strict_maxima(oracle):
    """lib.rs:1187-1191: neighbours `<` target; plateaus and borders never qualify."""
    img = np.zeros((7, 7), np.uint8)
    img[3, 3] = 9
    img[1, 1] = img[1, 2] = 5            # a two-pixel plateau: neither is strictly greater
    img[0, 5] = 200                      # border
    assert oracle.find_local_minima(img).tolist() == [[3, 3]]
    assert oracle.find_local_minima(np.zeros((2, 9), np.uint8)).shape == (0, 2)


def test_seed_out_of_bounds_is_an_error(oracle):
    img = fieldgen.uniform(10, 10, 1)
    with pytest.raises(IndexError):
        oracle.transform(oracle.SEGMENTING, img, [(10, 0)])
    # with edge correction the output is 12x12, so (11, 11) is legal (lib.rs:1365-1367)
    oracle.transform(oracle.SEGMENTING, img, [(11, 11)], edge_correction=True)


def test_merging_transform_const(oracle):
    out = oracle.merging_transform_const(5, 6)
    assert out[0].sum() == 0 and out[:, 0].sum() == 0 and (out[1:-1, 1:-1] == 123).all()


def _pp_numpy(a, MAX=254):
    """Independent statement of lib.rs:1134-1173 in numpy (f64 arithmetic)."""
    f = a.astype(np.float64)
    fin = np.isfinite(f)
    mn = min(0.0, float(f[fin].min())) if fin.any() else 0.0
    mx = max(0.0, float(f[fin].max())) if fin.any() else 0.0
    out = np.full(a.shape, 255, np.uint8)
    normal = fin & (np.abs(f) >= np.finfo(np.float64).tiny)
    with np.errstate(all="ignore"):
        out[normal] = (((f[normal] - mn) / (mx - mn)) * MAX).astype(np.uint8)
    out[np.isposinf(f)] = 0
    return out


def test_pre_processor_quirks(oracle):
    """lib.rs:1147-1170: folds start at zero, 0.0 / subnormal / NaN / -inf -> 255, +inf -> 0."""
    a = np.array([[0.0, 1.0, -2.0, np.nan, np.inf, -np.inf, 5e-324, 3.5]])
    assert oracle.pre_processor(a).tolist() == [[255, 138, 0, 255, 0, 255, 255, 254]]
    assert oracle.pre_processor(np.array([5.0, 10.0])).tolist() == [127, 254]      # min folded from 0, not 5
    assert oracle.pre_processor(np.array([0, 1, -2, 7, 3])).tolist() == [255, 84, 0, 254, 141]
    assert oracle.pre_processor(np.array([1.0, 2.0]), 127).tolist() == [63, 127]
    rng = np.random.default_rng(3)
    for dt in (np.float64, np.float32):
        x = rng.normal(size=(57, 33)).astype(dt)
        x[rng.random(x.shape) < 0.05] = np.nan
        x[3, 4], x[5, 6], x[7, 8] = np.inf, -np.inf, 0.0
        assert np.array_equal(oracle.pre_processor(x), _pp_numpy(x))
        assert np.array_equal(oracle.pre_processor(x, 99), _pp_numpy(x, 99))
    with pytest.raises(AssertionError):
        oracle.pre_processor(a, 255)
    with pytest.raises(AssertionError):
        oracle.pre_processor(a, 0)
