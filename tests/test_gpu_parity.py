"""GPU parity: the CUDA path through the C ABI vs the CPU oracle, same seeded inputs.

Parity rules:
  * find_local_minima: identical list, row-major order;
  * segmenting: bit-exact labels at every level under the `first coloured neighbour
    in the order down,right,left,up` tie-break; arrival times (level, pass) identical;
  * merging: per level equal up to a renumbering of the non-zero labels; lake count,
    lake-size multiset and uncoloured count exact.
"""
import numpy as np
import pytest

import fieldgen
from wsb200_loader import load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ws():
    return load()


def _fields():
    return {
        "uniform_64": fieldgen.uniform(64, 64, 1),
        "uniform_odd": fieldgen.uniform(97, 131, 2),            # not a multiple of the 64x32 tile
        "smooth_128": fieldgen.smooth(128, 128, 4.0, 3),
        "smooth_wide": fieldgen.smooth(70, 300, 6.0, 4),
        "plateaus": fieldgen.plateaus(96, 160, 5, 3.0, 5),
        "obstacles": fieldgen.obstacles(120, 90, 6),
        "maze": fieldgen.maze(67, 71),
        "cgps": fieldgen.cgps_like(128, 192, 7),
        "tiny_3x3": fieldgen.uniform(3, 3, 8),
        "thin": fieldgen.uniform(3, 200, 9),
        "tall": fieldgen.uniform(150, 4, 10),
    }


FIELDS = _fields()


def _seeds_for(oracle, name, img):
    s = oracle.find_local_minima(img)
    if name == "maze":
        s = np.array([[1, 1]], np.uint64)
    extra = []
    if img.shape[0] > 4 and img.shape[1] > 6:       # a border seed, a corner seed, a duplicate position
        extra = [[0, 3], [img.shape[0] - 1, 5], [0, 0]]
    s = np.concatenate([s, np.array(extra, np.uint64).reshape(-1, 2), s[:1]])
    return s


@pytest.mark.parametrize("name", list(FIELDS))
def test_find_local_minima(ws, oracle, name):
    img = FIELDS[name]
    t = ws.TransformBuilder.default().build_segmenting()
    got = t.find_local_minima(img)
    exp = oracle.find_local_minima(img)
    assert got.shape == exp.shape
    assert np.array_equal(got, exp)


def test_find_local_minima_strided_views(ws, oracle):
    base = fieldgen.uniform(90, 120, 11)
    t = ws.TransformBuilder.default().build_merging()
    for view in (base.T, base[::2, ::3], base[::-1, :], base[5:60, 7:99]):
        assert np.array_equal(t.find_local_minima(view), oracle.find_local_minima(np.ascontiguousarray(view)))


def test_find_local_minima_large_rowmajor(ws, oracle):
    img = fieldgen.uniform(1030, 2051, 12)       # several 1024-column chunks per row
    t = ws.TransformBuilder.default().build_segmenting()
    assert np.array_equal(t.find_local_minima(img), oracle.find_local_minima(img))


@pytest.mark.parametrize("lmax", [254, 100, 1])
@pytest.mark.parametrize("name", list(FIELDS))
def test_segmenting_labels_and_levels(ws, oracle, name, lmax):
    img = FIELDS[name]
    seeds = _seeds_for(oracle, name, img)
    t = ws.TransformBuilder.default().set_max_water_lvl(lmax).build_segmenting()
    lab, lvl = t.transform_compact(img, seeds)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds, lmax)
    assert np.array_equal(lvl, ref.lvl)
    assert np.array_equal(lab.astype(np.uint64), ref.final)
    assert np.array_equal(t.transform(img, seeds), ref.final)
    assert oracle.check_valid_segmentation(lab, ref.lvl, ref.hop, seeds) is None


@pytest.mark.parametrize("name", ["uniform_odd", "smooth_128", "plateaus", "maze", "obstacles"])
def test_arrival_times_match_reference_loop(ws, oracle, name):
    """(level, flood pass) of every pixel == the counters of the reference's nested loops."""
    img = FIELDS[name]
    seeds = _seeds_for(oracle, name, img)
    ctx = ws.default_context()
    plan = ws.Plan(ctx, 1, img.shape[0], img.shape[1])
    d_img = ctx.dev_malloc(img.size)
    s32 = np.ascontiguousarray(seeds, dtype=np.uint32)
    d_seeds = ctx.dev_malloc(max(1, s32.nbytes))
    d_off = ctx.dev_malloc(8)
    try:
        ctx.h2d(d_img, img)
        ctx.h2d(d_seeds, s32)
        ctx.h2d(d_off, np.array([0, len(s32)], np.uint32))
        plan.run(0, 254, d_img, d_seeds, d_off, len(s32))
        T = ctx.d2h(plan.arrival_times_ptr, img.shape, np.uint32)
        stats = plan.stats()
    finally:
        for p in (d_img, d_seeds, d_off):
            ctx.dev_free(p)
        plan.close()
    ref = oracle.transform(oracle.SEGMENTING, img, seeds)
    exp = (ref.lvl.astype(np.uint32) << 24) | ref.hop
    exp[ref.lvl == 255] = 0xFF000000
    assert np.array_equal(np.minimum(T, 0xFF000000), exp)
    assert stats["tile_activations"] >= 1 and stats["kernel_launches"] >= 5


@pytest.mark.parametrize("name", ["uniform_64", "smooth_wide", "obstacles", "tiny_3x3"])
def test_segmenting_history_every_level(ws, oracle, name):
    img = FIELDS[name]
    seeds = _seeds_for(oracle, name, img)
    t = ws.TransformBuilder.default().build_segmenting()
    hist = t.transform_history(img, seeds)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds, want_history=True)
    assert [l for l, _ in hist] == list(range(255))
    for (l, snap), exp in zip(hist, ref.history):
        assert np.array_equal(snap, exp), f"level {l}"


@pytest.mark.parametrize("name", ["uniform_64", "smooth_128", "plateaus"])
def test_segmenting_to_list_exact(ws, oracle, name):
    img = FIELDS[name]
    seeds = _seeds_for(oracle, name, img)
    t = ws.TransformBuilder.default().set_max_water_lvl(200).build_segmenting()
    lst = t.transform_to_list(img, seeds)
    ref = oracle.transform(oracle.SEGMENTING, img, seeds, 200, want_sizes=True)
    assert len(lst) == 201
    for (l, sizes), exp in zip(lst, ref.sizes):
        assert sizes.shape == (img.size + 1,)              # find_lake_sizes: H*W + 1 (lib.rs:630)
        assert np.array_equal(sizes, exp), f"level {l}"


@pytest.mark.parametrize("name", ["uniform_64", "uniform_odd", "smooth_128", "obstacles", "cgps", "thin"])
def test_merging_history_partition(ws, oracle, name):
    img = FIELDS[name]
    seeds = _seeds_for(oracle, name, img)
    t = ws.TransformBuilder.default().build_merging()
    hist = t.transform_history(img, seeds)
    ref = oracle.transform(oracle.MERGING, img, seeds, want_history=True, fast_closure=True)
    lakes, unc = t.lake_counts(img, seeds)
    for (l, snap), exp in zip(hist, ref.history):
        assert oracle.same_partition(snap, exp), f"level {l}"
        assert lakes[l] == np.unique(exp[exp != 0]).size, f"lake count, level {l}"
        assert unc[l] == (exp == 0).sum()


def test_merging_literal_closure_small(ws, oracle):
    """Against the literal make_colour_map restatement (not the union-find shortcut)."""
    img = fieldgen.uniform(40, 48, 21)
    seeds = oracle.find_local_minima(img)
    t = ws.TransformBuilder.default().build_merging()
    hist = t.transform_history(img, seeds)
    ref = oracle.transform(oracle.MERGING, img, seeds, want_history=True, fast_closure=False)
    for (l, snap), exp in zip(hist, ref.history):
        assert oracle.same_partition(snap, exp), f"level {l}"


@pytest.mark.parametrize("name", ["uniform_64", "smooth_128"])
def test_merging_to_list_multiset(ws, oracle, name):
    img = FIELDS[name]
    seeds = _seeds_for(oracle, name, img)
    t = ws.TransformBuilder.default().build_merging()
    lst = t.transform_to_list(img, seeds)
    ref = oracle.transform(oracle.MERGING, img, seeds, want_sizes=True)
    for (l, sizes), exp in zip(lst, ref.sizes):
        assert sizes.shape == exp.shape
        assert sizes[0] == exp[0], f"uncoloured, level {l}"
        assert np.array_equal(np.sort(sizes[1:]), np.sort(exp[1:])), f"lake sizes, level {l}"


def test_merging_transform_is_constant(ws, oracle):
    """lib.rs:1524-1536"""
    img = fieldgen.uniform(20, 30, 3)
    t = ws.TransformBuilder.default().enable_edge_correction().build_merging()
    out = t.transform(img, [(5, 5)])
    assert np.array_equal(out, oracle.merging_transform_const(20, 30))


def test_with_hook_order_and_contents(ws, oracle):
    img = FIELDS["smooth_wide"]
    seeds = _seeds_for(oracle, "smooth_wide", img)
    ref = oracle.transform(oracle.MERGING, img, seeds, 60, want_history=True)
    seen = []

    def hook(ctx):
        assert ctx.max_water_level == 60
        assert np.array_equal(ctx.image, img)
        assert ctx.seeds[0] == (1, (int(seeds[0][0]), int(seeds[0][1])))
        seen.append(ctx.water_level)
        return int((ctx.colours != 0).sum()), np.unique(ctx.colours).size

    t = ws.TransformBuilder.new().set_max_water_lvl(60).set_wlvl_hook(hook).build_merging()
    res = t.transform_with_hook(img, seeds)
    assert seen == list(range(61))
    for l, (ncol, nuniq) in enumerate(res):
        assert ncol == (ref.history[l] != 0).sum()
        assert nuniq == np.unique(ref.history[l]).size
    # no hook configured -> empty Vec (lib.rs:1510-1521)
    assert ws.TransformBuilder.default().build_segmenting().transform_with_hook(img, seeds) == []


@pytest.mark.parametrize("kind", ["seg", "merge"])
def test_edge_correction(ws, oracle, kind):
    """Padded shape, unshifted seeds (lib.rs:1330-1367)."""
    img = fieldgen.uniform(50, 70, 31)
    seeds = oracle.find_local_minima(img)
    b = ws.TransformBuilder.default().enable_edge_correction()
    if kind == "seg":
        out = b.build_segmenting().transform(img, seeds)
        ref = oracle.transform(oracle.SEGMENTING, img, seeds, edge_correction=True)
        assert out.shape == (52, 72)
        assert np.array_equal(out, ref.final)
    else:
        hist = b.set_max_water_lvl(120).build_merging().transform_history(img, seeds)
        ref = oracle.transform(oracle.MERGING, img, seeds, 120, edge_correction=True, want_history=True)
        for (l, snap), exp in zip(hist, ref.history):
            assert snap.shape == (52, 72)
            assert oracle.same_partition(snap, exp), f"level {l}"


def test_strided_input(ws, oracle):
    base = fieldgen.smooth(100, 140, 3.0, 41)
    for view in (base.T, base[::2, 1::2], base[:, ::-1]):
        dense = np.ascontiguousarray(view)
        seeds = oracle.find_local_minima(dense)
        t = ws.TransformBuilder.default().build_segmenting()
        assert np.array_equal(t.transform(view, seeds), oracle.transform(oracle.SEGMENTING, dense, seeds).final)


def test_no_seeds_and_errors(ws):
    img = fieldgen.uniform(33, 47, 51)
    t = ws.TransformBuilder.default().build_segmenting()
    assert not t.transform(img, []).any()
    lakes, unc = ws.TransformBuilder.default().build_merging().lake_counts(img, [])
    assert not lakes.any() and (unc == img.size).all()
    with pytest.raises(ws.WatershedError) as e:
        t.transform(img, [(33, 0)])                      # the reference panics (lib.rs:1676)
    assert e.value.status == 4
    with pytest.raises(ws.WatershedError):
        t.transform(img, [(0, 47)])
    # still usable afterwards
    assert t.transform(img, [(5, 5)])[5, 5] == 1


def test_batch_equals_per_slice(ws, oracle):
    imgs = np.stack([fieldgen.uniform(96, 80, 60 + i) if i % 2 else fieldgen.smooth(96, 80, 3.0, 60 + i)
                     for i in range(5)])
    t = ws.TransformBuilder.default().build_merging()
    seeds, off = t.find_local_minima_batch(imgs)
    for i in range(5):
        assert np.array_equal(seeds[int(off[i]):int(off[i + 1])], oracle.find_local_minima(imgs[i]))
    labels, counts = t.transform_batch(imgs, seeds, off, want_labels=True, want_lake_counts=True)
    for i in range(5):
        s = seeds[int(off[i]):int(off[i + 1])]
        ref = oracle.transform(oracle.SEGMENTING, imgs[i], s)
        assert np.array_equal(labels[i], ref.final), f"slice {i}"
        refm = oracle.transform(oracle.MERGING, imgs[i], s, want_history=True)
        exp = [np.unique(h[h != 0]).size for h in refm.history]
        assert counts[i].tolist() == exp, f"slice {i}"


def test_medium_fields_against_oracle(ws, oracle):
    """512^2 (BASELINE.json configs[0] and [1]) and a 1024x768 smooth field."""
    for img in (fieldgen.uniform(512, 512, 0), fieldgen.smooth(768, 1024, 8.0, 0)):
        seg = ws.TransformBuilder.default().build_segmenting()
        seeds = seg.find_local_minima(img)
        assert np.array_equal(seeds, oracle.find_local_minima(img))
        lab, lvl = seg.transform_compact(img, seeds)
        ref = oracle.transform(oracle.SEGMENTING, img, seeds)
        assert np.array_equal(lvl, ref.lvl)
        assert np.array_equal(lab.astype(np.uint64), ref.final)
        lakes, unc = ws.TransformBuilder.default().build_merging().lake_counts(img, seeds)
        refm = oracle.transform(oracle.MERGING, img, seeds, want_sizes=True)
        assert lakes.tolist() == [int((s[1:] != 0).sum()) for s in refm.sizes]
        assert unc.tolist() == [int(s[0]) for s in refm.sizes]


@pytest.mark.parametrize("dtype", ["float64", "float32", "int32", "uint16", "int16", "uint8", "int64"])
def test_pre_processor_bit_exact(ws, oracle, dtype):
    """lib.rs:1081-1173 incl. its quirks (0 and subnormals -> 255, +inf -> 0, folds from zero)."""
    rng = np.random.default_rng(11)
    t = ws.TransformBuilder.default().build_segmenting()
    if dtype.startswith("float"):
        x = (rng.normal(size=(301, 517)) * 37.0).astype(dtype)
        x[rng.random(x.shape) < 0.03] = np.nan
        x[rng.random(x.shape) < 0.01] = 0.0
        x[1, 2], x[3, 4], x[5, 6] = np.inf, -np.inf, np.finfo(dtype).tiny / 4
        cube = (rng.poisson(0.85, size=(7, 40, 50)).astype(dtype))        # tests/integration.rs:189 style
    else:
        info = np.iinfo(dtype)
        x = rng.integers(max(info.min, -5000), min(info.max, 5000), size=(301, 517)).astype(dtype)
        cube = rng.integers(0, 100, size=(7, 40, 50)).astype(dtype)
    for arr in (x, cube, x[:1, :1], np.abs(x) + 1):
        assert np.array_equal(t.pre_processor(arr), oracle.pre_processor(arr))
    assert np.array_equal(t.pre_processor_with_max(x, 127), oracle.pre_processor(x, 127))
    for bad in (0, 255):
        with pytest.raises(AssertionError):
            t.pre_processor_with_max(x, bad)


def test_pre_processor_feeds_transform(ws, oracle):
    """README-style pipeline on real-valued data: pre_processor -> find_local_minima -> transform."""
    rng = np.random.default_rng(5)
    raw = rng.poisson(0.85, size=(200, 220)).astype(np.float64) + rng.normal(size=(200, 220))
    t = ws.TransformBuilder.default().build_segmenting()
    img = t.pre_processor(raw)
    assert np.array_equal(img, oracle.pre_processor(raw))
    seeds = t.find_local_minima(img)
    assert np.array_equal(t.transform(img, seeds), oracle.transform(oracle.SEGMENTING, img, seeds).final)
