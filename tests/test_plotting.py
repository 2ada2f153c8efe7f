"""The `plots` feature (lib.rs:698-834): colour maps, the PNG writer, file names.  CPU only."""
import numpy as np
import pytest

from wsb200_loader import load

ws = load()
from rustronomy_watershed_b200 import plotting as P  # noqa: E402


def test_grey_scale_follows_reference_arithmetic():
    # lib.rs:748-760: count <= min -> black, else ((255 * count + min) / max) as u8 on all three channels
    c = np.array([0, 1, 2, 5, 10], np.uint64)
    rgb = P.grey_scale(c, 0, 10)
    assert rgb.shape == (5, 3) and rgb.dtype == np.uint8
    assert rgb[:, 0].tolist() == [0, 25, 51, 127, 255]
    assert (rgb[:, 0] == rgb[:, 1]).all() and (rgb[:, 1] == rgb[:, 2]).all()


@pytest.mark.parametrize("name", ["viridis", "magma", "plasma", "inferno"])
def test_tables(name):
    f = getattr(P, name)
    ramp = f(np.arange(1, 256, dtype=np.uint64), 0, 255)
    assert ramp.shape == (255, 3) and ramp.dtype == np.uint8
    assert len(np.unique(ramp, axis=0)) > 200                   # a real colour table, not a constant
    assert (f(np.array([0], np.uint64), 0, 255) == 0).all()     # NAN_COL
    lum = ramp.astype(np.float64) @ np.array([0.2126, 0.7152, 0.0722])
    assert lum[-1] > lum[0] + 100                               # all four run dark -> bright


def test_known_table_ends():
    # matplotlib: viridis starts at (68, 1, 84) and ends at (253, 231, 37); magma ends near white-yellow
    assert P.viridis(np.array([255], np.uint64), 0, 255)[0].tolist() == [253, 231, 36] or \
        P.viridis(np.array([255], np.uint64), 0, 255)[0].tolist() == [253, 231, 37]
    lo = P.viridis(np.array([1], np.uint64), 0, 255 * 255)[0].tolist()      # index 0 without being <= min
    assert lo == [68, 1, 84]


def test_png_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    P.write_png(str(tmp_path / "a.png"), img)
    assert (P.read_png(str(tmp_path / "a.png")) == img).all()
    try:                                                        # an independent decoder, when the image has one
        import cv2
        got = cv2.imread(str(tmp_path / "a.png"), cv2.IMREAD_COLOR)[:, :, ::-1]
        assert (got == img).all()
    except ImportError:
        pass


def test_plot_slice_orientation(tmp_path):
    # lib.rs:726-741: picture shape[0] wide, shape[1] high; element (x, y) at abscissa x, ordinate y (y up)
    a = np.zeros((4, 6), np.uint64)
    a[3, 0] = 7                                                 # x = 3, y = 0 -> bottom row, 4th column
    a[0, 5] = 3                                                 # x = 0, y = 5 -> top row, 1st column
    f = tmp_path / "s.png"
    P.plot_slice(a, f, P.grey_scale)
    img = P.read_png(str(f))
    assert img.shape == (6, 4, 3)
    assert img[5, 3, 0] == 255 and img[0, 0, 0] == (255 * 3) // 7
    assert int((img[:, :, 0] > 0).sum()) == 2


def test_level_file(tmp_path):
    assert P.level_file(tmp_path, 17).endswith("ws_lvl17.png")  # lib.rs:1482


def test_builder_options():
    b = ws.TransformBuilder.default().set_plot_folder("/tmp/x").set_plot_colour_map(P.magma)
    assert b.plot_path == "/tmp/x" and b.plot_colour_map is P.magma
    assert ws.TransformBuilder.default().plot_path is None      # no folder, no plots (lib.rs:899-900)


@pytest.mark.gpu
@pytest.mark.parametrize("edge", [False, True])
def test_per_level_pictures(tmp_path, oracle, edge):
    """lib.rs:1472-1487 / 1758-1773: one `ws_lvl{level}.png` per water level, the level's label image without the
    edge-correction padding; the segmenting labels are bit-exact with the oracle's history, so the pictures are too."""
    import fieldgen
    img = fieldgen.smooth(60, 90, 4.0, 11)
    seeds = oracle.find_local_minima(img)
    b = ws.TransformBuilder.default().set_max_water_lvl(40).set_plot_folder(tmp_path).set_plot_colour_map(P.magma)
    if edge:
        b = b.enable_edge_correction()
    t = b.build_segmenting()
    assert t.transform_with_hook(img, seeds) == []              # pictures, and still no hook results
    ref = oracle.transform(oracle.SEGMENTING, img, seeds, 40, edge_correction=edge, want_history=True)
    for l in range(41):
        h = ref.history[l][1:-1, 1:-1] if edge else ref.history[l]
        assert h.shape == img.shape
        exp = P.magma(h.reshape(-1), 0, max(0, int(h.max()))).reshape(60, 90, 3).transpose(1, 0, 2)[::-1]
        assert (P.read_png(P.level_file(tmp_path, l)) == exp).all(), f"level {l}"
    # transform_history of a plotting transform makes the pictures as well (lib.rs:1538-1549 go through the hook)
    for f in tmp_path.iterdir():
        f.unlink()
    t2 = ws.TransformBuilder.default().set_max_water_lvl(5).set_plot_folder(tmp_path).build_merging()
    t2.transform_history(img, seeds)
    assert sorted(p.name for p in tmp_path.iterdir()) == sorted(f"ws_lvl{l}.png" for l in range(6))
