"""Per-level pictures of a transform: the reference's `plots` feature (lib.rs:698-834, 1472-1487, 1758-1773).

    TransformBuilder.default().set_plot_folder(path).set_plot_colour_map(plotting.magma).build_merging()

writes `ws_lvl{water_level}.png` for every water level: the label image of that level (without the edge-correction
padding, lib.rs:1476-1481) through a colour map.  Visualisation only -- not part of the hot path; the label
images come from the engine's per-level hook (ws_transform_with_hook), the colouring and the PNG encoding are
plain numpy / zlib on the host.

Colour maps have the reference's signature in vector form: `f(count, min, max) -> (n, 3) uint8`, with its
arithmetic: values <= min are black (NAN_COL), the others index 256 entries by `(255 * count + min) / max`
truncated (lib.rs:757-758).  The 256-entry viridis / magma / plasma / inferno tables are matplotlib's (CC0), stored
in data/colour_maps.npz by scripts/make_colour_maps.py (the reference ships them as a 1000-line source file).
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Callable

import numpy as np

ColourMap = Callable[[np.ndarray, float, float], np.ndarray]


def _index(count: np.ndarray, mn: float, mx: float) -> np.ndarray:
    """`((255.0 * count + min) / max) as usize`, clipped to the table (lib.rs:757)."""
    c = np.asarray(count, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (255.0 * c + float(mn)) / float(mx)
    return np.clip(np.nan_to_num(g, nan=0.0, posinf=255.0, neginf=0.0), 0, 255).astype(np.int64)


def grey_scale(count, mn, mx) -> np.ndarray:
    """lib.rs:748-760"""
    g = _index(count, mn, mx).astype(np.uint8)
    rgb = np.stack([g, g, g], axis=-1)
    rgb[np.asarray(count) <= mn] = 0
    return rgb


_TABLES = {}


def _table(name: str) -> np.ndarray:
    if not _TABLES:
        with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "colour_maps.npz")) as z:
            for k in z.files:
                _TABLES[k] = z[k]
    return _TABLES[name]


def _mapped(name: str) -> ColourMap:
    def f(count, mn, mx) -> np.ndarray:
        rgb = _table(name)[_index(count, mn, mx)]
        rgb = rgb.copy()
        rgb[np.asarray(count) <= mn] = 0                              # NAN_COL = BLACK (lib.rs:706)
        return rgb
    f.__name__ = name
    f.__doc__ = f"lib.rs:762-834 ({name})"
    return f


viridis, magma, plasma, inferno = (_mapped(n) for n in ("viridis", "magma", "plasma", "inferno"))


def write_png(path: str, rgb: np.ndarray) -> None:
    """8-bit RGB PNG, [height][width][3] (zlib only)."""
    a = np.ascontiguousarray(rgb, dtype=np.uint8)
    h, w, _ = a.shape
    raw = np.concatenate([np.zeros((h, 1), np.uint8), a.reshape(h, w * 3)], axis=1).tobytes()   # filter 0 per row

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def read_png(path: str) -> np.ndarray:
    """Inverse of write_png (filter type 0 only): used by the tests."""
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    assert not raw[:, 0].any()
    return raw[:, 1:].reshape(h, w, 3).copy()


def plot_slice(slice_, file_name, color_map: ColourMap = viridis) -> None:
    """lib.rs:713-745.  min / max are folded from the type's default (0) like the reference's; the picture is
    `shape[0]` wide and `shape[1]` high with element (x, y) at abscissa x, ordinate y of a cartesian chart (y up)."""
    a = np.asarray(slice_)
    mn = min(0, a.min()) if a.size else 0
    mx = max(0, a.max()) if a.size else 0
    rgb = color_map(a.reshape(-1), mn, mx).reshape(a.shape[0], a.shape[1], 3)
    write_png(str(file_name), np.ascontiguousarray(rgb.transpose(1, 0, 2)[::-1]))


def level_file(folder, water_level: int) -> str:
    return os.path.join(str(folder), f"ws_lvl{int(water_level)}.png")               # lib.rs:1482
