"""FITS cubes through the engine (rustronomy-watershed_b200/fits_cube.py): the reader / writer on the CPU, and on
the GPU the whole chain file bytes -> pre_processor -> find_local_minima -> transform against the oracle's
pre_processor + transform per channel (the reference's real-data pipeline, tests/integration.rs:72-94, 267-300)."""
import importlib

import numpy as np
import pytest

import fieldgen
from wsb200_loader import load


def _mod():
    load()
    return importlib.import_module("rustronomy_watershed_b200.fits_cube")


def _synthetic_cube(nchan=5, rows=150, cols=210, seed=0) -> np.ndarray:
    """CGPS-like channels: smooth emission + noise, NaN outside an elliptical mosaic, one channel all NaN
    (tests/integration.rs:344-356 picks such a NaN-heavy channel), one with infinities and exact zeros."""
    rng = np.random.default_rng(seed)
    cube = np.empty((nchan, rows, cols), np.float32)
    y = (np.arange(rows, dtype=np.float32)[:, None] - rows / 2) / (0.48 * rows)
    x = (np.arange(cols, dtype=np.float32)[None, :] - cols / 2) / (0.46 * cols)
    outside = (x * x + y * y) > 1.0
    for k in range(nchan):
        base = fieldgen.smooth(rows, cols, 4.0, seed + k).astype(np.float32)
        ch = base * np.float32(0.37) - np.float32(20.0 * (k % 2)) + rng.standard_normal((rows, cols), dtype=np.float32)
        ch[outside] = np.nan
        cube[k] = ch
    cube[2] = np.nan
    cube[3, 40:44, 50:60] = np.inf
    cube[3, 70:73, 80:90] = -np.inf
    cube[3, 90:95, 100:120] = 0.0
    return cube


def test_fits_round_trip_and_header_parsing(tmp_path):
    fc = _mod()
    rng = np.random.default_rng(1)
    for dt in ("uint8", "int16", "int32", "int64", "float32", "float64"):
        a = (rng.standard_normal((3, 7, 11)) * 100).astype(dt)
        p = str(tmp_path / f"c_{dt}.fits")
        fc.write_fits_cube(p, a, {"OBJECT": "it's a 'cube'", "BUNIT": "K", "CRVAL3": -1.5e3, "EXTEND": True})
        c = fc.read_fits_cube(p)
        assert c.data.shape == (3, 7, 11)
        assert np.array_equal(c.physical, a)
        assert c.header["OBJECT"] == "it's a 'cube'" and c.header["BUNIT"] == "K"
        assert c.header["CRVAL3"] == -1500.0 and c.header["EXTEND"] is True and c.header["NAXIS3"] == 3
        assert c.data.dtype.byteorder in (">", "|")           # the file's bytes, not a converted copy
    # BSCALE / BZERO (the unsigned-16 convention) and a degenerate fourth axis
    raw = rng.integers(-32768, 32767, (2, 5, 6)).astype(np.int16)
    p = str(tmp_path / "scaled.fits")
    fc.write_fits_cube(p, raw, {"BSCALE": 1.0, "BZERO": 32768.0})
    c = fc.read_fits_cube(p)
    assert np.array_equal(c.physical, raw.astype(np.float64) + 32768.0)
    txt = open(p, "rb").read()
    patched = txt.replace(b"NAXIS   =                    3", b"NAXIS   =                    4", 1)
    patched = patched.replace(b"END" + b" " * 77, ("NAXIS4  = " + "1".rjust(20)).ljust(80).encode() + b"END" + b" " * 77, 1)
    p4 = str(tmp_path / "four.fits")
    open(p4, "wb").write(patched)
    assert fc.read_fits_cube(p4).data.shape == (2, 5, 6)
    with pytest.raises(ValueError):
        open(str(tmp_path / "bad.fits"), "wb").write(b"XTENSION" + b" " * 2872)
        fc.read_fits_cube(str(tmp_path / "bad.fits"))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_cube_pipeline_matches_oracle_per_channel(tmp_path, oracle, dtype):
    fc = _mod()
    cube = _synthetic_cube().astype(dtype)
    path = str(tmp_path / "cube.fits")
    fc.write_fits_cube(path, cube)
    f = fc.read_fits_cube(path)
    res_m = fc.watershed_cube(f, kind=1, batch=2, want_quantised=True)      # merging: lakes per level
    res_s = fc.watershed_cube(f, kind=0, batch=5, want_labels=True)         # segmenting: final labels
    for k in range(cube.shape[0]):
        q = oracle.pre_processor(cube[k])
        assert np.array_equal(res_m.quantised[k], q), f"pre_processor, channel {k}"
        seeds = oracle.find_local_minima(q)
        assert np.array_equal(res_m.seeds[k].astype(np.uint64), seeds), f"seeds, channel {k}"
        assert np.array_equal(res_s.seeds[k].astype(np.uint64), seeds)
        lakes = []
        oracle.transform(oracle.MERGING, q, seeds, hook=lambda l, c: lakes.append(np.unique(c[c != 0]).size))
        assert np.array_equal(res_m.lake_counts[k], np.array(lakes, np.uint64)), f"lakes per level, channel {k}"
        seg = oracle.transform(oracle.SEGMENTING, q, seeds)
        assert np.array_equal(res_s.labels[k].astype(np.uint64), seg.final), f"labels, channel {k}"
    assert len(res_m.seeds[2]) == 0 and not res_m.lake_counts[2].any()       # the all-NaN channel floods nothing


@pytest.mark.gpu
def test_cube_pipeline_integer_cube_and_native_arrays(tmp_path, oracle):
    fc = _mod()
    rng = np.random.default_rng(5)
    raw = (fieldgen.smooth(90, 130, 3.0, 9).astype(np.int16) * 50 - 3000)[None].repeat(3, 0)
    raw[1] += rng.integers(-200, 200, raw[1].shape).astype(np.int16)
    p = str(tmp_path / "int.fits")
    fc.write_fits_cube(p, raw)
    a = fc.watershed_cube(fc.read_fits_cube(p), kind=1, want_quantised=True)   # BITPIX 16, swapped on the device
    b = fc.watershed_cube(raw, kind=1, want_quantised=True)                      # the same values, native array
    assert np.array_equal(a.quantised, b.quantised) and np.array_equal(a.lake_counts, b.lake_counts)
    for k in range(3):
        assert np.array_equal(a.quantised[k], oracle.pre_processor(raw[k]))
    fc.write_fits_cube(p, raw, {"BZERO": 32768.0, "BSCALE": 1.0})
    c = fc.watershed_cube(fc.read_fits_cube(p), kind=1, want_quantised=True, channels=[0, 2])
    for i, k in enumerate((0, 2)):
        assert np.array_equal(c.quantised[i], oracle.pre_processor(raw[k].astype(np.float64) + 32768.0))
