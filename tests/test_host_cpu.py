"""CPU-side checks: the C-ABI library loads and exports every declared symbol, and the
host mirror of the reference API validates like the reference (no compute, no GPU)."""
import os
import re
import subprocess

import numpy as np
import pytest

from wsb200_loader import build, load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ws():
    build()
    return load()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ws_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ws_[a-z0-9_]+)\s*\(", hdr)) - {"ws_level_hook"})


def test_library_exports_every_declared_symbol(ws):
    names = _declared_symbols()
    assert len(names) >= 30
    lib = ws.load_library()
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/ws_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(ws._native.SIGNATURES) == names
    assert lib.ws_abi_version() == 1


def test_library_is_sm100a_only(ws):
    out = subprocess.run(["cuobjdump", "--list-elf", ws._native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_builder_validation_matches_reference(ws):
    """lib.rs:998-1004, 1024-1030, 1051-1065"""
    TB = ws.TransformBuilder
    assert TB.default().max_water_level == 254 and TB.new().edge_correction is False
    for build_fn in ("build_merging", "build_segmenting"):
        for ok in (1, 127, 254):
            t = getattr(TB.default().set_max_water_lvl(ok), build_fn)()
            assert t.max_water_level == ok and t.levels == ok + 1
        with pytest.raises(ws.BuildErr) as e:
            getattr(TB.default().set_max_water_lvl(255), build_fn)()
        assert e.value.kind == "MaxToHigh" and e.value.value == 255 and "254" in str(e.value)
        with pytest.raises(ws.BuildErr) as e:
            getattr(TB.default().set_max_water_lvl(0), build_fn)()
        assert e.value.kind == "MaxToLow" and "255" in str(e.value)     # the reference's message bug
    with pytest.raises(OverflowError):
        TB.default().set_max_water_lvl(256)
    assert TB.default().enable_edge_correction().build_merging().edge_correction is True
    assert ws.prelude == ("MergingWatershed", "TransformBuilder", "Watershed", "WatershedUtils")
    assert (ws.UNCOLOURED, ws.NORMAL_MAX, ws.ALWAYS_FILL, ws.NEVER_FILL) == (0, 254, 0, 255)


def test_status_strings_and_output_shape(ws):
    import ctypes as C
    lib = ws.load_library()
    assert lib.ws_status_str(0) == b"ok"
    assert b"fallback" in lib.ws_status_str(5)
    cfg = ws._native.make_config(0, 254, True)
    r, c = C.c_size_t(), C.c_size_t()
    assert lib.ws_output_shape(C.byref(cfg), 10, 20, C.byref(r), C.byref(c)) == 0
    assert (r.value, c.value) == (12, 22)
    bad = ws._native.make_config(2, 254, False)
    assert lib.ws_config_validate(C.byref(bad)) == 1


def test_no_cpu_fallback_without_device(ws):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ws.WatershedError) as e:
        ws.Context(0)
    assert e.value.status == 5
    with pytest.raises(ws.WatershedError):
        ws.TransformBuilder.default().build_segmenting().transform(np.zeros((8, 8), np.uint8), [(1, 1)])


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rustronomy-watershed_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "ws_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the oracle on the host cores) needs no GPU: one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-crop", "96", "--size", "256"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "cpu_baseline", "e2e"):
        assert key in d
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
