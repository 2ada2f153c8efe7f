"""tests/c_abi_smoke.c: a plain C11 program that drives libws_b200.so through include/ws_b200.h the way the FFI of
the reference's language would (stack structs, caller buffers, callbacks).  Without a GPU it must compile and
link warning-free; on the GPU box it runs."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "rustronomy-watershed_b200")


def _build(tmp_path) -> str:
    import wsb200_loader
    wsb200_loader.build()
    exe = str(tmp_path / "c_abi_smoke")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi_smoke.c"), "-L", LIBDIR, "-lws_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_program_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_c_program_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_smoke: OK" in r.stdout
