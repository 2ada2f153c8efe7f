"""ctypes front-end of the CPU oracle (oracle/ws_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, by __graft_entry__.smoke() and by
the cpu_baseline / --impl reference legs of bench.py.  The product package
(rustronomy-watershed_b200/) never imports this module.

Parity pin: the reference's seven in-file unit tests (tests/golden/); whole
transforms are unpinned by the reference's own tests (see ws_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libws_oracle.so")

TIE_FIRST, TIE_RANDOM, TIE_LAST = 0, 1, 2
SEGMENTING, MERGING = 0, 1
UNCOLOURED, NORMAL_MAX, ALWAYS_FILL, NEVER_FILL = 0, 254, 0, 255


def build(force: bool = False) -> str:
    """Compile the oracle with the system gcc (OpenMP if libgomp is usable)."""
    src = os.path.join(_HERE, "ws_oracle.c")
    hdr = os.path.join(_HERE, "ws_oracle.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    base = ["-O3", "-fPIC", "-shared", "-std=c11", "-o", _SO, src]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc, *omp, *base], check=True, capture_output=True)
                return _SO
            except (OSError, subprocess.CalledProcessError) as e:  # try next
                last = e
    raise RuntimeError(f"could not build the oracle: {last}")


_lib = None
_HOOK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p,
                    C.c_size_t, C.c_size_t)


class _Stats(C.Structure):
    _fields_ = [("flood_passes", C.c_uint64), ("max_passes_lvl", C.c_uint64),
                ("contested_px", C.c_uint64), ("merge_pairs", C.c_uint64)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_find_local_minima.restype = C.c_size_t
        L.orc_find_local_minima.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_find_flooded_px.restype = C.c_size_t
        L.orc_find_flooded_px.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint8,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        for f in (L.orc_merge_eq, L.orc_sort_by_small_big, L.orc_sort_by_big_small):
            f.restype = C.c_int
            f.argtypes = [C.c_uint64] * 4
        L.orc_find_merge.restype = C.c_size_t
        L.orc_find_merge.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_make_colour_map.restype = None
        L.orc_make_colour_map.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_recolour.restype = None
        L.orc_recolour.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_find_lake_sizes.restype = None
        L.orc_find_lake_sizes.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_transform_with_hook.restype = C.c_int
        L.orc_transform_with_hook.argtypes = [
            C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint8,
            C.c_int, C.c_int, C.c_uint64, C.c_int, _HOOK, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p]
        L.orc_merging_transform_const.restype = None
        L.orc_merging_transform_const.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p]
        for f in (L.orc_pre_processor_f64, L.orc_pre_processor_f32, L.orc_pre_processor_i64):
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_size_t, C.c_uint8, C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _u8(img) -> np.ndarray:
    a = np.ascontiguousarray(img, dtype=np.uint8)
    assert a.ndim == 2
    return a


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(n: int) -> None:
    lib().orc_set_num_threads(int(n))


def find_local_minima(img) -> np.ndarray:
    """lib.rs:1178-1197 -> (n, 2) uint64 array of (row, col), row-major order."""
    img = _u8(img)
    cap = max(1, img.size)
    out = np.empty((cap, 2), dtype=np.uint64)
    n = lib().orc_find_local_minima(_p(img), img.shape[0], img.shape[1], _p(out), cap)
    return out[:n].copy()


def find_flooded_px(img, colours, lvl: int, tie: int = TIE_FIRST, rng_seed: int = 0):
    """lib.rs:196-257 -> (flat_idx[n], colour[n])."""
    img, colours = _u8(img), _u64(colours)
    idx = np.empty(max(1, img.size), dtype=np.uint64)
    col = np.empty(max(1, img.size), dtype=np.uint64)
    rng = np.array([rng_seed], dtype=np.uint64)
    n = lib().orc_find_flooded_px(_p(img), _p(colours), img.shape[0], img.shape[1], lvl, tie,
                                  _p(rng), _p(idx), _p(col))
    return idx[:n].copy(), col[:n].copy()


def merge_eq(a, b) -> bool:
    return bool(lib().orc_merge_eq(a[0], a[1], b[0], b[1]))


def sort_by_small_big(a, b) -> int:
    return lib().orc_sort_by_small_big(a[0], a[1], b[0], b[1])


def sort_by_big_small(a, b) -> int:
    return lib().orc_sort_by_big_small(a[0], a[1], b[0], b[1])


def find_merge(colours) -> np.ndarray:
    """lib.rs:393-445 -> (n, 2) uint64 unique (small, big) pairs, sorted."""
    colours = _u64(colours)
    cap = max(1, 4 * colours.size)
    out = np.empty((cap, 2), dtype=np.uint64)
    n = lib().orc_find_merge(_p(colours), colours.shape[0], colours.shape[1], _p(out), cap)
    return out[:n].copy()


def make_colour_map(base_map, pairs) -> np.ndarray:
    """lib.rs:467-542, literal.  Returns the updated map."""
    m = _u64(base_map).copy()
    pr = _u64(np.asarray(pairs, dtype=np.uint64).reshape(-1, 2))
    lib().orc_make_colour_map(_p(m), m.size, _p(pr), pr.shape[0])
    return m


def recolour(canvas, colour_map) -> np.ndarray:
    c = _u64(canvas).copy()
    cm = _u64(colour_map)
    lib().orc_recolour(_p(c), c.size, _p(cm))
    return c


def find_lake_sizes(colours) -> np.ndarray:
    c = _u64(colours)
    out = np.empty(c.size + 1, dtype=np.uint64)
    lib().orc_find_lake_sizes(_p(c), c.size, _p(out))
    return out


def merging_transform_const(rows: int, cols: int) -> np.ndarray:
    out = np.empty((rows, cols), dtype=np.uint64)
    lib().orc_merging_transform_const(rows, cols, _p(out))
    return out


def pre_processor(img, max_value: int = NORMAL_MAX) -> np.ndarray:
    """lib.rs:1081-1173 (pre_processor = pre_processor_with_max::<NORMAL_MAX>)."""
    a = np.ascontiguousarray(img)
    if a.dtype == np.float64:
        fn, b = lib().orc_pre_processor_f64, a
    elif a.dtype == np.float32:
        fn, b = lib().orc_pre_processor_f32, a
    elif np.issubdtype(a.dtype, np.integer):
        fn, b = lib().orc_pre_processor_i64, np.ascontiguousarray(a, dtype=np.int64)
    else:
        raise TypeError(a.dtype)
    out = np.empty(a.shape, dtype=np.uint8)
    rc = fn(_p(b), b.size, max_value, _p(out))
    if rc == -1:
        raise AssertionError("MAX must satisfy ALWAYS_FILL < MAX < NEVER_FILL (lib.rs:1143-1144)")
    if rc != 0:
        raise ValueError("to_u8() failed (the reference would panic on unwrap, lib.rs:1164)")
    return out


@dataclass
class Result:
    final: np.ndarray                       # labels after the last level
    lvl: np.ndarray                         # uint8, level of colouring, 255 = never
    hop: np.ndarray                         # uint32, flood iteration inside the level
    stats: dict
    history: List[np.ndarray] = field(default_factory=list)   # per level label images
    sizes: List[np.ndarray] = field(default_factory=list)     # per level find_lake_sizes


def transform(kind: int, img, seeds_rc, max_water_level: int = NORMAL_MAX, edge_correction: bool = False,
              tie: int = TIE_FIRST, rng_seed: int = 0, fast_closure: bool = True,
              want_history: bool = False, want_sizes: bool = False,
              hook: Optional[Callable[[int, np.ndarray], None]] = None) -> Result:
    """The reference's transform_with_hook (lib.rs:1328-1522 / 1638-1808) with
    the built-in hooks of transform_history (1545/1831) and transform_to_list
    (1557/1843) selectable.  Raises IndexError on an out-of-bounds seed."""
    img = _u8(img)
    seeds = _u64(np.asarray(seeds_rc, dtype=np.uint64).reshape(-1, 2))
    pad = 2 if edge_correction else 0
    shape = (img.shape[0] + pad, img.shape[1] + pad)
    final = np.zeros(shape, dtype=np.uint64)
    lvl = np.empty(shape, dtype=np.uint8)
    hop = np.empty(shape, dtype=np.uint32)
    st = _Stats()
    history: List[np.ndarray] = []
    sizes: List[np.ndarray] = []

    def _cb(_user, wl, _mx, _img, colp, r, c):
        arr = np.ctypeslib.as_array(C.cast(colp, C.POINTER(C.c_uint64)), shape=(r, c))
        if want_history:
            history.append(arr.copy())
        if want_sizes:
            sizes.append(find_lake_sizes(arr))
        if hook is not None:
            hook(int(wl), arr)

    need_cb = want_history or want_sizes or hook is not None
    cb = _HOOK(_cb) if need_cb else C.cast(None, _HOOK)
    rc = lib().orc_transform_with_hook(kind, _p(img), img.shape[0], img.shape[1], _p(seeds), seeds.shape[0],
                                       max_water_level, int(edge_correction), tie, rng_seed,
                                       int(fast_closure), cb, None, _p(final), _p(lvl), _p(hop),
                                       C.byref(st))
    if rc != 0:
        raise IndexError("seed out of bounds (the reference panics: lib.rs:1366/1676)")
    stats = {k: int(getattr(st, k)) for k, _ in _Stats._fields_}
    return Result(final, lvl, hop, stats, history, sizes)


# --------------------------------------------------------------------------
# Checkers used by the parity tests
# --------------------------------------------------------------------------

def same_partition(a: np.ndarray, b: np.ndarray) -> bool:
    """True when label images a and b are equal up to a bijection of the
    non-zero labels (0 <-> 0 fixed).  This is the merging parity rule: the
    reference's representative colour is region[0] after two unstable sorts
    with inconsistent comparators (lib.rs:440-443, 539), i.e. unspecified."""
    a = np.asarray(a).ravel().astype(np.int64)
    b = np.asarray(b).ravel().astype(np.int64)
    if a.shape != b.shape:
        return False
    if not np.array_equal(a == 0, b == 0):
        return False
    pairs = np.unique(np.stack([a, b], axis=1), axis=0)
    return (np.unique(pairs[:, 0]).size == pairs.shape[0]
            and np.unique(pairs[:, 1]).size == pairs.shape[0])


def check_valid_segmentation(labels, lvl, hop, seeds_rc) -> Optional[str]:
    """Validity of a segmenting result against ANY tie-break the reference
    could draw (lib.rs:246-254): every coloured non-seed pixel carries the
    label of at least one 4-neighbour coloured strictly earlier, and THE label
    when those neighbours agree.  Returns None or a message."""
    labels = np.asarray(labels).astype(np.int64)
    lvl = np.asarray(lvl).astype(np.int64)
    hop = np.asarray(hop).astype(np.int64)
    t = lvl * (1 << 32) + hop
    t[lvl == 255] = np.iinfo(np.int64).max
    H, W = labels.shape
    seed_mask = np.zeros((H, W), bool)
    s = np.asarray(seeds_rc, dtype=np.int64).reshape(-1, 2)
    seed_mask[s[:, 0], s[:, 1]] = True
    big = np.iinfo(np.int64).max
    tp = np.pad(t, 1, constant_values=big)
    lp = np.pad(labels, 1, constant_values=0)
    ok_any = np.zeros((H, W), bool)
    all_same = np.ones((H, W), bool)
    first = np.zeros((H, W), np.int64)
    have = np.zeros((H, W), bool)
    for dr, dc in ((1, 0), (0, 1), (0, -1), (-1, 0)):
        tq = tp[1 + dr:1 + dr + H, 1 + dc:1 + dc + W]
        lq = lp[1 + dr:1 + dr + H, 1 + dc:1 + dc + W]
        pred = tq < t
        ok_any |= pred & (lq == labels)
        first = np.where(pred & ~have, lq, first)
        all_same &= ~pred | ~have | (lq == first)
        have |= pred
    coloured = (lvl != 255) & ~seed_mask
    if np.any((labels != 0) != (lvl != 255)):
        return "coloured set differs from lvl != 255"
    if np.any(coloured & ~ok_any):
        return "a pixel carries a label none of its predecessors has"
    if np.any(coloured & all_same & have & (labels != first)):
        return "an uncontested pixel carries the wrong label"
    return None
