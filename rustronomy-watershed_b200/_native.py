"""ctypes binding of libws_b200.so (the C ABI of include/ws_b200.h).

There is no fallback of any kind: if the library is missing, or no sm_100
device is present, creating a Context raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WS_B200_LIB") or os.path.join(HERE, "libws_b200.so")   # override: debugging builds only

WS_OK = 0
STATUS_NAMES = {
    0: "WS_OK", 1: "WS_ERR_INVALID_ARG", 2: "WS_ERR_MAX_TOO_HIGH", 3: "WS_ERR_MAX_TOO_LOW",
    4: "WS_ERR_SEED_OOB", 5: "WS_ERR_NO_DEVICE", 6: "WS_ERR_CUDA", 7: "WS_ERR_OOM",
    8: "WS_ERR_TOO_LARGE", 9: "WS_ERR_HOP_OVERFLOW", 10: "WS_ERR_INTERNAL",
}
WS_ERR_MAX_TOO_HIGH, WS_ERR_MAX_TOO_LOW, WS_ERR_SEED_OOB, WS_ERR_NO_DEVICE = 2, 3, 4, 5
WS_SEGMENTING, WS_MERGING = 0, 1
WS_TIE_FIRST, WS_TIE_RANDOM = 0, 1
WS_SEEDS_AUTO = C.c_size_t(-1).value          # nseeds: find_local_minima(image) on the device
WS_OPT_PINNED_HOST_WIDEN = 1


class WsConfig(C.Structure):
    _fields_ = [("kind", C.c_uint8), ("max_water_level", C.c_uint8),
                ("edge_correction", C.c_uint8), ("tie_break", C.c_uint8)]


class WsImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_size_t), ("cols", C.c_size_t),
                ("row_stride", C.c_ssize_t), ("col_stride", C.c_ssize_t)]


class WsHookCtx(C.Structure):
    _fields_ = [("water_level", C.c_uint8), ("max_water_level", C.c_uint8),
                ("image", C.c_void_p), ("colours", C.c_void_p),
                ("rows", C.c_size_t), ("cols", C.c_size_t),
                ("seeds", C.c_void_p), ("nseeds", C.c_size_t)]


class WsStrip(C.Structure):
    _fields_ = [("global_rows", C.c_size_t), ("row_offset", C.c_size_t), ("halo_top", C.c_uint8),
                ("halo_bottom", C.c_uint8), ("colour_base", C.c_uint32)]


HOOK_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(WsHookCtx))

# name -> (restype, argtypes); every symbol include/ws_b200.h declares
_P = C.c_void_p
SIGNATURES = {
    "ws_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ws_ctx_destroy": (None, [_P]),
    "ws_last_error": (C.c_char_p, [_P]),
    "ws_status_str": (C.c_char_p, [C.c_int]),
    "ws_abi_version": (C.c_int, []),
    "ws_free": (None, [_P]),
    "ws_ctx_set_tie_seed": (C.c_int, [_P, C.c_uint64]),
    "ws_ctx_set_host_threads": (C.c_int, [_P, C.c_int]),
    "ws_ctx_set_option": (C.c_int, [_P, C.c_int, C.c_int]),
    "ws_config_validate": (C.c_int, [C.POINTER(WsConfig)]),
    "ws_output_shape": (C.c_int, [C.POINTER(WsConfig), C.c_size_t, C.c_size_t,
                                  C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "ws_find_local_minima": (C.c_int, [_P, C.POINTER(WsImage), C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "ws_transform": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsImage), _P, C.c_size_t, _P]),
    "ws_transform_history": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsImage), _P, C.c_size_t, _P, _P]),
    "ws_transform_to_list": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsImage), _P, C.c_size_t, _P, _P]),
    "ws_transform_with_hook": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsImage), _P, C.c_size_t,
                                         HOOK_FN, _P]),
    "ws_transform_lake_counts": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsImage), _P, C.c_size_t, _P, _P]),
    "ws_transform_lake_sizes_compact": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsImage), _P, C.c_size_t,
                                                  _P, _P]),
    "ws_transform_compact": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsImage), _P, C.c_size_t, _P, _P]),
    "ws_transform_batch": (C.c_int, [_P, C.POINTER(WsConfig), _P, C.c_size_t, C.c_size_t, C.c_size_t,
                                     _P, _P, _P, _P]),
    "ws_find_local_minima_batch": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(_P), _P]),
    "ws_pre_processor": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.c_uint8, _P]),
    "ws_dev_pre_processor": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.c_uint8, _P]),
    "ws_ctx_stream": (_P, [_P]),
    "ws_ctx_synchronize": (C.c_int, [_P]),
    "ws_plan_create": (C.c_int, [_P, C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(_P)]),
    "ws_plan_destroy": (None, [_P]),
    "ws_plan_find_local_minima": (C.c_int, [_P, _P, _P, C.c_size_t, _P, C.POINTER(C.c_size_t)]),
    "ws_plan_run": (C.c_int, [_P, C.POINTER(WsConfig), _P, _P, _P, C.c_size_t]),
    "ws_plan_labels": (_P, [_P]),
    "ws_plan_levels": (_P, [_P]),
    "ws_plan_lake_counts": (_P, [_P]),
    "ws_plan_arrival_times": (_P, [_P]),
    "ws_plan_snapshot": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_uint8, _P]),
    "ws_dev_malloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "ws_dev_free": (C.c_int, [_P, _P]),
    "ws_memcpy_h2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "ws_memcpy_d2h": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "ws_memcpy_d2d": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "ws_plan_strip_begin": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsStrip), _P, _P, C.c_size_t]),
    "ws_plan_strip_export_times": (C.c_int, [_P, _P, _P]),
    "ws_plan_strip_import_times": (C.c_int, [_P, _P, _P, C.POINTER(C.c_int)]),
    "ws_plan_strip_labels": (C.c_int, [_P]),
    "ws_plan_strip_export_labels": (C.c_int, [_P, _P, _P]),
    "ws_plan_strip_import_labels": (C.c_int, [_P, _P, _P, C.POINTER(C.c_size_t)]),
    "ws_plan_strip_edges": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_size_t),
                                      C.POINTER(C.c_uint32)]),
    "ws_plan_union_edges": (C.c_int, [_P, _P, _P, C.c_size_t, C.c_size_t, C.c_uint32, C.c_uint8]),
    "ws_plan_strip_begin_async": (C.c_int, [_P, C.POINTER(WsConfig), C.POINTER(WsStrip), _P, _P, C.c_size_t]),
    "ws_plan_strip_export_times_async": (C.c_int, [_P, _P, _P]),
    "ws_plan_strip_import_times_async": (C.c_int, [_P, _P, _P, _P]),
    "ws_plan_strip_labels_async": (C.c_int, [_P]),
    "ws_plan_strip_export_labels_async": (C.c_int, [_P, _P, _P]),
    "ws_plan_strip_import_labels_async": (C.c_int, [_P, _P, _P, _P]),
    "ws_plan_strip_labels_finish_async": (C.c_int, [_P]),
    "ws_plan_strip_check": (C.c_int, [_P]),
    "ws_strip_packet_bytes": (C.c_size_t, [C.c_size_t]),
    "ws_plan_strip_forest": (C.c_int, [_P, C.c_size_t, _P]),
    "ws_plan_forest_packets": (C.c_int, [_P, _P, C.c_size_t, C.c_size_t, C.c_uint8]),
    "ws_plan_stats": (C.c_int, [_P, C.POINTER(C.c_uint64 * 8)]),
    "ws_plan_phase_ms": (C.c_int, [_P, C.POINTER(C.c_float * 4)]),
    "ws_plan_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float * 10)]),
}

_lib = None


class WatershedError(RuntimeError):
    """A non-zero ws_status.  The Rust shim turns these into panics."""

    def __init__(self, status: int, message: str = ""):
        self.status = status
        name = STATUS_NAMES.get(status, str(status))
        super().__init__(f"{name}: {message}" if message else name)


def load_library(path: Optional[str] = None):
    """dlopen the CUDA library and bind every declared symbol (fails loudly)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build it with __graft_entry__.build() "
                          "(the engine has no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def make_config(kind: int, max_water_level: int, edge_correction: bool, tie_break: int = WS_TIE_FIRST) -> WsConfig:
    return WsConfig(kind, max_water_level, 1 if edge_correction else 0, tie_break)


def image_view(img: np.ndarray) -> WsImage:
    """ArrayView2<u8> of a numpy array, strides preserved (no copy)."""
    if img.dtype != np.uint8 or img.ndim != 2:
        raise TypeError("image must be a 2-D uint8 array")
    return WsImage(img.ctypes.data, img.shape[0], img.shape[1], img.strides[0], img.strides[1])


class AutoSeeds:
    """Stands for `seeds = find_local_minima(image)` computed on the device (WS_SEEDS_AUTO)."""
    shape = (WS_SEEDS_AUTO, 2)

    class _Null:
        data = None
    ctypes = _Null()


def seeds_array(seeds):
    """&[(usize, usize)] -> C-contiguous [n][2] uint64 (None: the engine finds the local maxima itself)."""
    if seeds is None:
        return AutoSeeds()
    a = np.asarray(seeds)
    if a.size == 0:
        return np.zeros((0, 2), dtype=np.uint64)
    if np.issubdtype(a.dtype, np.signedinteger) and (a < 0).any():
        raise OverflowError("seed coordinates are usize: negative values are not representable")
    return np.ascontiguousarray(a.reshape(-1, 2), dtype=np.uint64)


class Context:
    """Owns a ws_ctx (device, streams, workspaces).  Single-threaded use."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = _P()
        st = self.lib.ws_ctx_create(device, C.byref(h))
        if st != WS_OK:
            raise WatershedError(st, self.lib.ws_status_str(st).decode())
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.ws_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def check(self, st: int):
        if st != WS_OK:
            msg = self.lib.ws_last_error(self.handle)
            raise WatershedError(st, (msg or b"").decode() or self.lib.ws_status_str(st).decode())

    # -- plain device memory (for callers that do not bring torch / cuda-python) --
    def dev_malloc(self, nbytes: int) -> int:
        out = _P()
        self.check(self.lib.ws_dev_malloc(self.handle, nbytes, C.byref(out)))
        return int(out.value)

    def dev_free(self, ptr: int):
        self.check(self.lib.ws_dev_free(self.handle, ptr))

    def h2d(self, d_ptr: int, arr: np.ndarray):
        a = np.ascontiguousarray(arr)
        self.check(self.lib.ws_memcpy_h2d(self.handle, d_ptr, a.ctypes.data, a.nbytes))

    def d2d(self, d_dst: int, d_src: int, nbytes: int):
        self.check(self.lib.ws_memcpy_d2d(self.handle, d_dst, d_src, nbytes))

    def d2h(self, d_ptr: int, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype=dtype)
        self.check(self.lib.ws_memcpy_d2h(self.handle, out.ctypes.data, d_ptr, out.nbytes))
        return out

    @property
    def stream(self) -> int:
        """cudaStream_t of the compute stream as an integer (torch.cuda.ExternalStream)."""
        return int(self.lib.ws_ctx_stream(self.handle) or 0)

    def synchronize(self):
        self.check(self.lib.ws_ctx_synchronize(self.handle))

    def set_host_threads(self, n: int):
        """Threads that stage pageable caller memory (0 = default)."""
        self.check(self.lib.ws_ctx_set_host_threads(self.handle, int(n)))

    def set_option(self, opt: int, value: int):
        self.check(self.lib.ws_ctx_set_option(self.handle, int(opt), int(value)))

    def set_tie_seed(self, seed: int):
        """Key of the WS_TIE_RANDOM generator (default: drawn from the OS per context, like thread_rng)."""
        self.check(self.lib.ws_ctx_set_tie_seed(self.handle, int(seed) & 0xFFFFFFFFFFFFFFFF))


_default_ctx = {}


def default_context(device: int = 0) -> Context:
    """One lazily created Context per device (what the Rust shim keeps per thread)."""
    ctx = _default_ctx.get(device)
    if ctx is None or ctx.handle is None:
        ctx = Context(device)
        _default_ctx[device] = ctx
    return ctx


class Plan:
    """Device-resident pipeline (ws_plan_*): pointers are CUDA device pointers (ints)."""

    def __init__(self, ctx: Context, n_img: int, rows: int, cols: int):
        self.ctx, self.lib = ctx, ctx.lib
        self.n_img, self.rows, self.cols = n_img, rows, cols
        h = _P()
        ctx.check(self.lib.ws_plan_create(ctx.handle, n_img, rows, cols, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None) and self.ctx.handle:
            self.lib.ws_plan_destroy(self.handle)
        self.handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def find_local_minima(self, d_imgs: int, d_seeds_rc: int, cap: int, d_seed_off: int) -> int:
        total = C.c_size_t(0)
        self.ctx.check(self.lib.ws_plan_find_local_minima(self.handle, d_imgs, d_seeds_rc, cap, d_seed_off,
                                                          C.byref(total)))
        return int(total.value)

    def run(self, kind: int, max_water_level: int, d_imgs: int, d_seeds_rc: int, d_seed_off: int, nseeds: int):
        cfg = make_config(kind, max_water_level, False)
        self.ctx.check(self.lib.ws_plan_run(self.handle, C.byref(cfg), d_imgs, d_seeds_rc, d_seed_off, nseeds))

    @property
    def labels_ptr(self) -> int:
        return int(self.lib.ws_plan_labels(self.handle) or 0)

    @property
    def levels_ptr(self) -> int:
        return int(self.lib.ws_plan_levels(self.handle) or 0)

    @property
    def arrival_times_ptr(self) -> int:
        return int(self.lib.ws_plan_arrival_times(self.handle) or 0)

    @property
    def lake_counts_ptr(self) -> int:
        return int(self.lib.ws_plan_lake_counts(self.handle) or 0)

    # -- row strips of one large field (ws_plan_strip_*) -----------------------------------
    def strip_begin(self, kind: int, max_water_level: int, global_rows: int, row_offset: int, halo_top: bool,
                    halo_bottom: bool, colour_base: int, d_img: int, d_seeds_rc: int, nseeds: int):
        cfg = make_config(kind, max_water_level, False)
        st = WsStrip(global_rows, row_offset, 1 if halo_top else 0, 1 if halo_bottom else 0, colour_base)
        self.ctx.check(self.lib.ws_plan_strip_begin(self.handle, C.byref(cfg), C.byref(st), d_img, d_seeds_rc, nseeds))

    def strip_export_times(self, d_top: int, d_bottom: int):
        self.ctx.check(self.lib.ws_plan_strip_export_times(self.handle, d_top or None, d_bottom or None))

    def strip_import_times(self, d_top: int, d_bottom: int) -> bool:
        ch = C.c_int(0)
        self.ctx.check(self.lib.ws_plan_strip_import_times(self.handle, d_top or None, d_bottom or None, C.byref(ch)))
        return bool(ch.value)

    def strip_labels(self):
        self.ctx.check(self.lib.ws_plan_strip_labels(self.handle))

    def strip_export_labels(self, d_top: int, d_bottom: int):
        self.ctx.check(self.lib.ws_plan_strip_export_labels(self.handle, d_top or None, d_bottom or None))

    def strip_import_labels(self, d_top: int, d_bottom: int) -> int:
        n = C.c_size_t(0)
        self.ctx.check(self.lib.ws_plan_strip_import_labels(self.handle, d_top or None, d_bottom or None, C.byref(n)))
        return int(n.value)

    def strip_edges(self):
        """-> (device ptr of uint32 pairs, device ptr of level bytes, count, colours present)."""
        ab, w, n, nd = _P(), _P(), C.c_size_t(0), C.c_uint32(0)
        self.ctx.check(self.lib.ws_plan_strip_edges(self.handle, C.byref(ab), C.byref(w), C.byref(n), C.byref(nd)))
        return int(ab.value or 0), int(w.value or 0), int(n.value), int(nd.value)

    def union_edges(self, d_ab: int, d_w: int, n: int, ncolours: int, ndistinct: int, max_water_level: int):
        self.ctx.check(self.lib.ws_plan_union_edges(self.handle, d_ab or None, d_w or None, n, ncolours, ndistinct,
                                                    max_water_level))

    # -- the same without host synchronisation (several exchange rounds in flight) ----------
    def strip_begin_async(self, kind: int, max_water_level: int, global_rows: int, row_offset: int, halo_top: bool,
                          halo_bottom: bool, colour_base: int, d_img: int, d_seeds_rc: int, nseeds: int):
        cfg = make_config(kind, max_water_level, False)
        st = WsStrip(global_rows, row_offset, 1 if halo_top else 0, 1 if halo_bottom else 0, colour_base)
        self.ctx.check(self.lib.ws_plan_strip_begin_async(self.handle, C.byref(cfg), C.byref(st), d_img, d_seeds_rc,
                                                          nseeds))

    def strip_export_times_async(self, d_top: int, d_bottom: int):
        self.ctx.check(self.lib.ws_plan_strip_export_times_async(self.handle, d_top or None, d_bottom or None))

    def strip_import_times_async(self, d_top: int, d_bottom: int, d_changed: int):
        self.ctx.check(self.lib.ws_plan_strip_import_times_async(self.handle, d_top or None, d_bottom or None, d_changed))

    def strip_labels_async(self):
        self.ctx.check(self.lib.ws_plan_strip_labels_async(self.handle))

    def strip_export_labels_async(self, d_top: int, d_bottom: int):
        self.ctx.check(self.lib.ws_plan_strip_export_labels_async(self.handle, d_top or None, d_bottom or None))

    def strip_import_labels_async(self, d_top: int, d_bottom: int, d_pending: int):
        self.ctx.check(self.lib.ws_plan_strip_import_labels_async(self.handle, d_top or None, d_bottom or None,
                                                                  d_pending))

    def strip_labels_finish_async(self):
        self.ctx.check(self.lib.ws_plan_strip_labels_finish_async(self.handle))

    def strip_check(self):
        self.ctx.check(self.lib.ws_plan_strip_check(self.handle))

    def strip_packet_bytes(self) -> int:
        return int(self.lib.ws_strip_packet_bytes(self.cols))

    def strip_forest(self, ncolours_total: int, d_packet: int):
        self.ctx.check(self.lib.ws_plan_strip_forest(self.handle, ncolours_total, d_packet))

    def forest_packets(self, d_packets: int, n_packets: int, ncolours_total: int, max_water_level: int):
        self.ctx.check(self.lib.ws_plan_forest_packets(self.handle, d_packets, n_packets, ncolours_total,
                                                       max_water_level))

    def snapshot(self, kind: int, i: int, level: int, d_out: int):
        self.ctx.check(self.lib.ws_plan_snapshot(self.handle, kind, i, level, d_out))

    def phase_ms(self) -> dict:
        arr = (C.c_float * 4)()
        self.ctx.check(self.lib.ws_plan_phase_ms(self.handle, C.byref(arr)))
        return {"init": arr[0], "flood": arr[1], "labels": arr[2], "merge": arr[3]}

    KERNELS = ("fill_state", "seed_init", "flood", "label_tile", "rim_jump", "label_finish", "merge_reduce",
               "forest_init", "forest_boruvka", "lake_counts")

    def kernel_ms(self) -> dict:
        arr = (C.c_float * 10)()
        self.ctx.check(self.lib.ws_plan_kernel_ms(self.handle, C.byref(arr)))
        return {k: float(arr[i]) for i, k in enumerate(self.KERNELS)}

    def stats(self) -> dict:
        arr = (C.c_uint64 * 8)()
        self.ctx.check(self.lib.ws_plan_stats(self.handle, C.byref(arr)))
        return {"stale_entries": arr[0], "tile_activations": arr[1], "jump_rounds": arr[2],
                "merge_edges": arr[3], "kernel_launches": arr[4], "flood_phases": arr[5],
                "flood_wait_kcycles": arr[6], "flood_busy_kcycles": arr[7]}
