"""Host-side mirror of the rustronomy-watershed public API over the C ABI.

Same names, argument meaning and error behaviour as the reference crate
(v0.4.1, src/lib.rs), so a test written against the crate reads the same here:

    ws   = TransformBuilder.default().build_segmenting()          # lib.rs:925, 1024
    mins = ws.find_local_minima(img)                              # lib.rs:1178
    out  = ws.transform(img, mins)                                # lib.rs:1208

Arrays are numpy: images uint8 (any strides), labels uint64 ("usize"), seeds a
sequence of (row, col).  All compute happens in libws_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Callable, Generic, List, Optional, Sequence, Tuple, TypeVar

import numpy as np

from . import _native as N
from . import plotting

T = TypeVar("T")

# lib.rs:138-141
UNCOLOURED = 0
NORMAL_MAX = 254
ALWAYS_FILL = 0
NEVER_FILL = 255


class BuildErr(Exception):
    """lib.rs:1051-1065.  `kind` is "MaxToHigh" or "MaxToLow", `value` the offending level."""

    def __init__(self, kind: str, value: int):
        self.kind, self.value = kind, value
        if kind == "MaxToHigh":
            msg = (f"Maximum water level set to {value}, which is higher than the maximum "
                   f"allowed value {NORMAL_MAX}")
        else:  # the reference prints NEVER_FILL here (lib.rs:1062)
            msg = (f"Maximum water level set to {value}, which is lower than the minimum "
                   f"allowed value {NEVER_FILL}")
        super().__init__(msg)

    @staticmethod
    def MaxToHigh(v: int) -> "BuildErr":
        return BuildErr("MaxToHigh", v)

    @staticmethod
    def MaxToLow(v: int) -> "BuildErr":
        return BuildErr("MaxToLow", v)


@dataclass
class HookCtx:
    """lib.rs:844-850."""
    water_level: int
    max_water_level: int
    image: np.ndarray                        # uint8 view, padded when edge correction is on
    colours: np.ndarray                      # uint64 view, valid during the hook call only
    seeds: List[Tuple[int, Tuple[int, int]]]  # (colour, (row, col))


class TransformBuilder(Generic[T]):
    """lib.rs:908-1047, including the `plots` feature's two options (lib.rs:972-994)."""

    def __init__(self):
        self.max_water_level = NORMAL_MAX          # lib.rs:942
        self.edge_correction = False
        self.wlvl_hook: Optional[Callable[[HookCtx], T]] = None
        self.device = 0
        self.tie_break = N.WS_TIE_FIRST
        self.tie_seed: Optional[int] = None
        self.plot_path: Optional[str] = None       # lib.rs:939
        self.plot_colour_map = None                # lib.rs:941; viridis when unset (lib.rs:1011)

    @classmethod
    def new(cls) -> "TransformBuilder":
        return cls()

    @classmethod
    def default(cls) -> "TransformBuilder":
        return cls()

    def set_max_water_lvl(self, max_water_lvl: int) -> "TransformBuilder":
        if not 0 <= int(max_water_lvl) <= 255:
            raise OverflowError("max_water_lvl is a u8")
        self.max_water_level = int(max_water_lvl)
        return self

    def enable_edge_correction(self) -> "TransformBuilder":
        self.edge_correction = True
        return self

    def set_wlvl_hook(self, hook: Callable[[HookCtx], T]) -> "TransformBuilder":
        self.wlvl_hook = hook
        return self

    def set_plot_colour_map(self, colour_map) -> "TransformBuilder":
        """lib.rs:975-985: `colour_map(count, min, max)`, vectorised (see plotting.py)."""
        self.plot_colour_map = colour_map
        return self

    def set_plot_folder(self, path) -> "TransformBuilder":
        """lib.rs:991-994: with a folder set every water level writes `ws_lvl{level}.png` there."""
        self.plot_path = str(path)
        return self

    def set_device(self, device: int) -> "TransformBuilder":
        """Extension: CUDA device ordinal the transform runs on."""
        self.device = int(device)
        return self

    def set_tie_break(self, policy: str, seed: Optional[int] = None) -> "TransformBuilder":
        """Extension: "first" (default; the reference's `col0`, lib.rs:245) or "random" (the reference's own
        behaviour, lib.rs:250-253: uniform over the coloured neighbours; `seed` makes it reproducible)."""
        self.tie_break = {"first": N.WS_TIE_FIRST, "random": N.WS_TIE_RANDOM}[policy]
        self.tie_seed = seed
        return self

    def _validate(self, kind: int):
        cfg = N.make_config(kind, self.max_water_level, self.edge_correction, self.tie_break)
        st = N.load_library().ws_config_validate(C.byref(cfg))
        if st == N.WS_ERR_MAX_TOO_HIGH:
            raise BuildErr.MaxToHigh(self.max_water_level)      # lib.rs:1000-1001
        if st == N.WS_ERR_MAX_TOO_LOW:
            raise BuildErr.MaxToLow(self.max_water_level)       # lib.rs:1002-1003
        if st != N.WS_OK:
            raise N.WatershedError(st)

    def build_merging(self) -> "MergingWatershed":
        self._validate(N.WS_MERGING)
        return MergingWatershed(self.max_water_level, self.edge_correction, self.wlvl_hook, self.device,
                                self.tie_break, self.tie_seed, self.plot_path, self.plot_colour_map)

    def build_segmenting(self) -> "SegmentingWatershed":
        self._validate(N.WS_SEGMENTING)
        return SegmentingWatershed(self.max_water_level, self.edge_correction, self.wlvl_hook, self.device,
                                   self.tie_break, self.tie_seed, self.plot_path, self.plot_colour_map)


class WatershedUtils:
    """lib.rs:1069-1198."""

    device = 0

    def _ctx(self) -> N.Context:
        return N.default_context(self.device)

    _DTYPES = {"float32": 0, "float64": 1, "int32": 2, "uint16": 3, "int16": 4, "uint8": 5, "int64": 6}

    def pre_processor(self, img: np.ndarray) -> np.ndarray:
        """lib.rs:1081-1087: any numeric array (any dimension) -> u8 in [0, NORMAL_MAX]."""
        return self.pre_processor_with_max(img, NORMAL_MAX)

    def pre_processor_with_max(self, img: np.ndarray, MAX: int) -> np.ndarray:
        """lib.rs:1134-1173.  Raises AssertionError like the reference's asserts (1143-1144)."""
        a = np.asarray(img)
        code = self._DTYPES.get(a.dtype.name)
        if code is None:
            raise TypeError(f"unsupported element type {a.dtype}")
        if not 0 <= int(MAX) <= 255:
            raise OverflowError("MAX is a u8")
        a = np.ascontiguousarray(a)
        out = np.empty(a.shape, dtype=np.uint8)
        ctx = self._ctx()
        st = ctx.lib.ws_pre_processor(ctx.handle, code, a.ctypes.data, a.size, int(MAX), out.ctypes.data)
        if st in (N.WS_ERR_MAX_TOO_HIGH, N.WS_ERR_MAX_TOO_LOW):
            raise AssertionError("MAX must satisfy ALWAYS_FILL < MAX < NEVER_FILL")
        ctx.check(st)
        return out

    def find_local_minima(self, img: np.ndarray) -> np.ndarray:
        """Interior pixels strictly greater than all 8 neighbours, row-major, as an
        (n, 2) uint64 array of (row, col)  (lib.rs:1178-1197)."""
        ctx = self._ctx()
        view = N.image_view(img)
        out = C.c_void_p()
        n = C.c_size_t(0)
        ctx.check(ctx.lib.ws_find_local_minima(ctx.handle, C.byref(view), C.byref(out), C.byref(n)))
        try:
            if n.value == 0:
                return np.zeros((0, 2), dtype=np.uint64)
            buf = (C.c_uint64 * (2 * n.value)).from_address(out.value)
            return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 2).copy()
        finally:
            ctx.lib.ws_free(out)


class Watershed(WatershedUtils, Generic[T]):
    """lib.rs:1206-1238: the four trait methods, shared by both transforms."""

    KIND = N.WS_SEGMENTING

    def __init__(self, max_water_level: int, edge_correction: bool,
                 wlvl_hook: Optional[Callable[[HookCtx], T]], device: int = 0,
                 tie_break: int = N.WS_TIE_FIRST, tie_seed: Optional[int] = None,
                 plot_path: Optional[str] = None, plot_colour_map=None):
        self.plot_path = plot_path
        self.plot_colour_map = plot_colour_map if plot_colour_map is not None else plotting.viridis  # lib.rs:1011
        self.max_water_level = max_water_level
        self.edge_correction = edge_correction
        self.wlvl_hook = wlvl_hook
        self.device = device
        self.tie_break = tie_break
        self.tie_seed = tie_seed

    # -- helpers --------------------------------------------------------------
    def _cfg(self) -> N.WsConfig:
        if self.tie_break == N.WS_TIE_RANDOM and self.tie_seed is not None:
            self._ctx().set_tie_seed(self.tie_seed)
        return N.make_config(self.KIND, self.max_water_level, self.edge_correction, self.tie_break)

    def _out_shape(self, img: np.ndarray) -> Tuple[int, int]:
        pad = 2 if self.edge_correction else 0            # lib.rs:1330-1336
        return img.shape[0] + pad, img.shape[1] + pad

    @property
    def levels(self) -> int:
        return self.max_water_level + 1                   # 0..=max, lib.rs:1379 / 1689

    # -- Watershed::transform --------------------------------------------------
    def transform(self, input: np.ndarray, seeds: Sequence, out: Optional[np.ndarray] = None) -> np.ndarray:
        """`out` (extension): a caller-owned C-order uint64 array to fill, e.g. pinned memory."""
        self._plot_pass(input, seeds)
        ctx = self._ctx()
        cfg, view, s = self._cfg(), N.image_view(input), N.seeds_array(seeds)
        shape = tuple(input.shape) if self.KIND == N.WS_MERGING else self._out_shape(input)
        if out is None:
            out = np.empty(shape, dtype=np.uint64)
        elif out.shape != shape or out.dtype != np.uint64 or not out.flags.c_contiguous:
            raise ValueError(f"out must be a C-contiguous uint64 array of shape {shape}")
        ctx.check(ctx.lib.ws_transform(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data, s.shape[0],
                                       out.ctypes.data))
        return out

    # -- Watershed::transform_with_hook ----------------------------------------
    def _plot_level(self, water_level: int, colours: np.ndarray) -> None:
        """lib.rs:1472-1487 / 1758-1773: the level's label image without the edge-correction padding; a plot
        that fails is reported and the transform goes on."""
        try:
            view = colours[1:-1, 1:-1] if self.edge_correction else colours
            plotting.plot_slice(view, plotting.level_file(self.plot_path, water_level), self.plot_colour_map)
        except Exception as err:
            print(f"Could not make watershed plot. Error: {err}")

    def _plot_pass(self, input: np.ndarray, seeds: Sequence) -> None:
        """The reference routes transform / transform_history / transform_to_list through transform_with_hook
        (lib.rs:1524-1560), so they plot too; here they have their own device paths, and with a plot folder set
        one extra pass through the per-level hook makes the pictures."""
        if self.plot_path is not None:
            self._hook_run(input, seeds, None)

    def transform_with_hook(self, input: np.ndarray, seeds: Sequence) -> list:
        return self._hook_run(input, seeds, self.wlvl_hook)

    def _hook_run(self, input: np.ndarray, seeds: Sequence, hook) -> list:
        ctx = self._ctx()
        cfg, view, s = self._cfg(), N.image_view(input), N.seeds_array(seeds)
        plot = self.plot_path is not None
        results: list = []
        errors: list = []
        seed_cache: list = []

        def _cb(_user, hp):
            h = hp.contents
            try:
                if not seed_cache:      # (colour, (row, col)) as HookCtx carries them (lib.rs:849), from the C side
                    tri = np.frombuffer((C.c_uint64 * (3 * h.nseeds)).from_address(h.seeds), dtype=np.uint64) \
                        if h.nseeds else np.zeros(0, np.uint64)
                    seed_cache.append([(int(a), (int(r), int(c))) for a, r, c in tri.reshape(-1, 3)])
                seed_list = seed_cache[0]
                n = h.rows * h.cols
                image = np.frombuffer((C.c_uint8 * n).from_address(h.image), dtype=np.uint8).reshape(h.rows, h.cols)
                colours = np.frombuffer((C.c_uint64 * n).from_address(h.colours), dtype=np.uint64).reshape(h.rows, h.cols)
                if plot:
                    self._plot_level(h.water_level, colours)
                if hook is not None:
                    results.append(hook(HookCtx(h.water_level, h.max_water_level, image, colours, seed_list)))
            except BaseException as e:  # never unwind through C
                errors.append(e)

        cb = N.HOOK_FN(_cb) if (hook is not None or plot) else C.cast(None, N.HOOK_FN)
        ctx.check(ctx.lib.ws_transform_with_hook(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data,
                                                 s.shape[0], cb, None))
        if errors:
            raise errors[0]
        return results                                   # empty without a hook (lib.rs:1510, 1520)

    # -- Watershed::transform_to_list -------------------------------------------
    def transform_to_list(self, input: np.ndarray, seeds: Sequence) -> List[Tuple[int, np.ndarray]]:
        self._plot_pass(input, seeds)
        ctx = self._ctx()
        cfg, view, s = self._cfg(), N.image_view(input), N.seeds_array(seeds)
        r, c = self._out_shape(input)
        sizes = np.empty((self.levels, r * c + 1), dtype=np.uint64)
        lv = np.empty(self.levels, dtype=np.uint8)
        ctx.check(ctx.lib.ws_transform_to_list(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data, s.shape[0],
                                               lv.ctypes.data, sizes.ctypes.data))
        return [(int(lv[i]), sizes[i]) for i in range(self.levels)]

    # -- Watershed::transform_history --------------------------------------------
    def transform_history(self, input: np.ndarray, seeds: Sequence) -> List[Tuple[int, np.ndarray]]:
        self._plot_pass(input, seeds)
        ctx = self._ctx()
        cfg, view, s = self._cfg(), N.image_view(input), N.seeds_array(seeds)
        r, c = self._out_shape(input)
        hist = np.empty((self.levels, r, c), dtype=np.uint64)
        lv = np.empty(self.levels, dtype=np.uint8)
        ctx.check(ctx.lib.ws_transform_history(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data, s.shape[0],
                                               lv.ctypes.data, hist.ctypes.data))
        return [(int(lv[i]), hist[i]) for i in range(self.levels)]

    # -- extensions (compact results of the same computation) --------------------
    def lake_counts(self, input: np.ndarray, seeds: Sequence) -> Tuple[np.ndarray, np.ndarray]:
        """Per level: (number of lakes, number of uncoloured pixels)."""
        ctx = self._ctx()
        cfg, view, s = self._cfg(), N.image_view(input), N.seeds_array(seeds)
        lakes = np.empty(self.levels, dtype=np.uint64)
        unc = np.empty(self.levels, dtype=np.uint64)
        ctx.check(ctx.lib.ws_transform_lake_counts(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data,
                                                   s.shape[0], lakes.ctypes.data, unc.ctypes.data))
        return lakes, unc

    def lake_sizes_compact(self, input: np.ndarray, seeds: Sequence) -> Tuple[np.ndarray, np.ndarray]:
        """(lakes per level [levels], sizes [levels][nseeds+1]): the first nseeds+1 entries of every
        find_lake_sizes row of transform_to_list (the rest of those rows is zero)."""
        ctx = self._ctx()
        cfg, view, s = self._cfg(), N.image_view(input), N.seeds_array(seeds)
        lakes = np.empty(self.levels, dtype=np.uint64)
        sizes = np.empty((self.levels, s.shape[0] + 1), dtype=np.uint64)
        ctx.check(ctx.lib.ws_transform_lake_sizes_compact(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data,
                                                          s.shape[0], lakes.ctypes.data, sizes.ctypes.data))
        return lakes, sizes

    def transform_compact(self, input: np.ndarray, seeds: Sequence) -> Tuple[np.ndarray, np.ndarray]:
        """(uint32 final segmenting labels, uint8 level of colouring; 255 = never)."""
        ctx = self._ctx()
        cfg, view, s = self._cfg(), N.image_view(input), N.seeds_array(seeds)
        shape = self._out_shape(input)
        lab = np.empty(shape, dtype=np.uint32)
        lvl = np.empty(shape, dtype=np.uint8)
        ctx.check(ctx.lib.ws_transform_compact(ctx.handle, C.byref(cfg), C.byref(view), s.ctypes.data, s.shape[0],
                                               lab.ctypes.data, lvl.ctypes.data))
        return lab, lvl

    def find_local_minima_batch(self, imgs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Seeds of a stack of slices [n][rows][cols] -> (seeds [total][2], offsets [n+1])."""
        ctx = self._ctx()
        a = np.ascontiguousarray(imgs, dtype=np.uint8)
        assert a.ndim == 3
        out = C.c_void_p()
        off = np.zeros(a.shape[0] + 1, dtype=np.uint64)
        ctx.check(ctx.lib.ws_find_local_minima_batch(ctx.handle, a.ctypes.data, a.shape[0], a.shape[1], a.shape[2],
                                                     C.byref(out), off.ctypes.data))
        try:
            n = int(off[-1])
            if n == 0:
                return np.zeros((0, 2), dtype=np.uint64), off
            buf = (C.c_uint64 * (2 * n)).from_address(out.value)
            return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 2).copy(), off
        finally:
            ctx.lib.ws_free(out)

    def transform_batch(self, imgs: np.ndarray, seeds: np.ndarray, seed_offsets: np.ndarray,
                        want_labels: bool = True, want_lake_counts: bool = False):
        """A stack of equally shaped slices in one launch set."""
        ctx = self._ctx()
        a = np.ascontiguousarray(imgs, dtype=np.uint8)
        assert a.ndim == 3
        cfg, s = self._cfg(), N.seeds_array(seeds)
        off = np.ascontiguousarray(seed_offsets, dtype=np.uint64).ravel()
        if off.shape[0] != a.shape[0] + 1 or int(off[0]) != 0 or int(off[-1]) != s.shape[0] or np.any(np.diff(off.astype(np.int64)) < 0):
            raise ValueError("seed_offsets must have n_img + 1 ascending entries, starting at 0 and ending at len(seeds)")
        pad = 2 if self.edge_correction else 0
        labels = np.empty((a.shape[0], a.shape[1] + pad, a.shape[2] + pad), dtype=np.uint64) if want_labels else None
        counts = np.empty((a.shape[0], self.levels), dtype=np.uint64) if want_lake_counts else None
        ctx.check(ctx.lib.ws_transform_batch(ctx.handle, C.byref(cfg), a.ctypes.data, a.shape[0], a.shape[1],
                                             a.shape[2], s.ctypes.data, off.ctypes.data,
                                             labels.ctypes.data if want_labels else None,
                                             counts.ctypes.data if want_lake_counts else None))
        return labels, counts


class SegmentingWatershed(Watershed[T]):
    """lib.rs:1609-1848."""
    KIND = N.WS_SEGMENTING


class MergingWatershed(Watershed[T]):
    """lib.rs:1297-1562."""
    KIND = N.WS_MERGING
