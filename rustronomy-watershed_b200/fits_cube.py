"""FITS data cubes through the engine without a host pass over the pixels.

The reference's real-data pipeline (tests/integration.rs:72-94, 267-300, 344-356) is, per channel of a cube,

    img  = watershed.pre_processor(cube.slice(channel))      # f64 -> u8, lib.rs:1081-1173
    mins = watershed.find_local_minima(img.view())           # lib.rs:1178-1197
    watershed.transform_to_list(img.view(), &mins)           # or transform(...)

with every step a separate pass over host memory.  Here a batch of channels goes up ONCE, as the bytes of the
file (big-endian, swapped by the quantisation kernel), and `pre_processor` -> `find_local_minima` -> the transform
run back to back on the device (ws_dev_pre_processor, ws_plan_find_local_minima, ws_plan_run of the C ABI).

The FITS reader below handles what a radio data cube needs: the primary HDU, BITPIX 8 / 16 / 32 / 64 / -32 / -64,
NAXIS up to 4 with degenerate leading axes, BSCALE / BZERO.  No astropy in the image -- and none is needed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as N

BLOCK = 2880
_BITPIX = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}
# dtype codes of ws_pre_processor for data as it is stored in the file / for native arrays
_FILE_DTYPE = {">u1": 5, ">i2": 9, ">i4": 10, ">i8": 11, ">f4": 7, ">f8": 8}
_NATIVE_DTYPE = {"float32": 0, "float64": 1, "int32": 2, "uint16": 3, "int16": 4, "uint8": 5, "int64": 6}


@dataclass
class FitsCube:
    header: Dict[str, object]
    data: np.ndarray          # [channel][row][col], big-endian as stored (a view of the file's bytes)
    bscale: float = 1.0
    bzero: float = 0.0

    @property
    def physical(self) -> np.ndarray:
        """Native-endian values with BSCALE / BZERO applied (a host pass: tests and small cubes only)."""
        a = self.data.astype(self.data.dtype.newbyteorder("="))
        if self.bscale != 1.0 or self.bzero != 0.0:
            a = self.bzero + self.bscale * a.astype(np.float64)
        return a


def _parse_card(card: str) -> Optional[Tuple[str, object]]:
    key = card[:8].strip()
    if not key or key in ("COMMENT", "HISTORY") or card[8:10] != "= ":
        return None
    val = card[10:]
    if "'" in val[:2] or val.lstrip().startswith("'"):
        v = val.lstrip()
        end = 1
        while True:                     # a doubled quote is an escaped quote
            end = v.find("'", end)
            if end < 0 or v[end + 1:end + 2] != "'":
                break
            end += 2
        return key, v[1:end if end >= 0 else None].replace("''", "'").rstrip()
    val = val.split("/")[0].strip()
    if val in ("T", "F"):
        return key, val == "T"
    try:
        return key, int(val)
    except ValueError:
        try:
            return key, float(val.replace("D", "E"))
        except ValueError:
            return key, val


def read_fits_cube(path: str) -> FitsCube:
    """Primary HDU of a FITS file as a cube [channel][row][col] (NAXIS3, NAXIS2, NAXIS1); the data stay a
    memory map of the file, in the file's byte order."""
    header: Dict[str, object] = {}
    with open(path, "rb") as f:
        offset, done = 0, False
        while not done:
            block = f.read(BLOCK)
            if len(block) < BLOCK:
                raise ValueError("truncated FITS header")
            offset += BLOCK
            for i in range(0, BLOCK, 80):
                card = block[i:i + 80].decode("ascii", "replace")
                if card.startswith("END") and card[3:].strip() == "":
                    done = True
                    break
                kv = _parse_card(card)
                if kv:
                    header[kv[0]] = kv[1]
    if header.get("SIMPLE") is not True:
        raise ValueError("not a standard FITS file (SIMPLE != T)")
    bitpix = int(header["BITPIX"])
    if bitpix not in _BITPIX:
        raise ValueError(f"unsupported BITPIX {bitpix}")
    naxis = int(header.get("NAXIS", 0))
    dims = [int(header[f"NAXIS{k}"]) for k in range(1, naxis + 1)]      # NAXIS1 varies fastest
    while len(dims) > 3 and dims[-1] == 1:                               # degenerate Stokes / frequency axes
        dims.pop()
    if len(dims) == 2:
        dims.append(1)
    if len(dims) != 3:
        raise ValueError(f"expected an image or a cube, got NAXIS = {naxis} with axes {dims}")
    nx, ny, nchan = dims
    dt = np.dtype(_BITPIX[bitpix])
    data = np.memmap(path, dtype=dt, mode="r", offset=offset, shape=(nchan, ny, nx))
    return FitsCube(header, data, float(header.get("BSCALE", 1.0)), float(header.get("BZERO", 0.0)))


def write_fits_cube(path: str, cube: np.ndarray, extra: Optional[Dict[str, object]] = None) -> None:
    """A minimal standard-conforming primary HDU (tests; the engine only reads cubes)."""
    a = np.asarray(cube)
    if a.ndim != 3:
        raise ValueError("cube must be [channel][row][col]")
    bitpix = {"uint8": 8, "int16": 16, "int32": 32, "int64": 64, "float32": -32, "float64": -64}[a.dtype.name]
    cards = [("SIMPLE", True), ("BITPIX", bitpix), ("NAXIS", 3), ("NAXIS1", a.shape[2]), ("NAXIS2", a.shape[1]),
             ("NAXIS3", a.shape[0])] + list((extra or {}).items())

    def fmt(k, v):
        if isinstance(v, bool):
            s = f"{'T' if v else 'F':>20}"
        elif isinstance(v, (int, np.integer)):
            s = f"{int(v):>20}"
        elif isinstance(v, float):
            s = f"{v:>20.12E}"
        else:
            s = "'" + str(v).replace("'", "''").ljust(8) + "'"
        return f"{k:<8}= {s}".ljust(80)[:80]
    hdr = "".join(fmt(k, v) for k, v in cards) + "END".ljust(80)
    hdr = hdr.ljust((len(hdr) + BLOCK - 1) // BLOCK * BLOCK)
    raw = np.ascontiguousarray(a.astype(a.dtype.newbyteorder(">"))).tobytes()
    with open(path, "wb") as f:
        f.write(hdr.encode("ascii"))
        f.write(raw)
        f.write(b"\0" * ((-len(raw)) % BLOCK))


@dataclass
class CubeResult:
    seeds: List[np.ndarray]                   # per channel: (n, 2) uint32 (row, col) of find_local_minima
    lake_counts: Optional[np.ndarray]         # merging: [channel][max + 1]
    labels: Optional[np.ndarray]              # segmenting final labels [channel][rows][cols] uint32, if asked for
    quantised: Optional[np.ndarray]           # the pre_processor output [channel][rows][cols] uint8, if asked for


def watershed_cube(cube, kind: int = N.WS_MERGING, max_water_level: int = 254, max_value: int = 254,
                   channels: Optional[List[int]] = None, batch: int = 32, device: int = 0, want_labels: bool = False,
                   want_quantised: bool = False) -> CubeResult:
    """pre_processor -> find_local_minima -> transform for the channels of a cube, `batch` channels per launch set.

    `cube`: a FitsCube (data go up as stored) or a numpy array [channel][row][col] of a supported element type.
    BSCALE / BZERO other than 1 / 0 are applied on the host in f64 first (the quantisation is not affine-invariant
    in its last bit)."""
    if isinstance(cube, FitsCube):
        if cube.bscale != 1.0 or cube.bzero != 0.0:
            data, code = np.ascontiguousarray(cube.physical, dtype=np.float64), 1
        else:
            data, code = cube.data, _FILE_DTYPE[cube.data.dtype.str.replace("|", ">") if cube.data.dtype.itemsize == 1
                                               else cube.data.dtype.str]
    else:
        data = np.asarray(cube)
        code = _NATIVE_DTYPE.get(data.dtype.name)
        if code is None:
            raise TypeError(f"unsupported element type {data.dtype}")
    if data.ndim != 3:
        raise ValueError("cube must be [channel][row][col]")
    nchan, rows, cols = data.shape
    todo = list(range(nchan)) if channels is None else list(channels)
    ctx = N.default_context(device)
    lib = ctx.lib
    npx = rows * cols
    es = data.dtype.itemsize
    nlev = max_water_level + 1
    seeds_out: List[np.ndarray] = []
    counts_out = np.zeros((len(todo), nlev), np.uint64) if kind == N.WS_MERGING else None
    labels_out = np.zeros((len(todo), rows, cols), np.uint32) if want_labels else None
    quant_out = np.zeros((len(todo), rows, cols), np.uint8) if want_quantised else None
    nb = max(1, min(batch, len(todo)))
    d_raw = ctx.dev_malloc(nb * npx * es)
    d_u8 = ctx.dev_malloc(nb * npx)
    d_off = ctx.dev_malloc((nb + 1) * 4)
    plan = None
    try:
        for b0 in range(0, len(todo), nb):
            chans = todo[b0:b0 + nb]
            n = len(chans)
            if plan is None or plan.n_img != n:
                if plan is not None:
                    plan.close()
                plan = N.Plan(ctx, n, rows, cols)
            for k, ch in enumerate(chans):                     # the file's bytes, one channel after the other
                ctx.h2d(d_raw + k * npx * es, np.ascontiguousarray(data[ch]))
            for k in range(n):                                  # per-slice min / max like the reference's per-slice call
                ctx.check(lib.ws_dev_pre_processor(ctx.handle, code, d_raw + k * npx * es, npx, max_value, d_u8 + k * npx))
            total = plan.find_local_minima(d_u8, 0, 0, d_off)
            d_seeds = ctx.dev_malloc(max(total, 1) * 8)
            try:
                plan.find_local_minima(d_u8, d_seeds, total, d_off)
                plan.run(kind, max_water_level, d_u8, d_seeds, d_off, total)
                off = ctx.d2h(d_off, (n + 1,), np.uint32).astype(np.int64)
                s_all = ctx.d2h(d_seeds, (max(total, 1), 2), np.uint32)[:total]
                for k in range(n):
                    seeds_out.append(s_all[off[k]:off[k + 1]].copy())
                if counts_out is not None:
                    c = ctx.d2h(plan.lake_counts_ptr, (n, 256), np.uint32)
                    counts_out[b0:b0 + n] = c[:, :nlev]
                if labels_out is not None:
                    labels_out[b0:b0 + n] = ctx.d2h(plan.labels_ptr, (n, rows, cols), np.uint32) & 0x7FFFFFFF
                if quant_out is not None:
                    quant_out[b0:b0 + n] = ctx.d2h(d_u8, (n, rows, cols), np.uint8)
            finally:
                ctx.dev_free(d_seeds)
    finally:
        if plan is not None:
            plan.close()
        for p in (d_raw, d_u8, d_off):
            ctx.dev_free(p)
    return CubeResult(seeds_out, counts_out, labels_out, quant_out)
