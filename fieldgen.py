"""Synthetic u8 input fields for tests and bench.

Generators are deterministic for a given (shape, seed).  numpy only, so the
same bytes can be produced on the build container and on the GPU box.
"""
from __future__ import annotations

import numpy as np


def uniform(rows: int, cols: int, seed: int = 0) -> np.ndarray:
    """i.i.d. uniform in [0, 254) like `Uniform::new(0, 254)` in the reference's
    README example and tests/core_bench.rs:29 (values 0..=253)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 254, size=(rows, cols), dtype=np.uint8)


def _rfft_filter(rows: int, cols: int, seed: int, transfer) -> np.ndarray:
    """White Gaussian noise filtered in Fourier space (periodic)."""
    import scipy.fft as sfft
    rng = np.random.default_rng(seed)
    white = rng.standard_normal((rows, cols), dtype=np.float32)
    f = sfft.rfft2(white, workers=-1)
    ky = np.fft.fftfreq(rows).astype(np.float32)[:, None]
    kx = np.fft.rfftfreq(cols).astype(np.float32)[None, :]
    f *= transfer(ky, kx)
    return sfft.irfft2(f, s=(rows, cols), workers=-1).astype(np.float32)


def _quantise(x: np.ndarray, lo: int, hi: int) -> np.ndarray:
    mn, mx = float(x.min()), float(x.max())
    y = (x - mn) / max(mx - mn, 1e-30) * (hi - lo) + lo
    return np.clip(y, lo, hi).astype(np.uint8)      # truncation, like `as u8`


def smooth(rows: int, cols: int, sigma: float = 8.0, seed: int = 0) -> np.ndarray:
    """White noise -> periodic Gaussian blur (sigma px) -> affine map to 1..=254."""
    def tr(ky, kx):
        return np.exp(-2.0 * (np.pi * sigma) ** 2 * (ky * ky + kx * kx)).astype(np.float32)
    return _quantise(_rfft_filter(rows, cols, seed, tr), 1, 254)


def cgps_like(rows: int, cols: int, seed: int = 0, noise: float = 0.10, masked: bool = True) -> np.ndarray:
    """Power-law Gaussian random field P(k) ~ k^-3 plus white noise at `noise`
    of the signal sigma, quantised to 0..=254; outside an ellipse the "mosaic"
    is NaN -> 255 = NEVER_FILL (cf. the NaN-heavy channel of
    tests/integration.rs:344-356)."""
    def tr(ky, kx):
        k = np.sqrt(ky * ky + kx * kx)
        k[0, 0] = 1.0
        t = k ** (-1.5)                                   # amplitude = sqrt(P)
        t[0, 0] = 0.0
        return t.astype(np.float32)
    g = _rfft_filter(rows, cols, seed, tr)
    g /= max(float(g.std()), 1e-30)
    rng = np.random.default_rng(seed + 0x5EED)
    g += noise * rng.standard_normal((rows, cols), dtype=np.float32)
    q = _quantise(g, 0, 254)
    if masked:
        y = (np.arange(rows, dtype=np.float32)[:, None] - rows / 2) / (0.49 * rows)
        x = (np.arange(cols, dtype=np.float32)[None, :] - cols / 2) / (0.47 * cols)
        q[(x * x + y * y) > 1.0] = 255
    return q


def plateaus(rows: int, cols: int, levels: int = 6, sigma: float = 3.0, seed: int = 0) -> np.ndarray:
    """Smooth field quantised to a handful of values: large flat plateaus."""
    s = smooth(rows, cols, sigma, seed).astype(np.float32)
    q = np.floor(s / 255.0 * levels)
    return (q * (254 // max(levels, 1))).astype(np.uint8)


def obstacles(rows: int, cols: int, seed: int = 0) -> np.ndarray:
    """Uniform field with NEVER_FILL (255) blocks and ALWAYS_FILL (0) lines."""
    rng = np.random.default_rng(seed)
    a = uniform(rows, cols, seed)
    for _ in range(max(1, rows * cols // 400)):
        r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
        h, w = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        a[r:r + h, c:c + w] = 255
    for _ in range(max(1, rows // 8)):
        r = int(rng.integers(0, rows))
        c0 = int(rng.integers(0, cols))
        a[r, c0:c0 + int(rng.integers(2, max(3, cols // 2)))] = 0
    return a


def maze(rows: int, cols: int, seed: int = 0) -> np.ndarray:
    """Serpentine corridor of zeros between 255 walls: one very long geodesic."""
    a = np.full((rows, cols), 255, np.uint8)
    for r in range(1, rows - 1, 2):
        a[r, 1:cols - 1] = 0
        if r + 1 < rows - 1:
            a[r + 1, (cols - 2) if ((r // 2) % 2 == 0) else 1] = 0
    return a
