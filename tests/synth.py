"""Seeded synthetic inputs shared by tests, bench.py and the golden generator.

Nothing here touches /root/reference (it does not exist on the GPU box).  The
analytic surface is SURVEY.md 8(d)'s: 30 + 5 sin(x/60) + 3 cos(y/45) + 0.01 x.
"""
from __future__ import annotations

import numpy as np


def clean_surface(h: int, w: int, r0: int = 0, c0: int = 0) -> np.ndarray:
    y, x = np.meshgrid(np.arange(r0, r0 + h, dtype=np.float64),
                       np.arange(c0, c0 + w, dtype=np.float64), indexing="ij")
    return (30.0 + 5.0 * np.sin(x / 60.0) + 3.0 * np.cos(y / 45.0) + 0.01 * x).astype(np.float32)


def noisy_tile(h: int, w: int, seed: int = 0, sigma: float = 0.2, spike_p: float = 0.005,
               r0: int = 0, c0: int = 0) -> np.ndarray:
    """Surface + gaussian noise everywhere + sparse spikes of +-U(1,5) m."""
    rng = np.random.default_rng(seed)
    d = clean_surface(h, w, r0, c0).astype(np.float64)
    d += rng.normal(0.0, sigma, size=d.shape)
    spikes = rng.random(d.shape) < spike_p
    d += spikes * rng.uniform(1.0, 5.0, size=d.shape) * rng.choice([-1.0, 1.0], size=d.shape)
    return d.astype(np.float32)


def blob_mask(h: int, w: int, seed: int = 0, invalid_frac: float = 0.02, corner: int = 0) -> np.ndarray:
    """Valid mask with ~invalid_frac invalid cells in random blobs (+ optional
    invalid corner block of side `corner`)."""
    rng = np.random.default_rng(seed + 1000)
    mask = np.ones((h, w), bool)
    target = int(invalid_frac * h * w)
    guard = 0
    while (~mask).sum() < target and guard < 10000:
        guard += 1
        r, c = rng.integers(0, h), rng.integers(0, w)
        rad = int(rng.integers(1, max(2, min(h, w) // 16)))
        rr, cc = np.ogrid[:h, :w]
        mask &= ~(((rr - r) ** 2 + (cc - c) ** 2) <= rad * rad)
    if corner:
        mask[:corner, :corner] = False
    return mask


def scattered_mask(h: int, w: int, seed: int = 0, invalid_frac: float = 0.05) -> np.ndarray:
    rng = np.random.default_rng(seed + 2000)
    return rng.random((h, w)) >= invalid_frac


def uncertainty_plane(h: int, w: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed + 3000)
    return (0.1 + 0.05 * rng.random((h, w))).astype(np.float32)


def ragged_grids(n: int, seed: int = 7, lo: int = 3, hi: int = 50, nodata: float = 1.0e6,
                 nodata_frac: float = 0.10, all_invalid_every: int = 0):
    """n grids with rows, cols ~ U{lo..hi}, resolution in
    {0.5,1,2,4}, `nodata` on ~10 % of the cells.  Returns list of
    (depth f32 [r,c], uncertainty f32 [r,c], (rx, ry))."""
    rng = np.random.default_rng(seed)
    out = []
    for g in range(n):
        r, c = int(rng.integers(lo, hi + 1)), int(rng.integers(lo, hi + 1))
        res = float(rng.choice([0.5, 1.0, 2.0, 4.0]))
        d = (20.0 + 0.05 * np.arange(c)[None, :] + 0.03 * np.arange(r)[:, None]
             + rng.normal(0, 0.15, size=(r, c))).astype(np.float32)
        spikes = rng.random((r, c)) < 0.01
        d = (d + spikes * rng.uniform(1, 4, size=(r, c)).astype(np.float32)).astype(np.float32)
        inval = rng.random((r, c)) < nodata_frac
        if all_invalid_every and g % all_invalid_every == all_invalid_every - 1:
            inval[:] = True
        d[inval] = nodata
        u = (0.1 + 0.05 * rng.random((r, c))).astype(np.float32)
        out.append((d, u, (res, res)))
    return out
