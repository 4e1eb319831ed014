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


def sidecar_scene(seed=5, n=60):
    """A synthetic VR-BAG layout: 6 x 8 base cells of 16 m, raster at 0.5 m (192 x 256), refinement grids at 0.5 / 1 / 2 / 4 m
    sitting at sw_corner offsets inside their base cell; a few grids deliberately spill over their base cell (overlap with the
    neighbour: later grids overwrite) and over the raster edge (clipping)."""
    rng = np.random.default_rng(seed)
    base_shape, cell = (6, 8), 16.0
    bounds = (1000.0, 5000.0, 1000.0 + base_shape[1] * cell, 5000.0 + base_shape[0] * cell)
    res = 0.5
    shape = (int(round((bounds[3] - bounds[1]) / res)), int(round((bounds[2] - bounds[0]) / res)))
    gt = (bounds[0], res, 0.0, bounds[3], 0.0, -res)
    grids = []
    for i in range(n):
        r = float(rng.choice([0.5, 1.0, 2.0, 4.0]))
        rows, cols = int(rng.integers(2, int(cell / r) + (4 if i % 7 == 0 else 1))), int(rng.integers(2, int(cell / r) + (4 if i % 5 == 0 else 1)))
        br, bc = int(rng.integers(0, base_shape[0])), int(rng.integers(0, base_shape[1]))
        sw = (float(rng.integers(0, 4)) * r, float(rng.integers(0, 4)) * r)
        d = (20.0 + rng.normal(0, 0.3, size=(rows, cols))).astype(np.float32)
        d[rng.random((rows, cols)) < 0.15] = 1.0e6
        if i % 11 == 0:
            d[0, 0] = np.nan
        u = (0.1 + 0.05 * rng.random((rows, cols))).astype(np.float32)
        grids.append(dict(base_row=br, base_col=bc, depth=d, uncertainty=u, resolution=(r, r), dimensions=(rows, cols),
                          sw_corner=sw, start_index=i))
    results = []
    for g in grids:
        rows, cols = g["dimensions"]
        cls = rng.integers(0, 3, size=(rows, cols)).astype(np.float32)
        conf = rng.random((rows, cols)).astype(np.float32)
        conf[rng.random((rows, cols)) < 0.1] = np.float32(0.85)          # exactly on the >= threshold
        corr = rng.normal(0, 0.5, size=(rows, cols)).astype(np.float32)
        results.append((cls, conf, corr))
    return dict(shape=shape, geotransform=gt, bounds=bounds, base_shape=base_shape), grids, results


def multiscale_case(seed=21):
    """37 x 50 grid (not a multiple of 2 or 4: the crop of the block reduction is exercised), ~12 % holes marked NaN in the
    depth AND False in the mask, uncertainty plane."""
    rng = np.random.default_rng(seed)
    d = noisy_tile(37, 50, seed=seed)
    m = rng.random((37, 50)) > 0.12
    m[10:14, 20:27] = False
    d = d.copy()
    d[~m] = np.nan
    u = (0.1 + 0.05 * rng.random((37, 50))).astype(np.float32)
    return d, m, u
