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


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[2]: the synthetic survey as a COUNTER-BASED function of (seed, row, col), so that any block of it can
# be regenerated on the CPU for spot parity.  Everything that is not exactly reproducible across
# CPU / GPU math libraries (sin, cos) lives in 1-D HOST tables; the per-cell part is integer hashing + a handful of
# IEEE fp32 adds / multiplies carried out in the same order by `survey_block` (NumPy) and `survey_block_torch`.
_M32 = 0xFFFFFFFF
SURVEY_NODATA = 1.0e6


def survey_tables(S: int):
    """1-D fp32 tables over rows / cols / (row + col) of an S x S survey: the analytic surface of SURVEY.md 8(d)
    (30 + 5 sin(x/60) + 3 cos(y/45) + 0.01 x, separable) and the smooth 'coast' field that marks ~5 % of the area invalid."""
    i = np.arange(S, dtype=np.float64)
    k = np.arange(2 * S, dtype=np.float64)
    return dict(col_surface=(30.0 + 5.0 * np.sin(i / 60.0) + 0.01 * i).astype(np.float32), row_surface=(3.0 * np.cos(i / 45.0)).astype(np.float32),
                coast_row=np.sin(i / 2900.0).astype(np.float32), coast_col=np.cos(i / 2300.0).astype(np.float32),
                coast_diag=(0.3 * np.sin(k / 700.0)).astype(np.float32))


def _mix32(h, xp):
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & _M32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & _M32
    return h ^ (h >> 16)


def _survey_cells(xp, f32, rows, cols, seed, tabs_rc, sigma, spike_p, where):
    """shared recipe; `rows` [h,1] / `cols` [1,w] int64, tabs_rc = the five tables already gathered to [h,1] / [1,w] / [h,w]"""
    col_s, row_s, c_row, c_col, c_diag = tabs_rc
    key = (rows * 0x85EBCA77 + cols * 0xC2B2AE3D + (int(seed) * 0x9E3779B1 & _M32)) & _M32
    h1, h2 = _mix32(key ^ 0x3C6EF372, xp), _mix32(key ^ 0x1B873593, xp)
    h3, h4 = _mix32(key ^ 0x27D4EB2F, xp), _mix32(key ^ 0x165667B1, xp)
    g = (h1 & 0xFFFF) + (h1 >> 16) + (h2 & 0xFFFF) + (h2 >> 16) - 131070          # Irwin-Hall(4): std 37837.2 counts
    d = (col_s + row_s) + f32(g) * np.float32(sigma / 37837.2)
    amp = np.float32(1.0) + f32(h4 & 0xFFFF) * np.float32(4.0 / 65536.0)
    amp = where(((h4 >> 16) & 1) == 1, -amp, amp)
    d = where(h3 < int(spike_p * 4294967296.0), d + amp, d)
    coast = (c_row + c_col) + c_diag
    return where(coast > np.float32(1.72), np.float32(SURVEY_NODATA), d)


def survey_block(S: int, r0: int, r1: int, c0: int, c1: int, seed: int = 1234, sigma: float = 0.2, spike_p: float = 0.005,
                 tables=None) -> np.ndarray:
    """Rows [r0, r1) x cols [c0, c1) of the S x S synthetic survey (fp32, nodata = 1e6), NumPy."""
    t = tables or survey_tables(S)
    rows = np.arange(r0, r1, dtype=np.int64)[:, None]
    cols = np.arange(c0, c1, dtype=np.int64)[None, :]
    tabs = (t["col_surface"][c0:c1][None, :], t["row_surface"][r0:r1][:, None], t["coast_row"][r0:r1][:, None],
            t["coast_col"][c0:c1][None, :], t["coast_diag"][(rows + cols)])
    with np.errstate(over="ignore"):
        out = _survey_cells(np, lambda a: a.astype(np.float32), rows, cols, seed, tabs, sigma, spike_p, np.where)
    return np.ascontiguousarray(out, dtype=np.float32)


def survey_block_torch(S: int, r0: int, r1: int, c0: int, c1: int, device, seed: int = 1234, sigma: float = 0.2,
                       spike_p: float = 0.005, tables=None, out=None, chunk_rows: int = 512):
    """The same block generated on `device` (bit-identical to `survey_block`); `tables` = survey_tables(S) moved to the
    device once by the caller (dict of tensors) or None."""
    import torch
    if tables is None or not isinstance(tables["col_surface"], torch.Tensor):
        tables = {k: torch.from_numpy(v).to(device) for k, v in (tables or survey_tables(S)).items()}
    if out is None:
        out = torch.empty((r1 - r0, c1 - c0), dtype=torch.float32, device=device)
    cols = torch.arange(c0, c1, dtype=torch.int64, device=device)[None, :]
    t = tables
    where = lambda c, a, b: torch.where(c, torch.as_tensor(a, device=device) if not isinstance(a, torch.Tensor) else a,
                                        torch.as_tensor(b, device=device) if not isinstance(b, torch.Tensor) else b)
    for a in range(r0, r1, chunk_rows):
        b = min(a + chunk_rows, r1)
        rows = torch.arange(a, b, dtype=torch.int64, device=device)[:, None]
        tabs = (t["col_surface"][c0:c1][None, :], t["row_surface"][a:b][:, None], t["coast_row"][a:b][:, None],
                t["coast_col"][c0:c1][None, :], t["coast_diag"][(rows + cols)])
        out[a - r0:b - r0] = _survey_cells(torch, lambda v: v.to(torch.float32), rows, cols, seed, tabs, sigma, spike_p, where)
    return out
