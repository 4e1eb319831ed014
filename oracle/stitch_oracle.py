"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

NumPy restatement of the reference's overlapping-tile stitch
(/root/reference/data/tiling.py) and of the tile post-processing around it in
/root/reference/models/pipeline.py.  Pinned against the unmodified reference
(`TileManager`/`TileMerger`) by tests/golden/stitch_*.npz (tests/test_oracle_golden.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def tile_specs(shape, tile_size, overlap) -> Tuple[int, int, List[Tuple[int, int, int, int, int, int]]]:
    """tiling.py:87-138 -> (n_rows, n_cols, [(r0, c0, r1, c1, tile_row, tile_col)])."""
    stride = tile_size - overlap
    if stride <= 0:
        raise ValueError("Tile size must be larger than overlap")          # :84-85
    H, W = shape
    ntr = max(1, (H - overlap) // stride + (1 if (H - overlap) % stride > 0 else 0))
    ntc = max(1, (W - overlap) // stride + (1 if (W - overlap) % stride > 0 else 0))
    specs = []
    for tr in range(ntr):
        for tc in range(ntc):
            r0, c0 = tr * stride, tc * stride
            r1, c1 = min(r0 + tile_size, H), min(c0 + tile_size, W)
            if r1 - r0 < tile_size and r0 > 0:
                r0 = max(0, r1 - tile_size)
            if c1 - c0 < tile_size and c0 > 0:
                c0 = max(0, c1 - tile_size)
            specs.append((r0, c0, r1, c1, tr, tc))
    return ntr, ntc, specs


def blend_1d(size: int, overlap: int) -> np.ndarray:
    """tiling.py:313-330 (fp64 cosine ramp stored into an fp32 array)."""
    w = np.ones(size, dtype=np.float32)
    ramp = min(overlap, size // 4)
    if ramp > 0:
        w[:ramp] = 0.5 * (1 - np.cos(np.pi * np.linspace(0, 1, ramp)))
        w[-ramp:] = 0.5 * (1 - np.cos(np.pi * np.linspace(1, 0, ramp)))
    return w


def blend_2d(shape, overlap) -> np.ndarray:
    """tiling.py:296-311."""
    return np.outer(blend_1d(shape[0], overlap), blend_1d(shape[1], overlap)).astype(np.float32)


class Merger:
    """tiling.py:333-454 (`TileMerger`) + `merge_tile` weighted branch (:242-258)."""
    DISCRETE = {"classification"}

    def __init__(self, shape, channels, overlap):
        self.overlap = overlap
        self.out = {c: np.full(shape, np.nan, np.float32) for c in channels}
        self.w = {c: np.zeros(shape, np.float32) for c in channels}
        self.tracker = np.full(shape, -1.0, np.float32) if any(c in self.DISCRETE for c in channels) else None

    def add_tile(self, spec, channel_data: Dict[str, np.ndarray]):
        r0, c0, r1, c1 = spec[:4]
        conf = channel_data.get("confidence")
        for ch, data in channel_data.items():
            if ch not in self.out:
                raise ValueError(f"Unknown channel: {ch}")                # :405-406
            region = self.out[ch][r0:r1, c0:c1]
            if ch in self.DISCRETE and conf is not None and self.tracker is not None:
                creg = self.tracker[r0:r1, c0:c1]
                upd = np.isfinite(data) & ((conf > creg) | np.isnan(region))
                region[upd] = data[upd]
                creg[upd] = conf[upd]
            else:
                wts = blend_2d((r1 - r0, c1 - c0), self.overlap)
                wreg = self.w[ch][r0:r1, c0:c1]
                valid = np.isfinite(data)
                region[np.isnan(region) & valid] = 0.0
                wreg[valid] += wts[valid]
                region[valid] += (data * wts)[valid]

    def finalize(self) -> Dict[str, np.ndarray]:
        res = {}
        for ch, o in self.out.items():
            if ch not in self.DISCRETE:
                v = self.w[ch] > 0
                o[v] /= self.w[ch][v]                                      # :289-292
            res[ch] = o
        return res


def tile_outputs_to_grids(shape, valid_rows, valid_cols, predicted_class, confidence,
                          correction, local_std, floor=0.01):
    """pipeline.py:277-314: class as float (fill 0), confidence (fill 0),
    correction = norm_corr_grid * max(local_std_grid, CORRECTION_NORM_FLOOR)."""
    def grid(v):
        g = np.zeros(shape, np.float32)
        g[valid_rows, valid_cols] = v
        return g
    cls = grid(np.asarray(predicted_class, np.float32))
    conf = grid(np.asarray(confidence, np.float32))
    corr = grid(np.asarray(correction, np.float32)) * np.maximum(grid(np.asarray(local_std, np.float32)),
                                                                np.float32(floor))
    return cls, conf, corr


def finish_survey(depth, valid_mask, results, auto_correct_threshold=0.85):
    """pipeline.py:192-211 + 316-349: unprocessed valid cells -> class 0 / conf 0 /
    corr 0; cleaned = depth - corr where class==2 & conf > thr & valid."""
    res = {k: v.copy() for k, v in results.items()}
    unproc = valid_mask & np.isnan(res["classification"])
    res["classification"][unproc] = 0.0
    res["confidence"][unproc] = 0.0
    res["correction"][unproc] = 0.0
    cleaned = depth.copy()
    with np.errstate(invalid="ignore"):
        apply = (res["classification"] == 2) & (res["confidence"] > auto_correct_threshold) & valid_mask
    cleaned[apply] = depth[apply] - res["correction"][apply]
    res["cleaned_depth"] = cleaned
    res["valid_mask"] = valid_mask.astype(np.float32)
    return res
