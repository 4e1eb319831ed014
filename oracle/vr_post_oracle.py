"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

NumPy restatement of the native VR-BAG post-processing right behind the ragged hot path:

* `apply_results`  -- the `apply_results` closure of /root/reference/scripts/inference_native.py:480-503 (nested in
  `main()`, so it cannot be imported: restated line by line; ** parity unpinned ** beyond that reading -- the arithmetic
  is three NumPy expressions).
* `SidecarOracle.add_refinement_results` -- SidecarBuilder.add_refinement_results, /root/reference/data/vr_bag.py:695-778,
  with the 4-deep Python loop vectorised.  PINNED bit-for-bit against the unmodified reference method in
  tests/golden/sidecar_case.npz (tests/golden/make_golden.py sidecar; the reference object is built without its GDAL
  georeferencing step, which is file I/O outside the path).
"""
from __future__ import annotations

import numpy as np

CLASS_NOISE = 2
NODATA = 1.0e6


def valid_mask(depth, nodata=NODATA):
    """RefinementGrid.valid_mask, vr_bag.py:90-93."""
    return (depth != nodata) & np.isfinite(depth)


def apply_results(depth, uncertainty, classification, confidence, correction, auto_correct_threshold=0.85, nodata=NODATA):
    """inference_native.py:485-502.  Returns (corrected_depth, corrected_uncertainty, stats)."""
    valid = valid_mask(depth, nodata)
    cd, cu = depth.copy(), uncertainty.copy()
    noise_mask = (classification == CLASS_NOISE) & valid                    # :487
    high_conf = confidence >= auto_correct_threshold                        # :488
    apply_mask = noise_mask & high_conf                                     # :489
    if np.any(apply_mask):
        cd[apply_mask] -= correction[apply_mask]                            # :492
        scale = 2.0 - confidence[apply_mask]                                # :494
        cu[apply_mask] *= scale                                             # :495
    stats = dict(cells_corrected=int(np.sum(apply_mask)), cells_classified_noise=int(np.sum(noise_mask)),
                 cells_processed=int(np.sum(valid)), total_confidence=float(np.sum(confidence[valid])))
    return cd, cu, stats


def placement(shape, geotransform, bounds, base_shape, base_row, base_col, sw_corner, dimensions, resolution, raster_resolution):
    """vr_bag.py:716-748: (out_row_start, out_col_start, scale) of one refinement grid, float64 + int(round(.))."""
    base_cell_width = (bounds[2] - bounds[0]) / base_shape[1]
    base_cell_height = (bounds[3] - bounds[1]) / base_shape[0]
    base_x = bounds[0] + base_col * base_cell_width
    base_y = bounds[1] + base_row * base_cell_height
    ref_x = base_x + sw_corner[0]
    ref_y = base_y + sw_corner[1]
    ref_height = dimensions[0] * resolution[1]
    gt = geotransform
    out_col_start = int(round((ref_x - gt[0]) / gt[1]))
    ref_top_y = ref_y + ref_height
    out_row_start = int(round((gt[3] - ref_top_y) / abs(gt[5])))
    scale = max(1, int(round(resolution[0] / raster_resolution)))
    return out_row_start, out_col_start, scale


class SidecarOracle:
    """SidecarBuilder state (vr_bag.py:628-632) + add_refinement_results."""

    def __init__(self, shape, geotransform, bounds, base_shape):
        self.shape, self.geotransform, self.bounds, self.base_shape = tuple(shape), tuple(geotransform), tuple(bounds), tuple(base_shape)
        self.resolution = abs(self.geotransform[1])
        self.classification = np.full(self.shape, np.nan, dtype=np.float32)
        self.confidence = np.full(self.shape, np.nan, dtype=np.float32)
        self.correction = np.full(self.shape, np.nan, dtype=np.float32)
        self.valid_mask = np.zeros(self.shape, dtype=np.float32)

    def add_refinement_results(self, grid, classification, confidence, correction):
        rows, cols = grid.dimensions
        orow, ocol, scale = placement(self.shape, self.geotransform, self.bounds, self.base_shape, grid.base_row, grid.base_col,
                                      grid.sw_corner, grid.dimensions, grid.resolution, self.resolution)
        r = np.arange(rows)[:, None, None, None]
        c = np.arange(cols)[None, :, None, None]
        dr = np.arange(scale)[None, None, :, None]
        dc = np.arange(scale)[None, None, None, :]
        out_r = np.broadcast_to(orow + (rows - 1 - r) * scale + dr, (rows, cols, scale, scale))
        out_c = np.broadcast_to(ocol + c * scale + dc, (rows, cols, scale, scale))
        ok = (out_r >= 0) & (out_r < self.shape[0]) & (out_c >= 0) & (out_c < self.shape[1])
        bc = lambda a: np.broadcast_to(np.asarray(a)[:, :, None, None], (rows, cols, scale, scale))
        rr, cc = out_r[ok], out_c[ok]
        self.classification[rr, cc] = bc(classification)[ok]
        self.confidence[rr, cc] = bc(confidence)[ok]
        self.correction[rr, cc] = bc(correction)[ok]
        v = bc(grid.valid_mask)[ok]
        self.valid_mask[rr[v], cc[v]] = 1.0
