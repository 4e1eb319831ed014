"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Plain-Python restatement of the VR-BAG container walk right in front of the ragged hot path,
operating on the two arrays the reference reads from the HDF5 container (`varres_metadata` [R, C] structured,
`varres_refinements` [1, N] structured) instead of on an h5py file:

* `iterate_refinements` -- VRBagHandler.iterate_refinements, /root/reference/data/vr_bag.py:243-298 (row-major walk of the
  base grid, one slice of `varres_refinements[0]` per refined cell, valid-ratio filter).
* `update_refinement_batch` -- VRBagWriter.update_refinement_batch, /root/reference/data/vr_bag.py:550-588.

PINNED bit-for-bit against the UNMODIFIED reference classes run on the h5py stand-in of tests/fake_h5py.py:
tests/golden/vr_bag_case.npz (tests/golden/make_golden.py vr_bag).
"""
from __future__ import annotations

import numpy as np

NODATA = 1.0e6


def valid_mask(depth):
    """RefinementGrid.valid_mask, vr_bag.py:88-92."""
    return (depth != NODATA) & np.isfinite(depth)


def iterate_refinements(meta, ref, min_valid_ratio=0.0):
    """vr_bag.py:257-298.  Yields dicts with the RefinementGrid fields."""
    ref_data = ref[0, :]                                                       # :262
    for row in range(meta.shape[0]):                                            # :265
        for col in range(meta.shape[1]):                                        # :266
            m = meta[row, col]
            dims_x, dims_y = int(m["dimensions_x"]), int(m["dimensions_y"])     # :269-270
            if dims_x == 0 or dims_y == 0:                                      # :272
                continue
            start = int(m["index"])                                             # :275
            sl = ref_data[start:start + dims_x * dims_y]                        # :279
            depth = sl["depth"].reshape(dims_y, dims_x).copy()                  # :282
            unc = sl["depth_uncrt"].reshape(dims_y, dims_x).copy()              # :283
            num_valid = int(np.sum(valid_mask(depth)))
            if num_valid / depth.size >= min_valid_ratio:                       # :295-297
                yield dict(base_row=row, base_col=col, depth=depth, uncertainty=unc,
                           resolution=(float(m["resolution_x"]), float(m["resolution_y"])), dimensions=(dims_y, dims_x),
                           sw_corner=(float(m["sw_corner_x"]), float(m["sw_corner_y"])), start_index=start,
                           num_valid=num_valid)


def update_refinement_batch(ref, grid, corrected_depth, corrected_uncertainty=None):
    """vr_bag.py:566-588 on the in-memory record array; returns the number of changed valid cells (:586-587)."""
    a = grid["start_index"]
    b = a + grid["dimensions"][0] * grid["dimensions"][1]
    cur = ref[0, a:b].copy()                                                    # :573
    cur["depth"] = corrected_depth.flatten()                                    # :576
    if corrected_uncertainty is not None:
        cur["depth_uncrt"] = corrected_uncertainty.flatten()                    # :580
    ref[0, a:b] = cur                                                           # :583
    return int(np.sum((corrected_depth != grid["depth"]) & valid_mask(grid["depth"])))
