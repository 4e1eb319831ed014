"""Native BAG containers -> device ragged descriptor.

Mirrors the class surface of the reference's data/vr_bag.py (`detect_bag_type` :29-63,
`RefinementGrid` :66-97, `VRBagHandler` :100-311, `SRBagHandler` :315-428, `SRBagWriter`
:431-475, `VRBagWriter` :478-606, `process_vr_bag_native` :837-924) so that
`from data import VRBagHandler, VRBagWriter, SRBagHandler, SRBagWriter` and
`from data.vr_bag import SidecarBuilder, detect_bag_type`
(scripts/inference_native.py:35-36) resolve under the drop-in.

What is new: the reference walks the base grid in a Python double loop and slices
`varres_refinements` once per refinement grid (:264-298); `VRBagHandler.refinement_table()`
reads the container ONCE and returns a `RefinementTable` -- one descriptor row per grid
(`cell_off, rows, cols, res, base_row, base_col, sw_corner, start_index`) plus the packed
depth / uncertainty buffers -- which is exactly what `GraphBuilder.build_graph_packed`
consumes (`bgnn_cells.cell_off / grid_rows / grid_cols / grid_res`, include/bgnn.h).
`VRBagWriter.update_refinements_packed()` is the bulk counterpart of
`update_refinement_batch` (:550-588).  `iterate_refinements` is kept, implemented on the
table, and yields the reference's `RefinementGrid` objects in the reference's order.

HDF5 access is duck-typed: the module named `h5py` is looked up when a file is opened
(`File(path, mode)` with groups / datasets indexed like h5py's), so any object with that
surface works; without one the constructors raise ImportError like the reference.
"""
from __future__ import annotations

import importlib
import logging
import re
import shutil
import sys
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Generator, List, Optional, Tuple

import numpy as np

logger = logging.getLogger(__name__)

NODATA = 1.0e6


def _h5():
    """The HDF5 binding in use: whatever module is registered as `h5py`."""
    mod = sys.modules.get("h5py")
    if mod is None:
        try:
            mod = importlib.import_module("h5py")
        except ImportError as e:
            raise ImportError("h5py is required for native BAG processing") from e
    return mod


def _valid(depth: np.ndarray) -> np.ndarray:
    return (depth != NODATA) & np.isfinite(depth)


def detect_bag_type(path) -> str:
    """'VR' when the BAG carries `varres_metadata` + `varres_refinements` with at least one
    refined base cell, else 'SR' (data/vr_bag.py:29-63)."""
    with _h5().File(Path(path), "r") as f:
        root = f["BAG_root"]
        if "varres_metadata" in root and "varres_refinements" in root:
            if np.any(root["varres_metadata"][:]["dimensions_x"] > 0):
                return "VR"
    return "SR"


@dataclass
class RefinementGrid:
    """One refinement grid (data/vr_bag.py:66-97)."""
    base_row: int
    base_col: int
    depth: np.ndarray
    uncertainty: np.ndarray
    resolution: Tuple[float, float]
    dimensions: Tuple[int, int]
    sw_corner: Tuple[float, float]
    start_index: int

    @property
    def shape(self) -> Tuple[int, int]:
        return self.depth.shape

    @property
    def valid_mask(self) -> np.ndarray:
        return _valid(self.depth)

    @property
    def num_valid(self) -> int:
        return int(np.sum(self.valid_mask))


@dataclass
class RefinementTable:
    """All refinement grids of a container as one ragged descriptor + packed cell buffers.

    Row g describes grid g in PROCESSING ORDER (base cells row-major, data/vr_bag.py:264-266);
    its cells are `depth[cell_off[g]:cell_off[g+1]]` in row-major (rows, cols) order."""
    base_row: np.ndarray        # int32 [G]
    base_col: np.ndarray        # int32 [G]
    rows: np.ndarray            # int32 [G]   (dimensions_y)
    cols: np.ndarray            # int32 [G]   (dimensions_x)
    cell_off: np.ndarray        # int64 [G+1]
    res: np.ndarray             # float64 [G, 2]  (resolution_x, resolution_y)
    sw_corner: np.ndarray       # float64 [G, 2]
    start_index: np.ndarray     # int64 [G]   offset of the grid's first cell in varres_refinements
    depth: np.ndarray           # float32 [cells]
    uncertainty: np.ndarray     # float32 [cells]

    def __len__(self) -> int:
        return int(self.rows.shape[0])

    @property
    def n_cells(self) -> int:
        return int(self.cell_off[-1])

    def source_index(self) -> np.ndarray:
        """int64 [cells]: position of every packed cell in `varres_refinements[0]`."""
        n = np.diff(self.cell_off)
        return np.repeat(self.start_index - self.cell_off[:-1], n) + np.arange(self.n_cells, dtype=np.int64)

    def grid(self, g: int, copy: bool = True) -> RefinementGrid:
        a, b = int(self.cell_off[g]), int(self.cell_off[g + 1])
        shape = (int(self.rows[g]), int(self.cols[g]))
        d, u = self.depth[a:b].reshape(shape), self.uncertainty[a:b].reshape(shape)
        return RefinementGrid(base_row=int(self.base_row[g]), base_col=int(self.base_col[g]),
                              depth=d.copy() if copy else d, uncertainty=u.copy() if copy else u,
                              resolution=(float(self.res[g, 0]), float(self.res[g, 1])), dimensions=shape,
                              sw_corner=(float(self.sw_corner[g, 0]), float(self.sw_corner[g, 1])),
                              start_index=int(self.start_index[g]))

    def grids(self, copy: bool = True) -> List[RefinementGrid]:
        return [self.grid(g, copy) for g in range(len(self))]

    def select(self, keep: np.ndarray) -> "RefinementTable":
        """Sub-table of the grids flagged in `keep` (bool [G]), cells re-packed."""
        keep = np.asarray(keep, bool)
        if keep.all():
            return self
        n = np.diff(self.cell_off)
        cell_keep = np.repeat(keep, n)
        off = np.zeros(int(keep.sum()) + 1, np.int64)
        np.cumsum(n[keep], out=off[1:])
        return RefinementTable(self.base_row[keep], self.base_col[keep], self.rows[keep], self.cols[keep], off,
                               self.res[keep], self.sw_corner[keep], self.start_index[keep],
                               self.depth[cell_keep], self.uncertainty[cell_keep])

    def slice(self, g0: int, g1: int) -> "RefinementTable":
        """Grids [g0, g1) as a table of views (chunking a big container by node budget)."""
        a, b = int(self.cell_off[g0]), int(self.cell_off[g1])
        return RefinementTable(self.base_row[g0:g1], self.base_col[g0:g1], self.rows[g0:g1], self.cols[g0:g1],
                               self.cell_off[g0:g1 + 1] - a, self.res[g0:g1], self.sw_corner[g0:g1],
                               self.start_index[g0:g1], self.depth[a:b], self.uncertainty[a:b])

    def chunks(self, max_cells: int) -> Generator["RefinementTable", None, None]:
        """Consecutive slices of at most `max_cells` cells each (at least one grid per slice)."""
        G, g0 = len(self), 0
        while g0 < G:
            g1 = int(np.searchsorted(self.cell_off, self.cell_off[g0] + max_cells, side="right")) - 1
            g1 = min(max(g1, g0 + 1), G)
            yield self.slice(g0, g1)
            g0 = g1

    def to_device(self, device=None, pin: bool = True):
        """Upload the packed buffers (one H2D copy each from pinned memory); returns (depth, uncertainty)
        CUDA tensors ready for `GraphBuilder.build_graph_packed(depth, cell_off, rows, cols, res, uncertainty)`."""
        import torch
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

        def up(a):
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
            if pin and t.numel():
                t = t.pin_memory()
            return t.to(dev, non_blocking=True)
        return up(self.depth), up(self.uncertainty)


def _table_from_arrays(meta: np.ndarray, ref_depth: np.ndarray, ref_unc: np.ndarray, min_valid_ratio: float) -> RefinementTable:
    """Descriptor + packed buffers from `varres_metadata` [R, C] and the two fields of `varres_refinements[0]`."""
    dx, dy = meta["dimensions_x"].astype(np.int64), meta["dimensions_y"].astype(np.int64)
    br, bc = np.nonzero((dx != 0) & (dy != 0))                  # row-major = the reference's loop order
    rows, cols = dy[br, bc], dx[br, bc]
    n = rows * cols
    start = meta["index"][br, bc].astype(np.int64)
    off = np.zeros(n.shape[0] + 1, np.int64)
    np.cumsum(n, out=off[1:])
    total = int(off[-1])
    if total and np.array_equal(start, start[0] + off[:-1]):    # grids stored back to back in processing order
        a = int(start[0])
        depth, unc = ref_depth[a:a + total], ref_unc[a:a + total]
    else:
        src = np.repeat(start - off[:-1], n) + np.arange(total, dtype=np.int64)
        depth, unc = ref_depth[src], ref_unc[src]
    f64 = lambda name: meta[name][br, bc].astype(np.float64)
    tab = RefinementTable(br.astype(np.int32), bc.astype(np.int32), rows.astype(np.int32), cols.astype(np.int32), off,
                          np.stack([f64("resolution_x"), f64("resolution_y")], axis=1),
                          np.stack([f64("sw_corner_x"), f64("sw_corner_y")], axis=1), start,
                          np.ascontiguousarray(depth, dtype=np.float32), np.ascontiguousarray(unc, dtype=np.float32))
    if len(tab) and min_valid_ratio > 0.0:
        nvalid = np.add.reduceat(_valid(tab.depth).astype(np.int64), off[:-1])      # n >= 1 for every listed grid
        tab = tab.select(nvalid / n >= min_valid_ratio)                            # :295-297
    return tab


class VRBagHandler:
    """Reader of a variable-resolution BAG (data/vr_bag.py:100-311)."""

    NODATA = NODATA
    INVALID_INDEX = 4294967295

    def __init__(self, path):
        h5 = _h5()
        self.path = Path(path)
        with h5.File(str(self.path), "r") as f:
            if "BAG_root" not in f:
                raise ValueError(f"Not a valid BAG file: {self.path}")
            root = f["BAG_root"]
            for name in ("varres_refinements", "varres_metadata"):
                if name not in root:
                    raise ValueError(f"Not a VR BAG (no {name}): {self.path}")
            elevation = root["elevation"]
            self.base_shape = tuple(elevation.shape)
            self.varres_metadata = root["varres_metadata"][:]
            self.min_depth = elevation.attrs.get("Minimum Elevation Value", None)
            self.max_depth = elevation.attrs.get("Maximum Elevation Value", None)
        self.geotransform, self.crs = _gdal_georef(self.path)

    # ---- geometry helpers (:155-199)
    def _refined(self):
        return self.varres_metadata["dimensions_x"] > 0

    @property
    def base_cell_size(self) -> Tuple[float, float]:
        if self.geotransform:
            return (abs(self.geotransform[1]), abs(self.geotransform[5]))
        m = self._refined()
        if np.any(m):
            s = float(np.max(self.varres_metadata["resolution_x"][m] * self.varres_metadata["dimensions_x"][m]))
            return (s, s)
        return (50.0, 50.0)

    @property
    def finest_resolution(self) -> float:
        rx = self.varres_metadata["resolution_x"]
        return float(np.min(rx[rx > 0])) if np.any(rx > 0) else 1.0

    @property
    def bounds(self) -> Tuple[float, float, float, float]:
        gt = self.geotransform
        if gt:
            return (gt[0], gt[3] + self.base_shape[0] * gt[5], gt[0] + self.base_shape[1] * gt[1], gt[3])
        return (0, 0, self.base_shape[1] * 50, self.base_shape[0] * 50)

    @property
    def resampled_shape(self) -> Tuple[int, int]:
        b, r = self.bounds, self.finest_resolution
        return (int(np.ceil((b[3] - b[1]) / r)), int(np.ceil((b[2] - b[0]) / r)))

    @property
    def num_refinement_cells(self) -> int:
        return int(np.sum(self._refined()))

    @property
    def total_refinement_nodes(self) -> int:
        md = self.varres_metadata
        return int(np.sum(md["dimensions_x"].astype(np.int64) * md["dimensions_y"].astype(np.int64)))

    def get_refinement_info(self) -> Dict:
        md, m = self.varres_metadata, self._refined()
        return {
            "base_shape": self.base_shape,
            "num_refined_cells": int(np.sum(m)),
            "total_refinement_nodes": self.total_refinement_nodes,
            "unique_dimensions": sorted(set(zip(md["dimensions_x"][m].flatten(), md["dimensions_y"][m].flatten()))),
            "unique_resolutions": sorted(set(md["resolution_x"][m].flatten())),
        }

    # ---- the bulk path
    def refinement_table(self, min_valid_ratio: float = 0.0) -> RefinementTable:
        with _h5().File(str(self.path), "r") as f:
            ref = f["BAG_root"]["varres_refinements"][0, :]
        return _table_from_arrays(self.varres_metadata, ref["depth"], ref["depth_uncrt"], min_valid_ratio)

    def iterate_refinements(self, min_valid_ratio: float = 0.0) -> Generator[RefinementGrid, None, None]:
        """:243-298, same grids in the same order (each with its own copies of depth / uncertainty)."""
        tab = self.refinement_table(min_valid_ratio)
        for g in range(len(tab)):
            yield tab.grid(g)

    def copy_and_open_for_writing(self, output_path) -> "VRBagWriter":
        logger.info("Copying VR BAG: %s -> %s", self.path, output_path)
        shutil.copy(str(self.path), str(output_path))
        return VRBagWriter(output_path)


def _gdal_georef(path):
    """(geotransform, crs) through GDAL when it is installed, else (None, None) (:140-153)."""
    try:
        from osgeo import gdal
    except ImportError:
        return None, None
    ds = gdal.Open(str(path))
    if not ds:
        return None, None
    return ds.GetGeoTransform(), ds.GetProjection()


class VRBagWriter:
    """Write-back into (a copy of) a VR BAG (data/vr_bag.py:478-606)."""

    NODATA = NODATA

    def __init__(self, path):
        self.path = Path(path)
        self._file = _h5().File(str(self.path), "r+")
        self._root = self._file["BAG_root"]
        self._refinements = self._root["varres_refinements"]
        self._uncertainty = self._root["uncertainty"] if "uncertainty" in self._root else None
        self._metadata = self._root["varres_metadata"][:]
        self._corrections_applied = 0
        self._uncertainty_updates = 0

    @staticmethod
    def _check(grid, corrected_depth):
        if corrected_depth.shape != grid.shape:
            raise ValueError(f"Shape mismatch: corrected {corrected_depth.shape} vs grid {grid.shape}")

    def update_refinement(self, grid: RefinementGrid, corrected_depth: np.ndarray, corrected_uncertainty: Optional[np.ndarray] = None):
        """:505-548 (the reference writes cell by cell; one slice write stores the same values)."""
        self.update_refinement_batch(grid, corrected_depth, corrected_uncertainty)
        if corrected_uncertainty is not None:
            self._uncertainty_updates += int(np.sum((corrected_uncertainty != grid.uncertainty) & grid.valid_mask))

    def update_refinement_batch(self, grid: RefinementGrid, corrected_depth: np.ndarray,
                                corrected_uncertainty: Optional[np.ndarray] = None):
        """:550-588: read-modify-write of the grid's record slice."""
        self._check(grid, corrected_depth)
        a = grid.start_index
        b = a + grid.dimensions[0] * grid.dimensions[1]
        cur = self._refinements[0, a:b]
        cur["depth"] = corrected_depth.reshape(-1)
        if corrected_uncertainty is not None:
            cur["depth_uncrt"] = corrected_uncertainty.reshape(-1)
        self._refinements[0, a:b] = cur
        self._corrections_applied += int(np.sum((corrected_depth != grid.depth) & grid.valid_mask))

    def update_refinements_packed(self, table: RefinementTable, corrected_depth, corrected_uncertainty=None):
        """Bulk write-back of a whole table (or chunk): `corrected_*` are packed like `table.depth` (NumPy arrays or
        tensors).  Equivalent to `update_refinement_batch` for every grid of the table, with one read-modify-write of
        the record range the table touches instead of one per grid."""
        if len(table) == 0:
            return
        host = lambda a: np.ascontiguousarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float32).reshape(-1)
        d = host(corrected_depth)
        if d.shape[0] != table.n_cells:
            raise ValueError(f"Shape mismatch: corrected {d.shape[0]} cells vs table {table.n_cells}")
        src = table.source_index()
        lo, hi = int(src.min()), int(src.max()) + 1
        cur = self._refinements[0, lo:hi]
        cur["depth"][src - lo] = d
        if corrected_uncertainty is not None:
            cur["depth_uncrt"][src - lo] = host(corrected_uncertainty)
        self._refinements[0, lo:hi] = cur
        self._corrections_applied += int(np.sum((d != table.depth) & _valid(table.depth)))

    def close(self):
        if self._file is not None:
            self._file.close()
            self._file = None
        logger.info("VR BAG modifications complete:")
        logger.info("  - Depth corrections applied: %s", f"{self._corrections_applied:,}")
        if self._uncertainty_updates > 0:
            logger.info("  - Uncertainty updates: %s", f"{self._uncertainty_updates:,}")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class SRBagHandler:
    """Single-resolution BAG behind the VR handler's interface: the elevation grid is the one
    "refinement" (data/vr_bag.py:315-428)."""

    NODATA = NODATA
    _RES = re.compile(r"<gmd:resolution>.*?<gco:Measure[^>]*>([0-9.]+)</gco:Measure>", re.DOTALL)

    def __init__(self, path):
        h5 = _h5()
        self.path = Path(path)
        with h5.File(self.path, "r") as f:
            root = f["BAG_root"]
            self._depth = root["elevation"][:].astype(np.float32)
            self._uncertainty = (root["uncertainty"][:].astype(np.float32) if "uncertainty" in root
                                 else np.zeros_like(self._depth))
            self._shape = self._depth.shape
            xml = root["metadata"][()]
            if hasattr(xml, "tobytes"):
                xml = xml.tobytes()
            if isinstance(xml, bytes):
                xml = xml.decode("utf-8")
        m = self._RES.search(xml)
        if m:
            self._resolution = float(m.group(1))
        else:
            gt, _ = _gdal_georef(self.path)
            self._resolution = abs(gt[1]) if gt else 1.0
        self.geotransform, self.crs = None, None
        logger.info("Opened SR BAG: %s  shape %s  resolution %sm", self.path, self._shape, self._resolution)

    @property
    def base_shape(self) -> Tuple[int, int]:
        return self._shape

    def get_refinement_info(self) -> Dict:
        return {"base_shape": self._shape, "num_refined_cells": 1,
                "total_refinement_nodes": int(np.sum(_valid(self._depth))), "unique_resolutions": [self._resolution]}

    def refinement_table(self, min_valid_ratio: float = 0.0) -> RefinementTable:
        H, W = self._shape
        keep = np.sum(_valid(self._depth)) / self._depth.size >= min_valid_ratio
        G = 1 if keep else 0
        r = self._resolution
        z = lambda dt: np.zeros(G, dt)
        return RefinementTable(z(np.int32), z(np.int32), np.full(G, H, np.int32), np.full(G, W, np.int32),
                               np.array([0, H * W][:G + 1], np.int64), np.full((G, 2), r, np.float64), np.zeros((G, 2)),
                               z(np.int64), self._depth.reshape(-1)[:G * H * W].copy(),
                               self._uncertainty.reshape(-1)[:G * H * W].copy())

    def iterate_refinements(self, min_valid_ratio: float = 0.0) -> Generator[RefinementGrid, None, None]:
        if np.sum(_valid(self._depth)) / self._depth.size >= min_valid_ratio:
            yield RefinementGrid(base_row=0, base_col=0, depth=self._depth.copy(), uncertainty=self._uncertainty.copy(),
                                 resolution=(self._resolution, self._resolution), dimensions=self._shape,
                                 sw_corner=(0.0, 0.0), start_index=0)

    def copy_and_open_for_writing(self, output_path) -> "SRBagWriter":
        logger.info("Copying SR BAG: %s -> %s", self.path, output_path)
        shutil.copy(str(self.path), str(output_path))
        return SRBagWriter(output_path)


class SRBagWriter:
    """data/vr_bag.py:431-475."""

    def __init__(self, path):
        self.path = Path(path)
        self.file = _h5().File(self.path, "r+")
        self.root = self.file["BAG_root"]

    def update_refinement_batch(self, grid, corrected_depth: np.ndarray, corrected_uncertainty: np.ndarray):
        self.root["elevation"][:] = corrected_depth
        if "uncertainty" in self.root:
            self.root["uncertainty"][:] = corrected_uncertainty

    def update_refinements_packed(self, table: RefinementTable, corrected_depth, corrected_uncertainty=None):
        if len(table) == 0:
            return
        host = lambda a: np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float32)
        shape = (int(table.rows[0]), int(table.cols[0]))
        self.update_refinement_batch(None, host(corrected_depth).reshape(shape),
                                     None if corrected_uncertainty is None else host(corrected_uncertainty).reshape(shape))

    def close(self):
        if self.file is not None:
            self.file.close()
            self.file = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def process_vr_bag_native(input_path, output_path, process_func, min_valid_ratio: float = 0.01,
                          confidence_for_uncertainty: bool = True) -> Dict:
    """Legacy per-grid driver (data/vr_bag.py:837-924): `process_func(depth, uncertainty, resolution)` ->
    (classification, confidence, correction).  Kept with its own conventions (`confidence > 0.5`, `depth += correction`)."""
    handler = VRBagHandler(input_path)
    info = handler.get_refinement_info()
    writer = handler.copy_and_open_for_writing(output_path)
    stats = dict(grids_processed=0, grids_skipped=0, cells_processed=0, cells_corrected=0, total_grids=info["num_refined_cells"])
    try:
        for grid in handler.iterate_refinements(min_valid_ratio):
            cls, conf, corr = process_func(grid.depth, grid.uncertainty, grid.resolution)
            depth, unc = grid.depth.copy(), grid.uncertainty.copy()
            apply = (cls == 2) & grid.valid_mask & (conf > 0.5)
            if np.any(apply):
                depth[apply] += corr[apply]
                stats["cells_corrected"] += int(np.sum(apply))
                if confidence_for_uncertainty:
                    unc[apply] *= 2.0 - conf[apply]
            writer.update_refinement_batch(grid, depth, unc)
            stats["grids_processed"] += 1
            stats["cells_processed"] += grid.num_valid
    finally:
        writer.close()
    return stats
