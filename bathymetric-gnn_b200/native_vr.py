"""`NativeVRProcessor`: ragged batching of native VR-BAG refinement grids
(scripts/inference_native.py:117-342) on the device.

Same public surface (`process_grid`, `add_to_batch`, `batch_ready`,
`batch_pending`, `flush_batch`, `BATCH_NODE_BUDGET`).  Where the reference builds
one graph per grid on the CPU and concatenates them with PyG
`Batch.from_data_list`, a flush here packs the pending grids into ONE ragged cell
space, builds the block-diagonal graph with the same kernels as a tile (per-grid
boundary conditions), runs one forward and scatters the results back into the
packed cell space in a single pass; per-grid results are views of one host copy.
"""
from __future__ import annotations

import logging
from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .config import CORRECTION_NORM_FLOOR
from .gnn import BathymetricGNN
from .graph import GraphBuilder

logger = logging.getLogger(__name__)


class NativeVRProcessor:
    CLASS_NOISE = 2
    REFERENCE_BATCH_NODE_BUDGET = 50000    # scripts/inference_native.py:128
    # Nodes per flush.  The reference's 50 000 keeps a CPU / small-GPU forward in cache; on a B200 a flush of that size is
    # ~0.3 ms of kernels behind ~1 ms of per-flush host work, so the default here is 40x larger (eval-mode BatchNorm makes
    # the per-grid results independent of how grids are batched; 2 M nodes = ~16 MB of pending host arrays).  It stays a
    # plain class attribute like the reference's: scripts that set it keep working.
    BATCH_NODE_BUDGET = 2_000_000

    def __init__(self, model: BathymetricGNN, graph_builder: GraphBuilder, device=None, auto_correct_threshold: float = 0.85):
        self.model, self.graph_builder = model, graph_builder
        _lib.require_cuda()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise RuntimeError("NativeVRProcessor (B200 build) runs on CUDA only (no CPU fallback)")
        self.auto_correct_threshold = auto_correct_threshold
        self.model.eval()
        try:
            self.expected_in_channels = model.feature_extractor.mlp[0].in_features      # :147
            logger.info("Model expects %d input features", self.expected_in_channels)
        except (AttributeError, IndexError):
            self.expected_in_channels = None
        self._pending: List[Tuple[np.ndarray, Optional[np.ndarray], tuple]] = []
        self._batch_node_count = 0
        self._nodata = 1.0e6

    # ------------------------------------------------------------------ helpers
    def _valid(self, depth, nodata):
        return (depth != nodata) & np.isfinite(depth)                                  # :160

    def _use_unc(self, uncertainty):
        return None if self.expected_in_channels == 7 else uncertainty                  # :165-167

    def _run(self, grids, nodata, packed: bool = False):
        """Build + predict + scatter for a list of (depth, unc, res); returns per-grid result tuples
        (`packed=True`: the device-resident batch instead, see flush_batch_packed)."""
        b = self.graph_builder
        with torch.cuda.device(self.device):
            use_unc = all(g[1] is not None for g in grids)
            graph = b.build_graph_batch(grids, nodata=nodata, use_uncertainty=use_unc)
            shapes = [g[0].shape for g in grids]
            if graph.num_nodes == 0 and not packed:
                return [tuple(np.zeros(s, np.float32) for _ in range(3)) for s in shapes]
            if packed:
                return self._run_packed(graph, grids, shapes, nodata)
            out = self.model.predict(graph, auto_correct_threshold=self.auto_correct_threshold)
            dg = graph._dev
            planes = torch.empty((3, dg.n_cells), dtype=torch.float32, device=self.device)
            corr = out.get("correction")
            check(lib.bgnn_scatter_tile(ptr(dg.node_id), dg.n_cells, ptr(out["predicted_class"]), ptr(out["confidence"]),
                                        ptr(corr), ptr(dg.local_std), float(CORRECTION_NORM_FLOOR), ptr(planes[0]),
                                        ptr(planes[1]), ptr(planes[2]) if corr is not None else None, stream_ptr()))
            if corr is None:
                planes[2].zero_()
            host = planes.cpu().numpy()
        res, off = [], 0
        for s in shapes:
            n = s[0] * s[1]
            res.append(tuple(host[k, off:off + n].reshape(s) for k in range(3)))
            off += n
        return res

    def _run_packed(self, graph, grids, shapes, nodata):
        off = np.zeros(len(shapes) + 1, np.int64)
        np.cumsum([r * c for r, c in shapes], out=off[1:])
        n = int(off[-1])
        planes = torch.zeros((3, n), dtype=torch.float32, device=self.device)
        if graph.num_nodes > 0:
            out = self.model.predict(graph, auto_correct_threshold=self.auto_correct_threshold)
            dg = graph._dev
            corr = out.get("correction")
            check(lib.bgnn_scatter_tile(ptr(dg.node_id), dg.n_cells, ptr(out["predicted_class"]), ptr(out["confidence"]),
                                        ptr(corr), ptr(dg.local_std), float(CORRECTION_NORM_FLOOR), ptr(planes[0]),
                                        ptr(planes[1]), ptr(planes[2]) if corr is not None else None, stream_ptr()))
        cat = lambda k: torch.from_numpy(np.concatenate([np.asarray(g[k], np.float32).ravel() for g in grids])).to(self.device)
        depth = cat(0)
        raw = lambda g: g[3] if len(g) > 3 else g[1]
        unc = (torch.from_numpy(np.concatenate([np.asarray(raw(g), np.float32).ravel() for g in grids])).to(self.device)
               if all(raw(g) is not None for g in grids) else None)
        return dict(cell_off=off, shapes=shapes, depth=depth, uncertainty=unc, classification=planes[0], confidence=planes[1],
                    correction=planes[2], nodata=nodata)

    def flush_batch_packed(self):
        """Like flush_batch, but the results STAY ON THE DEVICE as packed planes over the batch's cells (dict with `cell_off`,
        `shapes`, `depth`, `uncertainty`, `classification`, `confidence`, `correction`): the hand-over to
        `apply_results_packed` and `SidecarBuilder.add_batch`, which replace the per-grid host loop of
        scripts/inference_native.py:480-503 / data/vr_bag.py:695-778."""
        if not self._pending:
            return None
        grids, self._pending, self._batch_node_count = self._pending, [], 0
        # the processor drops the uncertainty plane for 7-feature models (:165-167); post-processing still needs it
        return self._run(grids, self._nodata, packed=True)

    # ------------------------------------------------------------------ reference API
    def process_grid(self, depth, uncertainty, resolution, nodata: float = 1.0e6):
        """:206-247 (unbatched)."""
        depth = np.asarray(depth)
        if not np.any(self._valid(depth, nodata)):
            return tuple(np.zeros_like(depth, dtype=np.float32) for _ in range(3))
        return self._run([(depth, self._use_unc(uncertainty), resolution)], nodata)[0]

    def add_to_batch(self, depth, uncertainty, resolution, nodata=1.0e6):
        """:249-269: returns empty results immediately for grids without valid data, else None."""
        depth = np.asarray(depth)
        valid = self._valid(depth, nodata)
        nvalid = int(np.count_nonzero(valid))
        if nvalid == 0:
            return tuple(np.zeros_like(depth, dtype=np.float32) for _ in range(3))
        if min(depth.shape) < 2:
            # the reference raises here, inside build_graph -> np.gradient (graph_construction.py:284)
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least (edge_order + 1) elements are required.")
        if self._pending and nodata != self._nodata:
            raise ValueError("all grids of a batch must share the nodata value")
        self._nodata = nodata
        self._pending.append((depth, self._use_unc(uncertainty), resolution, uncertainty))   # [3]: raw plane for post-processing
        self._batch_node_count += nvalid
        return None

    @property
    def batch_ready(self) -> bool:
        return self._batch_node_count >= self.BATCH_NODE_BUDGET

    @property
    def batch_pending(self) -> bool:
        return len(self._pending) > 0

    def flush_batch(self):
        """:281-342: one ragged forward for everything accumulated; results in insertion order."""
        if not self._pending:
            return []
        grids, self._pending, self._batch_node_count = self._pending, [], 0
        return self._run(grids, self._nodata)


def apply_results_packed(depth: torch.Tensor, uncertainty: Optional[torch.Tensor], classification: torch.Tensor,
                         confidence: torch.Tensor, correction: torch.Tensor, auto_correct_threshold: float = 0.85,
                         nodata: float = 1.0e6):
    """`apply_results` of scripts/inference_native.py:480-503 for a whole packed batch ON THE DEVICE (one kernel,
    bgnn_vr_apply_results): returns (corrected_depth, corrected_uncertainty or None, stats) with the reference's stats keys
    (`cells_corrected`, `cells_classified_noise`, `cells_processed`, `total_confidence`).  All tensors are f32 CUDA tensors
    over the packed cells of the batch."""
    _lib.require_cuda()
    dev = depth.device
    with torch.cuda.device(dev):
        f = lambda t: None if t is None else t.to(device=dev, dtype=torch.float32).contiguous().view(-1)
        d, u, cl, cf, cr = f(depth), f(uncertainty), f(classification), f(confidence), f(correction)
        n = int(d.numel())
        if not (cl.numel() == n and cf.numel() == n and cr.numel() == n and (u is None or u.numel() == n)):
            raise ValueError("apply_results_packed: all planes must cover the same packed cells")
        od = torch.empty_like(d)
        ou = torch.empty_like(u) if u is not None else None
        counts = torch.zeros(3, dtype=torch.int64, device=dev)
        csum = torch.zeros(1, dtype=torch.float64, device=dev)
        check(lib.bgnn_vr_apply_results(ptr(d), ptr(u), n, float(nodata), ptr(cl), ptr(cf), ptr(cr), float(auto_correct_threshold),
                                        ptr(od), ptr(ou), ptr(counts), ptr(csum), stream_ptr()))
        c = counts.cpu().tolist()
        stats = dict(cells_corrected=int(c[0]), cells_classified_noise=int(c[1]), cells_processed=int(c[2]),
                     total_confidence=float(csum.cpu()[0]))
    return od, ou, stats


class SidecarBuilder:
    """Device-resident counterpart of data/vr_bag.py:608-790 (`SidecarBuilder`): accumulates classification / confidence /
    correction of processed refinement grids at the raster resolution.

    The reference takes its georeferencing from GDAL's resampled view of the BAG (`_init_from_gdal_resampled`, :638-693:
    file I/O, outside the path); here it is passed explicitly (`shape`, `geotransform`, `bounds`, `base_shape`) or read from a
    handler object exposing those attributes.  `add_refinement_results(grid, classification, confidence, correction)` keeps
    the reference signature; `add_batch` paints a whole packed batch with one kernel pair (bgnn_sidecar_add).  Later grids
    overwrite earlier ones exactly as in the reference's sequential loop (owner plane, include/bgnn.h).  `save` (GeoTIFF
    export through GDAL, :780-) is not built: read the planes and hand them to any writer."""

    def __init__(self, handler=None, *, shape=None, geotransform=None, bounds=None, base_shape=None, crs=None, device=None):
        _lib.require_cuda()
        g = lambda name, v: v if v is not None else getattr(handler, name, None)
        self.handler = handler
        self.shape, self.geotransform, self.bounds = g("shape", shape), g("geotransform", geotransform), g("bounds", bounds)
        self.base_shape, self.crs = g("base_shape", base_shape), g("crs", crs)
        if handler is not None and (self.shape is None or self.geotransform is None):
            self._georef_from_handler(handler)
        if self.shape is None or self.geotransform is None or self.bounds is None or self.base_shape is None:
            raise RuntimeError("SidecarBuilder (B200 build) needs shape, geotransform, bounds and base_shape "
                               "(pass them, or a handler that carries them)")
        self.shape = (int(self.shape[0]), int(self.shape[1]))
        self.resolution = abs(self.geotransform[1])                                     # :681
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            self._planes = torch.empty((4,) + self.shape, dtype=torch.float32, device=self.device)
            self._owner = torch.empty(self.shape, dtype=torch.int32, device=self.device)
            self._sc = _lib.Sidecar()
            self._sc.rows, self._sc.cols = self.shape
            (self._sc.classification, self._sc.confidence, self._sc.correction, self._sc.valid_mask) = (self._planes[k].data_ptr() for k in range(4))
            self._sc.owner = self._owner.data_ptr()
            check(lib.bgnn_sidecar_init(self._sc, stream_ptr()))
        self._seq = 0
        logger.info("SidecarBuilder initialized: %s at %sm", self.shape, self.resolution)

    def _georef_from_handler(self, handler):
        """`SidecarBuilder(handler)` as the entry script calls it (scripts/inference_native.py:461).  With GDAL installed
        the raster geometry comes from GDAL's resampled view of the container like in the reference (data/vr_bag.py:638-693:
        file I/O); without it, from the handler's own extent at its finest refinement resolution (the reference has no
        GDAL-free mode and raises ImportError there)."""
        try:
            from osgeo import gdal
            ds = gdal.OpenEx(str(handler.path), gdal.OF_RASTER | gdal.OF_READONLY, open_options=["MODE=RESAMPLED_GRID"]) \
                or gdal.Open(str(handler.path))
        except ImportError:
            ds = None
        if ds is not None:
            self.shape, self.geotransform, self.crs = (ds.RasterYSize, ds.RasterXSize), ds.GetGeoTransform(), ds.GetProjection()
            gt = self.geotransform
            self.bounds = (gt[0], gt[3] + self.shape[0] * gt[5], gt[0] + self.shape[1] * gt[1], gt[3])
            return
        b = tuple(float(v) for v in handler.bounds)
        res = float(getattr(handler, "finest_resolution", None) or getattr(handler, "_resolution", 1.0))
        shape = getattr(handler, "resampled_shape", None) or handler.base_shape
        self.shape, self.bounds = (int(shape[0]), int(shape[1])), b
        self.geotransform = (b[0], res, 0.0, b[3], 0.0, -res)

    # placement integers exactly as the reference computes them (float64, int(round(.))): vr_bag.py:716-748
    def _placement(self, base_row, base_col, sw_corner, dimensions, resolution):
        b, gt = self.bounds, self.geotransform
        base_cell_width = (b[2] - b[0]) / self.base_shape[1]
        base_cell_height = (b[3] - b[1]) / self.base_shape[0]
        base_x = b[0] + base_col * base_cell_width
        base_y = b[1] + base_row * base_cell_height
        ref_x = base_x + sw_corner[0]
        ref_y = base_y + sw_corner[1]
        ref_height = dimensions[0] * resolution[1]
        out_col_start = int(round((ref_x - gt[0]) / gt[1]))
        ref_top_y = ref_y + ref_height
        out_row_start = int(round((gt[3] - ref_top_y) / abs(gt[5])))
        scale = max(1, int(round(resolution[0] / self.resolution)))
        return out_row_start, out_col_start, scale

    def add_batch(self, grids, cell_off, depth, classification, confidence, correction, nodata: float = 1.0e6):
        """Paint a packed batch: `grids` = objects with the RefinementGrid location attributes (base_row, base_col, sw_corner,
        dimensions, resolution) in PROCESSING ORDER; `cell_off` int64 [G+1]; the planes are f32 CUDA tensors over the packed
        cells (e.g. from `NativeVRProcessor.flush_batch_packed`), `depth` the packed ORIGINAL depths (valid mask)."""
        G = len(grids)
        if G == 0:
            return
        desc = np.empty((G, _lib.SIDECAR_DESC), np.int32)
        for i, g in enumerate(grids):
            orow, ocol, scale = self._placement(g.base_row, g.base_col, g.sw_corner, g.dimensions, g.resolution)
            desc[i] = (g.dimensions[0], g.dimensions[1], orow, ocol, scale, self._seq + i)
        self._seq += G
        off = np.ascontiguousarray(cell_off.cpu().numpy() if isinstance(cell_off, torch.Tensor) else cell_off, dtype=np.int64)
        if off.shape[0] != G + 1 or np.any(np.diff(off) != desc[:, 0].astype(np.int64) * desc[:, 1]):
            raise ValueError("add_batch: cell_off does not match the grid dimensions")
        with torch.cuda.device(self.device):
            f = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous().view(-1)
            d, cl, cf, cr = f(depth), f(classification), f(confidence), f(correction)
            n = int(off[-1])
            if not (d.numel() == n and cl.numel() == n and cf.numel() == n and cr.numel() == n):
                raise ValueError("add_batch: planes do not cover the packed cells")
            t_off, t_desc = torch.from_numpy(off).to(self.device), torch.from_numpy(desc).to(self.device)
            check(lib.bgnn_sidecar_add(self._sc, G, ptr(t_off), ptr(t_desc), n, ptr(cl), ptr(cf), ptr(cr), ptr(d), float(nodata),
                                       stream_ptr()))

    def add_refinement_results(self, grid, classification, confidence, correction):
        """data/vr_bag.py:695-778 (one grid, host arrays)."""
        rows, cols = grid.dimensions
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).reshape(-1)).to(self.device)
        self.add_batch([grid], np.array([0, rows * cols], np.int64), up(grid.depth), up(classification), up(confidence), up(correction))

    classification = property(lambda self: self._planes[0].cpu().numpy())
    confidence = property(lambda self: self._planes[1].cpu().numpy())
    correction = property(lambda self: self._planes[2].cpu().numpy())
    valid_mask = property(lambda self: self._planes[3].cpu().numpy())

    def save(self, path):
        """data/vr_bag.py:780-834: four float32 bands (classification, confidence, correction, valid_mask).  GeoTIFF through
        GDAL when it is installed; `.npz` paths (and any path when GDAL is missing, where the reference only logs an error)
        get the same four planes + the geotransform as an .npz next to the requested name."""
        from pathlib import Path
        path = Path(path)
        bands = [("classification", self.classification), ("confidence", self.confidence),
                 ("correction", self.correction), ("valid_mask", self.valid_mask)]
        gdal = None
        if path.suffix.lower() != ".npz":
            try:
                from osgeo import gdal
            except ImportError:
                logger.error("GDAL required for GeoTIFF export; writing %s instead", path.with_suffix(".npz"))
        if gdal is None:
            np.savez_compressed(path.with_suffix(".npz"), geotransform=np.array(self.geotransform, np.float64), **dict(bands))
            return
        ds = gdal.GetDriverByName("GTiff").Create(str(path), self.shape[1], self.shape[0], len(bands), gdal.GDT_Float32,
                                                  options=["COMPRESS=LZW", "TILED=YES"])
        ds.SetGeoTransform(self.geotransform)
        if self.crs:
            ds.SetProjection(self.crs)
        for i, (name, plane) in enumerate(bands, 1):
            band = ds.GetRasterBand(i)
            band.WriteArray(plane)
            band.SetDescription(name)
        ds.FlushCache()


def apply_results(depth, uncertainty, valid_mask, classification, confidence, correction, auto_correct_threshold=0.85):
    """Post-processing of scripts/inference_native.py:480-503 (`>=` threshold, unlike the tiled path):
    returns (corrected_depth, corrected_uncertainty, n_corrected).  Host NumPy: it runs on the tiny
    per-grid arrays right before the HDF5 write-back, outside the device path."""
    cd, cu = depth.copy(), uncertainty.copy()
    apply = (classification == NativeVRProcessor.CLASS_NOISE) & valid_mask & (confidence >= auto_correct_threshold)
    if np.any(apply):
        cd[apply] -= correction[apply]
        cu[apply] *= (2.0 - confidence[apply])
    return cd, cu, int(apply.sum())
