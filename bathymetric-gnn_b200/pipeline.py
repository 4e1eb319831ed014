"""`BathymetricPipeline` with the reference's constructor / `load_model` /
`process` / `_process_tile` signatures (models/pipeline.py:36-382), driving the
device hot path: per tile  graph build -> predict -> scatter -> stitch, all on
one CUDA stream with the survey and its accumulators resident in HBM.

Differences that are NOT behaviour changes: the reference's Blackwell "no kernel
image -> CPU" fallback (pipeline.py:72-88) is deliberately absent (no CPU
fallback on this path); tiles are consumed as strided device views instead of
host copies; the dead `cleaned_depth` blend (merged then overwritten,
pipeline.py:210-211) is skipped.  File I/O stays behind a loader / writer object
(GDAL is outside the hot path): any object with the reference's
`load(path) -> BathymetricGrid` / `save(grid, path, additional_bands=)` plugs in;
the built-in pair reads / writes .npz.
"""
from __future__ import annotations

import ctypes as C
import logging
from pathlib import Path
from typing import Dict, Optional, Union

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .config import Config, CORRECTION_NORM_FLOOR
from .gnn import BathymetricGNN
from .graph import GraphBuilder
from .tiling import BathymetricGrid, TileManager, TileMerger, TileSpec

logger = logging.getLogger(__name__)


class NpzLoader:
    """Minimal loader: .npz with `depth` (+ optional `uncertainty`, `nodata`, `resolution`)."""

    def __init__(self, vr_bag_mode: str = "resampled"):
        self.vr_bag_mode = vr_bag_mode

    def load(self, path) -> BathymetricGrid:
        z = np.load(Path(path))
        depth = np.asarray(z["depth"], np.float32)
        res = tuple(float(v) for v in z["resolution"]) if "resolution" in z.files else (1.0, 1.0)
        return BathymetricGrid(depth=depth, uncertainty=np.asarray(z["uncertainty"], np.float32) if "uncertainty" in z.files else None,
                               nodata_value=float(z["nodata"]) if "nodata" in z.files else 1.0e6, resolution=res,
                               bounds=(0.0, 0.0, depth.shape[1] * res[0], depth.shape[0] * res[1]), source_path=Path(path))


class NpzWriter:
    def save(self, grid: BathymetricGrid, path, additional_bands: Optional[Dict[str, np.ndarray]] = None):
        out = dict(depth=grid.depth, nodata=np.float64(grid.nodata_value), resolution=np.array(grid.resolution))
        if grid.uncertainty is not None:
            out["uncertainty"] = grid.uncertainty
        out.update(additional_bands or {})
        np.savez_compressed(Path(path), **out)


def _reference_io(name: str):
    """The reference's own GDAL `BathymetricLoader` / `BathymetricWriter` (data/loaders.py:93-822), loaded unmodified from
    a reference checkout (file formats are outside the accelerated path and are not re-implemented)."""
    from .dropin import _passthrough
    ns = {"__name__": "data._reference_loaders", "__package__": "data"}
    _passthrough.exec_reference("data/loaders.py", ns)
    if not ns.get("GDAL_AVAILABLE", False):
        raise ImportError("GDAL is required for loading bathymetric data. Install with: conda install -c conda-forge gdal")
    return ns[name]


class FormatDispatchLoader:
    """`BathymetricLoader(vr_bag_mode)` of the drop-in (data/loaders.py:98-118 signature): .npz is read here, every other
    format (BAG / GeoTIFF / ASC) goes to the reference's GDAL loader."""

    def __init__(self, vr_bag_mode: str = "resampled"):
        self.vr_bag_mode = vr_bag_mode

    def load(self, path) -> BathymetricGrid:
        if Path(path).suffix.lower() == ".npz":
            return NpzLoader(self.vr_bag_mode).load(path)
        return _reference_io("BathymetricLoader")(vr_bag_mode=self.vr_bag_mode).load(path)


class FormatDispatchWriter:
    def save(self, grid: BathymetricGrid, path, additional_bands: Optional[Dict[str, np.ndarray]] = None):
        if Path(path).suffix.lower() == ".npz":
            return NpzWriter().save(grid, path, additional_bands)
        return _reference_io("BathymetricWriter")().save(grid, path, additional_bands=additional_bands)


class BathymetricPipeline:
    def __init__(self, config: Config, vr_bag_mode: str = "resampled", loader=None, writer=None):
        self.config = config
        self.loader = loader if loader is not None else FormatDispatchLoader(vr_bag_mode)
        self.writer = writer if writer is not None else FormatDispatchWriter()
        self.tile_manager = TileManager(config.tile.tile_size, config.tile.overlap, config.tile.min_valid_ratio)
        self.graph_builder = GraphBuilder(connectivity=config.graph.connectivity, edge_features=config.graph.edge_features)
        self.model: Optional[BathymetricGNN] = None
        if config.device not in ("cuda",) and not str(config.device).startswith("cuda"):
            raise RuntimeError(f"config.device={config.device!r}: the B200 pipeline runs on CUDA only (no CPU fallback)")
        _lib.require_cuda()
        self.device = torch.device(config.device if ":" in str(config.device) else f"cuda:{torch.cuda.current_device()}")
        logger.info("Pipeline initialized, device: %s", self.device)

    # ------------------------------------------------------------------ model
    def load_model(self, model_path: Union[str, Path]):
        """models/pipeline.py:92-132 (checkpoint layout: training/trainer.py:811-822)."""
        model_path = Path(model_path)
        if not model_path.exists():
            raise FileNotFoundError(f"Model not found: {model_path}")
        ckpt = torch.load(model_path, map_location="cpu", weights_only=False)
        mc = ckpt.get("model_config", self.config.model)
        self.model = BathymetricGNN(in_channels=ckpt.get("in_channels", 7), hidden_channels=mc.gnn_hidden_channels,
                                    num_gnn_layers=mc.gnn_num_layers, gnn_type=mc.gnn_type, heads=mc.gnn_heads,
                                    num_classes=mc.num_classes, predict_correction=mc.predict_correction, dropout=0.0,
                                    edge_dim=ckpt.get("edge_dim", 3))
        self.model.load_state_dict(ckpt["model_state_dict"])
        self.model.to(self.device)
        self.model.eval()
        logger.info("Model loaded from %s", model_path)

    # ------------------------------------------------------------------ per tile
    def _tile_on_device(self, depth_view, unc_view, mask, resolution, nodata, min_valid_ratio=None, into=None, sink=None):
        """graph build -> predict -> scatter; returns (device tile grids [class, conf, corr] or None, N).
        `into=(merger, spec)`: the scatter is fused into the stitch accumulate instead (returns (True, N));
        `sink(device_graph, outputs)`: the caller consumes the per-node outputs itself (survey.py; returns (True, N))."""
        graph = self.graph_builder.build_graph(depth_view, mask, unc_view, resolution, nodata=nodata)
        shape = tuple(depth_view.shape)
        if graph.num_nodes == 0:
            return None, 0
        if min_valid_ratio is not None and graph.num_nodes / (shape[0] * shape[1]) < min_valid_ratio:
            return None, graph.num_nodes                     # tile will be skipped (tiling.py:200)
        out = self.model.predict(graph, auto_correct_threshold=self.config.inference.auto_correct_threshold,
                                 review_threshold=self.config.inference.review_threshold)
        dg = graph._dev
        if sink is not None:
            sink(dg, out)
            return True, graph.num_nodes
        if into is not None:
            merger, spec = into
            merger.add_tile_nodes(spec, dg.node_id, out["predicted_class"], out["confidence"], out.get("correction"),
                                  dg.local_std, float(CORRECTION_NORM_FLOOR))
            return True, graph.num_nodes
        dev = dg.device
        n = shape[0] * shape[1]
        grids = torch.empty((3, shape[0], shape[1]), dtype=torch.float32, device=dev)
        corr = out.get("correction")
        with torch.cuda.device(dev):
            check(lib.bgnn_scatter_tile(ptr(dg.node_id), n, ptr(out["predicted_class"]), ptr(out["confidence"]), ptr(corr),
                                        ptr(dg.local_std), float(CORRECTION_NORM_FLOOR), ptr(grids[0]), ptr(grids[1]),
                                        ptr(grids[2]) if corr is not None else None, stream_ptr()))
            if corr is None:
                grids[2].zero_()
        return grids, graph.num_nodes

    def _process_tile(self, tile, grid: BathymetricGrid) -> Dict[str, np.ndarray]:
        """models/pipeline.py:243-314: host tile in, host result grids out (drop-in signature)."""
        if self.model is None:
            raise RuntimeError("Model not loaded. Call load_model() first.")
        with torch.cuda.device(self.device):
            grids, n = self._tile_on_device(torch.from_numpy(np.ascontiguousarray(tile.data, dtype=np.float32)).to(self.device),
                                            tile.uncertainty, tile.valid_mask, grid.resolution, None)
        if grids is None:
            z = lambda: np.zeros(tile.shape, dtype=np.float32)
            return {"cleaned_depth": tile.data, "classification": z(), "confidence": z(), "correction": z()}
        host = grids.cpu().numpy()
        return {"cleaned_depth": tile.data, "classification": host[0], "confidence": host[1], "correction": host[2]}

    # ------------------------------------------------------------------ survey
    def process_grid(self, grid: BathymetricGrid, specs=None, keep_on_device: bool = False):
        """The body of `process` without file I/O: survey resident in HBM, tiles as strided views,
        results stitched on the device.  `specs` (optional) restricts the run to a subset of the
        tile list (multi-GPU sharding, see parallel.py); returns the result dict."""
        if self.model is None:
            raise RuntimeError("Model not loaded. Call load_model() first.")
        dev = self.device
        H, W = grid.shape
        tm = self.tile_manager
        if specs is None and not keep_on_device and not isinstance(grid.depth, torch.Tensor):
            return self._process_grid_streamed(grid)
        with torch.cuda.device(dev):
            def to_dev(a):          # host arrays are uploaded once; a survey already resident in HBM is used in place
                if isinstance(a, torch.Tensor):
                    return a.to(device=dev, dtype=torch.float32)
                return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
            depth = to_dev(grid.depth)
            unc = None if grid.uncertainty is None else to_dev(grid.uncertainty)
            nodata = grid.nodata_value
            use_nodata = nodata is not None and not np.isnan(nodata)
            valid = torch.isfinite(depth)
            if use_nodata:
                valid &= depth != float(nodata)
            valid_u8 = valid.view(torch.uint8)
            merger = TileMerger(tm, device=dev)
            merger.initialize((H, W), ["classification", "confidence", "correction"])
            if specs is None:
                _, _, specs = tm.compute_tile_grid((H, W))
            num_tiles = 0
            for spec in specs:
                sl = (slice(spec.row_start, spec.row_end), slice(spec.col_start, spec.col_end))
                dv = depth[sl]
                # valid = isfinite & != nodata is derived in-kernel (== grid.valid_mask, loaders.py:58-69)
                grids, n = self._tile_on_device(dv, None if unc is None else unc[sl], None, grid.resolution,
                                                float(nodata) if use_nodata else None, tm.min_valid_ratio,
                                                into=(merger, spec))
                size = dv.shape[0] * dv.shape[1]
                if n / size < tm.min_valid_ratio:            # tiling.py:200 (skip_empty)
                    continue
                if grids is None:                            # pipeline.py:257-264 (only when min_valid_ratio <= 0)
                    grids = torch.zeros((3,) + tuple(dv.shape), dtype=torch.float32, device=dev)
                    merger.add_tile(spec, {"classification": grids[0], "confidence": grids[1], "correction": grids[2]})
                num_tiles += 1
            logger.info("Processed %d tiles (%d skipped below min_valid_ratio)", num_tiles, len(specs) - num_tiles)
            if keep_on_device == "partial":                  # sharded runs gather unfinalised band planes (parallel.py)
                return merger.partial_state()
            res = merger.finalize_device()
            cleaned = torch.empty_like(depth)
            check(lib.bgnn_apply_corrections(ptr(depth), ptr(valid_u8), H * W, ptr(res["classification"]), ptr(res["confidence"]),
                                             ptr(res["correction"]), ptr(cleaned), float(self.config.inference.auto_correct_threshold),
                                             stream_ptr()))
            res["cleaned_depth"] = cleaned
            res["valid_mask"] = valid.to(torch.float32)
            if keep_on_device:
                return res
            return {k: v.cpu().numpy() for k, v in res.items()}

    def _process_grid_streamed(self, grid: BathymetricGrid) -> Dict[str, np.ndarray]:
        """Host grid in, host planes out through the streaming survey driver (survey.py): rows go up tile-row by
        tile-row, finalised rows come back while later tiles are still computing.  Same results, bit for bit, as the
        resident path above (tests/test_gpu_survey.py)."""
        from . import survey as sv
        H, W = grid.shape
        pin = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        plan, res = sv.process_survey(self, pin(grid.depth), pin(grid.uncertainty), grid.nodata_value, grid.resolution)
        logger.info("Processed %d tiles (%d skipped below min_valid_ratio)", res["stats"]["tiles"], res["stats"]["skipped"])
        out = {"classification": sv.classification_as_float(res["classification"]), "confidence": res["confidence"].numpy(),
               "correction": res["correction"].numpy(), "cleaned_depth": res["cleaned_depth"].numpy()}
        out["valid_mask"] = grid.valid_mask.astype(np.float32)
        return out

    def process(self, input_path, output_path, export_extras: bool = True) -> Dict[str, np.ndarray]:
        """models/pipeline.py:134-241."""
        if self.model is None:
            raise RuntimeError("Model not loaded. Call load_model() first.")
        input_path, output_path = Path(input_path), Path(output_path)
        grid = self.loader.load(input_path)
        logger.info("Loaded grid: %s, resolution: %s", grid.shape, grid.resolution)
        results = self.process_grid(grid)
        out_grid = BathymetricGrid(depth=results["cleaned_depth"], uncertainty=grid.uncertainty, nodata_value=grid.nodata_value,
                                   transform=grid.transform, crs=grid.crs, resolution=grid.resolution, bounds=grid.bounds,
                                   source_path=input_path)
        bands = None
        if export_extras:
            bands = {k: results[k] for k in ("classification", "confidence", "correction", "valid_mask")}
        self.writer.save(out_grid, output_path, additional_bands=bands)
        self._log_summary(grid, results)
        return results

    def _apply_corrections(self, grid: BathymetricGrid, results: Dict[str, np.ndarray]) -> np.ndarray:
        """models/pipeline.py:316-349 on host arrays (kept for API parity; `process` fuses it on device)."""
        dev = self.device
        with torch.cuda.device(dev):
            d = torch.from_numpy(np.ascontiguousarray(grid.depth, dtype=np.float32)).to(dev)
            v = torch.from_numpy(grid.valid_mask).to(dev).view(torch.uint8)
            c, f, r = (torch.from_numpy(np.ascontiguousarray(results[k], dtype=np.float32)).to(dev)
                       for k in ("classification", "confidence", "correction"))
            # NaN classes stay NaN here: the fill of unprocessed cells is part of process() in the reference
            cn = torch.where(torch.isnan(c), torch.full_like(c, -1.0), c)
            out = torch.empty_like(d)
            check(lib.bgnn_apply_corrections(ptr(d), ptr(v), d.numel(), ptr(cn), ptr(f), ptr(r), ptr(out),
                                             float(self.config.inference.auto_correct_threshold), stream_ptr()))
        return out.cpu().numpy()

    def _log_summary(self, grid: BathymetricGrid, results: Dict[str, np.ndarray]):
        vm = grid.valid_mask
        n = max(int(vm.sum()), 1)
        cl, cf = results["classification"], results["confidence"]
        logger.info("Total valid cells: %d", n)
        for k, name in enumerate(("Smooth seafloor", "Features", "Noise")):
            c = int(((cl == k) & vm).sum())
            logger.info("  - %s: %d (%.1f%%)", name, c, 100.0 * c / n)
        logger.info("Mean confidence: %.3f", float(cf[vm].mean()) if vm.any() else 0.0)
