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
    BATCH_NODE_BUDGET = 50000          # scripts/inference_native.py:128 (raise it on B200: see DESIGN.md)

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

    def _run(self, grids, nodata):
        """Build + predict + scatter for a list of (depth, unc, res); returns per-grid result tuples."""
        b = self.graph_builder
        with torch.cuda.device(self.device):
            use_unc = all(g[1] is not None for g in grids)
            graph = b.build_graph_batch(grids, nodata=nodata, use_uncertainty=use_unc)
            shapes = [g[0].shape for g in grids]
            if graph.num_nodes == 0:
                return [tuple(np.zeros(s, np.float32) for _ in range(3)) for s in shapes]
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
        self._pending.append((depth, self._use_unc(uncertainty), resolution))
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
