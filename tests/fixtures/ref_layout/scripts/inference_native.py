#!/usr/bin/env python3
"""Stand-in for the reference's scripts/inference_native.py (same import block and the same calls into the packages:
per-grid GraphBuilder.build_graph, PyG-style Batch.from_data_list, model.predict, graph_to_grid, handler / writer /
sidecar; see ../README.md).  Like the shipped script, the processor logs through a module-level `logger` that the
script itself never defines."""
import torch
import sys
import argparse
import logging
from pathlib import Path

import numpy as np

# Add parent to path for imports
sys.path.insert(0, str(Path(__file__).parent.parent))

try:
    from torch_geometric.data import Batch as PyGBatch
    PYG_BATCH_AVAILABLE = True
except ImportError:
    PYG_BATCH_AVAILABLE = False

from config import Config
from config.constants import CORRECTION_NORM_FLOOR
from models.gnn import BathymetricGNN
from data import GraphBuilder, VRBagHandler, VRBagWriter, SRBagHandler, SRBagWriter
from data.vr_bag import SidecarBuilder, detect_bag_type


class ScriptSideProcessor:
    """The calls the reference's in-script NativeVRProcessor makes (inference_native.py:158-342), nothing more."""
    CLASS_NOISE = 2
    BATCH_NODE_BUDGET = 50000

    def __init__(self, model, graph_builder, device, auto_correct_threshold=0.85):
        self.model, self.builder, self.device, self.thr = model, graph_builder, device, auto_correct_threshold
        self.model.eval()
        self.expected_in_channels = model.feature_extractor.mlp[0].in_features
        logger.info(f"Model expects {self.expected_in_channels} input features")     # noqa: F821 (pre-seeded by the runner)
        self.graphs, self.shapes, self.nodes = [], [], 0

    def add_to_batch(self, depth, uncertainty, resolution, nodata=1.0e6):
        valid = (depth != nodata) & np.isfinite(depth)
        graph = None
        if np.any(valid):
            graph = self.builder.build_graph(depth=depth, valid_mask=valid,
                                             uncertainty=None if self.expected_in_channels == 7 else uncertainty,
                                             resolution=resolution)
        if graph is None or graph.num_nodes == 0:
            return tuple(np.zeros_like(depth, dtype=np.float32) for _ in range(3))
        self.graphs.append(graph); self.shapes.append(depth.shape); self.nodes += graph.num_nodes
        return None

    batch_ready = property(lambda self: self.nodes >= self.BATCH_NODE_BUDGET)

    def _to_grids(self, graph, out, shape):
        g2g = self.builder.graph_to_grid
        cls = g2g(graph, out["predicted_class"].float(), fill_value=0.0)
        conf = g2g(graph, out["confidence"], fill_value=0.0)
        corr = np.zeros(shape, dtype=np.float32)
        if "correction" in out:
            std = g2g(graph, graph.local_std if hasattr(graph, "local_std") else torch.zeros(graph.num_nodes), fill_value=0.0)
            corr = g2g(graph, out["correction"], fill_value=0.0) * np.maximum(std, CORRECTION_NORM_FLOOR)
        return cls, conf, corr

    def flush_batch(self):
        graphs, shapes, self.graphs, self.shapes, self.nodes = self.graphs, self.shapes, [], [], 0
        if not graphs:
            return []
        if len(graphs) == 1 or not PYG_BATCH_AVAILABLE:
            res = []
            for g, s in zip(graphs, shapes):
                g = g.to(self.device)
                with torch.no_grad():
                    out = self.model.predict(g, auto_correct_threshold=self.thr)
                res.append(self._to_grids(g.cpu(), {k: v.cpu() for k, v in out.items()}, s))
            return res
        batch = PyGBatch.from_data_list(graphs).to(self.device)
        with torch.no_grad():
            out = self.model.predict(batch, auto_correct_threshold=self.thr)
        out = {k: v.cpu() if isinstance(v, torch.Tensor) else v for k, v in out.items()}
        ids = batch.cpu().batch.numpy()
        res = []
        for i, (g, s) in enumerate(zip(graphs, shapes)):
            sel = ids == i
            res.append(self._to_grids(g, {k: (v[sel] if isinstance(v, torch.Tensor) and v.shape[0] == len(ids) else v)
                                          for k, v in out.items()}, s))
        return res


def parse_args():
    p = argparse.ArgumentParser()
    for flag in ("--input", "--model", "--output"):
        p.add_argument(flag, type=Path, required=True)
    p.add_argument("--config", type=Path, default=None)
    p.add_argument("--min-valid-ratio", type=float, default=0.01)
    p.add_argument("--auto-correct-threshold", type=float, default=0.85)
    p.add_argument("--device", type=str, default="cuda")
    p.add_argument("--export-sidecar", action="store_true", default=True)
    p.add_argument("--no-export-sidecar", action="store_false", dest="export_sidecar")
    p.add_argument("--log-level", type=str, default="INFO")
    return p.parse_args()


def main():
    args = parse_args()
    logging.basicConfig(level=getattr(logging, args.log_level.upper()))
    args.output.parent.mkdir(parents=True, exist_ok=True)
    config = Config()
    device = torch.device("cuda" if args.device == "cuda" and torch.cuda.is_available() else "cpu")
    checkpoint = torch.load(args.model, map_location=device, weights_only=False)
    mc = checkpoint["config"].model
    model = BathymetricGNN(in_channels=checkpoint.get("in_channels", 7), hidden_channels=mc.gnn_hidden_channels,
                           num_gnn_layers=mc.gnn_num_layers, gnn_type=mc.gnn_type, heads=mc.gnn_heads,
                           num_classes=mc.num_classes, predict_correction=mc.predict_correction, dropout=0.0,
                           edge_dim=checkpoint.get("edge_dim", 3))
    model.load_state_dict(checkpoint["model_state_dict"])
    model.to(device)
    model.eval()
    builder = GraphBuilder(connectivity=config.graph.connectivity, edge_features=config.graph.edge_features)
    processor = ScriptSideProcessor(model=model, graph_builder=builder, device=device,
                                    auto_correct_threshold=args.auto_correct_threshold)
    handler = VRBagHandler(args.input) if detect_bag_type(args.input) == "VR" else SRBagHandler(args.input)
    info = handler.get_refinement_info()
    writer = handler.copy_and_open_for_writing(args.output)
    assert isinstance(writer, (VRBagWriter, SRBagWriter))
    sidecar = SidecarBuilder(handler) if args.export_sidecar else None
    stats = dict(grids_processed=0, cells_processed=0, cells_classified_noise=0, cells_corrected=0, total_confidence=0.0)
    nodata = getattr(handler, "NODATA", 1.0e6)
    pending = []

    def apply_results(grid, classification, confidence, correction):
        if sidecar is not None:
            sidecar.add_refinement_results(grid, classification, confidence, correction)
        depth, unc = grid.depth.copy(), grid.uncertainty.copy()
        noise = (classification == processor.CLASS_NOISE) & grid.valid_mask
        apply = noise & (confidence >= args.auto_correct_threshold)
        if np.any(apply):
            depth[apply] -= correction[apply]
            unc[apply] *= 2.0 - confidence[apply]
            stats["cells_corrected"] += int(np.sum(apply))
        writer.update_refinement_batch(grid, depth, unc)
        stats["grids_processed"] += 1
        stats["cells_processed"] += grid.num_valid
        stats["cells_classified_noise"] += int(np.sum(noise))
        stats["total_confidence"] += float(np.sum(confidence[grid.valid_mask]))

    def flush_pending():
        results = iter(processor.flush_batch())
        for grid, immediate in pending:
            apply_results(grid, *(immediate if immediate is not None else next(results)))
        pending.clear()

    try:
        for grid in handler.iterate_refinements(args.min_valid_ratio):
            pending.append((grid, processor.add_to_batch(grid.depth, grid.uncertainty, grid.resolution, nodata=nodata)))
            if processor.batch_ready:
                flush_pending()
        flush_pending()
    finally:
        writer.close()
    if sidecar is not None:
        sidecar.save(args.output.with_name(args.output.stem + "_gnn_outputs.tif"))
    stats["info"] = info
    return stats, sidecar


if __name__ == "__main__":
    STATS, SIDECAR = main()
