#!/usr/bin/env python3
"""Stand-in for the reference's scripts/test_pipeline.py (BASELINE configs[0]'s driver): the same in-function imports
and the calls of its test_tiling / test_graph_construction / test_model_forward bodies; see ../README.md."""
import torch
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).parent.parent))

import numpy as np


def check_imports():
    import torch_geometric
    return torch_geometric.__version__


def test_tiling(grid):
    from data import TileManager
    tile_manager = TileManager(tile_size=1024, overlap=128, min_valid_ratio=0.1)
    n_rows, n_cols, specs = tile_manager.compute_tile_grid(grid.shape)
    sample = None
    for tile in tile_manager.iterate_tiles(grid, skip_empty=True):
        if tile.valid_ratio >= 0.1:
            sample = tile
            break
    return sample, len(specs)


def test_graph_construction(tile, resolution):
    import torch
    from data import GraphBuilder
    graph_builder = GraphBuilder(connectivity="8-connected", edge_features=["distance", "depth_difference", "slope"])
    start = time.time()
    graph = graph_builder.build_graph(depth=tile.data, valid_mask=tile.valid_mask, uncertainty=tile.uncertainty,
                                      resolution=resolution)
    elapsed = time.time() - start
    assert not torch.isnan(graph.x).any() and not torch.isinf(graph.x).any()
    return graph, dict(nodes=graph.num_nodes, edges=graph.num_edges, x=tuple(graph.x.shape),
                       edge_attr=tuple(graph.edge_attr.shape), seconds=elapsed)


def test_model_forward(graph, device_override=None):
    import torch
    from models import BathymetricGNN
    device = torch.device(device_override) if device_override else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    in_channels = graph.x.shape[1]
    edge_dim = graph.edge_attr.shape[1] if graph.edge_attr is not None else None
    model = BathymetricGNN(in_channels=in_channels, hidden_channels=64, num_gnn_layers=4, gnn_type="GAT", heads=4,
                           num_classes=3, predict_correction=True, dropout=0.1, edge_dim=edge_dim)
    model.to(device)
    num_params = sum(p.numel() for p in model.parameters())
    graph = graph.to(device)
    model.eval()
    start = time.time()
    with torch.no_grad():
        outputs = model(graph)
    elapsed = time.time() - start
    classes, counts = torch.unique(outputs["predicted_class"], return_counts=True)
    return model, outputs, dict(params=num_params, seconds=elapsed, classes=classes.tolist(), counts=counts.tolist(),
                                mean_confidence=float(outputs["confidence"].mean()))
