"""Drop-in for data/graph_construction.py."""
from bathymetric_gnn_b200.graph import GraphBuilder, MultiScaleGraphBuilder, Data, Batch  # noqa: F401
