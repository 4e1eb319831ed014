"""Drop-in for data/tiling.py."""
from bathymetric_gnn_b200.tiling import Tile, TileSpec, TileManager, TileMerger  # noqa: F401
