"""Drop-in for data/tiling.py."""
from . import Tile, TileSpec, TileManager, TileMerger  # noqa: F401
