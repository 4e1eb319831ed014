"""Drop-in for the reference's top-level `data` package (data/__init__.py): the classes on the
inference hot path.  File I/O classes (BathymetricLoader / Writer, VRBagHandler, ...) are outside the
path: use the reference's own loaders next to this package, or the built-in .npz pair."""
import importlib.util as _u, os as _o
_s = _u.spec_from_file_location("_bgnn_dropin_path", _o.path.join(_o.path.dirname(_o.path.dirname(_o.path.abspath(__file__))), "_path.py"))
_s.loader.exec_module(_u.module_from_spec(_s))
from bathymetric_gnn_b200.tiling import BathymetricGrid, Tile, TileSpec, TileManager, TileMerger  # noqa: E402,F401
from bathymetric_gnn_b200.graph import GraphBuilder, Data, Batch  # noqa: E402,F401
from bathymetric_gnn_b200.pipeline import NpzLoader as BathymetricLoader, NpzWriter as BathymetricWriter  # noqa: E402,F401
