"""Drop-in for the reference's top-level `data` package (data/__init__.py:1-33): every name it exports.
Hot-path classes are the B200 implementations; `BathymetricLoader` / `BathymetricWriter` (GDAL) and the synthetic noise
classes are outside the path: the former dispatch on the file type (.npz built in, other formats through the
reference's own loaders when a checkout + GDAL are present), the latter are passed through from the checkout."""
import importlib.util as _u, os as _o
_s = _u.spec_from_file_location("_bgnn_dropin_path", _o.path.join(_o.path.dirname(_o.path.dirname(_o.path.abspath(__file__))), "_path.py"))
_s.loader.exec_module(_u.module_from_spec(_s))
from .loaders import BathymetricGrid, BathymetricLoader, BathymetricWriter  # noqa: E402,F401
from .tiling import Tile, TileSpec, TileManager, TileMerger  # noqa: E402,F401
from .graph_construction import GraphBuilder, MultiScaleGraphBuilder, Data, Batch  # noqa: E402,F401
from .vr_bag import (VRBagHandler, VRBagWriter, SRBagHandler, SRBagWriter, RefinementGrid, RefinementTable,  # noqa: E402,F401
                     SidecarBuilder, process_vr_bag_native, detect_bag_type)

_NOISE = ("SyntheticNoiseGenerator", "NoiseAugmentor", "NoiseLabel")
__all__ = ["BathymetricGrid", "BathymetricLoader", "BathymetricWriter", "Tile", "TileSpec", "TileManager", "TileMerger",
           "GraphBuilder", "MultiScaleGraphBuilder", *_NOISE, "VRBagHandler", "VRBagWriter", "SRBagHandler", "SRBagWriter",
           "RefinementGrid", "SidecarBuilder", "process_vr_bag_native", "detect_bag_type"]


def __getattr__(name):          # data/synthetic_noise.py is loaded from the reference checkout on first use
    if name in _NOISE:
        from . import synthetic_noise
        return getattr(synthetic_noise, name)
    raise AttributeError(f"module 'data' has no attribute {name!r}")
