"""Drop-in for the reference's top-level `models` package (models/__init__.py)."""
import importlib.util as _u, os as _o
_s = _u.spec_from_file_location("_bgnn_dropin_path", _o.path.join(_o.path.dirname(_o.path.dirname(_o.path.abspath(__file__))), "_path.py"))
_s.loader.exec_module(_u.module_from_spec(_s))
from .gnn import (BathymetricGNN, LocalFeatureExtractor, GNNBackbone, ClassificationHead,  # noqa: E402,F401
                  ConfidenceHead, CorrectionHead)
from .pipeline import BathymetricPipeline  # noqa: E402,F401
from bathymetric_gnn_b200.native_vr import NativeVRProcessor  # noqa: E402,F401

__all__ = ["BathymetricGNN", "LocalFeatureExtractor", "GNNBackbone", "ClassificationHead", "ConfidenceHead",
           "CorrectionHead", "BathymetricPipeline"]
