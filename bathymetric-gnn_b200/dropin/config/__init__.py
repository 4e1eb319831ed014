"""Drop-in for the reference's top-level `config` package (config/__init__.py:1-22)."""
import importlib.util as _u, os as _o
_s = _u.spec_from_file_location("_bgnn_dropin_path", _o.path.join(_o.path.dirname(_o.path.dirname(_o.path.abspath(__file__))), "_path.py"))
_s.loader.exec_module(_u.module_from_spec(_s))
from .config import (Config, TileConfig, GraphConfig, ModelConfig, TrainingConfig,  # noqa: E402,F401
                     SyntheticNoiseConfig, InferenceConfig)
from .constants import CORRECTION_NORM_FLOOR, CORRECTION_NORM_CAP  # noqa: E402,F401

__all__ = ["Config", "TileConfig", "GraphConfig", "ModelConfig", "TrainingConfig", "SyntheticNoiseConfig",
           "InferenceConfig", "CORRECTION_NORM_FLOOR", "CORRECTION_NORM_CAP"]
