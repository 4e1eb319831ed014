"""Drop-in for config/config.py."""
from bathymetric_gnn_b200.config import (Config, TileConfig, GraphConfig, ModelConfig, TrainingConfig,  # noqa: F401
                                         SyntheticNoiseConfig, InferenceConfig)
