"""Drop-in for config/config.py."""
from . import (Config, TileConfig, GraphConfig, ModelConfig, TrainingConfig, SyntheticNoiseConfig,  # noqa: F401
               InferenceConfig)
