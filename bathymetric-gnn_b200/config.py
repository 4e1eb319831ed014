"""Configuration tree with the reference's field names and defaults
(config/config.py:12-222, config/constants.py).  Entry points read
`config.tile.*`, `config.graph.*`, `config.model.*`, `config.inference.*`,
`config.device`, so the names are part of the drop-in boundary.
Additive knob: graph.connectivity also accepts "24-connected" (BASELINE config 5).
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Optional, Tuple

import yaml

CORRECTION_NORM_FLOOR = 0.01      # config/constants.py:12
CORRECTION_NORM_CAP = 50.0        # config/constants.py:18


@dataclass
class TileConfig:
    tile_size: int = 1024
    overlap: int = 128
    min_valid_ratio: float = 0.1


@dataclass
class GraphConfig:
    connectivity: str = "8-connected"
    max_edge_distance: float = 2.0
    include_self_loops: bool = False
    edge_features: List[str] = field(default_factory=lambda: ["distance", "depth_difference", "slope"])


@dataclass
class ModelConfig:
    local_feature_channels: int = 32
    local_feature_layers: int = 3
    local_kernel_size: int = 5
    gnn_type: str = "GAT"
    gnn_hidden_channels: int = 64
    gnn_num_layers: int = 4
    gnn_heads: int = 4
    gnn_dropout: float = 0.1
    num_classes: int = 3
    predict_correction: bool = True


@dataclass
class TrainingConfig:
    learning_rate: float = 1e-3
    weight_decay: float = 1e-4
    batch_size: int = 4
    epochs: int = 100
    scheduler: str = "cosine"
    warmup_epochs: int = 5
    patience: int = 15
    min_delta: float = 1e-4
    classification_weight: float = 1.0
    correction_weight: float = 0.5
    confidence_weight: float = 0.2
    class_weights: Optional[List[float]] = None
    augment_rotations: bool = True
    augment_flips: bool = True
    augment_noise_intensity: bool = True


@dataclass
class SyntheticNoiseConfig:
    enable_gaussian: bool = True
    enable_spikes: bool = True
    enable_blobs: bool = True
    enable_systematic: bool = True
    gaussian_std_range: Tuple[float, float] = (0.1, 0.5)
    spike_magnitude_range: Tuple[float, float] = (1.0, 5.0)
    spike_density_range: Tuple[float, float] = (0.001, 0.01)
    blob_size_range: Tuple[int, int] = (3, 15)
    blob_count_range: Tuple[int, int] = (5, 50)
    noise_complexity_correlation: float = 0.3


@dataclass
class InferenceConfig:
    auto_correct_threshold: float = 0.85
    review_threshold: float = 0.6
    export_classification: bool = True
    export_confidence: bool = True
    export_correction_magnitude: bool = True
    export_review_priority: bool = True


_SUB = dict(tile=TileConfig, graph=GraphConfig, model=ModelConfig, training=TrainingConfig,
            noise=SyntheticNoiseConfig, inference=InferenceConfig)


@dataclass
class Config:
    tile: TileConfig = field(default_factory=TileConfig)
    graph: GraphConfig = field(default_factory=GraphConfig)
    model: ModelConfig = field(default_factory=ModelConfig)
    training: TrainingConfig = field(default_factory=TrainingConfig)
    noise: SyntheticNoiseConfig = field(default_factory=SyntheticNoiseConfig)
    inference: InferenceConfig = field(default_factory=InferenceConfig)
    data_dir: Optional[str] = None
    output_dir: Optional[str] = None
    model_path: Optional[str] = None
    device: str = "cuda"
    num_workers: int = 4
    pin_memory: bool = True
    log_level: str = "INFO"
    wandb_project: Optional[str] = None
    wandb_entity: Optional[str] = None

    def __post_init__(self):
        assert self.tile.tile_size >= self.tile.overlap * 2, "Tile size must be at least 2x overlap"
        assert self.model.gnn_type in ["GCN", "GAT", "GraphSAGE", "GIN"], f"Unknown GNN type: {self.model.gnn_type}"
        assert self.graph.connectivity in ["4-connected", "8-connected", "24-connected"], \
            f"Unknown connectivity: {self.graph.connectivity}"

    def save(self, path):
        def plain(o):
            if isinstance(o, dict):
                return {k: plain(v) for k, v in o.items()}
            if isinstance(o, (list, tuple)):
                return [plain(v) for v in o]
            return o
        with open(Path(path), "w") as f:
            yaml.dump(plain(dataclasses.asdict(self)), f, default_flow_style=False)

    @classmethod
    def load(cls, path) -> "Config":
        with open(Path(path)) as f:
            data = yaml.safe_load(f) or {}
        kw = {}
        for name, klass in _SUB.items():
            sub = dict(data.get(name) or {})
            for fld in dataclasses.fields(klass):       # YAML has no tuples: restore them from the field type
                if fld.name in sub and isinstance(sub[fld.name], list) and str(fld.type).startswith("Tuple"):
                    sub[fld.name] = tuple(sub[fld.name])
            known = {f.name for f in dataclasses.fields(klass)}
            kw[name] = klass(**{k: v for k, v in sub.items() if k in known})
        for k in ("data_dir", "output_dir", "model_path", "device", "num_workers", "pin_memory", "log_level",
                  "wandb_project", "wandb_entity"):
            if k in data:
                kw[k] = data[k]
        return cls(**kw)
