"""bathymetric-gnn_b200: B200 (sm_100a) implementation of the Bathymetric-GNN
per-tile inference hot path (graph build -> GAT forward -> tile stitch).

The directory name carries a hyphen (task layout); import it as
`bathymetric_gnn_b200` (the sibling shim package extends its __path__ here).
"""
from .graph import Data, Batch, GridGraph, HostGridGraph, GraphBuilder, MultiScaleGraphBuilder, OFFSETS_4, OFFSETS_8, OFFSETS_24  # noqa: F401
from .gnn import (BathymetricGNN, LocalFeatureExtractor, GNNBackbone, ClassificationHead,  # noqa: F401
                  ConfidenceHead, CorrectionHead)

from .config import (Config, TileConfig, GraphConfig, ModelConfig, TrainingConfig, SyntheticNoiseConfig,  # noqa: F401
                     InferenceConfig, CORRECTION_NORM_FLOOR, CORRECTION_NORM_CAP)
from .tiling import BathymetricGrid, Tile, TileSpec, TileManager, TileMerger  # noqa: F401
from .pipeline import BathymetricPipeline, NpzLoader, NpzWriter  # noqa: F401
from .native_vr import NativeVRProcessor, SidecarBuilder, apply_results, apply_results_packed  # noqa: F401

__version__ = "0.1.0"
