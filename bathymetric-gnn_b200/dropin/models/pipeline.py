"""Drop-in for models/pipeline.py."""
from bathymetric_gnn_b200.pipeline import BathymetricPipeline  # noqa: F401
