"""Drop-in for models/pipeline.py."""
from . import BathymetricPipeline  # noqa: F401
