"""Drop-in for models/gnn.py."""
from . import (BathymetricGNN, LocalFeatureExtractor, GNNBackbone, ClassificationHead, ConfidenceHead,  # noqa: F401
               CorrectionHead)
