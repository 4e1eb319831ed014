"""Drop-in for models/gnn.py."""
from bathymetric_gnn_b200.gnn import (BathymetricGNN, LocalFeatureExtractor, GNNBackbone, ClassificationHead,  # noqa: F401
                                      ConfidenceHead, CorrectionHead)
