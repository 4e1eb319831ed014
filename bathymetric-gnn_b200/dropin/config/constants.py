"""Drop-in for config/constants.py."""
from bathymetric_gnn_b200.config import CORRECTION_NORM_FLOOR, CORRECTION_NORM_CAP  # noqa: F401
