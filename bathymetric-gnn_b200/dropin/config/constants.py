"""Drop-in for config/constants.py."""
from . import CORRECTION_NORM_FLOOR, CORRECTION_NORM_CAP  # noqa: F401
