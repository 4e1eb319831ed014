"""Drop-in for the in-memory container of data/loaders.py (file formats are out of scope)."""
from . import BathymetricGrid, BathymetricLoader, BathymetricWriter  # noqa: F401
