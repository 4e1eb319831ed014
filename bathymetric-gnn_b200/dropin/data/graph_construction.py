"""Drop-in for data/graph_construction.py."""
from . import GraphBuilder, Data  # noqa: F401
