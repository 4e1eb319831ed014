"""Makes `bathymetric_gnn_b200` importable when only this drop-in directory is on sys.path."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.insert(0, _root)
