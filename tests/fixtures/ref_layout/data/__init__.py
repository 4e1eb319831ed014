"""Decoy: stands where the reference checkout keeps its own `data` package.  The entry scripts put this directory
first on sys.path (scripts/inference.py:22); if this module is ever imported the drop-in overlay did not win."""
raise ImportError("the reference checkout's own 'data' package was imported: bathymetric_gnn_b200.dropin.install() did not take effect")
