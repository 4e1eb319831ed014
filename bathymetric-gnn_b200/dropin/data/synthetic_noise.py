"""data/synthetic_noise.py is a test-input generator outside the accelerated path:
the reference's own file runs here unmodified, under this module name."""
from bathymetric_gnn_b200.dropin._passthrough import exec_reference as _exec

_exec("data/synthetic_noise.py", globals())
