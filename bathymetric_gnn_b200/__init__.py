"""Import shim: the product package lives in the hyphenated directory
`bathymetric-gnn_b200/` (not a valid Python identifier); this package points its
__path__ there so `import bathymetric_gnn_b200` works from the repo root."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "bathymetric-gnn_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
