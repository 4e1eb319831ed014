"""Loads a module of the reference checkout, unmodified, under a drop-in module name: for the parts of the reference
that sit outside the accelerated path (GDAL file I/O, the synthetic noise generator used as a test-input generator)."""
import os
import sys


def _root():
    m = sys.modules.get("bathymetric_gnn_b200.dropin")
    r = m.reference_root() if m is not None else None
    if r is None and os.environ.get("BGNN_REFERENCE_ROOT"):
        r = os.environ["BGNN_REFERENCE_ROOT"]
    return None if r is None else str(r)


def available(relpath: str) -> bool:
    r = _root()
    return r is not None and os.path.exists(os.path.join(r, relpath))


def exec_reference(relpath: str, namespace: dict) -> None:
    r = _root()
    if r is None or not os.path.exists(os.path.join(r, relpath)):
        raise ImportError(f"{relpath} is outside the B200 hot path and is passed through from a reference checkout: "
                          "set BGNN_REFERENCE_ROOT or call bathymetric_gnn_b200.dropin.install(reference_root=...)")
    full = os.path.join(r, relpath)
    with open(full) as f:
        code = compile(f.read(), full, "exec")
    namespace["__file_reference__"] = full
    exec(code, namespace)
