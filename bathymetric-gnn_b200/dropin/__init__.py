"""Drop-in overlay for the reference's entry points.

The reference scripts (`scripts/inference.py:22-25`, `scripts/inference_native.py:23-36`,
`scripts/test_pipeline.py:30,93,137,180,227,276`) put THEIR repository root first on `sys.path`
and then import the top-level packages `config`, `data`, `models` (and, optionally,
`torch_geometric`).  A `PYTHONPATH` entry therefore cannot win; `install()` pre-seeds
`sys.modules` instead, so that every later `from config import Config`,
`from data import GraphBuilder, VRBagHandler, ...`, `from data.vr_bag import SidecarBuilder`,
`from models import BathymetricPipeline`, `from models.gnn import BathymetricGNN` and
`from torch_geometric.data import Batch` resolves to the B200 implementation no matter what the
script does to `sys.path` afterwards.

    python -m bathymetric_gnn_b200.dropin scripts/inference.py --input s.npz --model m.pt --output o.npz

runs an UNMODIFIED reference script that way (`run_script`).  Parts of the reference that are
outside the accelerated path and need no replacement (GDAL loaders / writers, the synthetic
noise generator) are passed through from the reference checkout the script belongs to when it
is available (`reference_root`, or `$BGNN_REFERENCE_ROOT`).
"""
from __future__ import annotations

import importlib
import logging
import os
import runpy
import sys
from pathlib import Path
from typing import Optional, Sequence

HERE = os.path.dirname(os.path.abspath(__file__))
_PACKAGES = ("config", "data", "models", "torch_geometric")
_SUBMODULES = ("config.config", "config.constants", "data.loaders", "data.tiling", "data.graph_construction",
               "data.vr_bag", "models.gnn", "models.pipeline", "torch_geometric.data", "torch_geometric.nn")
_state = {"reference_root": None, "installed": False}


def reference_root() -> Optional[Path]:
    r = _state["reference_root"] or os.environ.get("BGNN_REFERENCE_ROOT")
    return Path(r) if r else None


def _ours(name: str) -> bool:
    m = sys.modules.get(name)
    f = getattr(m, "__file__", None) or ""
    return m is not None and os.path.abspath(f).startswith(HERE + os.sep)


def install(reference_root: Optional[os.PathLike] = None, force: bool = True) -> None:
    """Register the drop-in packages in `sys.modules` under the reference's top-level names.

    `force=True` (default) replaces modules of those names that were imported from elsewhere (e.g. the reference's own
    `config`); with `force=False` a foreign module of the same name raises ImportError instead.  Idempotent."""
    if reference_root is not None:
        _state["reference_root"] = str(reference_root)
    foreign = [n for n in sys.modules if (n in _PACKAGES or n.split(".")[0] in _PACKAGES) and not _ours(n)]
    if foreign and not force:
        raise ImportError(f"modules already imported from elsewhere: {sorted(foreign)} (install(force=True) replaces them)")
    for n in foreign:
        del sys.modules[n]
    root = os.path.dirname(os.path.dirname(HERE))          # repository root: makes `bathymetric_gnn_b200` importable
    for p in (root, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        for name in _PACKAGES + _SUBMODULES:
            importlib.import_module(name)
    finally:
        # the overlay lives in sys.modules from here on; leave sys.path the way a script would expect it
        if HERE in sys.path:
            sys.path.remove(HERE)
    missing = [n for n in _PACKAGES + _SUBMODULES if not _ours(n)]
    if missing:
        raise ImportError(f"drop-in modules did not register: {missing}")
    _state["installed"] = True


def uninstall() -> None:
    for n in [n for n in sys.modules if n.split(".")[0] in _PACKAGES and _ours(n)]:
        del sys.modules[n]
    _state["installed"] = False


def installed() -> bool:
    return _state["installed"] and all(_ours(n) for n in _PACKAGES)


def run_script(path: os.PathLike, argv: Sequence[str] = (), run_name: str = "__main__") -> dict:
    """Run a reference entry script, unmodified, on the drop-in.  The script's checkout (its parent's parent) becomes
    the pass-through root.  A module-level `logger` is pre-seeded because `scripts/inference_native.py:148` uses one
    that the shipped script only defines inside `main()`."""
    path = Path(path).resolve()
    install(reference_root=_state["reference_root"] or path.parent.parent)
    old_argv = sys.argv
    sys.argv = [str(path), *argv]
    try:
        return runpy.run_path(str(path), init_globals={"logger": logging.getLogger(path.stem)}, run_name=run_name)
    finally:
        sys.argv = old_argv
