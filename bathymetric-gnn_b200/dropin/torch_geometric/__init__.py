"""Stand-in for the `torch_geometric` names the reference touches on its inference path
(`scripts/inference_native.py:26-30`, `data/graph_construction.py:17-22`, `models/gnn.py:17-29`,
`scripts/test_pipeline.py:56-57`): the containers are the device-backed `Data` / `Batch`, the
layers are parameter holders whose arithmetic runs in libbgnn.so.  No PyG code is involved."""
import importlib.util as _u, os as _o
_s = _u.spec_from_file_location("_bgnn_dropin_path", _o.path.join(_o.path.dirname(_o.path.dirname(_o.path.abspath(__file__))), "_path.py"))
_s.loader.exec_module(_u.module_from_spec(_s))

__version__ = "2.5.0+bgnn.b200"
from . import data, nn  # noqa: E402,F401
