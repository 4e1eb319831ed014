"""torch_geometric.data: `Data`, `Batch` (block-diagonal `from_data_list`, on the device for built graphs)."""
from bathymetric_gnn_b200.graph import Data, Batch  # noqa: F401
