"""torch_geometric.nn: the layer classes models/gnn.py:17-29 imports, as parameter holders."""
import torch

from bathymetric_gnn_b200.gnn import (GATConvParams as GATConv, BatchNormParams as BatchNorm,  # noqa: F401
                                      GCNConvParams as GCNConv, SAGEConvParams as SAGEConv, GINConvParams as GINConv)


def global_mean_pool(x, batch, size=None):
    n = int(batch.max()) + 1 if size is None else size
    out = torch.zeros((n, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    return out / torch.bincount(batch, minlength=n).clamp(min=1).unsqueeze(1).to(x.dtype)
