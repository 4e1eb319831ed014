"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Restatement (pure torch, CPU) of the pieces of PyTorch Geometric that the
reference's hot path calls.  PyG (`torch_geometric`, un-pinned dependency of the
reference: environment.yml:52, install.sh:58) is NOT installed in this image and
is not vendored under /root/reference, so its published algorithm is restated
here from the public PyG 2.3-2.6 definition.

    ** parity unpinned ** for the attention layer: the reference ships no tests,
    golden vectors or checkpoints for `GATConv`; the only pins are structural
    (parameter count 182 469 / 182 405, state-dict key names, layer wiring in
    models/gnn.py:107-154).  See DESIGN.md "Oracle".

Call sites in the reference that these classes stand in for:
  * `GATConv(layer_in, layer_out, heads=..., dropout=..., edge_dim=..., concat=...)`
    models/gnn.py:125-132, called as `conv(x, edge_index, edge_attr=edge_attr)`
    models/gnn.py:175-176
  * `BatchNorm(channels)`            models/gnn.py:151-154, 181
  * `Data(x=, edge_index=, edge_attr=, pos=)`   data/graph_construction.py:150-155, 462-467
  * `Batch.from_data_list(graphs)`   scripts/inference_native.py:312

`install()` registers the classes as `torch_geometric.{data,nn}` in sys.modules so
the UNMODIFIED reference modules can be imported in the build container (used by
tests/golden/make_golden.py only; /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import math
import sys
import types
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- data
class Data:
    """Attribute bag with the PyG `Data` surface the reference touches."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, pos=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.pos = pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        return 0 if self.x is None else int(self.x.shape[0])

    @property
    def num_edges(self) -> int:
        return 0 if self.edge_index is None else int(self.edge_index.shape[1])

    def _apply(self, fn):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = fn(v) if isinstance(v, torch.Tensor) else v
        return out

    def to(self, device):
        return self._apply(lambda t: t.to(device))

    def cpu(self):
        return self._apply(lambda t: t.cpu())

    def keys(self):
        return list(self.__dict__.keys())


class Batch(Data):
    """`Batch.from_data_list`: block-diagonal concatenation (PyG collate rule:
    tensors are concatenated along dim 0, `edge_index` along dim 1 with the
    running node offset added; non-tensor attributes become lists)."""

    @classmethod
    def from_data_list(cls, data_list: List[Data]) -> "Batch":
        out = cls()
        offs = 0
        cat = {}
        batch_vec = []
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k, v in d.__dict__.items():
                if isinstance(v, torch.Tensor):
                    if k == "edge_index":
                        v = v + offs
                    cat.setdefault(k, []).append(v)
                else:
                    cat.setdefault(k, []).append(v)
            batch_vec.append(torch.full((n,), g, dtype=torch.long))
            offs += n
        for k, vs in cat.items():
            if isinstance(vs[0], torch.Tensor):
                dim = 1 if k == "edge_index" else 0
                setattr(out, k, torch.cat(vs, dim=dim))
            else:
                setattr(out, k, vs)
        out.batch = torch.cat(batch_vec) if batch_vec else torch.zeros(0, dtype=torch.long)
        out.num_graphs = len(data_list)
        return out


# --------------------------------------------------------------------------- utils
def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    if edge_attr is not None:
        edge_attr = edge_attr[mask]
    return edge_index, edge_attr


def scatter_mean(src, index, num_nodes):
    """torch_geometric.utils.scatter(..., reduce='mean'): sum / clamp(count, 1)."""
    out = torch.zeros((num_nodes,) + tuple(src.shape[1:]), dtype=src.dtype)
    out.index_add_(0, index, src)
    cnt = torch.zeros(num_nodes, dtype=src.dtype)
    cnt.index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
    cnt = cnt.clamp(min=1)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


def add_self_loops_mean(edge_index, edge_attr, num_nodes):
    """add_self_loops(fill_value='mean'): N loops appended AFTER the existing edges;
    loop attribute = mean of the attributes of the edges pointing at the node
    (grouped by edge_index[1]); 0 for nodes without incoming edges."""
    loop = torch.arange(num_nodes, dtype=edge_index.dtype)
    loop_index = torch.stack([loop, loop])
    if edge_attr is not None:
        loop_attr = scatter_mean(edge_attr, edge_index[1], num_nodes)
        edge_attr = torch.cat([edge_attr, loop_attr], dim=0)
    return torch.cat([edge_index, loop_index], dim=1), edge_attr


def segment_softmax(src, index, num_nodes):
    """torch_geometric.utils.softmax: per-target max subtraction, exp, sum + 1e-16."""
    smax = torch.full((num_nodes,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    smax.scatter_reduce_(0, index.view(-1, 1).expand_as(src), src, reduce="amax", include_self=True)
    out = (src - smax.index_select(0, index)).exp()
    ssum = torch.zeros((num_nodes,) + tuple(src.shape[1:]), dtype=src.dtype)
    ssum.index_add_(0, index, out)
    ssum = ssum + 1e-16
    return out / ssum.index_select(0, index)


def glorot_(t: torch.Tensor):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


# --------------------------------------------------------------------------- nn
class _Lin(nn.Module):
    """torch_geometric.nn.dense.linear.Linear(bias=False, weight_initializer='glorot')."""

    def __init__(self, fin, fout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        glorot_(self.weight)

    def forward(self, x):
        return F.linear(x, self.weight)


class GATConv(nn.Module):
    """Restated `torch_geometric.nn.GATConv` (int in_channels, add_self_loops=True,
    fill_value='mean', negative_slope=0.2, bias=True, flow source_to_target)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True,
                 negative_slope=0.2, dropout=0.0, add_self_loops=True,
                 edge_dim=None, fill_value="mean", bias=True):
        super().__init__()
        assert add_self_loops and fill_value == "mean"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.heads, self.concat = heads, concat
        self.negative_slope, self.dropout, self.edge_dim = negative_slope, dropout, edge_dim
        self.lin = _Lin(in_channels, heads * out_channels)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        if edge_dim is not None:
            self.lin_edge = _Lin(edge_dim, heads * out_channels)
            self.att_edge = nn.Parameter(torch.empty(1, heads, out_channels))
        else:
            self.lin_edge = None
            self.register_parameter("att_edge", None)
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels)) if bias else None
        glorot_(self.att_src)
        glorot_(self.att_dst)
        if self.att_edge is not None:
            glorot_(self.att_edge)

    # chunk of targets processed at once (keeps the [E', H, C] intermediates small)
    CHUNK_EDGES = 1 << 21

    def forward(self, x, edge_index, edge_attr=None):
        H, C = self.heads, self.out_channels
        N = x.shape[0]
        h = self.lin(x).view(N, H, C)
        a_src = (h * self.att_src).sum(-1)
        a_dst = (h * self.att_dst).sum(-1)
        edge_index, edge_attr = remove_self_loops(edge_index, edge_attr)
        edge_index, edge_attr = add_self_loops_mean(edge_index, edge_attr, N)
        src, dst = edge_index[0], edge_index[1]
        alpha = a_src.index_select(0, src) + a_dst.index_select(0, dst)
        if edge_attr is not None and self.lin_edge is not None:
            if edge_attr.dim() == 1:
                edge_attr = edge_attr.view(-1, 1)
            a_edge = torch.empty(src.numel(), H, dtype=x.dtype)
            for s in range(0, src.numel(), self.CHUNK_EDGES):
                e = self.lin_edge(edge_attr[s:s + self.CHUNK_EDGES]).view(-1, H, C)
                a_edge[s:s + self.CHUNK_EDGES] = (e * self.att_edge).sum(-1)
            alpha = alpha + a_edge
        alpha = F.leaky_relu(alpha, self.negative_slope)
        alpha = segment_softmax(alpha, dst, N)
        if self.training and self.dropout > 0:
            alpha = F.dropout(alpha, p=self.dropout, training=True)
        out = torch.zeros(N, H, C, dtype=x.dtype)
        for s in range(0, src.numel(), self.CHUNK_EDGES):
            sl = slice(s, s + self.CHUNK_EDGES)
            out.index_add_(0, dst[sl], alpha[sl].unsqueeze(-1) * h.index_select(0, src[sl]))
        out = out.view(N, H * C) if self.concat else out.mean(dim=1)
        if self.bias is not None:
            out = out + self.bias
        return out


class BatchNorm(nn.Module):
    """torch_geometric.nn.BatchNorm: `self.module = torch.nn.BatchNorm1d(c, eps=1e-5, momentum=0.1)`."""

    def __init__(self, in_channels, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
        super().__init__()
        self.module = nn.BatchNorm1d(in_channels, eps, momentum, affine, track_running_stats)

    def forward(self, x):
        return self.module(x)


class _LinBias(nn.Module):
    """PyG `Linear(in, out, bias=True)`: kaiming-uniform weight / uniform bias like torch.nn.Linear."""

    def __init__(self, fin, fout, bias=True):
        super().__init__()
        ref = nn.Linear(fin, fout, bias=bias)
        self.weight = ref.weight
        if bias:
            self.bias = ref.bias
        else:
            self.register_parameter("bias", None)

    def forward(self, x):
        y = x @ self.weight.t()
        return y if self.bias is None else y + self.bias


class GCNConv(nn.Module):
    """Public PyG `GCNConv(in, out)` with its defaults (improved=False, add_self_loops=True, normalize=True, bias=True), as
    constructed by models/gnn.py:120-121 and called as `conv(x, edge_index)` (gnn.py:177-178):
      lin = Linear(in, out, bias=False) (glorot); `gcn_norm`: self loops of weight 1 appended behind the edges
      (add_remaining_self_loops), deg[i] = sum of the weights of the edges INTO i (incl. the loop), norm_e =
      deg^-1/2[src] * w_e * deg^-1/2[dst] (inf -> 0); out = scatter_add_over_targets(norm_e * lin(x)[src]) + bias.
    Source: Kipf & Welling 2017, eq. 2, and the PyG 2.3-2.6 `gcn_norm` / `GCNConv.forward`."""

    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index, edge_weight=None):
        N = x.shape[0]
        ei, _ = remove_self_loops(edge_index)                     # add_remaining_self_loops: existing loops would keep their
        loops = torch.arange(N, dtype=ei.dtype)                   # weight; the grid graphs have none
        src = torch.cat([ei[0], loops])
        dst = torch.cat([ei[1], loops])
        w = torch.ones(src.numel(), dtype=x.dtype)
        deg = torch.zeros(N, dtype=x.dtype).index_add_(0, dst, w)
        dis = deg.pow(-0.5)
        dis[torch.isinf(dis)] = 0
        norm = dis[src] * w * dis[dst]
        h = self.lin(x)
        out = torch.zeros(N, self.out_channels, dtype=x.dtype).index_add_(0, dst, norm.unsqueeze(-1) * h[src])
        return out + self.bias


class SAGEConv(nn.Module):
    """Public PyG `SAGEConv(in, out)` with its defaults (aggr='mean', normalize=False, root_weight=True, project=False,
    bias=True; models/gnn.py:134-135): out = lin_l(mean_{j in N(i)} x_j) + lin_r(x_i); lin_l has the bias, lin_r none; a node
    without in-edges aggregates to 0 (sum / max(count, 1)).  Hamilton et al. 2017 (GraphSAGE-mean)."""

    def __init__(self, in_channels, out_channels, **kw):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = _LinBias(in_channels, out_channels, bias=True)
        self.lin_r = _LinBias(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        return self.lin_l(scatter_mean(x[edge_index[0]], edge_index[1], x.shape[0])) + self.lin_r(x)


class GINConv(nn.Module):
    """Public PyG `GINConv(nn, eps=0.0, train_eps=False)` (models/gnn.py:137-143): out = nn((1 + eps) * x_i + sum_{j in N(i)} x_j);
    `eps` is a buffer when it is not trained.  Xu et al. 2019, eq. 4.1."""

    def __init__(self, nn_module, eps: float = 0.0, train_eps: bool = False, **kw):
        super().__init__()
        self.nn = nn_module
        self.initial_eps = eps
        if train_eps:
            self.eps = nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer("eps", torch.tensor([float(eps)]))

    def forward(self, x, edge_index):
        out = torch.zeros_like(x).index_add_(0, edge_index[1], x[edge_index[0]])
        return self.nn(out + (1 + self.eps) * x)


def global_mean_pool(x, batch):
    n = int(batch.max()) + 1 if batch.numel() else 0
    return scatter_mean(x, batch, n)


def install():
    """Register the restated classes as `torch_geometric` so the unmodified
    reference imports.  Also stubs `osgeo.gdal` (loaders.py:108-118, 482-484 only
    need UseExceptions/SetConfigOption to construct Loader/Writer)."""
    if "torch_geometric" in sys.modules and not getattr(sys.modules["torch_geometric"], "_restated", False):
        return
    tg = types.ModuleType("torch_geometric")
    tg._restated = True
    tgd = types.ModuleType("torch_geometric.data")
    tgd.Data, tgd.Batch = Data, Batch
    tgn = types.ModuleType("torch_geometric.nn")
    for name in ("GATConv", "BatchNorm", "GCNConv", "SAGEConv", "GINConv", "global_mean_pool"):
        setattr(tgn, name, globals()[name])
    tgl = types.ModuleType("torch_geometric.loader")
    tgl.DataLoader = object
    tg.data, tg.nn, tg.loader = tgd, tgn, tgl
    sys.modules.update({"torch_geometric": tg, "torch_geometric.data": tgd,
                        "torch_geometric.nn": tgn, "torch_geometric.loader": tgl})
    if "osgeo" not in sys.modules:
        osgeo = types.ModuleType("osgeo")
        gdal = types.ModuleType("osgeo.gdal")
        gdal.UseExceptions = lambda: None
        gdal.SetConfigOption = lambda *a, **k: None
        gdal.Open = lambda *a, **k: None               # "could not open": VRBagHandler falls back to no georeferencing (vr_bag.py:140-153)
        osgeo.gdal = gdal
        osgeo.osr = types.ModuleType("osgeo.osr")
        osgeo.ogr = types.ModuleType("osgeo.ogr")
        sys.modules.update({"osgeo": osgeo, "osgeo.gdal": gdal, "osgeo.osr": osgeo.osr, "osgeo.ogr": osgeo.ogr})
