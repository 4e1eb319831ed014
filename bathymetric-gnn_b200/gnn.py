"""`BathymetricGNN` with the reference's module tree, constructor, `forward` and
`predict` (models/gnn.py:263-451) whose arithmetic runs in libbgnn.so.

The sub-module names are the reference's, so a reference checkpoint's
`model_state_dict` loads with `strict=True` (SURVEY.md 8(b) "State-dict
compatibility"); GATConv parameters use the PyG >= 2.4 names and the PyG <= 2.3
spelling (`lin_src` / `lin_dst`) is accepted on load.  The GAT backbone is the hot
path (fused tensor-core / tiled kernels behind one model handle); GCN / GraphSAGE / GIN
(models/gnn.py:120-143) run on the same in-edge CSR through `bgnn_csr_aggregate` +
`bgnn_linear` + `bgnn_heads_finish` (`_run_plain`).
Inference only: there is no autograd through the kernels and no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
import threading
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .graph import Data

logger = logging.getLogger(__name__)


def _glorot(t: torch.Tensor):
    a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    with torch.no_grad():
        t.uniform_(-a, a)


class _Lin(nn.Module):
    """Parameter holder for PyG's bias-free `Linear` (glorot init)."""

    def __init__(self, fin: int, fout: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin))
        _glorot(self.weight)


class GATConvParams(nn.Module):
    """Holds the parameters of one PyG `GATConv(in, out, heads, concat, edge_dim)`
    (models/gnn.py:125-132) under PyG's names; the arithmetic lives in the
    gat_aggregate / gemm_nt kernels."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, edge_dim=None):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat, self.edge_dim = in_channels, out_channels, heads, concat, edge_dim
        self.lin = _Lin(in_channels, heads * out_channels)
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        _glorot(self.att_src)
        _glorot(self.att_dst)
        if edge_dim is not None:
            self.lin_edge = _Lin(edge_dim, heads * out_channels)
            self.att_edge = nn.Parameter(torch.empty(1, heads, out_channels))
            _glorot(self.att_edge)
        else:
            self.lin_edge = None
            self.register_parameter("att_edge", None)
        self.bias = nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # PyG <= 2.3 stored the shared projection twice, as lin_src and lin_dst
        src, dst, new = prefix + "lin_src.weight", prefix + "lin_dst.weight", prefix + "lin.weight"
        if src in state_dict and new not in state_dict:
            state_dict[new] = state_dict.pop(src)
            state_dict.pop(dst, None)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class _LinB(nn.Module):
    """Parameter holder for a PyG `Linear` with bias (torch.nn.Linear's default init)."""

    def __init__(self, fin: int, fout: int):
        super().__init__()
        ref = nn.Linear(fin, fout)
        self.weight, self.bias = ref.weight, ref.bias


class GCNConvParams(nn.Module):
    """PyG `GCNConv(in, out)` (models/gnn.py:120-121): `lin` without bias (glorot) + `bias` (zeros); defaults
    add_self_loops=True, normalize=True:  out_i = sum_{j in N(i) + i} h_j / sqrt(d_i d_j) + b,  d = in-degree + 1."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)
        self.bias = nn.Parameter(torch.zeros(out_channels))


class SAGEConvParams(nn.Module):
    """PyG `SAGEConv(in, out)` (models/gnn.py:134-135): aggr='mean', root_weight=True:
    out_i = lin_l(mean_{j in N(i)} x_j) + lin_r(x_i); `lin_l` carries the bias, `lin_r` has none."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin_l = _LinB(in_channels, out_channels)
        self.lin_r = _LinB(in_channels, out_channels)
        self.lin_r.bias = None


class GINConvParams(nn.Module):
    """PyG `GINConv(nn)` (models/gnn.py:137-143), eps = 0 (not trained; kept as the buffer PyG registers):
    out_i = nn((1 + eps) x_i + sum_{j in N(i)} x_j)."""

    def __init__(self, mlp: nn.Module, eps: float = 0.0):
        super().__init__()
        self.nn = mlp
        self.register_buffer("eps", torch.tensor([float(eps)]))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        if prefix + "eps" not in state_dict:           # PyG versions that keep eps as a plain attribute
            state_dict[prefix + "eps"] = self.eps.detach().clone()
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class BatchNormParams(nn.Module):
    """`torch_geometric.nn.BatchNorm` wrapper layout: `.module` is a BatchNorm1d."""

    def __init__(self, channels: int):
        super().__init__()
        self.module = nn.BatchNorm1d(channels, eps=1e-5, momentum=0.1)


def _mlp(fin, hid, fout, dropout, final=None):
    layers = [nn.Linear(fin, hid), nn.ReLU(), nn.Dropout(dropout), nn.Linear(hid, fout)]
    if final is not None:
        layers.append(final)
    return nn.Sequential(*layers)


class LocalFeatureExtractor(nn.Module):
    """models/gnn.py:34-71 (instantiated by the model with num_layers=2)."""

    def __init__(self, in_channels, hidden_channels, out_channels, num_layers: int = 2, dropout: float = 0.1):
        super().__init__()
        if num_layers != 2:
            raise NotImplementedError("the reference model always uses num_layers=2 (models/gnn.py:313-319)")
        self.mlp = _mlp(in_channels, hidden_channels, out_channels, dropout)


class GNNBackbone(nn.Module):
    """models/gnn.py:74-188: the module tree (parameter holders) of every `gnn_type` the reference accepts."""

    def __init__(self, in_channels, hidden_channels, num_layers, gnn_type="GAT", heads=4, dropout=0.1, edge_dim=None):
        super().__init__()
        if gnn_type not in ("GAT", "GCN", "GraphSAGE", "GIN"):
            raise ValueError(f"Unknown GNN type: {gnn_type}")                   # models/gnn.py:145-146
        self.gnn_type, self.num_layers, self.dropout = gnn_type, num_layers, dropout
        self.convs, self.norms = nn.ModuleList(), nn.ModuleList()
        for i in range(num_layers):
            last = i == num_layers - 1
            if gnn_type == "GAT":
                layer_in = in_channels if i == 0 else hidden_channels * heads
                self.convs.append(GATConvParams(layer_in, hidden_channels, heads=1 if last else heads,
                                                concat=not last, edge_dim=edge_dim))
                self.norms.append(BatchNormParams(hidden_channels if last else hidden_channels * heads))
                continue
            layer_in = in_channels if i == 0 else hidden_channels                # models/gnn.py:109-114
            if gnn_type == "GCN":
                self.convs.append(GCNConvParams(layer_in, hidden_channels))
            elif gnn_type == "GraphSAGE":
                self.convs.append(SAGEConvParams(layer_in, hidden_channels))
            else:
                self.convs.append(GINConvParams(nn.Sequential(nn.Linear(layer_in, hidden_channels), nn.ReLU(),
                                                              nn.Linear(hidden_channels, hidden_channels))))
            self.norms.append(BatchNormParams(hidden_channels))


class ClassificationHead(nn.Module):
    def __init__(self, in_channels, hidden_channels, num_classes, dropout=0.1):
        super().__init__()
        self.mlp = _mlp(in_channels, hidden_channels, num_classes, dropout)


class ConfidenceHead(nn.Module):
    def __init__(self, in_channels, hidden_channels, dropout=0.1):
        super().__init__()
        self.mlp = _mlp(in_channels, hidden_channels, 1, dropout, nn.Sigmoid())


class CorrectionHead(nn.Module):
    def __init__(self, in_channels, hidden_channels, dropout=0.1):
        super().__init__()
        self.mlp = _mlp(in_channels, hidden_channels, 1, dropout)


class BathymetricGNN(nn.Module):
    CLASS_SEAFLOOR = 0
    CLASS_FEATURE = 1
    CLASS_NOISE = 2

    def __init__(self, in_channels: int, hidden_channels: int = 64, num_gnn_layers: int = 4, gnn_type: str = "GAT",
                 heads: int = 4, num_classes: int = 3, predict_correction: bool = True, dropout: float = 0.1,
                 edge_dim: Optional[int] = None):
        super().__init__()
        self.predict_correction, self.num_classes = predict_correction, num_classes
        self.in_channels, self.hidden_channels, self.num_gnn_layers = in_channels, hidden_channels, num_gnn_layers
        self.heads, self.edge_dim, self.dropout = heads, edge_dim, dropout
        self.feature_extractor = LocalFeatureExtractor(in_channels, hidden_channels, hidden_channels, 2, dropout)
        self.gnn = GNNBackbone(hidden_channels, hidden_channels, num_gnn_layers, gnn_type, heads, dropout, edge_dim)
        self.classification_head = ClassificationHead(hidden_channels, hidden_channels // 2, num_classes, dropout)
        self.confidence_head = ConfidenceHead(hidden_channels, hidden_channels // 2, dropout)
        self.correction_head = CorrectionHead(hidden_channels, hidden_channels // 2, dropout) if predict_correction else None
        self._handle = None
        self._handle_key = None
        self._handle_lock = threading.RLock()      # handle (re)creation / destruction vs a concurrent forward
        self._options: Dict[str, int] = {}         # kernel-selection switches of THIS model's handle (set_option)
        self.use_grid_kernel = True      # tiled aggregation when the graph carries its grid topology
        self.precision = "fp32"          # "fp32" | "bf16" (set_precision)
        logger.info("Created BathymetricGNN: %s with %d layers, %d hidden channels", gnn_type, num_gnn_layers, hidden_channels)

    # ------------------------------------------------------------------ packing
    def _desc(self) -> _lib.ModelDesc:
        d = _lib.ModelDesc()
        d.in_channels, d.hidden, d.layers, d.heads = self.in_channels, self.hidden_channels, self.num_gnn_layers, self.heads
        d.num_classes, d.predict_correction = self.num_classes, int(self.predict_correction)
        d.edge_dim = int(self.edge_dim or 0)
        d.bn_eps, d.negative_slope = 1e-5, 0.2
        return d

    def _flat_params(self) -> torch.Tensor:
        """Blob in the order include/bgnn.h documents for bgnn_model_create."""
        fe = self.feature_extractor.mlp
        parts = [fe[0].weight, fe[0].bias, fe[3].weight, fe[3].bias]
        for conv, norm in zip(self.gnn.convs, self.gnn.norms):
            parts += [conv.lin.weight, conv.att_src, conv.att_dst]
            if conv.edge_dim is not None:
                parts += [conv.lin_edge.weight, conv.att_edge]
            bn = norm.module
            parts += [conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        heads = [self.classification_head, self.confidence_head] + ([self.correction_head] if self.predict_correction else [])
        for h in heads:
            parts += [h.mlp[0].weight, h.mlp[0].bias, h.mlp[3].weight, h.mlp[3].bias]
        return torch.cat([p.detach().reshape(-1).to("cpu", torch.float32) for p in parts]).contiguous()

    def set_precision(self, precision: str = "fp32") -> "BathymetricGNN":
        """Compute precision of forward / predict (no counterpart in the reference, which has no reduced-precision path;
        BASELINE configs[4] asks for "bf16 vs fp32 tolerance").  "fp32" is the parity path.  "bf16" keeps the big per-node
        streams (projection operands and outputs, layer outputs) in HBM as bf16 while every sum, the softmax, BatchNorm and
        the heads stay fp32 (include/bgnn.h: BGNN_PREC_BF16); parameters stay fp32 master copies.  Returns self."""
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"Unknown precision: {precision}")
        self.precision = precision
        return self

    def set_option(self, name: str, value) -> "BathymetricGNN":
        """Kernel-selection switch of this model's device handle (include/bgnn.h: bgnn_model_set_option): tensor_cores,
        presplit, window, tile5, tile5_window.  Per model, not per process; used by the tests to cross-check variants."""
        if name not in _lib.OPTIONS:
            raise ValueError(f"Unknown option: {name}")
        with self._handle_lock:
            self._options[name] = int(bool(value))
            if self._handle is not None:
                check(lib.bgnn_model_set_option(self._handle, _lib.OPTIONS[name], self._options[name]))
        return self

    def _invalidate_plain(self):
        self.__dict__.pop("_plain_cache", None)

    def invalidate(self) -> "BathymetricGNN":
        """Drop the folded device weights; the next forward re-packs them from the module's parameters.  Called by
        load_state_dict / _apply (.to, .cuda, .float) / train; call it yourself after editing parameters through `.data`
        or an external view (those edits do not bump the tensors' version counters that the cache key watches)."""
        with self._handle_lock:
            self._release()
            self._invalidate_plain()
        return self

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        r = super().load_state_dict(state_dict, strict=strict, assign=assign)
        self.invalidate()
        return r

    def _apply(self, fn, recurse=True):
        r = super()._apply(fn, recurse) if recurse is True else super()._apply(fn, recurse=recurse)
        if "_handle_lock" in self.__dict__:
            self.invalidate()
        return r

    def train(self, mode: bool = True):
        r = super().train(mode)
        if "_handle_lock" in self.__dict__ and mode:
            self.invalidate()
        return r

    def _ensure_handle(self, device: torch.device):
        key = (device.index if device.index is not None else torch.cuda.current_device(),
               tuple(t._version for t in list(self.parameters()) + list(self.buffers())),
               tuple(t.data_ptr() for t in self.parameters()))
        with self._handle_lock:
            if self._handle is not None and self._handle_key == key:
                return self._handle
            self._release()
            flat = self._flat_params()
            d = self._desc()
            h = C.c_void_p()
            check(lib.bgnn_model_create(C.byref(d), C.c_void_p(flat.data_ptr()), flat.numel(), key[0], C.byref(h)))
            for name, v in self._options.items():
                check(lib.bgnn_model_set_option(h, _lib.OPTIONS[name], v))
            self._handle, self._handle_key = h, key
            return h

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            lib.bgnn_model_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def __getstate__(self):            # the device handle is never copied / pickled
        s = self.__dict__.copy()
        s["_handle"], s["_handle_key"] = None, None
        s.pop("_handle_lock", None)
        return s

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._handle_lock = threading.RLock()

    # ------------------------------------------------------------------ forward
    @staticmethod
    def _csr_from_coo(data: Data, device):
        """Foreign `Data` (not from the device GraphBuilder): in-edge CSR from the COO,
        existing self loops removed as GATConv does (plumbing, torch ops)."""
        ei = data.edge_index.to(device)
        ea = getattr(data, "edge_attr", None)
        keep = ei[0] != ei[1]
        ei = ei[:, keep]
        order = torch.sort(ei[1], stable=True).indices
        N = int(data.x.shape[0])
        rowptr = torch.zeros(N + 1, dtype=torch.int64, device=device)
        rowptr[1:] = torch.cumsum(torch.bincount(ei[1], minlength=N), 0)
        src = ei[0][order].to(torch.int32).contiguous()
        if ea is not None:
            ea = ea.to(device=device, dtype=torch.float32)
            if ea.dim() == 1:
                ea = ea.view(-1, 1)
            ea = ea[keep][order].contiguous()
        return rowptr.to(torch.int32), src, ea, int(src.numel())

    def _run(self, data: Data, auto_thr: float, review_thr: float, want_action: bool, return_hidden: bool = False):
        _lib.require_cuda()
        if self.training and self.dropout > 0:
            raise RuntimeError("BathymetricGNN (B200 build) is inference-only: call .eval() first")
        if "_host" in data.__dict__:            # a deferred small host grid (graph.py::HostGridGraph): build it on the device now
            data._materialise_host()
        dg = data.__dict__.get("_dev")
        if dg is not None and hasattr(data, "topology_overridden") and data.topology_overridden():
            dg = None               # caller-assigned edge_index / edge_attr: honour them (COO path)
        x = data.x
        device = dg.device if dg is not None else (x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        with torch.cuda.device(device):
            x = x.to(device=device, dtype=torch.float32).contiguous()
            N = int(x.shape[0])
            if x.dim() != 2 or x.shape[1] != self.in_channels:
                raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied ({N}x{x.shape[-1]} and "
                                   f"{self.in_channels}x{self.hidden_channels})")
            if self.gnn.gnn_type != "GAT":
                return self._run_plain(data, dg, x, device, auto_thr, review_thr, want_action, return_hidden)
            topo = None
            h = self._ensure_handle(device)
            want = _lib.PREC_BF16 if getattr(self, "precision", "fp32") == "bf16" else _lib.PREC_F32
            if lib.bgnn_model_precision(h) != want:
                check(lib.bgnn_model_set_precision(h, want))
            if dg is not None:
                E = dg.E
                ed_have = int(dg.cfg.n_edge_feats)
                if dg.grid_meta is None and int(dg.cells.n_grids) == 1 and self.use_grid_kernel:
                    topo = _lib.GridTopo()
                    topo.rows, topo.cols, topo.node_id = int(dg.cells.rows), int(dg.cells.cols), dg.node_id.data_ptr()
                    topo.n_offsets = int(dg.cfg.n_offsets)
                    for k in range(topo.n_offsets):
                        topo.off_dr[k], topo.off_dc[k] = dg.cfg.off_dr[k], dg.cfg.off_dc[k]
                    ge = getattr(dg, "grid_eattr", None)
                    topo.grid_eattr = ge.data_ptr() if ge is not None else None
                elif self.use_grid_kernel and int(dg.cfg.n_offsets) <= 8:
                    # ragged batch: no single grid, but the stencil bounds the in-degree (windowed CSR aggregation)
                    topo = _lib.GridTopo()
                    topo.rows, topo.cols, topo.node_id, topo.grid_eattr = 0, 0, None, None
                    topo.n_offsets = int(dg.cfg.n_offsets)
                    for k in range(topo.n_offsets):
                        topo.off_dr[k], topo.off_dc[k] = dg.cfg.off_dr[k], dg.cfg.off_dc[k]
                    if getattr(dg, "rag_patches", None) is not None and dg.node_id is not None and dg.grid_eattr is not None:
                        # band patches planned by the builder: tiled aggregation per band of grid rows
                        topo.node_id, topo.grid_eattr = dg.node_id.data_ptr(), dg.grid_eattr.data_ptr()
                        topo.patches, topo.n_patches = dg.rag_patches.data_ptr(), int(dg.n_rag_patches)
                rowptr = dg.rowptr
                if topo is not None and (topo.rows > 0 or topo.n_patches > 0) and lib.bgnn_gat_grid_path(h, C.byref(topo), E) == 1:
                    src, ea = None, None            # every layer runs the tiled kernel: the CSR payload is never read
                else:
                    src, ea = dg.csr_src, dg.csr_eattr
            else:
                rowptr, src, ea, E = self._csr_from_coo(data, device)
                ed_have = 0 if ea is None else int(ea.shape[1])
            ED = int(self.edge_dim or 0)
            if ED > 0 and ed_have != ED:
                raise RuntimeError(f"model expects edge_dim={ED} but the graph carries {ed_have} edge features")
            out = dict(class_logits=torch.empty((N, self.num_classes), dtype=torch.float32, device=device),
                       class_probs=torch.empty((N, self.num_classes), dtype=torch.float32, device=device),
                       predicted_class=torch.empty(N, dtype=torch.int64, device=device),
                       confidence=torch.empty(N, dtype=torch.float32, device=device))
            if self.predict_correction:
                out["correction"] = torch.empty(N, dtype=torch.float32, device=device)
            action = torch.empty(N, dtype=torch.int64, device=device) if want_action else None
            hidden = torch.empty((N, self.hidden_channels), dtype=torch.float32, device=device) if return_hidden else None
            if N > 0:
                wbytes = int(lib.bgnn_gat_workspace_bytes(h, N))
                ws = torch.empty(wbytes, dtype=torch.uint8, device=device)
                check(lib.bgnn_gat_forward(h, N, E, ptr(x), ptr(rowptr), ptr(src) if (E > 0 and src is not None) else None,
                                           ptr(ea) if (ED > 0 and E > 0 and ea is not None) else None, ptr(ws), wbytes,
                                           ptr(out["class_logits"]), ptr(out["class_probs"]), ptr(out["predicted_class"]),
                                           ptr(out["confidence"]), ptr(out.get("correction")), ptr(action),
                                           float(auto_thr), float(review_thr), ptr(hidden),
                                           C.byref(topo) if topo is not None else None, stream_ptr()))
            if want_action:
                out["action"] = action
            if return_hidden:
                out["hidden"] = hidden
        return out

    # ------------------------------------------------------------------ GCN / GraphSAGE / GIN (models/gnn.py:120-143)
    def _plain_weights(self, device):
        """Device copies of the folded weights of the attention-free backbones: BatchNorm(eval) becomes a per-channel
        scale / shift folded into the layer's last Linear (or applied by the aggregation epilogue for GCN); cached on the
        parameters' version counters like the GAT handle."""
        key = (str(device), tuple(t._version for t in list(self.parameters()) + list(self.buffers())),
               tuple(t.data_ptr() for t in self.parameters()))
        cache = self.__dict__.get("_plain_cache")
        if cache is not None and cache[0] == key:
            return cache[1]
        f = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
        fe = self.feature_extractor.mlp
        w = dict(fe=[(f(fe[0].weight), f(fe[0].bias)), (f(fe[3].weight), f(fe[3].bias))], layers=[])
        for conv, norm in zip(self.gnn.convs, self.gnn.norms):
            bn = norm.module
            s_ = (bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps))
            t_ = bn.bias.detach().double() - bn.running_mean.detach().double() * s_
            if isinstance(conv, GCNConvParams):
                w["layers"].append(dict(W=f(conv.lin.weight), pre=f(conv.bias), mul=f(s_.float()), add=f(t_.float())))
            elif isinstance(conv, SAGEConvParams):          # BN(lin_l(mean) + lin_r(x)) = [mean | x] . (s * [Wl | Wr])^T + (s * bl + t)
                Wc = torch.cat([conv.lin_l.weight.detach().double(), conv.lin_r.weight.detach().double()], dim=1) * s_[:, None]
                w["layers"].append(dict(W=f(Wc.float()), b=f((conv.lin_l.bias.detach().double() * s_ + t_).float())))
            else:
                l1, l2 = conv.nn[0], conv.nn[2]
                w["layers"].append(dict(W1=f(l1.weight), b1=f(l1.bias), W2=f((l2.weight.detach().double() * s_[:, None]).float()),
                                        b2=f((l2.bias.detach().double() * s_ + t_).float()), eps=float(conv.eps)))
        heads = [self.classification_head, self.confidence_head] + ([self.correction_head] if self.predict_correction else [])
        w["hW0"] = f(torch.cat([h.mlp[0].weight.detach() for h in heads], dim=0))
        w["hb0"] = f(torch.cat([h.mlp[0].bias.detach() for h in heads], dim=0))
        w["h1"] = [(f(h.mlp[3].weight), f(h.mlp[3].bias)) for h in heads]
        # deg^-1/2 exactly as PyG's gcn_norm computes it (torch float32 pow), indexed by in-degree + 1
        w["dis"] = f(torch.cat([torch.zeros(1), torch.arange(1, 64, dtype=torch.float32).pow(-0.5)]))
        self.__dict__["_plain_cache"] = (key, w)
        return w

    def _run_plain(self, data, dg, x, device, auto_thr, review_thr, want_action, return_hidden):
        """forward / predict for gnn_type in {GCN, GraphSAGE, GIN}: conv(x, edge_index) -> BatchNorm(eval) -> ReLU (all but the
        last layer), models/gnn.py:173-186; edge attributes are not used by these layers (gnn.py:177-178)."""
        if getattr(self, "precision", "fp32") != "fp32":
            raise NotImplementedError("the bf16 mode covers the GAT backbone only")
        N, Cc, L = int(x.shape[0]), self.hidden_channels, self.num_gnn_layers
        if dg is not None:
            rowptr, src = dg.rowptr, (dg.csr_src if dg.E > 0 else None)
            max_deg = int(dg.cfg.n_offsets)
        else:
            rowptr, src, _, E = self._csr_from_coo(data, device)
            max_deg = int((rowptr[1:] - rowptr[:-1]).max()) if N else 0
        w = self._plain_weights(device)
        if max_deg + 1 >= w["dis"].numel():
            raise NotImplementedError(f"in-degree {max_deg} exceeds the GCN normalisation table")
        new = lambda c: torch.empty((N, c), dtype=torch.float32, device=device)
        st = stream_ptr()
        lin = lambda a, K, W, b, relu, o, Nout: check(lib.bgnn_linear(ptr(a), a.stride(0), N, K, ptr(W), ptr(b), int(relu), Nout, ptr(o), o.stride(0), st))
        agg = lambda mode, a, o, Cin, eps=0.0, pre=None, mul=None, add=None, relu=0, copy_self=0: check(lib.bgnn_csr_aggregate(
            mode, N, Cin, ptr(a), a.stride(0), ptr(rowptr), ptr(src), ptr(w["dis"]), int(w["dis"].numel()), float(eps), ptr(pre), ptr(mul),
            ptr(add), int(relu), ptr(o), o.stride(0), int(copy_self), st))
        out = dict(class_logits=torch.empty((N, self.num_classes), dtype=torch.float32, device=device),
                   class_probs=torch.empty((N, self.num_classes), dtype=torch.float32, device=device),
                   predicted_class=torch.empty(N, dtype=torch.int64, device=device),
                   confidence=torch.empty(N, dtype=torch.float32, device=device))
        if self.predict_correction:
            out["correction"] = torch.empty(N, dtype=torch.float32, device=device)
        action = torch.empty(N, dtype=torch.int64, device=device) if want_action else None
        if N > 0:
            t0, cur = new(Cc), new(Cc)
            lin(x, self.in_channels, *w["fe"][0], True, t0, Cc)                     # LocalFeatureExtractor (gnn.py:34-71)
            lin(t0, Cc, *w["fe"][1], False, cur, Cc)
            kind = self.gnn.gnn_type
            for l, lw in enumerate(w["layers"]):
                relu = l < L - 1
                nxt = new(Cc)
                if kind == "GCN":                                                   # BN(A_norm (x W^T) + b)
                    lin(cur, Cc, lw["W"], None, False, t0, Cc)
                    agg(_lib.AGG_GCN, t0, nxt, Cc, pre=lw["pre"], mul=lw["mul"], add=lw["add"], relu=relu)
                elif kind == "GraphSAGE":
                    cat = new(2 * Cc)
                    agg(_lib.AGG_MEAN, cur, cat, Cc, copy_self=1)
                    lin(cat, 2 * Cc, lw["W"], lw["b"], relu, nxt, Cc)
                else:                                                               # GIN: BN(nn((1 + eps) x + sum x_j))
                    agg(_lib.AGG_GIN, cur, t0, Cc, eps=lw["eps"])
                    t1 = new(Cc)
                    lin(t0, Cc, lw["W1"], lw["b1"], True, t1, Cc)
                    lin(t1, Cc, lw["W2"], lw["b2"], relu, nxt, Cc)
                cur = nxt
            h2, nblk = Cc // 2, (3 if self.predict_correction else 2)
            hh = new(nblk * h2)
            lin(cur, Cc, w["hW0"], w["hb0"], True, hh, nblk * h2)
            hc = w["h1"]
            check(lib.bgnn_heads_finish(ptr(hh), N, h2, nblk, self.num_classes, ptr(hc[0][0]), ptr(hc[0][1]), ptr(hc[1][0]), ptr(hc[1][1]),
                                        ptr(hc[2][0]) if nblk > 2 else None, ptr(hc[2][1]) if nblk > 2 else None,
                                        ptr(out["class_logits"]), ptr(out["class_probs"]), ptr(out["predicted_class"]),
                                        ptr(out["confidence"]), ptr(out.get("correction")), ptr(action), float(auto_thr),
                                        float(review_thr), st))
            if return_hidden:
                out["hidden"] = cur
        elif return_hidden:
            out["hidden"] = new(Cc)
        if want_action:
            out["action"] = action
        return out

    def forward(self, data: Data) -> Dict[str, torch.Tensor]:
        """models/gnn.py:360-408."""
        return self._run(data, 0.85, 0.6, want_action=False)

    def predict(self, data: Data, auto_correct_threshold: float = 0.85, review_threshold: float = 0.6) -> Dict[str, torch.Tensor]:
        """models/gnn.py:410-451."""
        self.eval()
        with torch.no_grad():
            out = self._run(data, auto_correct_threshold, review_threshold, want_action=True)
        out["needs_review"] = out["action"] == 2
        out["auto_correct"] = out["action"] == 1
        return out
