"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

Functional (state-dict driven) restatement of the reference forward pass
`BathymetricGNN.forward/predict` (/root/reference/models/gnn.py:360-451) with the
GAT backbone wiring of `GNNBackbone` (gnn.py:107-154, 173-186) on top of the
restated PyG ops in oracle/pyg_restated.py.

** parity unpinned ** for the attention arithmetic (PyG absent, no reference
golden vectors).  What IS pinned, in tests/test_oracle_golden.py:
  * this functional form == the UNMODIFIED reference `models/gnn.py` module run on
    the restated PyG classes (fixtures tests/golden/model_*.npz made in the build
    container by tests/golden/make_golden.py);
  * parameter count 182 469 (F=8) / 182 405 (F=7)  (docs/QUICK_REFERENCE.md:185);
  * state-dict key set of the reference module.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pyg_restated as pyg


@dataclass
class ModelCfg:
    in_channels: int = 7
    hidden: int = 64
    layers: int = 4
    heads: int = 4
    num_classes: int = 3
    predict_correction: bool = True
    edge_dim: Optional[int] = 3
    gnn_type: str = "GAT"              # "GAT" | "GCN" | "GraphSAGE" | "GIN" (models/gnn.py:120-146)


def layer_dims(cfg: ModelCfg):
    """(in, heads, concat) per GAT layer: gnn.py:107-132."""
    out = []
    for i in range(cfg.layers):
        last = i == cfg.layers - 1
        fin = cfg.hidden if i == 0 else cfg.hidden * cfg.heads
        out.append((fin, 1 if last else cfg.heads, not last))
    return out


def init_state_dict(cfg: ModelCfg, seed: int = 0, randomize_bn: bool = True,
                    head_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random-init weights with the reference's key names (gnn.py module paths +
    PyG >= 2.4 GATConv parameter names).  nn.Linear uses torch's default init,
    GATConv uses glorot (pyg_restated).  BN running stats / affine are randomised
    so that BN folding is exercised; `head_gain` scales the last
    classification layer to thin out near-tie logits."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def linear(prefix, fin, fout):
        m = nn.Linear(fin, fout)
        sd[prefix + ".weight"] = m.weight.detach().clone()
        sd[prefix + ".bias"] = m.bias.detach().clone()

    linear("feature_extractor.mlp.0", cfg.in_channels, cfg.hidden)          # gnn.py:55
    linear("feature_extractor.mlp.3", cfg.hidden, cfg.hidden)               # gnn.py:66
    for i, (fin, H, concat) in enumerate(layer_dims(cfg)):
        if cfg.gnn_type == "GAT":
            conv = pyg.GATConv(fin, cfg.hidden, heads=H, concat=concat, edge_dim=cfg.edge_dim)
        elif cfg.gnn_type == "GCN":                                          # gnn.py:120-121 (layer_in = hidden for every layer)
            conv, H, concat = pyg.GCNConv(cfg.hidden, cfg.hidden), 1, False
        elif cfg.gnn_type == "GraphSAGE":                                    # gnn.py:134-135
            conv, H, concat = pyg.SAGEConv(cfg.hidden, cfg.hidden), 1, False
        elif cfg.gnn_type == "GIN":                                          # gnn.py:137-143
            conv = pyg.GINConv(nn.Sequential(nn.Linear(cfg.hidden, cfg.hidden), nn.ReLU(), nn.Linear(cfg.hidden, cfg.hidden)))
            H, concat = 1, False
        else:
            raise ValueError(f"Unknown GNN type: {cfg.gnn_type}")
        for k, v in conv.state_dict().items():
            sd[f"gnn.convs.{i}.{k}"] = v.detach().clone()
        C = cfg.hidden * H if concat else cfg.hidden
        if randomize_bn:
            if f"gnn.convs.{i}.bias" in sd:
                sd[f"gnn.convs.{i}.bias"] = 0.1 * torch.randn(C, generator=g)
            sd[f"gnn.norms.{i}.module.weight"] = 1.0 + 0.2 * torch.randn(C, generator=g)
            sd[f"gnn.norms.{i}.module.bias"] = 0.1 * torch.randn(C, generator=g)
            sd[f"gnn.norms.{i}.module.running_mean"] = 0.1 * torch.randn(C, generator=g)
            sd[f"gnn.norms.{i}.module.running_var"] = 0.5 + torch.rand(C, generator=g)
        else:
            sd[f"gnn.norms.{i}.module.weight"] = torch.ones(C)
            sd[f"gnn.norms.{i}.module.bias"] = torch.zeros(C)
            sd[f"gnn.norms.{i}.module.running_mean"] = torch.zeros(C)
            sd[f"gnn.norms.{i}.module.running_var"] = torch.ones(C)
        sd[f"gnn.norms.{i}.module.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    h2 = cfg.hidden // 2
    linear("classification_head.mlp.0", cfg.hidden, h2)                     # gnn.py:203-208
    linear("classification_head.mlp.3", h2, cfg.num_classes)
    linear("confidence_head.mlp.0", cfg.hidden, h2)                         # gnn.py:226-232
    linear("confidence_head.mlp.3", h2, 1)
    if cfg.predict_correction:
        linear("correction_head.mlp.0", cfg.hidden, h2)                     # gnn.py:250-256
        linear("correction_head.mlp.3", h2, 1)
    if head_gain != 1.0:
        sd["classification_head.mlp.3.weight"] *= head_gain
        sd["classification_head.mlp.3.bias"] *= head_gain
    return sd


def calibrate_heads(sd, cfg: "ModelCfg", x, edge_index, edge_attr, conf_target: float = 0.72):
    """Shift the last-layer biases of the classification / confidence heads so that
    random-init weights give a MIX of the three classes and confidences on both
    sides of the 0.6 / 0.85 thresholds (label checks need all three
    classes).  Returns the two adjusted bias tensors (stored in the fixtures)."""
    out = forward(sd, cfg, x, edge_index, edge_attr)
    sd["classification_head.mlp.3.bias"] = (sd["classification_head.mlp.3.bias"]
                                            - out["class_logits"].median(dim=0).values).contiguous()
    c = out["confidence"].double().clamp(1e-6, 1 - 1e-6)
    z_med = torch.log(c / (1 - c)).median()
    import math
    shift = math.log(conf_target / (1 - conf_target)) - float(z_med)
    sd["confidence_head.mlp.3.bias"] = sd["confidence_head.mlp.3.bias"] + torch.tensor([shift], dtype=torch.float32)
    return sd["classification_head.mlp.3.bias"].clone(), sd["confidence_head.mlp.3.bias"].clone()


def param_count(sd) -> int:
    return sum(v.numel() for k, v in sd.items()
               if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 (round to nearest even) -> fp32: the storage rounding of the bf16 mode."""
    return t.to(torch.bfloat16).to(t.dtype)


def gat_layer(sd, i, x, edge_index, edge_attr, H, C, concat, chunk=1 << 21, bf16_storage=False):
    """One restated GATConv (see pyg_restated.GATConv.forward) driven by `sd`.

    `bf16_storage` restates the CUDA path's bf16 mode (include/bgnn.h, bgnn_model_set_precision): the
    projection multiplies bf16-rounded operands with fp32 accumulation, the attention logits come from
    the fp32 accumulators, and `h` is rounded to bf16 before the aggregation reads it."""
    p = f"gnn.convs.{i}."
    dt = x.dtype
    W = sd[p + "lin.weight"].to(dt) if p + "lin.weight" in sd else sd[p + "lin_src.weight"].to(dt)
    N = x.shape[0]
    if bf16_storage:
        W = bf16_round(W)
    h = (x @ W.t()).view(N, H, C)
    a_src = (h * sd[p + "att_src"].to(dt)).sum(-1)
    a_dst = (h * sd[p + "att_dst"].to(dt)).sum(-1)
    if bf16_storage:
        h = bf16_round(h)
    ei, ea = pyg.remove_self_loops(edge_index, edge_attr)
    ei, ea = pyg.add_self_loops_mean(ei, ea, N)
    src, dst = ei[0], ei[1]
    alpha = a_src.index_select(0, src) + a_dst.index_select(0, dst)
    if ea is not None and p + "lin_edge.weight" in sd:
        We, ae = sd[p + "lin_edge.weight"].to(dt), sd[p + "att_edge"].to(dt)
        a_edge = torch.empty(src.numel(), H, dtype=dt)
        for s in range(0, src.numel(), chunk):
            e = (ea[s:s + chunk] @ We.t()).view(-1, H, C)
            a_edge[s:s + chunk] = (e * ae).sum(-1)
        alpha = alpha + a_edge
    alpha = F.leaky_relu(alpha, 0.2)
    alpha = pyg.segment_softmax(alpha, dst, N)
    out = torch.zeros(N, H, C, dtype=dt)
    for s in range(0, src.numel(), chunk):
        sl = slice(s, s + chunk)
        out.index_add_(0, dst[sl], alpha[sl].unsqueeze(-1) * h.index_select(0, src[sl]))
    out = out.view(N, H * C) if concat else out.mean(dim=1)
    return out + sd[p + "bias"].to(dt)


def plain_layer(sd, i, x, edge_index, gnn_type: str):
    """GCNConv / SAGEConv / GINConv of layer i driven by `sd` (pyg_restated classes; gnn.py:177-178 calls them without
    edge attributes)."""
    p = f"gnn.convs.{i}."
    dt = x.dtype
    C = x.shape[1]
    if gnn_type == "GCN":
        conv = pyg.GCNConv(C, sd[p + "lin.weight"].shape[0])
    elif gnn_type == "GraphSAGE":
        conv = pyg.SAGEConv(C, sd[p + "lin_l.weight"].shape[0])
    elif gnn_type == "GIN":
        out = sd[p + "nn.0.weight"].shape[0]
        conv = pyg.GINConv(nn.Sequential(nn.Linear(C, out), nn.ReLU(), nn.Linear(out, out)))
    else:
        raise ValueError(f"Unknown GNN type: {gnn_type}")
    conv.load_state_dict({k[len(p):]: v for k, v in sd.items() if k.startswith(p)}, strict=True)
    return conv.to(dt)(x, edge_index)


def bn_eval(sd, i, x):
    p = f"gnn.norms.{i}.module."
    dt = x.dtype
    return F.batch_norm(x, sd[p + "running_mean"].to(dt), sd[p + "running_var"].to(dt),
                        sd[p + "weight"].to(dt), sd[p + "bias"].to(dt), False, 0.1, 1e-5)


def forward(sd, cfg: ModelCfg, x, edge_index, edge_attr, dtype=torch.float32,
            return_hidden: bool = False, bf16_storage: bool = False) -> Dict[str, torch.Tensor]:
    """gnn.py:381-408 in eval mode (dropout = identity).

    `bf16_storage=True` (with dtype fp32) restates the rounding points of the CUDA path's bf16 mode:
    the feature-extractor output, the projection weights, every `h = x W^T` and every layer output
    except the last are rounded to bf16; all sums, the softmax, BN and the heads stay in fp32."""
    with torch.no_grad():
        x = torch.as_tensor(x).to(dtype)
        edge_index = torch.as_tensor(edge_index).long()
        edge_attr = None if edge_attr is None else torch.as_tensor(edge_attr).to(dtype)
        lin = lambda p, t: t @ sd[p + ".weight"].to(dtype).t() + sd[p + ".bias"].to(dtype)
        x = lin("feature_extractor.mlp.3", F.relu(lin("feature_extractor.mlp.0", x)))
        if bf16_storage:
            x = bf16_round(x)
        hidden = [x]
        for i, (_, H, concat) in enumerate(layer_dims(cfg)):
            if cfg.gnn_type == "GAT":
                x = gat_layer(sd, i, x, edge_index, edge_attr, H, cfg.hidden, concat, bf16_storage=bf16_storage)
            else:
                x = plain_layer(sd, i, x, edge_index, cfg.gnn_type)
            x = bn_eval(sd, i, x)
            if i < cfg.layers - 1:
                x = F.relu(x)
                if bf16_storage:
                    x = bf16_round(x)
            hidden.append(x)
        head = lambda p: lin(p + ".mlp.3", F.relu(lin(p + ".mlp.0", x)))
        logits = head("classification_head")
        probs = F.softmax(logits, dim=-1)
        out = dict(class_logits=logits, class_probs=probs,
                   predicted_class=torch.argmax(probs, dim=-1),
                   confidence=torch.sigmoid(head("confidence_head")).squeeze(-1))
        if cfg.predict_correction:
            out["correction"] = head("correction_head").squeeze(-1)
        if return_hidden:
            out["hidden"] = hidden
        return out


def predict(sd, cfg, x, edge_index, edge_attr, auto_correct_threshold=0.85,
            review_threshold=0.6, dtype=torch.float32):
    """gnn.py:410-451."""
    out = forward(sd, cfg, x, edge_index, edge_attr, dtype)
    conf, cls = out["confidence"], out["predicted_class"]
    action = torch.zeros_like(cls)
    action[(cls == 2) & (conf > auto_correct_threshold)] = 1
    action[conf < review_threshold] = 2
    out["action"] = action
    out["needs_review"] = action == 2
    out["auto_correct"] = action == 1
    return out


def forward_row_bands(sd, cfg: ModelCfg, x, edge_index, edge_attr, node_rows, radius: int = 1, band_rows: int = 128,
                      keys=("class_logits", "confidence", "correction", "predicted_class"), **kw):
    """`forward` evaluated band by band over DESTINATION rows, for tiles whose `[E+N, H*C]` per-edge tensors do not fit
    in host memory (gnn.py:381-408 on 4.2 M nodes materialises 2 x 38.6 GB per layer).

    Nodes are in raster order, `node_rows[i]` = grid row of node i, `radius` = the stencil radius in rows (1 for the
    4-/8-connected lists, 2 for 5x5).  A band of rows [a, b) is evaluated on the sub-graph induced by the rows
    [a - L*radius, b + L*radius) (L = number of GAT layers): every node of the band then sees its complete L-hop
    receptive field, so the band's outputs are those of the whole-graph forward (same per-destination edge order =>
    same fp32 sums); halo rows are discarded."""
    x = np.asarray(x)
    ei = np.asarray(edge_index)
    ea = None if edge_attr is None else np.asarray(edge_attr)
    node_rows = np.asarray(node_rows)
    assert np.all(np.diff(node_rows) >= 0), "nodes must be in raster order"
    n_rows = int(node_rows.max()) + 1 if node_rows.size else 0
    halo = cfg.layers * radius
    out = {k: [] for k in keys}
    src, dst = ei[0], ei[1]
    for a in range(0, n_rows, band_rows):
        b = min(a + band_rows, n_rows)
        lo, hi = np.searchsorted(node_rows, [a - halo, b + halo], side="left")
        own_lo, own_hi = np.searchsorted(node_rows, [a, b], side="left")
        sel = (src >= lo) & (src < hi) & (dst >= lo) & (dst < hi)
        o = forward(sd, cfg, x[lo:hi], ei[:, sel] - lo, None if ea is None else ea[sel], **kw)
        for k in keys:
            if k in o:
                out[k].append(o[k][own_lo - lo:own_hi - lo])
    return {k: torch.cat(v) for k, v in out.items() if v}
