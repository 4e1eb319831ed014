"""TEST INFRASTRUCTURE ONLY -- CPU oracle, never imported by the product path.

NumPy/SciPy restatement of the reference's spatial-graph builder
(`/root/reference/data/graph_construction.py`).  Pinned (bit-for-bit on integer
outputs and on every float column) against the UNMODIFIED reference run in the
build container: tests/golden/make_golden.py -> tests/golden/graph_*.npz, checked by
tests/test_oracle_golden.py.

The one deliberate difference from the reference is speed: the reference builds
edge features in a pure-Python per-edge loop (graph_construction.py:342-369,
95 % of its build time); here the same arithmetic is vectorised (same dtypes,
same operation order per element), which is what lets the oracle finish at test
sizes in seconds.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy import ndimage

# graph_construction.py:79-87
OFFSETS_4 = [(-1, 0), (1, 0), (0, -1), (0, 1)]
OFFSETS_8 = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
# BASELINE config 5 "k=24" has no reference code path; the faithful
# oracle is the reference loop fed a 24-entry offset list (5x5 minus centre, raster order).
OFFSETS_24 = [(dr, dc) for dr in range(-2, 3) for dc in range(-2, 3) if (dr, dc) != (0, 0)]

DEFAULT_NODE_FEATURES = ["depth", "local_mean", "local_std", "gradient_x",
                         "gradient_y", "gradient_magnitude", "curvature"]   # :60-68
DEFAULT_EDGE_FEATURES = ["distance", "depth_difference", "slope"]          # :71-75


def offsets_for(connectivity) -> List[Tuple[int, int]]:
    if connectivity == "4-connected":
        return list(OFFSETS_4)
    if connectivity == "8-connected":
        return list(OFFSETS_8)
    if connectivity in ("24-connected", "5x5"):
        return list(OFFSETS_24)
    if isinstance(connectivity, (list, tuple)):
        return [tuple(o) for o in connectivity]
    raise ValueError(f"Unknown connectivity: {connectivity}")          # :89


def masked_local_stats(depth: np.ndarray, valid_mask: np.ndarray, size: int = 5):
    """graph_construction.py:378-432 (same SciPy calls, same dtypes)."""
    depth_masked = np.where(valid_mask, depth, 0.0).astype(np.float64)
    valid_float = valid_mask.astype(np.float64)
    area = float(size * size)
    sum_vals = ndimage.uniform_filter(depth_masked, size=size, mode="constant", cval=0.0) * area
    count = ndimage.uniform_filter(valid_float, size=size, mode="constant", cval=0.0) * area
    safe = np.maximum(count, 1.0)
    local_mean = (sum_vals / safe).astype(np.float32)
    with np.errstate(all="ignore"):
        depth_sq = np.where(valid_mask, depth.astype(np.float64) ** 2, 0.0)
    sum_sq = ndimage.uniform_filter(depth_sq, size=size, mode="constant", cval=0.0) * area
    var = np.maximum(sum_sq / safe - (sum_vals / safe) ** 2, 0.0)
    return local_mean, np.sqrt(var).astype(np.float32), count.astype(np.float32)


def curvature(depth_filled: np.ndarray, valid_mask: np.ndarray) -> np.ndarray:
    """graph_construction.py:434-458."""
    c = ndimage.laplace(depth_filled)
    cnt = ndimage.convolve(valid_mask.astype(np.float64), np.ones((3, 3)), mode="constant", cval=0.0)
    c[cnt < 3] = 0.0
    return c


def build_edges(valid_rows, valid_cols, node_index_grid, shape, offsets):
    """graph_construction.py:176-243: offset-major, node-minor COO (src = centre,
    tgt = neighbour).  Returns (edge_index int64 [2,E], src_r, src_c, tgt_r, tgt_c)."""
    H, W = shape
    ids = np.arange(len(valid_rows))
    srcs, tgts, sr, sc, tr, tc = [], [], [], [], [], []
    for dr, dc in offsets:
        nr, nc = valid_rows + dr, valid_cols + dc
        inb = (nr >= 0) & (nr < H) & (nc >= 0) & (nc < W)
        nidx = node_index_grid[np.clip(nr, 0, H - 1), np.clip(nc, 0, W - 1)]
        ok = inb & (nidx >= 0)
        srcs.append(ids[ok]); tgts.append(nidx[ok])
        sr.append(valid_rows[ok]); sc.append(valid_cols[ok]); tr.append(nr[ok]); tc.append(nc[ok])
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)
    return (np.stack([cat(srcs), cat(tgts)]).astype(np.int64), cat(sr), cat(sc), cat(tr), cat(tc))


def edge_features(depth, sr, sc, tr, tc, resolution, names: Sequence[str]):
    """graph_construction.py:329-376, vectorised with the reference's dtypes:
    dx, dy are (int64 * python float) = float64; dz is a float32 subtract
    (numpy scalar arithmetic keeps float32), divided by a float64 distance."""
    E = len(sr)
    if E == 0:
        return np.zeros((0, len(names)), np.float32)
    rx, ry = resolution
    dx = (tc - sc) * rx
    dy = (tr - sr) * ry
    dist = np.sqrt(dx ** 2 + dy ** 2)
    with np.errstate(all="ignore"):
        dz = depth[tr, tc] - depth[sr, sc]                       # dtype of depth (fp32)
        cols = []
        for n in names:
            if n == "distance":
                v = dist
            elif n == "depth_difference":
                v = dz
            elif n == "slope":
                v = np.where(dist > 0, np.degrees(np.arctan(dz / np.where(dist > 0, dist, 1.0))), 0.0)
            else:
                v = np.zeros(E)
            # reference: list of scalars -> np.nan_to_num(list) promotes to float64
            cols.append(np.nan_to_num(np.asarray(v, dtype=np.float64), nan=0.0))
    with np.errstate(over="ignore"):
        return np.stack(cols, axis=1).astype(np.float32)


def build_graph(depth: np.ndarray, valid_mask: Optional[np.ndarray] = None,
                uncertainty: Optional[np.ndarray] = None,
                resolution: Tuple[float, float] = (1.0, 1.0),
                connectivity="8-connected",
                node_features: Optional[Sequence[str]] = None,
                edge_feature_names: Optional[Sequence[str]] = None) -> Dict[str, np.ndarray]:
    """graph_construction.py:91-174.  Returns a dict of numpy arrays with the
    reference `Data` field names."""
    node_features = list(node_features or DEFAULT_NODE_FEATURES)
    edge_feature_names = list(edge_feature_names or DEFAULT_EDGE_FEATURES)
    offsets = offsets_for(connectivity)
    if valid_mask is None:
        valid_mask = np.isfinite(depth)
    valid_rows, valid_cols = np.where(valid_mask)
    N = len(valid_rows)
    if N == 0:                                                  # :119-121, 460-469
        return dict(x=np.zeros((0, len(node_features)), np.float32),
                    edge_index=np.zeros((2, 0), np.int64),
                    edge_attr=np.zeros((0, len(edge_feature_names)), np.float32),
                    pos=np.zeros((0, 2), np.float32), local_std=np.zeros(0, np.float32),
                    valid_rows=np.zeros(0, np.int64), valid_cols=np.zeros(0, np.int64),
                    num_valid_cells=0, grid_shape=None)
    nig = np.full(depth.shape, -1, np.int64)
    nig[valid_rows, valid_cols] = np.arange(N)
    edge_index, sr, sc, tr, tc = build_edges(valid_rows, valid_cols, nig, depth.shape, offsets)

    # :245-327
    local_mean, local_std, _ = masked_local_stats(depth, valid_mask, 5)
    filled = np.nan_to_num(np.where(valid_mask, depth, local_mean), nan=0.0)
    with np.errstate(all="ignore"):
        grad_y, grad_x = np.gradient(filled)
        grad_mag = np.sqrt(grad_x ** 2 + grad_y ** 2)
        curv = curvature(filled, valid_mask)
    planes = dict(depth=depth, local_mean=local_mean, local_std=local_std, gradient_x=grad_x,
                  gradient_y=grad_y, gradient_magnitude=grad_mag, curvature=curv)
    feats = []
    for n in node_features:
        if n in planes:
            f = planes[n][valid_rows, valid_cols]
        elif n == "uncertainty" and uncertainty is not None:
            f = uncertainty[valid_rows, valid_cols]
        else:
            continue
        feats.append(np.nan_to_num(f, nan=0.0))
    if uncertainty is not None and "uncertainty" not in node_features:
        feats.append(np.nan_to_num(uncertainty[valid_rows, valid_cols], nan=0.0))
    x = np.stack(feats, axis=1).astype(np.float32)
    node_std = np.nan_to_num(local_std[valid_rows, valid_cols], nan=0.0).astype(np.float32)

    ea = edge_features(depth, sr, sc, tr, tc, resolution, edge_feature_names)
    pos = np.stack([valid_cols, valid_rows], axis=1).astype(np.float32)
    return dict(x=x, edge_index=edge_index, edge_attr=ea, pos=pos, local_std=node_std,
                valid_rows=valid_rows.astype(np.int64), valid_cols=valid_cols.astype(np.int64),
                num_valid_cells=N, grid_shape=tuple(depth.shape))


def graph_to_grid(grid_shape, valid_rows, valid_cols, node_values, fill_value=np.nan) -> np.ndarray:
    """graph_construction.py:471-505."""
    node_values = np.asarray(node_values)
    if node_values.ndim != 1:
        raise ValueError("For multi-channel node values, call graph_to_grid for each channel")
    grid = np.full(grid_shape, fill_value, dtype=np.float32)
    grid[valid_rows, valid_cols] = node_values
    return grid


def coo_to_in_csr(edge_index: np.ndarray, edge_attr: np.ndarray, num_nodes: int):
    """Canonical comparison form: stable sort of the reference
    COO by target = CSR of in-edges, each row in reference edge order."""
    order = np.argsort(edge_index[1], kind="stable")
    rowptr = np.zeros(num_nodes + 1, np.int64)
    np.add.at(rowptr, edge_index[1] + 1, 1)
    return np.cumsum(rowptr), edge_index[0][order], edge_attr[order], order


def build_multiscale_graph(depth, valid_mask=None, uncertainty=None, resolution=(1.0, 1.0), scales=(1, 2, 4), **kw):
    """MultiScaleGraphBuilder.build_multiscale_graph, graph_construction.py:508-607: {scale: build_graph dict}.  The block
    reductions are the reference's own NumPy calls (np.nanmean / np.mean over the [h/s, s, w/s, s] view)."""
    if valid_mask is None:
        valid_mask = np.isfinite(depth)
    out = {}
    for s in sorted(scales):
        if s == 1:
            d, m, u, res = depth, valid_mask, uncertainty, resolution
        else:
            blk = lambda a: a[:a.shape[0] // s * s, :a.shape[1] // s * s].reshape(a.shape[0] // s, s, a.shape[1] // s, s)
            d = np.nanmean(blk(depth), axis=(1, 3))                                   # :582-594
            m = np.mean(blk(valid_mask), axis=(1, 3)) >= 0.5                          # :596-606
            u = np.nanmean(blk(uncertainty), axis=(1, 3)) if uncertainty is not None else None
            res = (resolution[0] * s, resolution[1] * s)
        out[s] = build_graph(d, m, u, res, **kw)
    return out
