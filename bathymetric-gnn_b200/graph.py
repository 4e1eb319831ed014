"""Graph containers and the device GraphBuilder.

Mirrors the reference interface of data/graph_construction.py (`GraphBuilder`,
:29-505) and the `torch_geometric.data.{Data,Batch}` surface the reference touches
(SURVEY.md 8(b) "Data container"), with every array produced by the sm_100a
kernels behind include/bgnn.h.  No torch_geometric, no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import logging
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

logger = logging.getLogger(__name__)

OFFSETS_4 = [(-1, 0), (1, 0), (0, -1), (0, 1)]                                   # graph_construction.py:79-81
OFFSETS_8 = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]   # :83-87
OFFSETS_24 = [(dr, dc) for dr in range(-2, 3) for dc in range(-2, 3) if (dr, dc) != (0, 0)]


class Data:
    """Attribute bag with the PyG `Data` surface used by the reference."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, pos=None, **kwargs):
        if x is not None:
            self.x = x
        if edge_index is not None:
            self.edge_index = edge_index
        if edge_attr is not None:
            self.edge_attr = edge_attr
        if pos is not None:
            self.pos = pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def num_nodes(self) -> int:
        x = self.__dict__.get("x")
        return 0 if x is None else int(x.shape[0])

    @property
    def num_edges(self) -> int:
        ei = getattr(self, "edge_index", None)
        return 0 if ei is None else int(ei.shape[1])

    def _apply(self, fn):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = fn(v) if isinstance(v, torch.Tensor) and not k.startswith("_") else v
        return out

    def to(self, device, *a, **k):
        return self._apply(lambda t: t.to(device, *a, **k))

    def cpu(self):
        return self._apply(lambda t: t.cpu())

    def cuda(self, device=None):
        return self._apply(lambda t: t.cuda(device))

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]


class _DeviceGraph:
    """Device-resident lean graph (what the forward kernels consume).  Shared by a
    GridGraph and its `.cpu()` / `.to()` copies.  The in-edge CSR payload (`csr_src`,
    `csr_eattr`) is emitted on first access: the tiled aggregation of a single tile reads
    the dense per-cell `grid_eattr` instead and never needs it."""
    __slots__ = ("cells", "cfg", "scratch", "keep", "node_id", "node_cell", "rowptr", "_csr_src",
                 "_csr_eattr", "grid_eattr", "x", "local_std", "N", "E", "n_cells", "device", "grid_meta", "keep_versions", "scratch_borrowed", "rag_patches", "n_rag_patches")

    def check_sources_unchanged(self):
        """`build_graph` consumes CUDA planes IN PLACE (no tile copy) and the lazily emitted arrays (`edge_index`, `edge_attr`,
        the CSR payload) re-read them: an in-place edit of the caller's depth / mask / uncertainty tensor between the build and
        the first access would yield edge features inconsistent with `x` / `rowptr`.  The tensors' version counters catch it."""
        if getattr(self, "scratch_borrowed", False):
            raise RuntimeError("this graph was built on a caller-owned scratch buffer (pipelined survey tiles): edge_index / edge_attr "
                               "/ the CSR payload are not available; use GraphBuilder.build_graph for a self-contained graph")
        vs = getattr(self, "keep_versions", None)
        if vs is not None and any(t._version != v for t, v in zip(self.keep, vs)):
            raise RuntimeError("the depth / mask / uncertainty tensor this graph was built from has been modified in place; "
                               "access edge_index / edge_attr before editing the survey, or build the graph from a copy")

    def ensure_csr(self):
        if getattr(self, "_csr_src", None) is not None:
            return
        self.check_sources_unchanged()
        ED = int(self.cfg.n_edge_feats)
        with torch.cuda.device(self.device):
            src = torch.empty(max(self.E, 1), dtype=torch.int32, device=self.device)
            ea = torch.empty((max(self.E, 1), max(ED, 1)), dtype=torch.float32, device=self.device)
            check(lib.bgnn_graph_emit(C.byref(self.cells), C.byref(self.cfg), ptr(self.scratch), self.scratch.numel(),
                                      ptr(self.node_id), None, None, None, ptr(self.rowptr), ptr(src),
                                      ptr(ea) if ED > 0 else None, None, stream_ptr()))
        self._csr_src, self._csr_eattr = src, ea

    @property
    def csr_src(self):
        self.ensure_csr()
        return self._csr_src

    @property
    def csr_eattr(self):
        self.ensure_csr()
        return self._csr_eattr


class GridGraph(Data):
    """`Data` built by the device GraphBuilder.  `x`, `local_std` are materialised;
    `edge_index` (int64 COO, reference order), `edge_attr`, `pos`, `valid_rows`,
    `valid_cols` are emitted by a kernel on first access (the forward pass itself
    consumes the in-edge CSR and never needs them)."""

    _LAZY = ("edge_index", "edge_attr", "pos", "valid_rows", "valid_cols")

    def __getattr__(self, name):
        if name in GridGraph._LAZY and "_dev" in self.__dict__:
            self._materialise(name)
            return self.__dict__[name]
        raise AttributeError(name)

    @property
    def num_edges(self) -> int:
        return int(self._dev.E) + (int(self._dev.N) if self.__dict__.get("_self_loops") else 0)

    def _materialise(self, name):
        dg = self._dev
        tgt = self.__dict__.get("_target_device", dg.device)
        with torch.cuda.device(dg.device):
            if name in ("edge_index", "edge_attr"):
                dg.check_sources_unchanged()
                ED = dg.cfg.n_edge_feats
                ei = torch.empty((2, dg.E), dtype=torch.int64, device=dg.device)
                ea = torch.empty((dg.E, ED), dtype=torch.float32, device=dg.device)
                check(lib.bgnn_graph_coo(C.byref(dg.cells), C.byref(dg.cfg), ptr(dg.scratch), dg.scratch.numel(),
                                         ptr(dg.node_id), dg.E, ptr(ei), ptr(ea), stream_ptr()))
                if self.__dict__.get("_self_loops"):
                    # include_self_loops=True (graph_construction.py:226-229): N (i, i) edges appended AFTER the stencil
                    # edges; their features are distance 0, depth_difference 0, slope 0 (:344-364)
                    loops = torch.arange(dg.N, dtype=torch.int64, device=dg.device)
                    ei = torch.cat([ei, torch.stack([loops, loops])], dim=1)
                    ea = torch.cat([ea, torch.zeros((dg.N, ED), dtype=torch.float32, device=dg.device)])
                self.__dict__["edge_index"] = ei.to(tgt)
                self.__dict__["edge_attr"] = ea.to(tgt)
                self.__dict__["_emitted"] = dict(edge_index=self.__dict__["edge_index"], edge_attr=self.__dict__["edge_attr"])
            else:
                cell = dg.node_cell.to(torch.int64)
                if dg.grid_meta is None:
                    W = int(dg.cells.cols)
                    rows, cols = torch.div(cell, W, rounding_mode="floor"), cell % W
                else:                               # ragged batch: rows/cols local to each grid
                    off, gcols = dg.grid_meta["cell_off_dev"], dg.grid_meta["cols_dev"]
                    g = torch.searchsorted(off, cell, right=True) - 1
                    local = cell - off[g]
                    W = gcols[g].to(torch.int64)
                    rows, cols = torch.div(local, W, rounding_mode="floor"), local % W
                self.__dict__["valid_rows"] = rows.to(tgt)
                self.__dict__["valid_cols"] = cols.to(tgt)
                self.__dict__["pos"] = torch.stack([cols, rows], dim=1).to(torch.float32).to(tgt)

    def _apply(self, fn):
        out = super()._apply(fn)
        probe = fn(torch.empty(0, device=self._dev.device))
        out.__dict__["_target_device"] = probe.device
        em = self.__dict__.get("_emitted")
        if em is not None:          # moved copies of kernel-emitted tensors still count as kernel-emitted
            out.__dict__["_emitted"] = {k: out.__dict__[k] for k, v in em.items() if self.__dict__.get(k) is v}
        return out

    def topology_overridden(self) -> bool:
        """True when `edge_index` / `edge_attr` were assigned by the caller after the build: the cached device
        topology no longer describes this object and the model falls back to the COO path."""
        em = self.__dict__.get("_emitted", {})
        return any(k in self.__dict__ and self.__dict__[k] is not em.get(k) for k in ("edge_index", "edge_attr"))


class HostGridGraph(GridGraph):
    """A SMALL host grid (NumPy depth, <= GraphBuilder.DEFER_MAX_CELLS cells) whose device build is DEFERRED.

    `scripts/inference_native.py` builds one graph per refinement grid (`_build_graph_for_grid`, :158-179), reads only
    `graph.num_nodes`, and hands the list to `Batch.from_data_list` (:312).  One device build per 3x3 ... 50x50 grid costs a
    host round trip each (~150 us per grid, far more than its kernels), so `build_graph` returns this object instead: the node
    count comes from the mask on the host, and `Batch.from_data_list` then builds the WHOLE list as one ragged device graph
    (`GraphBuilder.build_graph_batch`).  Anything else that is asked of it (`x`, `edge_index`, `edge_attr`, `local_std`, the
    model's forward ...) triggers the ordinary single-grid device build first, so the object behaves exactly like the eager
    `GridGraph`; `valid_rows` / `valid_cols` / `pos` come from the host mask (what the reference's CPU graph holds)."""

    _DEV_NAMES = ("_dev", "x", "local_std", "edge_index", "edge_attr")

    def _materialise_host(self):
        h = self.__dict__.pop("_host")
        g = h["builder"]._build_now(h["depth"], h["mask"], h["unc"], h["res"])
        self.__dict__.update(g.__dict__)         # _dev, x, local_std, ... (a local_std slice handed over by a batch build is the same data)

    def __getattr__(self, name):
        d = self.__dict__
        if "_host" in d:
            if name in HostGridGraph._DEV_NAMES:
                self._materialise_host()
                if name in self.__dict__:
                    return self.__dict__[name]
            elif name in ("valid_rows", "valid_cols", "pos"):
                rr, cc = np.nonzero(d["_host"]["mask"])
                d["valid_rows"], d["valid_cols"] = torch.from_numpy(rr.astype(np.int64)), torch.from_numpy(cc.astype(np.int64))
                d["pos"] = torch.from_numpy(np.stack([cc, rr], axis=1).astype(np.float32))
                return d[name]
        return super().__getattr__(name)

    @property
    def num_nodes(self) -> int:
        h = self.__dict__.get("_host")
        return int(h["n"]) if h is not None else super().num_nodes

    @property
    def num_edges(self) -> int:
        if "_host" in self.__dict__:
            self._materialise_host()
        return super().num_edges

    def _apply(self, fn):
        if "_host" in self.__dict__:                 # nothing lives on a device yet: .to() / .cpu() are the identity
            return self
        return super()._apply(fn)

    def keys(self):
        if "_host" in self.__dict__:
            self._materialise_host()
        return super().keys()

    def topology_overridden(self) -> bool:
        if "_host" in self.__dict__:
            return any(k in self.__dict__ for k in ("edge_index", "edge_attr"))
        return super().topology_overridden()


class Batch(Data):
    """`torch_geometric.data.Batch.from_data_list` replacement
    (scripts/inference_native.py:312): block-diagonal concatenation."""

    _LAZY = GridGraph._LAZY

    def __getattr__(self, name):
        rag = self.__dict__.get("_ragged")
        if rag is not None and name in Batch._LAZY:          # built as ONE ragged device graph: its lazily emitted arrays
            tgt = self.__dict__.get("_target_device", self.__dict__["_dev"].device)
            self.__dict__[name] = getattr(rag, name).to(tgt)
            return self.__dict__[name]
        parts = self.__dict__.get("_parts")
        if parts is not None and name in Batch._LAZY:
            tgt = self.__dict__.get("_target_device", self.__dict__["_dev"].device)
            vs, offs = [], 0
            for d in parts:
                v = getattr(d, name).to(tgt)
                vs.append(v + offs if name == "edge_index" else v)
                offs += d.num_nodes
            self.__dict__[name] = torch.cat(vs, dim=1 if name == "edge_index" else 0)
            return self.__dict__[name]
        raise AttributeError(name)

    def _apply(self, fn):
        out = super()._apply(fn)
        if "_dev" in self.__dict__:
            out.__dict__["_target_device"] = fn(torch.empty(0, device=self._dev.device)).device
        return out

    @property
    def num_edges(self) -> int:
        dg = self.__dict__.get("_dev")
        return int(dg.E) if dg is not None else super().num_edges

    @classmethod
    def _from_grid_graphs(cls, graphs: List["GridGraph"]) -> "Batch":
        """All members come from the device GraphBuilder: concatenate their device arrays (node features, in-edge CSR
        with node / edge offsets) into one block-diagonal device graph, so that `model.predict(batch)` runs the CSR /
        windowed aggregation kernels directly; the int64 COO is still available lazily."""
        dgs = [g._dev for g in graphs]
        dev = dgs[0].device
        with torch.cuda.device(dev):
            n = torch.tensor([d.N for d in dgs], dtype=torch.int64)
            e = torch.tensor([d.E for d in dgs], dtype=torch.int64)
            noff = torch.zeros(len(dgs) + 1, dtype=torch.int64); noff[1:] = torch.cumsum(n, 0)
            eoff = torch.zeros(len(dgs) + 1, dtype=torch.int64); eoff[1:] = torch.cumsum(e, 0)
            N, E = int(noff[-1]), int(eoff[-1])
            if N >= 2 ** 31 or E >= 2 ** 31:
                raise ValueError("batch exceeds the int32 node / edge id range")
            dg = _DeviceGraph()
            dg.cells, dg.cfg, dg.scratch, dg.device = _lib.Cells(), dgs[0].cfg, None, dev
            dg.cells.n_grids, dg.cells.n_cells = len(dgs), sum(d.n_cells for d in dgs)
            dg.keep = [graphs]                       # the members own the inputs
            dg.grid_meta = dict(batched=True)
            dg.N, dg.E, dg.n_cells = N, E, int(dg.cells.n_cells)
            dg.x = torch.cat([d.x for d in dgs])
            dg.local_std = torch.cat([d.local_std for d in dgs])
            rp = [d.rowptr[:-1] + int(eoff[i]) for i, d in enumerate(dgs)]
            dg.rowptr = torch.cat(rp + [torch.tensor([E], dtype=torch.int32, device=dev)]).to(torch.int32)
            dg._csr_src = torch.cat([d.csr_src[:d.E] + int(noff[i]) for i, d in enumerate(dgs)]
                                    + [torch.zeros(1 if E == 0 else 0, dtype=torch.int32, device=dev)]).to(torch.int32)
            ED = max(int(dg.cfg.n_edge_feats), 1)
            dg._csr_eattr = torch.cat([d.csr_eattr[:d.E] for d in dgs]
                                      + [torch.zeros((1 if E == 0 else 0, ED), dtype=torch.float32, device=dev)])
            dg.node_id = dg.node_cell = dg.grid_eattr = None
            out = cls()
            out.__dict__["_dev"], out.__dict__["_parts"] = dg, list(graphs)
            out.x, out.local_std = dg.x, dg.local_std
            out.batch = torch.repeat_interleave(torch.arange(len(dgs), device=dev), n.to(dev))
            out.num_graphs = len(graphs)
            out.num_valid_cells = [g.num_valid_cells for g in graphs]
            out.grid_shape = [getattr(g, "grid_shape", None) for g in graphs]
        return out

    @classmethod
    def _from_host_graphs(cls, graphs: List["HostGridGraph"]) -> "Batch":
        """All members are deferred host grids of one builder: ONE ragged device build for the whole list (per-grid boundary
        conditions, explicit masks), the windowed / band-patch aggregation runs on it directly."""
        hs = [g.__dict__["_host"] for g in graphs]
        b = hs[0]["builder"]
        use_unc = all(h["unc"] is not None for h in hs)
        rag = b.build_graph_batch([(h["depth"], h["unc"], h["res"]) for h in hs], nodata=None, use_uncertainty=use_unc,
                                  masks=[h["mask"] for h in hs])
        out = cls()
        counts = np.array([h["n"] for h in hs], np.int64)
        out.__dict__["_dev"], out.__dict__["_ragged"] = rag._dev, rag
        out.x, out.local_std = rag.x, rag.local_std
        out.batch = torch.from_numpy(np.repeat(np.arange(len(hs), dtype=np.int64), counts))
        out.num_graphs = len(graphs)
        out.num_valid_cells = [int(c) for c in counts]
        out.grid_shape = [tuple(h["depth"].shape) for h in hs]
        # the per-grid correction scale the script reads after unbatching (`graph.local_std`, inference_native.py:181-204):
        # one read-back for the whole batch, handed to the members as views
        ls = rag.local_std.cpu()
        off = np.concatenate([[0], np.cumsum(counts)])
        for i, g in enumerate(graphs):
            g.__dict__["local_std"] = ls[int(off[i]):int(off[i + 1])]
        return out

    @classmethod
    def from_data_list(cls, data_list: List[Data]) -> "Batch":
        if data_list and all(isinstance(d, HostGridGraph) and "_host" in d.__dict__ and not d.topology_overridden()
                             for d in data_list):
            h0 = data_list[0].__dict__["_host"]
            if all(d.__dict__["_host"]["builder"] is h0["builder"] for d in data_list):
                return cls._from_host_graphs(list(data_list))
        if data_list and all(isinstance(d, GridGraph) and "_dev" in d.__dict__ and not d.topology_overridden()
                             and d.x is d._dev.x for d in data_list):
            d0 = data_list[0]._dev
            same = lambda d: (d.device == d0.device and bytes(d.cfg) == bytes(d0.cfg))
            if all(same(d._dev) for d in data_list):
                return cls._from_grid_graphs(list(data_list))
        out = cls()
        offs, cat, batch_vec = 0, {}, []
        for g, d in enumerate(data_list):
            n = d.num_nodes
            keys = list(d.keys())
            if isinstance(d, GridGraph):
                keys += [k for k in GridGraph._LAZY if k not in keys]
            for k in keys:
                v = getattr(d, k)
                if isinstance(v, torch.Tensor) and k == "edge_index":
                    v = v + offs
                cat.setdefault(k, []).append(v)
            batch_vec.append(torch.full((n,), g, dtype=torch.long, device=d.x.device))
            offs += n
        for k, vs in cat.items():
            if isinstance(vs[0], torch.Tensor):
                setattr(out, k, torch.cat(vs, dim=1 if k == "edge_index" else 0))
            else:
                setattr(out, k, vs)
        out.batch = torch.cat(batch_vec) if batch_vec else torch.zeros(0, dtype=torch.long)
        out.num_graphs = len(data_list)
        return out


RAG_PATCH_DTYPE = np.dtype([("cell_off", np.int64), ("gh", np.int32), ("gw", np.int32), ("r0", np.int32), ("nr", np.int32)])


def _ragged_patches(cell_off: np.ndarray, rows: np.ndarray, cols: np.ndarray) -> Optional[np.ndarray]:
    """Band patches of a ragged batch (include/bgnn.h: bgnn_grid_topo::patches): every grid is cut into bands of `nr` rows over
    its full width with (nr + 2)(cols + 2) <= 324 halo slots, nr * cols <= 256 destinations and nr <= 62.  None when a grid is
    too wide for one band row (cols > 106): the forward then uses the windowed CSR kernel."""
    rows = np.asarray(rows, np.int64); cols = np.asarray(cols, np.int64)
    if rows.size == 0 or int(cols.max()) > 106:
        return None
    ph = np.minimum(np.minimum(256 // cols, 324 // (cols + 2) - 2), 62)
    ph = np.maximum(np.minimum(ph, rows), 1)
    nb = (rows + ph - 1) // ph
    total = int(nb.sum())
    g = np.repeat(np.arange(rows.size), nb)
    first = np.zeros(rows.size, np.int64); np.cumsum(nb[:-1], out=first[1:])
    band = np.arange(total, dtype=np.int64) - first[g]
    t = np.empty(total, RAG_PATCH_DTYPE)
    t["cell_off"] = np.asarray(cell_off, np.int64)[:-1][g]
    t["gh"], t["gw"] = rows[g], cols[g]
    t["r0"] = band * ph[g]
    t["nr"] = np.minimum(ph[g], rows[g] - band * ph[g])
    return t


def _graph_cfg(offsets, node_feats: Sequence[str], has_unc: bool, edge_feats: Sequence[str]) -> Tuple[_lib.GraphCfg, int]:
    cfg = _lib.GraphCfg()
    if len(offsets) > _lib.MAX_OFFSETS:
        raise ValueError(f"at most {_lib.MAX_OFFSETS} neighbour offsets are supported")
    cfg.n_offsets = len(offsets)
    for k, (dr, dc) in enumerate(offsets):
        cfg.off_dr[k], cfg.off_dc[k] = int(dr), int(dc)
    # node feature columns exactly as the reference assembles them (:288-316)
    cols = []
    for n in node_feats:
        if n in _lib.NF_CODES and n != "uncertainty":
            cols.append(_lib.NF_CODES[n])
        elif n == "uncertainty" and has_unc:
            cols.append(_lib.NF_CODES["uncertainty"])
    if has_unc and "uncertainty" not in node_feats:
        cols.append(_lib.NF_CODES["uncertainty"])
    if len(cols) > _lib.MAX_NODE_FEATS:
        raise ValueError(f"at most {_lib.MAX_NODE_FEATS} node features are supported")
    cfg.n_node_feats = len(cols)
    for j, c in enumerate(cols):
        cfg.node_feats[j] = c
    if len(edge_feats) > _lib.MAX_EDGE_FEATS:
        raise ValueError(f"at most {_lib.MAX_EDGE_FEATS} edge features are supported")
    cfg.n_edge_feats = len(edge_feats)
    for j, n in enumerate(edge_feats):
        cfg.edge_feats[j] = _lib.EF_CODES.get(n, _lib.EF_ZERO)       # unknown name -> 0.0 (:366-367)
    return cfg, len(cols)


def _as_dev_f32(a, device) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device, non_blocking=True)
    return t


class GraphBuilder:
    """Device drop-in for the reference `GraphBuilder` (graph_construction.py:29-505):
    same constructor, `build_graph`, `graph_to_grid`; `neighbor_offsets` stays a
    plain list attribute that may be overridden (e.g. the 24-offset 5x5 stencil)."""

    def __init__(self, connectivity: str = "8-connected", include_self_loops: bool = False,
                 node_features: Optional[List[str]] = None, edge_features: Optional[List[str]] = None,
                 device: Optional[torch.device] = None):
        self.connectivity = connectivity
        self.include_self_loops = include_self_loops
        self.node_features = node_features or ["depth", "local_mean", "local_std", "gradient_x", "gradient_y",
                                               "gradient_magnitude", "curvature"]
        self.edge_features = edge_features or ["distance", "depth_difference", "slope"]
        if connectivity == "4-connected":
            self.neighbor_offsets = list(OFFSETS_4)
        elif connectivity == "8-connected":
            self.neighbor_offsets = list(OFFSETS_8)
        elif connectivity == "24-connected":          # additive: BASELINE config 5 (k=24)
            self.neighbor_offsets = list(OFFSETS_24)
        else:
            raise ValueError(f"Unknown connectivity: {connectivity}")
        self.device = device

    # ------------------------------------------------------------------ build
    def _staging(self, nbytes: int) -> torch.Tensor:
        """Pinned host staging bytes owned by this builder (grown geometrically, reused by every ragged batch)."""
        buf = self.__dict__.get("_stage")
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(int(nbytes * 1.5), 1 << 20), dtype=torch.uint8, pin_memory=True)
            self.__dict__["_stage"] = buf
        return buf

    def _device(self) -> torch.device:
        _lib.require_cuda()
        return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())

    def build_graph(self, depth, valid_mask=None, uncertainty=None,
                    resolution: Tuple[float, float] = (1.0, 1.0), nodata: Optional[float] = None) -> Data:
        """graph_construction.py:91-174.  `depth` may be a NumPy array or a CUDA
        tensor (a strided 2-D view of a survey is consumed in place).  `nodata`
        (additive) makes the kernel derive valid = isfinite & != nodata when no mask
        is given (scripts/inference_native.py:160)."""
        if (isinstance(depth, np.ndarray) and depth.ndim == 2 and 0 < depth.size <= self.DEFER_MAX_CELLS
                and (valid_mask is None or isinstance(valid_mask, np.ndarray))
                and (uncertainty is None or isinstance(uncertainty, np.ndarray))):
            return self._build_deferred(depth, valid_mask, uncertainty, resolution, nodata)
        dev = self._device()
        with torch.cuda.device(dev):
            return self._build(*self._tile_cells(depth, valid_mask, uncertainty, resolution, nodata, dev))

    DEFER_MAX_CELLS = 4096       # host grids up to 64 x 64 (native VR refinement grids are 3x3 ... 50x50) are built on first use

    def _build_deferred(self, depth, valid_mask, uncertainty, resolution, nodata) -> Data:
        """`build_graph` for a small host grid: everything the callers of scripts/inference_native.py:158-179, 249-269 read
        right away (node count, grid_shape) comes from the host mask; the device build happens on first use of a device
        array, or -- for a whole list at once -- in `Batch.from_data_list`.  Same errors as the eager path."""
        _lib.require_cuda()
        d = np.array(depth, dtype=np.float32)                      # private copies: the caller may reuse its buffers before the
        if valid_mask is not None:                                  # deferred device build runs (the eager path uploads right away)
            m = np.array(valid_mask, dtype=bool)
        else:
            m = np.isfinite(d)                                     # graph_construction.py:110-111
            if nodata is not None and not np.isnan(nodata):
                m &= d != np.float32(nodata)
        n = int(np.count_nonzero(m))
        if n == 0:
            logger.warning("No valid cells in grid")
            return self._empty()
        if min(d.shape) < 2:
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least (edge_order + 1) elements are required.")
        g = HostGridGraph()
        g.__dict__["_host"] = dict(depth=d, mask=m, unc=None if uncertainty is None else np.array(uncertainty, dtype=np.float32),
                                   res=(float(resolution[0]), float(resolution[1])), builder=self, n=n)
        g.grid_shape = tuple(d.shape)
        g.num_valid_cells = n
        if self.include_self_loops:
            g.__dict__["_self_loops"] = True
        return g

    def _build_now(self, depth, mask, unc, res) -> Data:
        """The eager single-grid device build (what `build_graph` does for large / device inputs)."""
        dev = self._device()
        with torch.cuda.device(dev):
            return self._build(*self._tile_cells(depth, mask, unc, res, None, dev))

    def count_tile(self, depth, valid_mask=None, uncertainty=None, resolution=(1.0, 1.0), nodata=None, stream=None, scratch=None):
        """`build_graph` in two steps for callers that pipeline tiles: `count_tile` (node / edge totals, host-synchronous on
        `stream`) now, `emit_tile(pending)` later on the current stream.  `pending["N"]` is the tile's valid-cell count."""
        dev = self._device()
        with torch.cuda.device(dev):
            return self._count(*self._tile_cells(depth, valid_mask, uncertainty, resolution, nodata, dev), stream=stream, scratch=scratch)

    def emit_tile(self, pending) -> Data:
        with torch.cuda.device(pending["dev"]):
            return self._emit(pending)

    def _tile_cells(self, depth, valid_mask, uncertainty, resolution, nodata, dev):
        """The `bgnn_cells` descriptor of one tile (argument tuple of `_build` / `_count`)."""
        if getattr(depth, "ndim", 0) != 2:
            raise ValueError("depth must be a 2-D array")
        H, W = int(depth.shape[0]), int(depth.shape[1])
        if True:
            d = _as_dev_f32(depth, dev)

            def same_view(t):       # a CUDA plane that shares the depth view's row stride is consumed in place
                return (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 2 and t.stride(1) == 1
                        and (H == 1 or t.stride(0) == d.stride(0)))

            others_ok = all(t is None or same_view(t) or d.is_contiguous() for t in (valid_mask, uncertainty))
            if d.stride(1) != 1 or not others_ok:
                d = d.contiguous()
            stride = int(d.stride(0)) if H > 1 else W
            if stride < W:
                d = d.contiguous(); stride = W
            keep = [d]
            cells = _lib.Cells()
            cells.n_grids, cells.n_cells, cells.rows, cells.cols, cells.row_stride = 1, H * W, H, W, stride
            cells.rx, cells.ry = float(resolution[0]), float(resolution[1])
            cells.depth = d.data_ptr()

            def plane(t, dtype):
                if same_view(t) and stride != W:
                    if t.dtype == dtype:
                        return t
                    if t.dtype == torch.bool and dtype == torch.uint8:
                        return t.view(torch.uint8)
                if not isinstance(t, torch.Tensor):
                    t = torch.from_numpy(np.ascontiguousarray(t))
                t = t.to(dev)
                if t.dtype == torch.bool and dtype == torch.uint8 and t.is_contiguous():
                    t = t.view(torch.uint8)
                t = t.to(dtype).contiguous()
                if stride != W:
                    raise ValueError("mask / uncertainty must be views with the same row stride as a strided depth view")
                return t

            if valid_mask is not None:
                m = plane(valid_mask, torch.uint8)
                keep.append(m)
                cells.mask = m.data_ptr()
            elif nodata is not None and not np.isnan(nodata):
                cells.use_nodata, cells.nodata = 1, float(nodata)
            if uncertainty is not None:
                u = plane(uncertainty, torch.float32)
                keep.append(u)
                cells.unc = u.data_ptr()
            return cells, keep, uncertainty is not None, (H, W), None, dev

    def build_graph_batch(self, grids: Sequence[Tuple[np.ndarray, Optional[np.ndarray], Tuple[float, float]]],
                          nodata: Optional[float] = 1.0e6, use_uncertainty: bool = True,
                          masks: Optional[Sequence[np.ndarray]] = None) -> Data:
        """Ragged batch of refinement grids as ONE block-diagonal graph: replaces the
        per-grid build_graph loop + Batch.from_data_list of
        scripts/inference_native.py:158-179, 249-269, 312.  valid = (depth != nodata)
        & isfinite(depth) (:160)."""
        dev = self._device()
        G = len(grids)
        if G == 0:
            return self._empty()
        rows = np.array([g[0].shape[0] for g in grids], np.int32)
        cols = np.array([g[0].shape[1] for g in grids], np.int32)
        off = np.zeros(G + 1, np.int64)
        np.cumsum(rows.astype(np.int64) * cols.astype(np.int64), out=off[1:])
        res = np.array([[float(g[2][0]), float(g[2][1])] for g in grids], np.float64)
        n = int(off[-1])
        has_unc = use_uncertainty and all(g[1] is not None for g in grids)
        with torch.cuda.device(dev):
            # ONE pinned staging buffer per builder (grown on demand) and ONE host->device copy per batch for the cells
            # (depth [+ uncertainty]) and one for the per-grid metadata (off | res | rows | cols): at the reference's stock
            # 50 000-node budget a flush is ~0.3 ms of kernels, so five pageable copies per flush were the larger cost.
            # The staging memory is free again when this call returns (the node count is read back synchronously).
            planes = 2 if has_unc else 1
            mbytes = (n + 7) // 8 * 8 if masks is not None else 0          # explicit validity masks (one byte per cell) behind the metadata
            stage = self._staging(planes * n * 4 + (G + 1) * 8 + 2 * G * 8 + 2 * G * 4 + 64 + mbytes)
            cells_h = stage[:planes * n * 4].view(torch.float32).numpy()
            if n:
                np.concatenate([np.asarray(g[0], np.float32).ravel() for g in grids], out=cells_h[:n])
                if has_unc:
                    np.concatenate([np.asarray(g[1], np.float32).ravel() for g in grids], out=cells_h[n:])
            m0 = (planes * n * 4 + 7) // 8 * 8
            meta_len = (G + 1) * 8 + 2 * G * 8 + 2 * G * 4
            meta_h = stage[m0:m0 + (meta_len + 7) // 8 * 8 + mbytes]
            mh = meta_h.numpy()
            if masks is not None and n:
                np.concatenate([np.asarray(mk, bool).ravel() for mk in masks], out=mh[(meta_len + 7) // 8 * 8:][:n].view(bool))
            mh[:(G + 1) * 8].view(np.int64)[:] = off
            mh[(G + 1) * 8:(G + 1) * 8 + 2 * G * 8].view(np.float64)[:] = res.ravel()
            mh[(G + 1) * 8 + 2 * G * 8:(G + 1) * 8 + 2 * G * 8 + G * 4].view(np.int32)[:] = rows
            mh[(G + 1) * 8 + 2 * G * 8 + G * 4:meta_len].view(np.int32)[:] = cols
            cells_d = torch.empty(max(planes * n, 1), dtype=torch.float32, device=dev)
            if n:
                cells_d[:planes * n].copy_(stage[:planes * n * 4].view(torch.float32), non_blocking=True)
            meta_d = torch.empty(meta_h.numel(), dtype=torch.uint8, device=dev)
            meta_d.copy_(meta_h, non_blocking=True)
            d = cells_d[:n]
            t_off = meta_d[:(G + 1) * 8].view(torch.int64)
            t_res = meta_d[(G + 1) * 8:(G + 1) * 8 + 2 * G * 8].view(torch.float64)
            t_rows = meta_d[(G + 1) * 8 + 2 * G * 8:(G + 1) * 8 + 2 * G * 8 + G * 4].view(torch.int32)
            t_cols = meta_d[(G + 1) * 8 + 2 * G * 8 + G * 4:meta_len].view(torch.int32)
            keep = [cells_d, meta_d]
            cells = _lib.Cells()
            cells.n_grids, cells.n_cells = G, n
            cells.cell_off, cells.grid_rows, cells.grid_cols, cells.grid_res = (t.data_ptr() for t in (t_off, t_rows, t_cols, t_res))
            cells.depth = d.data_ptr()
            if masks is not None:
                cells.mask = meta_d[(meta_len + 7) // 8 * 8:].data_ptr()
            elif nodata is not None and not np.isnan(nodata):
                cells.use_nodata, cells.nodata = 1, float(nodata)
            if has_unc:
                cells.unc = cells_d[n:].data_ptr()
            if G == 1:                                # degenerate: same as a single tile
                cells.n_grids, cells.rows, cells.cols, cells.row_stride = 1, int(rows[0]), int(cols[0]), int(cols[0])
                cells.rx, cells.ry = float(res[0, 0]), float(res[0, 1])
            meta = dict(cell_off=off, rows=rows, cols=cols, cell_off_dev=t_off, cols_dev=t_cols)
            return self._build(cells, keep, has_unc, None, meta if G > 1 else None, dev,
                               shapes=[(int(r), int(c)) for r, c in zip(rows, cols)])

    def build_graph_packed(self, depth: torch.Tensor, cell_off, rows, cols, res, uncertainty: Optional[torch.Tensor] = None,
                           nodata: Optional[float] = 1.0e6) -> Data:
        """Ragged batch from an ALREADY PACKED device buffer (the natural hand-over from a bulk VR-BAG reader:
        `varres_refinements` + per-grid metadata): `depth` f32 [sum rows*cols] on the device,
        `cell_off` int64 [G+1], `rows` / `cols` int32 [G], `res` f64 [G,2] (host arrays or tensors)."""
        dev = self._device()
        as_np = lambda a, dt: a.cpu().numpy().astype(dt) if isinstance(a, torch.Tensor) else np.ascontiguousarray(a, dtype=dt)
        off, rows_h, cols_h, res_h = as_np(cell_off, np.int64), as_np(rows, np.int32), as_np(cols, np.int32), as_np(res, np.float64)
        G = int(rows_h.shape[0])
        if G == 0:
            return self._empty()
        with torch.cuda.device(dev):
            d = depth.to(device=dev, dtype=torch.float32).contiguous()
            t_off, t_rows, t_cols, t_res = (torch.from_numpy(a).to(dev) for a in (off, rows_h, cols_h, res_h.reshape(G, 2)))
            keep = [d, t_off, t_rows, t_cols, t_res]
            cells = _lib.Cells()
            cells.n_grids, cells.n_cells = G, int(off[-1])
            cells.cell_off, cells.grid_rows, cells.grid_cols, cells.grid_res = (t.data_ptr() for t in (t_off, t_rows, t_cols, t_res))
            cells.depth = d.data_ptr()
            if nodata is not None and not np.isnan(nodata):
                cells.use_nodata, cells.nodata = 1, float(nodata)
            if uncertainty is not None:
                u = uncertainty.to(device=dev, dtype=torch.float32).contiguous()
                keep.append(u)
                cells.unc = u.data_ptr()
            if G == 1:
                cells.n_grids, cells.rows, cells.cols, cells.row_stride = 1, int(rows_h[0]), int(cols_h[0]), int(cols_h[0])
                cells.rx, cells.ry = float(res_h.reshape(G, 2)[0, 0]), float(res_h.reshape(G, 2)[0, 1])
            meta = dict(cell_off=off, rows=rows_h, cols=cols_h, cell_off_dev=t_off, cols_dev=t_cols)
            return self._build(cells, keep, uncertainty is not None, None, meta if G > 1 else None, dev,
                               shapes=[(int(r), int(c)) for r, c in zip(rows_h, cols_h)])

    def _empty(self) -> Data:
        """graph_construction.py:460-469 (note: no grid_shape attribute)."""
        dev = self._device()
        g = Data(x=torch.zeros((0, len(self.node_features)), dtype=torch.float32, device=dev),
                 edge_index=torch.zeros((2, 0), dtype=torch.long, device=dev),
                 edge_attr=torch.zeros((0, len(self.edge_features)), dtype=torch.float32, device=dev),
                 pos=torch.zeros((0, 2), dtype=torch.float32, device=dev))
        g.local_std = torch.zeros(0, dtype=torch.float32, device=dev)
        return g

    def _build(self, cells, keep, has_unc, grid_shape, meta, dev, shapes=None) -> Data:
        return self._emit(self._count(cells, keep, has_unc, grid_shape, meta, dev, shapes))

    def _count(self, cells, keep, has_unc, grid_shape, meta, dev, shapes=None, stream=None, scratch=None):
        """First half of a build: validity + neighbour counts + scans on `stream` (default: the current stream) and the
        host-synchronous read-back of (N, E) -- the only host round trip of the path.  The survey driver issues it for tile
        t + 1 on a side stream while tile t computes, so the GPU never waits for the host (survey.py)."""
        cfg, F = _graph_cfg(self.neighbor_offsets, self.node_features, has_unc, self.edge_features)
        n_cells = int(cells.n_cells)
        need = max(int(lib.bgnn_graph_scratch_bytes(n_cells, cfg.n_offsets)), 256)
        borrowed = scratch is not None
        if borrowed and scratch.numel() < need:
            raise ValueError(f"scratch buffer too small: {scratch.numel()} < {need} bytes")
        if scratch is None:
            scratch = torch.empty(need, dtype=torch.uint8, device=dev)
        totals = (C.c_int64 * 2)()
        st = stream_ptr() if stream is None else C.c_void_p(stream.cuda_stream)
        check(lib.bgnn_graph_count(C.byref(cells), C.byref(cfg), ptr(scratch), scratch.numel(), totals, st))
        return dict(cells=cells, keep=keep, has_unc=has_unc, grid_shape=grid_shape, meta=meta, dev=dev, shapes=shapes, cfg=cfg,
                    F=F, scratch=scratch, N=int(totals[0]), E=int(totals[1]), n_cells=n_cells, borrowed=borrowed)

    def _emit(self, pending) -> Data:
        """Second half of a build (current stream): ids, node features, edge features of a counted cell space."""
        cells, keep, grid_shape, meta, dev, shapes = (pending[k] for k in ("cells", "keep", "grid_shape", "meta", "dev", "shapes"))
        cfg, F, scratch, N, E, n_cells = (pending[k] for k in ("cfg", "F", "scratch", "N", "E", "n_cells"))
        st = stream_ptr()
        if N == 0:
            logger.warning("No valid cells in grid")
            g = self._empty()
            if shapes is not None:
                g._grid_shapes = shapes
            return g
        dims = shapes if shapes is not None else [grid_shape]
        if any(min(s) < 2 for s in dims):
            # the reference raises here from np.gradient (graph_construction.py:284)
            raise ValueError("Shape of array too small to calculate a numerical gradient, "
                             "at least (edge_order + 1) elements are required.")
        ED = cfg.n_edge_feats
        dg = _DeviceGraph()
        dg.cells, dg.cfg, dg.scratch, dg.keep, dg.device, dg.grid_meta = cells, cfg, scratch, keep, dev, meta
        dg.keep_versions = [t._version for t in keep]
        dg.scratch_borrowed = bool(pending.get("borrowed"))
        dg.N, dg.E, dg.n_cells = N, E, n_cells
        dg.node_id = torch.empty(n_cells, dtype=torch.int32, device=dev)
        dg.node_cell = torch.empty(N, dtype=torch.int32, device=dev)
        dg.x = torch.empty((N, F), dtype=torch.float32, device=dev)
        dg.local_std = torch.empty(N, dtype=torch.float32, device=dev)
        dg.rowptr = torch.empty(N + 1, dtype=torch.int32, device=dev)
        dg._csr_src = dg._csr_eattr = None
        # single tile with a radius-1 stencil: also emit the dense per-cell edge features the tiled aggregation reads
        radius1 = all(abs(cfg.off_dr[k]) <= 1 and abs(cfg.off_dc[k]) <= 1 for k in range(cfg.n_offsets))
        dg.grid_eattr = None
        dg.rag_patches, dg.n_rag_patches = None, 0
        if meta is None and int(cells.n_grids) == 1 and radius1 and cfg.n_offsets <= 8 and ED > 0:
            dg.grid_eattr = torch.empty((n_cells, cfg.n_offsets, ED), dtype=torch.float32, device=dev)
        elif meta is not None and radius1 and cfg.n_offsets in (4, 8) and ED > 0:
            # ragged batch: the same dense per-cell edge features + a table of band patches, so that the aggregation runs as
            # the tiled kernel per band of grid rows (gat_aggregate_ragged) and the CSR payload is never needed
            table = _ragged_patches(meta["cell_off"], meta["rows"], meta["cols"])
            if table is not None:
                dg.grid_eattr = torch.empty((n_cells, cfg.n_offsets, ED), dtype=torch.float32, device=dev)
                dg.rag_patches = torch.from_numpy(table.view(np.uint8)).to(dev, non_blocking=True)
                dg.n_rag_patches = int(table.shape[0])
        if dg.grid_eattr is None:           # no grid path for this graph: the forward needs the CSR payload right away
            dg._csr_src = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
            dg._csr_eattr = torch.empty((max(E, 1), max(ED, 1)), dtype=torch.float32, device=dev)
        if dg.rag_patches is not None and cfg.n_offsets == 8 and dg._csr_src is None and min(min(s_) for s_ in dims) >= 2:
            # ragged batch on the reference's 8 offsets: one CTA per band patch (shared-memory staging, 4 atans per cell)
            check(lib.bgnn_graph_emit_bands(C.byref(cells), C.byref(cfg), ptr(scratch), scratch.numel(), ptr(dg.rag_patches),
                                            dg.n_rag_patches, ptr(dg.node_id), ptr(dg.node_cell), ptr(dg.x), ptr(dg.local_std),
                                            ptr(dg.rowptr), ptr(dg.grid_eattr), st))
        else:
            check(lib.bgnn_graph_emit(C.byref(cells), C.byref(cfg), ptr(scratch), scratch.numel(), ptr(dg.node_id),
                                      ptr(dg.node_cell), ptr(dg.x), ptr(dg.local_std), ptr(dg.rowptr), ptr(dg._csr_src),
                                      ptr(dg._csr_eattr) if ED > 0 else None, ptr(dg.grid_eattr), st))
        g = GridGraph()
        g.__dict__["_dev"] = dg
        g.x = dg.x
        g.local_std = dg.local_std
        g.num_valid_cells = N
        if self.include_self_loops:
            # GATConv removes existing self loops before adding its own (PyG `remove_self_loops`), so the forward on the
            # cached device topology is unchanged; only the API-visible COO carries the extra N edges
            g.__dict__["_self_loops"] = True
        if grid_shape is not None:
            g.grid_shape = tuple(grid_shape)
        if shapes is not None:
            g._grid_shapes = shapes
        logger.debug("Built graph: %d nodes, %d edges, %d node features, %d edge features", N, E, F, ED)
        return g

    # ------------------------------------------------------------------ graph_to_grid
    def graph_to_grid(self, data: Data, node_values: torch.Tensor, fill_value: float = np.nan) -> np.ndarray:
        """graph_construction.py:471-505 (returns a host float32 array like the reference)."""
        if not hasattr(data, "grid_shape"):
            raise ValueError("Data object missing grid_shape metadata")
        if node_values.dim() != 1:
            raise ValueError("For multi-channel node values, call graph_to_grid for each channel")
        h = data.__dict__.get("_host")
        if h is not None and not node_values.is_cuda:
            # host values of a host grid asked back as a host grid (the per-grid unbatching of scripts/inference_native.py:
            # 181-204, 320-340): an index assignment through the grid's own mask, no arithmetic and no device round trip
            if int(node_values.shape[0]) != h["n"]:
                raise IndexError(f"node_values has {int(node_values.shape[0])} entries for {h['n']} nodes")
            grid = np.full(h["depth"].shape, fill_value, dtype=np.float32)
            grid[h["mask"]] = node_values.detach().to(torch.float32).numpy()
            return grid
        return self.graph_to_grid_device(data, node_values, fill_value).cpu().numpy()

    def graph_to_grid_device(self, data: Data, node_values: torch.Tensor, fill_value: float = np.nan) -> torch.Tensor:
        dev = self._device() if "_dev" not in data.__dict__ else data._dev.device
        H, W = data.grid_shape
        with torch.cuda.device(dev):
            if "_dev" in data.__dict__:
                node_id = data._dev.node_id
            else:                                    # foreign Data: rebuild the index grid (plumbing)
                node_id = torch.full((H * W,), -1, dtype=torch.int32, device=dev)
                cell = data.valid_rows.to(dev).long() * W + data.valid_cols.to(dev).long()
                node_id[cell] = torch.arange(cell.numel(), dtype=torch.int32, device=dev)
            vals = node_values.detach().to(device=dev, dtype=torch.float32).contiguous()
            grid = torch.empty((H, W), dtype=torch.float32, device=dev)
            check(lib.bgnn_graph_to_grid(ptr(node_id), H * W, ptr(vals), float(fill_value), ptr(grid), stream_ptr()))
        return grid


class MultiScaleGraphBuilder(GraphBuilder):
    """graph_construction.py:508-607: one graph per scale factor (1 = the grid itself, s > 1 = s x s block averages with a
    majority-valid mask and resolution * s).  The block reduction runs on the HOST with the reference's own NumPy calls
    (`np.nanmean` / `np.mean` over the [h/s, s, w/s, s] view -- their pairwise summation order is what makes the coarse depths
    bit-identical), each scale's graph is then built by the device kernels like any tile.  Not used by the reference's
    pipelines; kept so every public class of graph_construction.py has its drop-in."""

    def __init__(self, scales: List[int] = [1, 2, 4], **kwargs):
        super().__init__(**kwargs)
        self.scales = sorted(scales)

    def build_multiscale_graph(self, depth: np.ndarray, valid_mask: Optional[np.ndarray] = None,
                               uncertainty: Optional[np.ndarray] = None,
                               resolution: Tuple[float, float] = (1.0, 1.0)) -> Dict[int, Data]:
        depth = np.asarray(depth)
        if valid_mask is None:
            valid_mask = np.isfinite(depth)
        graphs: Dict[int, Data] = {}
        for s in self.scales:
            if s == 1:
                d, m, u, res = depth, valid_mask, uncertainty, resolution
            else:
                d, m = self._downsample(depth, s), self._downsample_mask(valid_mask, s)
                u = self._downsample(uncertainty, s) if uncertainty is not None else None
                res = (resolution[0] * s, resolution[1] * s)
            graphs[s] = self.build_graph(d, m, u, res)
            logger.debug("Scale %d: %d nodes", s, graphs[s].num_nodes)
        return graphs

    @staticmethod
    def _blocks(arr: np.ndarray, factor: int) -> np.ndarray:
        nh, nw = arr.shape[0] // factor, arr.shape[1] // factor
        return arr[:nh * factor, :nw * factor].reshape(nh, factor, nw, factor)

    def _downsample(self, arr: Optional[np.ndarray], factor: int) -> Optional[np.ndarray]:
        """:582-594 (block mean ignoring NaNs)."""
        return None if arr is None else np.nanmean(self._blocks(np.asarray(arr), factor), axis=(1, 3))

    def _downsample_mask(self, mask: np.ndarray, factor: int) -> np.ndarray:
        """:596-606 (valid when at least half of the block is)."""
        return np.mean(self._blocks(np.asarray(mask), factor), axis=(1, 3)) >= 0.5
