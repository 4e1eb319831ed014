"""Tile geometry and the overlapping-tile stitch, device-resident.

Mirrors data/tiling.py of the reference (`Tile`, `TileSpec`, `TileManager`,
`TileMerger`; :22-454).  Tile geometry is integer host logic; the accumulate /
arg-max-confidence / finalize arithmetic runs in the stitch kernels of libbgnn.so
on accumulators that live in HBM for the whole survey, so per-tile results never
leave the device (the reference round-trips every tile through host NumPy).
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass
from typing import Dict, Generator, List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr

logger = logging.getLogger(__name__)


@dataclass
class BathymetricGrid:
    """In-memory grid container (data/loaders.py:41-90).  File I/O (GDAL / HDF5) is
    outside the hot path; any loader that returns this dataclass plugs in."""
    depth: np.ndarray
    uncertainty: Optional[np.ndarray]
    nodata_value: float
    transform: Tuple[float, ...] = (0.0, 1.0, 0.0, 0.0, 0.0, -1.0)
    crs: Optional[str] = None
    resolution: Tuple[float, float] = (1.0, 1.0)
    bounds: Tuple[float, float, float, float] = (0.0, 0.0, 0.0, 0.0)
    source_path: Optional[object] = None

    @property
    def shape(self) -> Tuple[int, int]:
        return self.depth.shape

    @property
    def valid_mask(self) -> np.ndarray:
        mask = np.isfinite(self.depth)
        if self.nodata_value is not None and not np.isnan(self.nodata_value):
            mask &= (self.depth != self.nodata_value)
        return mask

    @property
    def valid_ratio(self) -> float:
        return np.sum(self.valid_mask) / self.depth.size


@dataclass
class Tile:
    data: np.ndarray
    uncertainty: Optional[np.ndarray]
    row_start: int
    col_start: int
    row_end: int
    col_end: int
    tile_row: int
    tile_col: int
    valid_mask: np.ndarray

    @property
    def shape(self) -> Tuple[int, int]:
        return self.data.shape

    @property
    def valid_ratio(self) -> float:
        return np.sum(self.valid_mask) / self.valid_mask.size


@dataclass
class TileSpec:
    row_start: int
    col_start: int
    row_end: int
    col_end: int
    tile_row: int
    tile_col: int


def _dev(t, device, dtype=torch.float32) -> torch.Tensor:
    if isinstance(t, torch.Tensor):
        return t.to(device=device, dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(t)).to(device=device, dtype=dtype)


class TileManager:
    def __init__(self, tile_size: int = 1024, overlap: int = 128, min_valid_ratio: float = 0.1):
        self.tile_size, self.overlap, self.min_valid_ratio = tile_size, overlap, min_valid_ratio
        self.stride = tile_size - overlap
        if self.stride <= 0:
            raise ValueError("Tile size must be larger than overlap")
        self._blend_cache: Dict[Tuple[int, str], torch.Tensor] = {}

    # ---- geometry (tiling.py:87-138): integer host logic
    def compute_tile_grid(self, grid_shape: Tuple[int, int]) -> Tuple[int, int, List[TileSpec]]:
        H, W = grid_shape
        span = lambda n: max(1, -(-(n - self.overlap) // self.stride))      # ceil((n - overlap) / stride), >= 1
        ntr, ntc = span(H), span(W)
        specs = []
        for tr in range(ntr):
            r1 = min(tr * self.stride + self.tile_size, H)
            r0 = tr * self.stride
            if r1 - r0 < self.tile_size and r0 > 0:
                r0 = max(0, r1 - self.tile_size)            # edge tiles slide back to stay full size
            for tc in range(ntc):
                c1 = min(tc * self.stride + self.tile_size, W)
                c0 = tc * self.stride
                if c1 - c0 < self.tile_size and c0 > 0:
                    c0 = max(0, c1 - self.tile_size)
                specs.append(TileSpec(r0, c0, r1, c1, tr, tc))
        logger.info("Grid %dx%d -> %dx%d = %d tiles", H, W, ntr, ntc, len(specs))
        return ntr, ntc, specs

    def extract_tile(self, grid: BathymetricGrid, spec: TileSpec, full_valid_mask: Optional[np.ndarray] = None) -> Tile:
        sl = (slice(spec.row_start, spec.row_end), slice(spec.col_start, spec.col_end))
        m = grid.valid_mask if full_valid_mask is None else full_valid_mask
        return Tile(data=grid.depth[sl].copy(), uncertainty=None if grid.uncertainty is None else grid.uncertainty[sl].copy(),
                    row_start=spec.row_start, col_start=spec.col_start, row_end=spec.row_end, col_end=spec.col_end,
                    tile_row=spec.tile_row, tile_col=spec.tile_col, valid_mask=m[sl].copy())

    def iterate_tiles(self, grid: BathymetricGrid, skip_empty: bool = True) -> Generator[Tile, None, None]:
        _, _, specs = self.compute_tile_grid(grid.shape)
        full = grid.valid_mask                        # computed once (the reference recomputes it per tile)
        for spec in specs:
            tile = self.extract_tile(grid, spec, full)
            if skip_empty and tile.valid_ratio < self.min_valid_ratio:
                continue
            yield tile

    def create_output_grid(self, grid_shape, dtype=np.float32, fill_value=np.nan) -> np.ndarray:
        return np.full(grid_shape, fill_value, dtype=dtype)

    # ---- blend weights (tiling.py:296-330): 1-D tables on the host in fp64 like the reference
    def _create_1d_blend(self, size: int) -> np.ndarray:
        w = np.ones(size, dtype=np.float32)
        ramp = min(self.overlap, size // 4)
        if ramp > 0:
            up = 0.5 * (1 - np.cos(np.pi * np.linspace(0, 1, ramp)))
            w[:ramp] = up
            w[-ramp:] = 0.5 * (1 - np.cos(np.pi * np.linspace(1, 0, ramp)))
        return w

    def _create_blend_weights(self, shape: Tuple[int, int]) -> np.ndarray:
        return np.outer(self._create_1d_blend(shape[0]), self._create_1d_blend(shape[1])).astype(np.float32)

    def blend_table_device(self, size: int, device) -> torch.Tensor:
        key = (size, str(device))
        if key not in self._blend_cache:
            self._blend_cache[key] = torch.from_numpy(self._create_1d_blend(size)).to(device)
        return self._blend_cache[key]

    # ---- merge_tile / finalize_output on caller arrays (tiling.py:218-294); NumPy arrays are
    #      round-tripped through the device kernels, CUDA tensors are updated in place
    def merge_tile(self, output, tile_data, spec: TileSpec, weight_grid=None):
        _lib.require_cuda()
        dev = output.device if isinstance(output, torch.Tensor) and output.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if weight_grid is None:                      # overwrite-with-blend branch (tiling.py:259-272)
            out_d, t = _dev(output, dev), _dev(tile_data, dev)
            h, w = spec.row_end - spec.row_start, spec.col_end - spec.col_start
            with torch.cuda.device(dev):
                check(lib.bgnn_stitch_blend_tile(ptr(out_d), int(out_d.shape[0]), int(out_d.shape[1]), spec.row_start,
                                                 spec.col_start, h, w, ptr(self.blend_table_device(h, dev)),
                                                 ptr(self.blend_table_device(w, dev)), ptr(t), stream_ptr()))
            if isinstance(output, np.ndarray):
                output[...] = out_d.cpu().numpy()
            elif out_d.data_ptr() != output.data_ptr():
                output.copy_(out_d)
            return
        out_d, w_d = _dev(output, dev), _dev(weight_grid, dev)
        s = _lib.Survey()
        s.rows, s.cols, s.n_cont = int(out_d.shape[0]), int(out_d.shape[1]), 1
        s.cont[0], s.wsum[0] = out_d.data_ptr(), w_d.data_ptr()
        h, w = spec.row_end - spec.row_start, spec.col_end - spec.col_start
        t = _dev(tile_data, dev)
        arr = (C.c_void_p * 1)(t.data_ptr())
        with torch.cuda.device(dev):
            check(lib.bgnn_stitch_add_tile(C.byref(s), spec.row_start, spec.col_start, h, w,
                                           ptr(self.blend_table_device(h, dev)), ptr(self.blend_table_device(w, dev)),
                                           arr, None, None, stream_ptr()))
        if isinstance(output, np.ndarray):
            output[...] = out_d.cpu().numpy()
            weight_grid[...] = w_d.cpu().numpy()
        elif out_d.data_ptr() != output.data_ptr():
            output.copy_(out_d); weight_grid.copy_(w_d)

    def finalize_output(self, output, weight_grid=None):
        if weight_grid is None:
            return output
        _lib.require_cuda()
        dev = output.device if isinstance(output, torch.Tensor) and output.is_cuda else torch.device("cuda", torch.cuda.current_device())
        out_d, w_d = _dev(output, dev), _dev(weight_grid, dev)
        s = _lib.Survey()
        s.rows, s.cols, s.n_cont = int(out_d.shape[0]), int(out_d.shape[1]), 1
        s.cont[0], s.wsum[0] = out_d.data_ptr(), w_d.data_ptr()
        with torch.cuda.device(dev):
            check(lib.bgnn_stitch_finalize(C.byref(s), stream_ptr()))
        if isinstance(output, np.ndarray):
            output[...] = out_d.cpu().numpy()
            return output
        if out_d.data_ptr() != output.data_ptr():
            output.copy_(out_d)
        return output


class TileMerger:
    """data/tiling.py:333-454 with the accumulators in HBM."""
    DISCRETE_CHANNELS = {"classification"}

    def __init__(self, tile_manager: TileManager, device=None):
        self.tile_manager = tile_manager
        self.device = device
        self.outputs: Dict[str, torch.Tensor] = {}
        self.weights: Dict[str, torch.Tensor] = {}
        self._confidence_tracker: Optional[torch.Tensor] = None
        self._shape = None

    def _dev(self):
        _lib.require_cuda()
        return torch.device(self.device) if self.device is not None else torch.device("cuda", torch.cuda.current_device())

    def initialize(self, grid_shape: Tuple[int, int], channels: List[str], dtypes: Optional[dict] = None):
        dev = self._dev()
        self._shape = (int(grid_shape[0]), int(grid_shape[1]))
        dtypes = dtypes or {}
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
        for c, v in dtypes.items():                  # tiling.py:370-373: NaN-initialised planes => floating dtypes only
            if np.dtype(v) not in tdt:
                raise ValueError(f"channel {c!r}: accumulator dtype must be float32 or float64 (NaN-initialised), got {np.dtype(v)}")
        cont = [c for c in channels if c not in self.DISCRETE_CHANNELS]
        for c in channels:
            self.outputs[c] = torch.empty(self._shape, dtype=tdt[np.dtype(dtypes.get(c, np.float32))], device=dev)
            self.weights[c] = torch.empty(self._shape, dtype=torch.float32, device=dev)
        disc = [c for c in channels if c in self.DISCRETE_CHANNELS]
        if disc:
            self._confidence_tracker = torch.empty(self._shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            for i in range(0, max(len(cont), 1), 4):
                d = disc[0] if (disc and i == 0) else None
                check(lib.bgnn_stitch_init(C.byref(self._survey(cont[i:i + 4], d)), stream_ptr()))
            for c in disc:                      # weight plane of a discrete channel is unused (stays 0)
                self.weights[c].zero_()
        logger.info("Initialized %d output channels for %s", len(channels), grid_shape)

    def _survey(self, cont: List[str], disc: Optional[str]) -> _lib.Survey:
        s = _lib.Survey()
        s.rows, s.cols, s.n_cont = self._shape[0], self._shape[1], len(cont)
        for k, c in enumerate(cont):
            s.cont[k], s.wsum[k] = self.outputs[c].data_ptr(), self.weights[c].data_ptr()
            if self.outputs[c].dtype == torch.float64:
                s.f64_mask |= 1 << k
        if disc is not None and self._confidence_tracker is not None:
            s.discrete, s.tracker = self.outputs[disc].data_ptr(), self._confidence_tracker.data_ptr()
            if self.outputs[disc].dtype == torch.float64:
                s.f64_mask |= 1 << 4
        return s

    def add_tile_nodes(self, spec: TileSpec, node_id: torch.Tensor, predicted_class: torch.Tensor, confidence: torch.Tensor,
                       correction: Optional[torch.Tensor], local_std: Optional[torch.Tensor], norm_floor: float):
        """`add_tile` with the per-tile scatter fused in (bgnn_stitch_add_tile_nodes): the model's per-node outputs go
        straight into the accumulators through the tile's node-id grid.  Needs exactly the pipeline's channel set
        (classification, confidence[, correction]); same results, bit for bit, as graph_to_grid + add_tile."""
        chans = set(self.outputs)
        if not ({"classification", "confidence"} <= chans <= {"classification", "confidence", "correction"}):
            raise ValueError(f"add_tile_nodes needs the channels classification, confidence[, correction]; have {sorted(chans)}")
        dev = self._dev()
        h, w = spec.row_end - spec.row_start, spec.col_end - spec.col_start
        tm = self.tile_manager
        cont = ["confidence"] + (["correction"] if "correction" in chans else [])
        with torch.cuda.device(dev):
            check(lib.bgnn_stitch_add_tile_nodes(C.byref(self._survey(cont, "classification")), spec.row_start, spec.col_start,
                                                 h, w, ptr(tm.blend_table_device(h, dev)), ptr(tm.blend_table_device(w, dev)),
                                                 ptr(node_id), ptr(predicted_class), ptr(confidence),
                                                 ptr(correction) if "correction" in chans else None,
                                                 ptr(local_std) if correction is not None else None, float(norm_floor),
                                                 stream_ptr()))

    def add_tile(self, spec: TileSpec, channel_data: dict):
        """Tile arrays may be NumPy (uploaded) or CUDA tensors (consumed in place).
        Must be called in the reference's tile order (fp32 accumulation order)."""
        for ch in channel_data:
            if ch not in self.outputs:
                raise ValueError(f"Unknown channel: {ch}")
        dev = self._dev()
        h, w = spec.row_end - spec.row_start, spec.col_end - spec.col_start
        conf = channel_data.get("confidence")
        tiles = {k: _dev(v, dev) for k, v in channel_data.items()}
        disc = [c for c in channel_data if c in self.DISCRETE_CHANNELS and conf is not None and self._confidence_tracker is not None]
        cont = [c for c in channel_data if c not in disc]
        tm = self.tile_manager
        with torch.cuda.device(dev):
            row_w, col_w = tm.blend_table_device(h, dev), tm.blend_table_device(w, dev)
            for i in range(0, max(len(cont), 1), 4):
                group = cont[i:i + 4]
                d = disc[0] if (disc and i == 0) else None
                if not group and d is None:
                    continue
                s = self._survey(group, d)
                arr = (C.c_void_p * 4)(*[tiles[c].data_ptr() for c in group], *([None] * (4 - len(group))))
                check(lib.bgnn_stitch_add_tile(C.byref(s), spec.row_start, spec.col_start, h, w, ptr(row_w), ptr(col_w),
                                               arr, ptr(tiles[d]) if d else None, ptr(tiles["confidence"]) if d else None,
                                               stream_ptr()))

    def finalize_device(self) -> Dict[str, torch.Tensor]:
        dev = self._dev()
        cont = [c for c in self.outputs if c not in self.DISCRETE_CHANNELS]
        with torch.cuda.device(dev):
            for i in range(0, len(cont), 4):
                check(lib.bgnn_stitch_finalize(C.byref(self._survey(cont[i:i + 4], None)), stream_ptr()))
        res = dict(self.outputs)
        self.outputs, self.weights, self._confidence_tracker = {}, {}, None
        return res

    def finalize(self) -> dict:
        return {k: v.cpu().numpy() for k, v in self.finalize_device().items()}

    def partial_state(self) -> Dict[str, torch.Tensor]:
        """Unfinalised planes of this merger (multi-GPU band gather, parallel.py)."""
        out = {}
        for c, v in self.outputs.items():
            if c in self.DISCRETE_CHANNELS:
                out["classification"] = v
            else:
                out["acc_" + c], out["wsum_" + c] = v, self.weights[c]
        if self._confidence_tracker is not None:
            out["tracker"] = self._confidence_tracker
        return out
