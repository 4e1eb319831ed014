"""Whole-survey driver (BASELINE configs[2]): host grid in -> host result planes out, tiles streamed through ONE GPU
per rank, tile-row BANDS per rank, bit-identical to the single-process reference order for any number of ranks.

What the reference does (models/pipeline.py:154-241, data/tiling.py:384-454): iterate the overlapping tiles in
row-major order, blend every tile's result grids into whole-survey host accumulators, finalize, fill unprocessed cells,
apply corrections.  Here, per rank:

  * the rank's band of survey rows goes host -> device ONCE, tile-row by tile-row on a copy stream (events gate the
    tile-rows that need them), so the first tiles start while the rest of the band is still in flight;
  * per tile: graph build -> predict -> scatter fused into the stitch accumulate (five band-resident planes:
    confidence sum, correction sum, ONE shared weight sum -- every tile cell is finite, so the reference's per-channel
    weight planes are identical --, class, confidence tracker);
  * as soon as every tile touching a range of rows is done, those rows are finalised + corrected in one pass
    (`bgnn_survey_emit`) and copied device -> host on a second copy stream, overlapped with the next tile-rows: the full
    result planes never exist on the device, "only the final grid is gathered" (SURVEY.md 8(e));
  * SEAMS between bands: the rows of a band that the previous band's tiles also touch are NOT accumulated at first --
    the tile strips covering them are parked as small tile grids; when the previous rank hands over its partial planes
    for those rows (one point-to-point message of 5 x seam_rows x W floats over NVLink), they are copied in and the
    parked strips are replayed on top in tile order.  Every cell therefore sees `((A + B) + C) + D` in exactly the
    reference's tile order: the sharded result is bit-identical to the one-GPU result, seams included.  There is no
    collective; ranks without a neighbour skip the hand-over.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, ptr, stream_ptr
from .config import CORRECTION_NORM_FLOOR
from .tiling import TileSpec

logger = logging.getLogger(__name__)

CLASS_NAN_U8 = 255          # `classification` leaves the device as uint8; 255 = NaN (never written and not valid)


# ------------------------------------------------------------------------------------------------ planning (host integers)
@dataclass
class BandPlan:
    """Rows and tiles of one rank.  Rows are survey rows; [r0, r1) = rows the rank's tiles touch."""
    rank: int
    world: int
    shape: Tuple[int, int]
    tile_rows: List[List[TileSpec]] = field(default_factory=list)   # reference order, grouped by tile_row
    r0: int = 0
    r1: int = 0
    top_end: int = 0        # rows [r0, top_end) are also touched by earlier bands: parked until their planes arrive
    bottom0: int = 0        # rows [bottom0, r1) are handed to the next band (== r1 when there is none)
    prev_rank: Optional[int] = None
    next_rank: Optional[int] = None

    @property
    def specs(self) -> List[TileSpec]:
        return [s for row in self.tile_rows for s in row]

    @property
    def own_rows(self) -> Tuple[int, int]:
        """Rows this rank emits: its band minus the seam it hands down."""
        return (self.r0, self.bottom0)

    @property
    def empty(self) -> bool:
        return not self.tile_rows


def plan_bands(specs: Sequence[TileSpec], num_tile_rows: int, world: int, shape: Tuple[int, int]) -> List[BandPlan]:
    """Contiguous, balanced bands of tile-rows (one per rank), their seams and the rows each rank owns.  The owned row
    ranges partition [0, H)."""
    from .parallel import band_partition
    H, W = int(shape[0]), int(shape[1])
    plans = []
    for rank, (lo, hi) in enumerate(band_partition(num_tile_rows, world)):
        p = BandPlan(rank=rank, world=world, shape=(H, W))
        for tr in range(lo, hi):
            row = [s for s in specs if s.tile_row == tr]
            if row:
                p.tile_rows.append(row)
        if p.tile_rows:
            p.r0 = min(s.row_start for s in p.specs)
            p.r1 = max(s.row_end for s in p.specs)
        plans.append(p)
    live = [p for p in plans if not p.empty]
    for i, p in enumerate(live):
        prev, nxt = (live[i - 1] if i > 0 else None), (live[i + 1] if i + 1 < len(live) else None)
        p.prev_rank = prev.rank if prev else None
        p.next_rank = nxt.rank if nxt else None
        p.top_end = min(max(prev.r1, p.r0), p.r1) if prev else p.r0
        p.bottom0 = max(min(nxt.r0, p.r1), p.r0) if nxt else p.r1
        if prev is not None:
            assert prev.r0 <= p.r0 and prev.r1 <= p.r1, "bands must be monotone in their row ranges"
    if live:
        assert live[0].r0 == 0 and live[-1].r1 == H, "tiles must cover the survey rows"
    return plans


# ------------------------------------------------------------------------------------------------ seam hand-over
class DistLink:
    """Point-to-point seam hand-over between neighbouring ranks over torch.distributed (NCCL: NVLink P2P; gloo on CPU)."""

    def __init__(self, dist, group=None):
        self.dist, self.group = dist, group

    def exchange(self, send: Optional[Tuple[int, List[torch.Tensor]]], recv: Optional[Tuple[int, List[torch.Tensor]]]):
        ops = []
        if send is not None:
            ops += [self.dist.P2POp(self.dist.isend, t, send[0], self.group) for t in send[1]]
        if recv is not None:
            ops += [self.dist.P2POp(self.dist.irecv, t, recv[0], self.group) for t in recv[1]]
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()


class LocalLink:
    """In-process mailbox: several bands executed one after the other on one device (tests, single-GPU debugging)."""

    def __init__(self):
        self.box: Dict[Tuple[int, int], List[torch.Tensor]] = {}
        self.rank = 0

    def exchange(self, send, recv):
        if recv is not None:
            parts = self.box.pop((recv[0], self.rank))
            for dst, src in zip(recv[1], parts):
                dst.copy_(src)
        if send is not None:
            self.box[(self.rank, send[0])] = [t.clone() for t in send[1]]


# ------------------------------------------------------------------------------------------------ band accumulators
class BandStitcher:
    """The five band-resident planes and the kernels that touch them.  Row indices are BAND-local."""
    PLANES = ("confidence", "correction", "wsum", "classification", "tracker")

    def __init__(self, tile_manager, rows: int, cols: int, device, predict_correction: bool = True):
        self.tm, self.rows, self.cols, self.device = tile_manager, int(rows), int(cols), device
        self.planes = torch.empty((5, self.rows, self.cols), dtype=torch.float32, device=device)
        self.conf, self.corr, self.wsum, self.cls, self.trk = self.planes.unbind(0)
        self.n_cont = 2
        check(lib.bgnn_stitch_init(C.byref(self._survey(0, self.rows)), stream_ptr()))
        self.parked: List[Tuple[TileSpec, int, torch.Tensor]] = []

    def _survey(self, a: int, b: int) -> _lib.Survey:
        s = _lib.Survey()
        s.rows, s.cols, s.n_cont = b - a, self.cols, self.n_cont
        off = a * self.cols * 4
        s.cont[0], s.cont[1] = self.conf.data_ptr() + off, self.corr.data_ptr() + off
        s.wsum[0] = s.wsum[1] = self.wsum.data_ptr() + off
        s.discrete, s.tracker = self.cls.data_ptr() + off, self.trk.data_ptr() + off
        return s

    def add_tile_nodes(self, spec: TileSpec, dg, out: Dict[str, torch.Tensor], park_rows: int = 0):
        """Accumulate one tile (spec in band-local rows) straight from its per-node outputs.  The first `park_rows` rows
        of the tile lie in the seam owned by the previous band: they are scattered into a small strip grid and parked."""
        h, w = spec.row_end - spec.row_start, spec.col_end - spec.col_start
        dev = self.device
        row_w, col_w = self.tm.blend_table_device(h, dev), self.tm.blend_table_device(w, dev)
        corr = out.get("correction")
        floor_ = float(CORRECTION_NORM_FLOOR)
        k = min(max(int(park_rows), 0), h)
        if k > 0:
            strip = torch.empty((3, k, w), dtype=torch.float32, device=dev)
            check(lib.bgnn_scatter_tile(ptr(dg.node_id), k * w, ptr(out["predicted_class"]), ptr(out["confidence"]), ptr(corr),
                                        ptr(dg.local_std), floor_, ptr(strip[0]), ptr(strip[1]),
                                        ptr(strip[2]) if corr is not None else None, stream_ptr()))
            if corr is None:
                strip[2].zero_()
            self.parked.append((spec, k, strip))
        if k < h:
            nid = dg.node_id.view(-1)[k * w:]
            check(lib.bgnn_stitch_add_tile_nodes(C.byref(self._survey(0, self.rows)), spec.row_start + k, spec.col_start, h - k, w,
                                                 ptr(row_w[k:]), ptr(col_w), ptr(nid), ptr(out["predicted_class"]),
                                                 ptr(out["confidence"]), ptr(corr), ptr(dg.local_std) if corr is not None else None,
                                                 floor_, stream_ptr()))

    def add_zero_tile(self, spec: TileSpec, park_rows: int = 0):
        """pipeline.py:257-264: a tile without valid cells that is not skipped contributes zeros."""
        h, w = spec.row_end - spec.row_start, spec.col_end - spec.col_start
        z = torch.zeros((3, h, w), dtype=torch.float32, device=self.device)
        k = min(max(int(park_rows), 0), h)
        if k > 0:
            self.parked.append((spec, k, z[:, :k].contiguous()))
        if k < h:
            self._add_grid(spec.row_start + k, spec.col_start, h - k, w, self.tm.blend_table_device(h, self.device)[k:],
                           self.tm.blend_table_device(w, self.device), z[:, k:].contiguous())

    def _add_grid(self, r0, c0, h, w, row_w, col_w, grids):
        arr = (C.c_void_p * 4)(grids[1].data_ptr(), grids[2].data_ptr(), None, None)
        check(lib.bgnn_stitch_add_tile(C.byref(self._survey(0, self.rows)), r0, c0, h, w, ptr(row_w), ptr(col_w), arr,
                                       ptr(grids[0]), ptr(grids[1]), stream_ptr()))

    def replay_parked(self):
        """Accumulate the parked strips in tile order (after the previous band's partial planes have been copied in)."""
        for spec, k, strip in self.parked:
            h, w = spec.row_end - spec.row_start, spec.col_end - spec.col_start
            self._add_grid(spec.row_start, spec.col_start, k, w, self.tm.blend_table_device(h, self.device),
                           self.tm.blend_table_device(w, self.device), strip)
        self.parked = []

    def seam_views(self, a: int, b: int) -> List[torch.Tensor]:
        """Contiguous row slices [a, b) of the five planes (what one band hands to the next)."""
        return [p[a:b] for p in (self.conf, self.corr, self.wsum, self.cls, self.trk)]

    def emit(self, a: int, b: int, c0: int, c1: int, depth: torch.Tensor, use_nodata: bool, nodata: float, thr: float,
             cls_u8: torch.Tensor):
        """Rows [a, b) x columns [c0, c1): finalize + fill + corrections in place; that block of `depth` becomes the cleaned depth."""
        if b <= a or c1 <= c0:
            return
        s = self._survey(a, b)
        s.cols = c1 - c0
        for k in range(2):
            s.cont[k] += c0 * 4
            s.wsum[k] += c0 * 4
        s.discrete += c0 * 4
        s.tracker += c0 * 4
        dptr = C.c_void_p(depth.data_ptr() + (a * self.cols + c0) * 4)
        uptr = C.c_void_p(cls_u8.data_ptr() + a * self.cols + c0)
        check(lib.bgnn_survey_emit(C.byref(s), dptr, 1 if use_nodata else 0, float(nodata), float(thr), dptr, uptr, self.cols, stream_ptr()))


# ------------------------------------------------------------------------------------------------ the band loop
def _copy_block(dst: torch.Tensor, src: torch.Tensor, r0: int, r1: int, c0: int, c1: int, stream):
    """dst[r0:r1, c0:c1] <- src[r0:r1, c0:c1] on `stream` (2-D planes of one dtype, one of them on the device): full-width
    blocks and device-to-device copies go through torch, column blocks between host and device through cudaMemcpy2DAsync."""
    if r1 <= r0 or c1 <= c0:
        return
    full = c0 == 0 and c1 == dst.shape[1] == src.shape[1] and dst.stride(0) == dst.shape[1] and src.stride(0) == src.shape[1]
    if full or (dst.is_cuda and src.is_cuda):
        with torch.cuda.stream(stream):
            dst[r0:r1, c0:c1].copy_(src[r0:r1, c0:c1], non_blocking=True)
        return
    esz = src.element_size()
    check(lib.bgnn_memcpy2d_async(C.c_void_p(dst.data_ptr() + (r0 * dst.stride(0) + c0) * esz), dst.stride(0) * esz,
                                  C.c_void_p(src.data_ptr() + (r0 * src.stride(0) + c0) * esz), src.stride(0) * esz,
                                  (c1 - c0) * esz, r1 - r0, 1 if dst.is_cuda else 2, C.c_void_p(stream.cuda_stream)))


def pinned_outputs(rows: int, cols: int) -> Dict[str, torch.Tensor]:
    mk = lambda dt: torch.empty((rows, cols), dtype=dt, pin_memory=bool(rows * cols))
    return dict(cleaned_depth=mk(torch.float32), classification=mk(torch.uint8), confidence=mk(torch.float32),
                correction=mk(torch.float32))


def process_band(pipe, plan: BandPlan, depth_host, uncertainty_host=None, nodata: Optional[float] = None,
                 resolution: Tuple[float, float] = (1.0, 1.0), out: Optional[Dict[str, torch.Tensor]] = None,
                 link=None) -> Dict[str, torch.Tensor]:
    """Run one rank's band.  `depth_host` / `uncertainty_host`: the band's rows [plan.r0, plan.r1) as host tensors / arrays
    (pinned for asynchronous copies) or CUDA tensors; `out`: host tensors for the OWNED rows [plan.r0, plan.bottom0)
    (allocated pinned when None).  Returns `out` (+ counters) once its copies have completed."""
    if pipe.model is None:
        raise RuntimeError("Model not loaded. Call load_model() first.")
    dev = pipe.device
    H, W = plan.shape
    own_rows = plan.bottom0 - plan.r0
    if out is None:
        out = pinned_outputs(own_rows if not plan.empty else 0, W)
    stats = dict(tiles=0, skipped=0, nodes=0)
    if plan.empty:
        out["stats"] = stats
        return out
    bh = plan.r1 - plan.r0
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    src_d = as_t(depth_host)
    src_u = None if uncertainty_host is None else as_t(uncertainty_host)
    assert tuple(src_d.shape) == (bh, W), f"band depth must be rows [{plan.r0}, {plan.r1}) x {W}"
    use_nodata = nodata is not None and not np.isnan(nodata)
    thr = float(pipe.config.inference.auto_correct_threshold)
    tm = pipe.tile_manager
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        depth = torch.empty((bh, W), dtype=torch.float32, device=dev)
        unc = None if src_u is None else torch.empty((bh, W), dtype=torch.float32, device=dev)
        cls_u8 = torch.empty((max(own_rows, 1), W), dtype=torch.uint8, device=dev)
        st = BandStitcher(tm, bh, W, dev)
        # ---- host -> device in BLOCKS (rows new to a tile-row x the columns between two tile starts), all queued now on the
        #      copy stream: the first tile starts after 16 MB have arrived, not after its whole tile-row (0.5 GB at 60 000 columns)
        col_blocks, ready, up_to = [], [], 0
        s_in.wait_stream(main)
        for row in plan.tile_rows:
            edges = [s.col_start for s in row] + [W]
            blk = [(edges[j], edges[j + 1]) for j in range(len(row))]
            need = max(s.row_end for s in row) - plan.r0
            evs = []
            for (c0, c1) in blk:
                _copy_block(depth, src_d, up_to, need, c0, c1, s_in)
                if unc is not None:
                    _copy_block(unc, src_u, up_to, need, c0, c1, s_in)
                ev = torch.cuda.Event()
                ev.record(s_in)
                evs.append(ev)
            up_to = max(up_to, need)
            col_blocks.append(blk)
            ready.append(evs)

        def ready_for(ti, spec):
            """the upload event of the last column block tile `spec` of tile-row ti reads (copies complete in issue order)"""
            blk = col_blocks[ti]
            j = max(i for i, (c0, _) in enumerate(blk) if c0 < spec.col_end)
            return ready[ti][j]

        def emit_block(a: int, b: int, c0: int, c1: int):
            """band-local rows [a, b) x columns [c0, c1) are final: finalise on the main stream, copy out on the output stream"""
            if b <= a or c1 <= c0:
                return
            st.emit(a, b, c0, c1, depth, use_nodata, nodata if use_nodata else 0.0, thr, cls_u8)
            ev = torch.cuda.Event()
            ev.record(main)
            s_out.wait_event(ev)
            _copy_block(out["cleaned_depth"], depth, a, b, c0, c1, s_out)
            _copy_block(out["classification"], cls_u8, a, b, c0, c1, s_out)
            _copy_block(out["confidence"], st.conf, a, b, c0, c1, s_out)
            _copy_block(out["correction"], st.corr, a, b, c0, c1, s_out)

        top = plan.top_end - plan.r0                      # band-local end of the parked seam
        bottom = plan.bottom0 - plan.r0
        emitted = top
        # ---- tiles, software-pipelined: the graph build's only host round trip (node / edge totals) of tile t + 1 runs on a
        #      side stream while tile t computes, so the host enqueues tile t + 1 before tile t has finished and the GPU never
        #      waits for Python (measured: 14.2 -> 13.4 ms per survey tile = the sum of the kernel times)
        s_cnt = torch.cuda.Stream()
        gb = pipe.graph_builder
        flat = [(ti, spec) for ti, row in enumerate(plan.tile_rows) for spec in row]
        n_cells_max = max((s.row_end - s.row_start) * (s.col_end - s.col_start) for _, s in flat)
        scratch = [torch.empty(max(int(lib.bgnn_graph_scratch_bytes(n_cells_max, len(gb.neighbor_offsets))), 256), dtype=torch.uint8,
                               device=dev) for _ in range(2)]
        tile_done = []

        def local(spec):
            return TileSpec(spec.row_start - plan.r0, spec.col_start, spec.row_end - plan.r0, spec.col_end, spec.tile_row, spec.tile_col)

        def count(k):
            ti, spec = flat[k]
            loc = local(spec)
            sl = (slice(loc.row_start, loc.row_end), slice(loc.col_start, loc.col_end))
            s_cnt.wait_event(ready_for(ti, spec))         # the tile's cells have arrived
            if k >= 2:
                s_cnt.wait_event(tile_done[k - 2])        # scratch[k % 2] is free again; keeps the host <= 2 tiles ahead
            return gb.count_tile(depth[sl], None, None if unc is None else unc[sl], resolution,
                                 float(nodata) if use_nodata else None, stream=s_cnt, scratch=scratch[k % 2])

        pending = count(0) if flat else None
        for k, (ti, spec) in enumerate(flat):
            loc = local(spec)
            park = max(0, min(top, loc.row_end) - loc.row_start)
            cur, n = pending, pending["N"]
            size = (loc.row_end - loc.row_start) * (loc.col_end - loc.col_start)
            main.wait_event(ready_for(ti, spec))
            if n / size < tm.min_valid_ratio:              # tiling.py:200 (skip_empty)
                stats["skipped"] += 1
            else:
                if n == 0:                                 # pipeline.py:257-264 (only when min_valid_ratio <= 0)
                    st.add_zero_tile(loc, park)
                else:
                    graph = gb.emit_tile(cur)
                    out_t = pipe.model.predict(graph, auto_correct_threshold=thr, review_threshold=pipe.config.inference.review_threshold)
                    st.add_tile_nodes(loc, graph._dev, out_t, park)
                    del graph, out_t
                stats["tiles"] += 1
                stats["nodes"] += n
            if (k + 1) % 10 == 0:                          # models/pipeline.py:184-185
                logger.info("Processed %d/%d tiles", k + 1, len(flat))
            ev = torch.cuda.Event()
            ev.record(main)
            tile_done.append(ev)
            pending = count(k + 1) if k + 1 < len(flat) else None      # host blocks here while tile k computes
            # rows that only this tile-row (and earlier ones) touch are final, column block by column block: once tile j is in,
            # no later tile writes the columns [col_start_j, col_start_j+1), so they are finalised and sent home now -- the
            # device->host stream runs one tile behind the compute instead of one tile-row behind
            nxt = plan.tile_rows[ti + 1] if ti + 1 < len(plan.tile_rows) else None
            final_to = min(min(s.row_start for s in nxt) - plan.r0, bottom) if nxt else bottom
            j = plan.tile_rows[ti].index(spec)
            emit_block(emitted, final_to, *col_blocks[ti][j])
            if k + 1 == len(flat) or flat[k + 1][0] != ti:
                emitted = max(emitted, final_to)
        # ---- seams: hand [bottom0, r1) down, take [r0, top_end) from above, replay the parked strips
        send = (plan.next_rank, st.seam_views(bottom, bh)) if plan.next_rank is not None and bh > bottom else None
        recv = (plan.prev_rank, st.seam_views(0, top)) if plan.prev_rank is not None and top > 0 else None
        if link is None and (send or recv):
            raise RuntimeError("a band with neighbours needs a seam link")
        if link is not None:
            if isinstance(link, LocalLink):
                link.rank = plan.rank
            if send and recv and top > bottom:             # thin band: its two seams overlap -> strictly in order
                link.exchange(None, recv)
                st.replay_parked()
                link.exchange(send, None)
            else:
                link.exchange(send, recv)
                if recv:
                    st.replay_parked()
        if top > 0:
            emit_block(0, min(top, bottom), 0, W)
        s_out.synchronize()
        main.synchronize()
    out["stats"] = stats
    return out


def process_survey(pipe, depth, uncertainty=None, nodata: Optional[float] = None, resolution=(1.0, 1.0), dist=None,
                   out: Optional[Dict[str, torch.Tensor]] = None):
    """Whole survey on `world` ranks (dist = torch.distributed module with an initialised group, or None for one rank).
    `depth`: the FULL host grid (each rank reads only its band's rows) -- a rank that holds just its band calls
    `process_band` directly.  Returns (plan, out) with `out` covering plan.own_rows."""
    H, W = int(depth.shape[0]), int(depth.shape[1])
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    ntr, _, specs = pipe.tile_manager.compute_tile_grid((H, W))
    plan = plan_bands(specs, ntr, world, (H, W))[rank]
    sl = slice(plan.r0, plan.r1)
    res = process_band(pipe, plan, depth[sl], None if uncertainty is None else uncertainty[sl], nodata, resolution, out,
                       DistLink(dist) if world > 1 else None)
    return plan, res


def classification_as_float(cls_u8) -> np.ndarray:
    """uint8 class plane -> the reference's float32 plane (NaN where never written and not valid)."""
    a = cls_u8.numpy() if isinstance(cls_u8, torch.Tensor) else np.asarray(cls_u8)
    f = a.astype(np.float32)
    f[a == CLASS_NAN_U8] = np.nan
    return f
