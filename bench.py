#!/usr/bin/env python3
"""Benchmark of the Bathymetric-GNN per-tile inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one tile: spatial-graph build
(count + scan + ids + emit of the lean in-edge CSR and node features) followed by
the GAT forward + heads (`predict`).  Workload at every N: BASELINE.json
configs[1] -- a 2048x2048 fp32 tile (8-connected, F=7, hidden 64, 4 heads, 4
layers) cut from a synthetic 3968^2 field; for N > 1 every rank processes its own
tile (tiles are independent: no collective on the data path; weak scaling).

The same line carries `survey` = BASELINE configs[2]: the 60 000 x 60 000 synthetic survey sharded as bands of tile-rows
over the N ranks (STRONG scaling), timed from pinned host input to pinned host output (H2D + tiles + stitch + seam hand-over +
finalize + corrections + D2H, see `survey_leg`), and at N = 1 `ragged` (configs[3], incl. the reference's own batching API)
and `wide` (configs[4], fp32 and bf16 side by side).

Prints ONE JSON line (rank 0).  `value` = depth cells/s with the tile resident in
HBM; `e2e` = the same through the public API (`GraphBuilder.build_graph` +
`BathymetricGNN.predict`) with pinned HOST buffers, H2D of depth+mask and D2H of
class / confidence / correction inside the timed region.  `roofline` is measured
live with CUDA events recorded on the launching stream around every kernel
(bgnn_profile_*), `cpu_baseline` is the CPU oracle (a port; the Python reference
cannot travel to the GPU box) timed on a bounded sample of the same tile.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TILE = 2048
FIELD_ORIGIN = (1920, 1920)        # second tile of the 2x2 tiling of a 3968^2 field (tile 2048, overlap 128)
RES = (0.5, 0.5)
METRIC = "depth cells/sec (graph build+GNN fwd)"
UNIT = "cells/s"

SLOT_NAMES = {0: "gb_count", 1: "gb_scan", 2: "gb_ids", 3: "gb_emit", 4: "gb_coo", 10: "mlp0", 11: "mlp1",
              60: "heads_gemm", 61: "heads_finish", 70: "scatter_tile", 71: "stitch_add_tile", 72: "stitch_finalize", 73: "apply_corrections",
              74: "graph_to_grid", 75: "vr_apply_results", 76: "sidecar_add", 77: "survey_emit", 78: "stitch_add_tile_nodes",
              80: "csr_aggregate", 81: "linear"}
for _l in range(16):
    SLOT_NAMES[20 + _l] = f"proj_l{_l}"
    SLOT_NAMES[40 + _l] = f"agg_l{_l}"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def make_tile(rank: int):
    from tests import synth
    r0, c0 = FIELD_ORIGIN
    depth = synth.noisy_tile(TILE, TILE, seed=1234 + rank, r0=r0, c0=c0)
    mask = synth.blob_mask(TILE, TILE, seed=1234 + rank, invalid_frac=0.02)
    depth[~mask] = 1.0e6
    return depth, mask


def make_weights(seed=0):
    from oracle import model_oracle as mo          # weights only (init with the reference's key names)
    cfg = mo.ModelCfg(in_channels=7)
    return cfg, mo.init_state_dict(cfg, seed=seed, randomize_bn=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Start streaming BEFORE the warm-up: nvidia-smi needs 0.1-1 s to print its first row, longer than a short timed
        region; rows are time-stamped on arrival and only those inside [begin(), end()] are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.end()
        time.sleep(0.1)                       # rows sampled inside the window may still be in the pipe
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t0 if self.t0 is not None else 0.0
        inside = [r for (t, r) in self.rows if t0 <= t <= self.t1 + 0.03]
        window = "timed region"
        if not inside:                        # region shorter than the sampling period: fall back to the rows under load around it
            inside = [r for (t, r) in self.rows if t0 - 0.25 <= t <= self.t1 + 0.1]
            window = "timed region +- 0.25 s (GPU under the same load: warm-up steps)"
        sm, mx, reasons = [], [], set()
        for r in inside:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window, "rows_total": len(self.rows)}


def ncu_traffic(slot_name):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernel behind `slot_name`, from the
    committed `ncu --set full` capture of this same workload (profiles/ncu_traffic.json, written by tools/ncu_summary.py);
    None when no capture is on record."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        ent = t["kernels"].get(slot_name)
        return (int(ent["dram_bytes"]), t.get("capture")) if ent else (None, None)
    except Exception:
        return None, None


def algorithmic_bytes(N, E, cells, F=7, hidden=64, heads=4, layers=4, ED=3, unc=False, fused_mlp=False, fused_heads=False):
    """SURVEY.md 8(d) per-unit figures x the units one launch processes (stated in DESIGN.md)."""
    b = {}
    K = 8                                                  # 8-connected
    # lean build of the grid path: depth + mask in, node_id grid out (per cell); x, local_std, node_cell, rowptr and the
    # dense per-cell edge features [K, ED] out (per node).  The CSR payload (csr_src / csr_eattr) is emitted lazily and
    # not at all on this path.
    b["graph_build"] = cells * (4 + 1 + 4) + N * (4 * F + 4 + 4 + 4) + N * K * ED * 4
    HC = hidden * heads
    for l in range(layers):
        last = l == layers - 1
        hc = hidden if last else HC
        H = 1 if last else heads
        fin = hidden if l == 0 else HC
        # h rows counted ONCE (neighbour re-reads must come from shared memory / L2), alpha, node ids, dense edge features
        b[f"agg_l{l}"] = N * (4 * hc + 8 * H + 4 * hc) + cells * 4 + N * K * ED * 4
        b[f"proj_l{l}"] = N * (4 * fin + 4 * hc + 8 * H)
    b["mlp0"] = N * (4 * F + 4 * hidden)
    b["mlp1"] = N * (8 * hidden)
    b["heads_gemm"] = N * (4 * hidden + 4 * 3 * (hidden // 2))
    b["heads_finish"] = N * (4 * 3 * (hidden // 2) + 4 * 3 * 2 + 8 + 4 + 4 + 8)
    if fused_mlp:           # feature extractor fused into the layer-0 projection: reads x instead of the MLP output
        b["proj_l0"] = N * (4 * F + 4 * HC + 8 * heads)
    if fused_heads:         # one kernel: reads the backbone output, writes logits, probs, class, conf, corr, action
        b["heads_gemm"] = N * (4 * hidden + 4 * 3 * 2 + 8 + 4 + 4 + 8)
    return b


def run_ours(args):
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200._lib import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = _init_dist(dev) if world > 1 else None

    depth_np, mask_np = make_tile(rank)
    cfg, sd = make_weights(0)
    model = bg.BathymetricGNN(cfg.in_channels, cfg.hidden, cfg.layers, "GAT", cfg.heads, cfg.num_classes,
                              cfg.predict_correction, 0.0, cfg.edge_dim)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    builder = bg.GraphBuilder("8-connected")

    depth_dev = torch.from_numpy(depth_np).to(dev)
    mask_dev = torch.from_numpy(mask_np).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    cells = TILE * TILE

    def step_resident():
        flush.fill_(1)                                                  # L2 flush, inside the timed region
        g = builder.build_graph(depth_dev, mask_dev, None, RES)
        out = model.predict(g)
        return g, out

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        g, out = step_resident()
    N, E = g.num_nodes, g.num_edges
    del g, out
    barrier()
    if rank == 0 and sampler.proc is not None:      # keep the GPU under the same load until nvidia-smi streams (<= 3 s, untimed)
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 3.0:
            g, out = step_resident()
            del g, out
            torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    ev0.record()
    for _ in range(args.steps):
        g, out = step_resident()
        del g, out
    ev1.record()
    barrier()
    sampler.end()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * cells / (ms_step * 1e-3)

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "ms_per_step": ms_step, "quick": True}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- end to end through the public API with pinned host buffers
    depth_pin = torch.from_numpy(depth_np).pin_memory()
    mask_pin = torch.from_numpy(mask_np).pin_memory()
    # Two sets of pinned result buffers and device input buffers: step t's H2D and step t-1's D2H run on a copy stream
    # while step t-1 / t computes (the way a caller streams tiles through the public API); every step's input upload and
    # result read-back is inside the timed region, and each step's result is consumed on the host (float(conf[0])).
    # class labels cross PCIe as uint8 (3 classes; the int64 tensor of the reference API stays on the device)
    host_out = [(torch.empty(N, dtype=torch.uint8).pin_memory(), torch.empty(N, dtype=torch.float32).pin_memory(),
                 torch.empty(N, dtype=torch.float32).pin_memory()) for _ in range(2)]
    dev_in = [(torch.empty_like(depth_dev), torch.empty_like(mask_dev)) for _ in range(2)]
    main_s, copy_s = torch.cuda.current_stream(), torch.cuda.Stream()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_cmp = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]

    def run_e2e(steps):
        keep, acc = [None, None], 0.0

        def upload(t):
            b = t % 2
            with torch.cuda.stream(copy_s):
                if t >= 2:
                    copy_s.wait_event(ev_cmp[b])               # the graph build of step t-2 has consumed this buffer
                dev_in[b][0].copy_(depth_pin, non_blocking=True)
                dev_in[b][1].copy_(mask_pin, non_blocking=True)
                ev_in[b].record(copy_s)
        upload(0)
        for t in range(steps):
            b = t % 2
            if t + 1 < steps:
                upload(t + 1)
            main_s.wait_event(ev_in[b])
            flush.fill_(1)
            g = builder.build_graph(dev_in[b][0], dev_in[b][1], None, RES)
            out = model.predict(g)
            cls_u8 = out["predicted_class"].to(torch.uint8)
            ev_cmp[b].record(main_s)
            keep[b] = (g, out, cls_u8)                          # alive until their read-back has finished
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(ev_cmp[b])
                host_out[b][0].copy_(cls_u8, non_blocking=True)
                host_out[b][1].copy_(out["confidence"], non_blocking=True)
                host_out[b][2].copy_(out["correction"], non_blocking=True)
                ev_out[b].record(copy_s)
            if t >= 1:
                ev_out[1 - b].synchronize()
                acc += float(host_out[1 - b][1][0])
                keep[1 - b] = None
        ev_out[(steps - 1) % 2].synchronize()
        acc += float(host_out[(steps - 1) % 2][1][0])
        main_s.wait_event(ev_out[(steps - 1) % 2])
        return acc

    run_e2e(2)
    barrier()
    ev0.record()
    run_e2e(args.steps)
    ev1.record()
    barrier()
    e2e_ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * cells / (e2e_ms / args.steps * 1e-3)
    h2d = depth_np.nbytes + mask_np.nbytes
    d2h = N * (1 + 4 + 4)

    # ---- per-kernel timing for the roofline (separate pass, same step, events inside the library)
    roof, kernels, launches_per_step = None, {}, 0
    if rank == 0:
        lib.bgnn_profile_enable(1)
        reps = min(args.steps, 5)
        for _ in range(reps):
            g, out = step_resident()
            del g, out
        torch.cuda.synchronize()
        n = lib.bgnn_profile_collect(0, None, None)
        slots = (C.c_int32 * n)()
        ms = (C.c_float * n)()
        lib.bgnn_profile_collect(n, slots, ms)
        lib.bgnn_profile_enable(0)
        launches_per_step = n // reps
        agg = {}
        for s_, m_ in zip(slots, ms):
            agg.setdefault(SLOT_NAMES.get(int(s_), str(int(s_))), []).append(float(m_))
        bytes_ = algorithmic_bytes(N, E, cells, fused_mlp="mlp0" not in agg, fused_heads="heads_finish" not in agg)
        peak, peak_src = peaks()
        gb_ms = sum(np.mean(agg[k]) for k in ("gb_count", "gb_scan", "gb_ids", "gb_emit") if k in agg)
        tot_ms = sum(np.mean(v) for v in agg.values())
        for k, v in agg.items():
            mean_ms = float(np.mean(v))
            ent = {"ms": round(mean_ms, 4), "share": round(mean_ms / tot_ms, 4)}
            if k in bytes_:
                gbs = bytes_[k] / (mean_ms * 1e-3) / 1e9
                ent.update({"alg_bytes": int(bytes_[k]), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)})
            kernels[k] = ent
        kernels["graph_build_total"] = {"ms": round(gb_ms, 4), "share": round(gb_ms / tot_ms, 4),
                                        "alg_bytes": int(bytes_["graph_build"]),
                                        "gbs": round(bytes_["graph_build"] / (gb_ms * 1e-3) / 1e9, 1),
                                        "frac": round(bytes_["graph_build"] / (gb_ms * 1e-3) / 1e9 / peak, 4)}
        cand = [k for k in kernels if k.startswith(("agg_", "proj_"))]
        dom = max(cand, key=lambda k: kernels[k]["ms"])
        traffic, capture = ncu_traffic(dom)
        roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac"], "traffic": traffic, "traffic_source": capture, "peak_source": peak_src,
                "alg_bytes_per_launch": kernels[dom]["alg_bytes"], "ms_per_launch": kernels[dom]["ms"],
                "message_passing": {k: kernels[k] for k in sorted(kernels) if k.startswith("agg_")}}

    # ---- CPU baseline (oracle port) on a bounded sample of the same tile, rank 0 / N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_oracle_sample(depth_np, mask_np, cfg, sd, crop=args.cpu_crop)

    # ---- BASELINE configs[2]: the 60 000^2 survey, strong scaling over the ranks (north_star's scaling claim)
    survey = None
    if not args.no_survey:
        del depth_dev, mask_dev, flush, dev_in, host_out
        torch.cuda.empty_cache()
        try:
            survey = survey_leg(args, dev, rank, world, dist, model)
        except Exception as e:                          # e.g. not enough pinned host memory on this box: say so, keep the line
            survey = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
            if dist is not None:                        # a rank that failed alone would leave the others in a barrier
                raise

    # ---- BASELINE configs[3] (ragged VR grids, incl. the reference's own batching API) and configs[4] (wide variant, fp32 and
    #      bf16 side by side): secondary objects of the same line, one GPU only (both are per-GPU workloads without exchange)
    extra = {}
    if world == 1 and not args.no_extra:
        torch.cuda.empty_cache()
        import copy
        for name, fn, over in (("ragged", run_ragged, dict(grids=100000, api_grids=10000)), ("wide", run_wide, dict(steps=5, warmup=2))):
            a2 = copy.copy(args)
            for k_, v_ in over.items():
                setattr(a2, k_, v_)
            try:
                extra[name] = fn(a2, embed=True)
            except Exception as e:                      # secondary objects never take the headline line down
                extra[name] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.empty_cache()

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "configs[1]: one 2048x2048 fp32 tile per GPU, 8-connected, F=7, GAT hidden 64 x 4 heads x 4 layers, "
                                       "graph build + predict", "tile": TILE, "valid_cells": N, "edges": E,
                           "cells_per_step_per_gpu": cells, "weights": "random-init (reference key names), BN stats randomised",
                           "l2": "256 MiB flush write before every step (inside the timed region)",
                           "parallelism": f"tiles x{world} (independent, no collective)"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
                "roofline": roof, "kernels": kernels, "cpu_baseline": cpu, "clocks": clocks, "survey": survey, "ragged": extra.get("ragged"),
                "wide": extra.get("wide")}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def _init_dist(dev):
    import torch.distributed as dist
    if dist.is_initialized():
        return dist
    # NCCL may print its version banner on stdout at communicator creation: keep stdout for the ONE JSON line
    sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
    finally:
        sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    return dist


def survey_leg(args, dev, rank, world, dist, model, S=None):
    """BASELINE.json configs[2]: the S x S synthetic survey (tests/synth.py: a counter-based function of (seed, row, col),
    ~5 % invalid 'coast'), tile 2048 / overlap 128, tile-row bands per rank (strong scaling), through
    `bathymetric_gnn_b200.survey.process_band`.  TIMED REGION (device events on the main stream, barrier on both sides,
    max over ranks): host->device copy of the rank's band from PINNED HOST memory + every tile (graph build, predict,
    scatter+stitch) + seam hand-over to the neighbour rank (point-to-point, NCCL over NVLink) + replay of the parked seam
    strips + finalize + unprocessed-cell fill + apply_corrections + device->host copy of the four final planes
    (cleaned depth f32, class u8, confidence f32, correction f32) into pinned host buffers."""
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200 import survey as sv
    from tests import synth
    S = S or args.survey_size
    cfg = bg.Config()
    cfg.tile.tile_size, cfg.tile.overlap, cfg.tile.min_valid_ratio = args.tile, args.overlap, 0.1
    pipe = bg.BathymetricPipeline(cfg)
    pipe.model = model
    link = sv.DistLink(dist) if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def make_band(size_rows, size_cols, plan, tabs_dev):
        """the band's rows generated on the device chunk by chunk, parked in pinned host memory (untimed set-up)"""
        host = torch.empty((plan.r1 - plan.r0, size_cols), dtype=torch.float32, pin_memory=True)
        for a in range(plan.r0, plan.r1, 2048):
            b_ = min(a + 2048, plan.r1)
            blk = synth.survey_block_torch(size_rows, a, b_, 0, size_cols, dev, tables=tabs_dev)
            host[a - plan.r0:b_ - plan.r0].copy_(blk)
        torch.cuda.synchronize()
        return host

    # ---- warm-up: one 2048^2 tile per rank with a seam between every pair of neighbours (module load, allocator, NCCL p2p)
    Hw, Ww = (args.tile - args.overlap) * world + args.overlap, args.tile
    tabs_w = {k: torch.from_numpy(v).to(dev) for k, v in synth.survey_tables(max(Hw, Ww)).items()}
    ntr_w, _, specs_w = pipe.tile_manager.compute_tile_grid((Hw, Ww))
    plan_w = sv.plan_bands(specs_w, ntr_w, world, (Hw, Ww))[rank]
    band_w = make_band(max(Hw, Ww), Ww, plan_w, tabs_w)
    for _ in range(2):
        sv.process_band(pipe, plan_w, band_w, None, synth.SURVEY_NODATA, RES, None, link)
        barrier()
    del band_w, tabs_w
    # ---- the survey
    ntr, ntc, specs = pipe.tile_manager.compute_tile_grid((S, S))
    plans = sv.plan_bands(specs, ntr, world, (S, S))
    plan = plans[rank]
    tabs = {k: torch.from_numpy(v).to(dev) for k, v in synth.survey_tables(S).items()}
    t_setup = time.perf_counter()
    band = make_band(S, S, plan, tabs)
    out = sv.pinned_outputs(plan.bottom0 - plan.r0, S)
    del tabs
    torch.cuda.empty_cache()
    t_setup = time.perf_counter() - t_setup
    barrier()
    diag = os.environ.get("BGNN_SURVEY_PROFILE") == "1" and rank == 0       # dev aid: device time per slot inside the timed pass
    if diag:
        from bathymetric_gnn_b200._lib import lib as _lib_
        _lib_.bgnn_profile_enable(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    res = sv.process_band(pipe, plan, band, None, synth.SURVEY_NODATA, RES, out, link)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    slot_ms = None
    if diag:
        n_ = _lib_.bgnn_profile_collect(0, None, None)
        sl_ = (C.c_int32 * n_)(); tm_ = (C.c_float * n_)()
        _lib_.bgnn_profile_collect(n_, sl_, tm_)
        _lib_.bgnn_profile_enable(0)
        slot_ms = {}
        for a_, b_ in zip(sl_, tm_):
            k_ = SLOT_NAMES.get(int(a_), str(int(a_)))
            slot_ms[k_] = slot_ms.get(k_, 0.0) + float(b_)
        nt_ = max(res["stats"]["tiles"], 1)
        slot_ms = {k_: round(v_ / nt_, 3) for k_, v_ in slot_ms.items()}
        slot_ms["device_sum_per_tile"] = round(sum(slot_ms.values()), 3)
        slot_ms["wall_per_tile_rank0"] = round(ms / nt_, 3)
    per_rank = [ms]
    if dist is not None:
        t = torch.zeros(world, device=dev, dtype=torch.float64)
        t[rank] = ms
        dist.all_reduce(t)
        per_rank = [float(v) for v in t.tolist()]
        ms = max(per_rank)
        cnt = torch.tensor([res["stats"]["tiles"], res["stats"]["skipped"], res["stats"]["nodes"]], device=dev, dtype=torch.float64)
        dist.all_reduce(cnt)
        tiles, skipped, nodes = (int(v) for v in cnt.tolist())
    else:
        tiles, skipped, nodes = res["stats"]["tiles"], res["stats"]["skipped"], res["stats"]["nodes"]
    # spot check on the host result (not timed): classes are 0/1/2/255 and corrected cells differ from the input
    cls = res["classification"]
    sample = cls[:: max(1, cls.shape[0] // 64), :: max(1, cls.shape[1] // 64)]
    assert bool(((sample <= 2) | (sample == 255)).all())
    # order-independent bit-pattern checksums of the host planes (untimed): equal across rank counts <=> bit-identical results
    sums = []
    for k in ("cleaned_depth", "classification", "confidence", "correction"):
        t_ = res[k]
        sums.append(int((t_.view(torch.int32) if t_.dtype == torch.float32 else t_).sum(dtype=torch.int64)) if t_.numel() else 0)
    if dist is not None:
        cs = torch.tensor(sums, device=dev, dtype=torch.int64)
        dist.all_reduce(cs)
        sums = [int(v) for v in cs.tolist()]
    seam_rows = sum(p.r1 - p.bottom0 for p in plans if not p.empty)
    return {"metric": "depth cells/sec (graph build+GNN fwd+stitch), whole survey, host in -> host out", "value": S * S / (ms * 1e-3),
            "unit": UNIT, "scaling": "strong", "n_gpus": world, "seconds": ms * 1e-3,
            "workload": f"configs[2]: {S}x{S} synthetic survey @0.5 m, tile {args.tile} / overlap {args.overlap}, {len(specs)} tiles "
                        f"({ntr} tile-rows as contiguous bands over {world} rank(s)), min_valid_ratio 0.1",
            "tiles_processed": tiles, "tiles_skipped": skipped, "valid_tile_cells": nodes,
            "tile_cells_per_s": tiles * args.tile * args.tile / (ms * 1e-3),
            "timed_region": "H2D of the band (pinned host) + tiles + stitch + seam hand-over (p2p) + finalize + fill + "
                            "apply_corrections + D2H of cleaned/class(u8)/confidence/correction to pinned host",
            "h2d_bytes": S * S * 4 + seam_rows * S * 4, "d2h_bytes": S * S * 13, "seam_p2p_bytes": seam_rows * S * 20,
            "checksum": dict(zip(("cleaned_depth", "classification", "confidence", "correction"), sums)),
            "per_rank_seconds": [round(v * 1e-3, 4) for v in per_rank], "host_cores": os.cpu_count(), "rank0_ms_per_tile_by_slot": slot_ms, "setup_seconds_rank0_untimed": round(t_setup, 2),
            "bit_identical_to_one_rank": "by construction (parked seam strips replayed in reference tile order; "
                                         "tests/test_gpu_survey.py, tests/test_host_logic.py::test_exact_seam_protocol_gloo)"}


def run_survey(args):
    """`--workload survey`: the configs[2] leg alone (same object the default line carries under "survey")."""
    import bathymetric_gnn_b200 as bg
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = _init_dist(dev) if world > 1 else None
    cfg_m, sd = make_weights(0)
    model = bg.BathymetricGNN(cfg_m.in_channels, cfg_m.hidden, cfg_m.layers, "GAT", cfg_m.heads, cfg_m.num_classes,
                              cfg_m.predict_correction, 0.0, cfg_m.edge_dim)
    model.load_state_dict(sd, strict=True)
    line = survey_leg(args, dev, rank, world, dist, model.to(dev).eval())
    if rank == 0:
        line.update({"higher_is_better": True, "dtype": "f32", "data": "synthetic"})
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_ragged(args, embed=False):
    """BASELINE.json configs[3]: native VR BAG refinement grids (sides 3..50, 10 % nodata) batched as ONE ragged
    block-diagonal graph per chunk of <= ~8 M cells; grids are sharded contiguously across ranks (no collective).
    Secondary workload; `value` = valid cells / slowest rank's device time (packed buffers resident in HBM)."""
    import bathymetric_gnn_b200 as bg
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    rng = np.random.default_rng(7 + rank)
    G = args.grids // world
    rows = rng.integers(3, 51, G).astype(np.int32); cols = rng.integers(3, 51, G).astype(np.int32)
    res = rng.choice([0.5, 1.0, 2.0, 4.0], G)
    res2 = np.stack([res, res], 1)
    size = rows.astype(np.int64) * cols
    cfg_m, sd = make_weights(0)
    model = bg.BathymetricGNN(cfg_m.in_channels, cfg_m.hidden, cfg_m.layers, "GAT", cfg_m.heads, cfg_m.num_classes,
                              cfg_m.predict_correction, 0.0, cfg_m.edge_dim)
    model.load_state_dict(sd, strict=True); model = model.to(dev).eval()
    builder = bg.GraphBuilder("8-connected")
    # chunks of consecutive grids with <= args.chunk_cells cells
    bounds, acc, start = [], 0, 0
    for g in range(G):
        if acc + size[g] > args.chunk_cells and g > start:
            bounds.append((start, g)); start, acc = g, 0
        acc += size[g]
    bounds.append((start, G))
    gen = torch.Generator(device=dev); gen.manual_seed(7 + rank)
    chunks = []
    for (a, b) in bounds:
        off = np.zeros(b - a + 1, np.int64); np.cumsum(size[a:b], out=off[1:])
        n = int(off[-1])
        d = 30.0 + 0.5 * torch.randn(n, device=dev, generator=gen)
        d[torch.rand(n, device=dev, generator=gen) < 0.10] = 1.0e6
        chunks.append((d, off, rows[a:b], cols[a:b], res2[a:b]))

    def one_pass():
        nodes = 0
        for (d, off, r, c, rs) in chunks:
            g = builder.build_graph_packed(d, off, r, c, rs, nodata=1.0e6)
            out = model.predict(g)
            nodes += g.num_nodes
            del g, out
        return nodes
    one_pass()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); nodes = one_pass(); ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    kernels, post = {}, {}
    if rank == 0:                # per-kernel device time of ONE chunk (events inside the library)
        from bathymetric_gnn_b200._lib import lib
        lib.bgnn_profile_enable(1)
        d, off, r, c, rs = chunks[0]
        g = builder.build_graph_packed(d, off, r, c, rs, nodata=1.0e6); out = model.predict(g)
        torch.cuda.synchronize()
        n = lib.bgnn_profile_collect(0, None, None)
        slots = (C.c_int32 * n)(); tms = (C.c_float * n)()
        lib.bgnn_profile_collect(n, slots, tms)
        lib.bgnn_profile_enable(0)
        for s_, m_ in zip(slots, tms):
            k = SLOT_NAMES.get(int(s_), str(int(s_)))
            kernels[k] = round(kernels.get(k, 0.0) + float(m_), 4)
        kernels["chunk_nodes"] = g.num_nodes
        post = vr_post_row(g, out, chunks[0], dev)
        del g, out
    api = {}
    if rank == 0 and not args.no_api:
        api = native_api_leg(args, model, builder, dev)
    if rank == 0:
        line = ({"metric": "depth cells/sec (graph build+GNN fwd), ragged VR grids", "value": world * nodes / (ms * 1e-3),
                          "unit": UNIT, "n_gpus": world, "ms_total": ms, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": f"configs[3]: {args.grids} refinement grids, sides U{{3..50}}, 10% nodata, "
                                                 f"{len(chunks)} ragged chunks of <= {args.chunk_cells} cells per rank",
                                     "valid_cells_rank0": nodes, "note": "rank-0 time (ranks are independent; no collective)"},
                          "kernels_ms_one_chunk": kernels, "vr_post_one_chunk": post, "reference_api": api})
        if embed:
            return line
        print(json.dumps(line))


def native_api_leg(args, model, builder, dev):
    """The REFERENCE's own batching API on the same kind of input (scripts/inference_native.py:249-342 as driven by its main
    loop, :505-560): host NumPy grids -> `NativeVRProcessor.add_to_batch` / `batch_ready` / `flush_batch` -> per-grid host
    result arrays, wall clock with a device synchronise on both sides.  Two budgets: the reference's stock
    BATCH_NODE_BUDGET = 50 000 nodes per flush and this build's default."""
    import bathymetric_gnn_b200 as bg
    from tests import synth
    G = args.api_grids
    grids = synth.ragged_grids(G, seed=7)
    valid = sum(int(((d != 1.0e6) & np.isfinite(d)).sum()) for d, _, _ in grids)
    out = {"grids": G, "valid_cells": valid, "note": "host arrays in, per-grid host arrays out; uncertainty dropped for the 7-feature model "
                                                   "(inference_native.py:165-167)"}
    for label, budget in (("stock_budget_50000", 50000), ("default_budget", None)):
        proc = bg.NativeVRProcessor(model, builder, dev)
        if budget is not None:
            proc.BATCH_NODE_BUDGET = budget
        for rep in range(2):                       # first pass = warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            flushes, acc = 0, 0.0
            for d, u, res in grids:
                r = proc.add_to_batch(d, u, res, 1.0e6)
                if proc.batch_ready:
                    for cls, conf, corr in proc.flush_batch():
                        acc += float(conf.flat[0])
                    flushes += 1
            if proc.batch_pending:
                for cls, conf, corr in proc.flush_batch():
                    acc += float(conf.flat[0])
                flushes += 1
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out[label] = {"budget_nodes": int(proc.BATCH_NODE_BUDGET), "seconds": round(dt, 4), "flushes": flushes,
                      "cells_per_s": valid / dt, "ms_per_flush": round(dt / max(flushes, 1) * 1e3, 3)}
    return out


def vr_post_row(g, out, chunk, dev):
    """SURVEY.md 8(f) rank 2 on one ragged chunk: results scattered into the packed cell space, then `apply_results`
    (scripts/inference_native.py:480-503) and the sidecar rasterisation (data/vr_bag.py:695-778) as device kernels, timed by
    the library's events; beside it the CPU oracle port of the same two steps on the first 300 grids (the reference's own
    sidecar loop is a 4-deep Python loop, far slower than this vectorised port)."""
    import types
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200._lib import lib, check, ptr, stream_ptr
    from bathymetric_gnn_b200.config import CORRECTION_NORM_FLOOR
    d, off, rows, cols, res2 = chunk
    dg = g._dev
    n = int(off[-1]); G = len(rows)
    planes = torch.zeros((3, n), dtype=torch.float32, device=dev)
    check(lib.bgnn_scatter_tile(ptr(dg.node_id), n, ptr(out["predicted_class"]), ptr(out["confidence"]), ptr(out["correction"]),
                                ptr(dg.local_std), float(CORRECTION_NORM_FLOOR), ptr(planes[0]), ptr(planes[1]), ptr(planes[2]),
                                stream_ptr()))
    unc = torch.full((n,), 0.1, dtype=torch.float32, device=dev)
    side = int(np.ceil(np.sqrt(G)))                       # one base cell of 50 x 50 raster pixels per grid, raster at 0.5 m
    geo = dict(shape=(side * 50, side * 50), geotransform=(0.0, 0.5, 0.0, side * 25.0, 0.0, -0.5), bounds=(0.0, 0.0, side * 25.0, side * 25.0),
               base_shape=(side, side))
    grids = [types.SimpleNamespace(base_row=i // side, base_col=i % side, sw_corner=(0.0, 0.0), dimensions=(int(rows[i]), int(cols[i])),
                                   resolution=(0.5, 0.5)) for i in range(G)]
    sb = bg.SidecarBuilder(**geo, device=dev)
    torch.cuda.synchronize()
    lib.bgnn_profile_enable(1)
    od, ou, st = bg.apply_results_packed(d, unc, planes[0], planes[1], planes[2], 0.85)
    sb.add_batch(grids, off, d, planes[0], planes[1], planes[2])
    torch.cuda.synchronize()
    k = lib.bgnn_profile_collect(0, None, None)
    slots = (C.c_int32 * k)(); tms = (C.c_float * k)()
    lib.bgnn_profile_collect(k, slots, tms)
    lib.bgnn_profile_enable(0)
    ms = {SLOT_NAMES.get(int(s_), str(int(s_))): round(float(m_), 4) for s_, m_ in zip(slots, tms)}
    # CPU port on a bounded sample
    from oracle import vr_post_oracle as vo
    ns = min(300, G)
    hp = planes[:, :int(off[ns])].cpu().numpy(); hd = d[:int(off[ns])].cpu().numpy(); hu = unc[:int(off[ns])].cpu().numpy()
    so = vo.SidecarOracle(**geo)
    t0 = time.perf_counter()
    for i in range(ns):
        sl = slice(int(off[i]), int(off[i + 1])); shp = (int(rows[i]), int(cols[i]))
        gd = hd[sl].reshape(shp)
        vo.apply_results(gd, hu[sl].reshape(shp), hp[0, sl].reshape(shp), hp[1, sl].reshape(shp), hp[2, sl].reshape(shp), 0.85)
        grids[i].depth = gd
        grids[i].valid_mask = vo.valid_mask(gd)
        so.add_refinement_results(grids[i], hp[0, sl].reshape(shp), hp[1, sl].reshape(shp), hp[2, sl].reshape(shp))
    cpu_s = time.perf_counter() - t0
    gpu_ms = sum(ms.values())
    return {"cells": n, "grids": G, "device_ms": ms, "cells_per_s_device": n / (gpu_ms * 1e-3) if gpu_ms > 0 else None,
            "alg_bytes": n * 28 + n * 48, "cells_corrected": st["cells_corrected"],
            "cpu_port": {"grids": ns, "cells": int(off[ns]), "seconds": round(cpu_s, 4), "cells_per_s": int(off[ns]) / cpu_s, "cores": 1}}


def run_wide(args, embed=False):
    """BASELINE.json configs[4]: wide variant -- 5x5 neighbourhood (k = 24), hidden 128 x 4 heads x 4 attention layers, one
    2048^2 tile per GPU, fp32 parity mode AND bf16 mode (bgnn_model_set_precision).  Secondary workload: one JSON line with
    both precisions, the per-kernel device times of one step each, and the bf16-vs-fp32 output differences on this tile."""
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200._lib import lib
    from oracle import model_oracle as mo          # weights only
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    T = args.tile
    from tests import synth
    depth = synth.noisy_tile(T, T, seed=1234 + rank, r0=FIELD_ORIGIN[0], c0=FIELD_ORIGIN[1])
    mask = synth.blob_mask(T, T, seed=1234 + rank, invalid_frac=0.02)
    depth[~mask] = 1.0e6
    cfg_m = mo.ModelCfg(in_channels=7, hidden=128)
    sd = mo.init_state_dict(cfg_m, seed=0, randomize_bn=True)
    model = bg.BathymetricGNN(7, 128, 4, "GAT", 4, 3, True, 0.0, 3)
    model.load_state_dict(sd, strict=True); model = model.to(dev).eval()
    builder = bg.GraphBuilder("24-connected")
    d_dev = torch.from_numpy(depth).to(dev); m_dev = torch.from_numpy(mask).to(dev)
    res, outs = {}, {}
    for prec in ("fp32", "bf16"):
        model.set_precision(prec)

        def step():
            g = builder.build_graph(d_dev, m_dev, None, RES)
            o = model.predict(g)
            return g, o
        for _ in range(max(args.warmup, 1)):
            g, o = step(); del g, o
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            g, o = step(); del g, o
        ev1.record(); torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / args.steps
        lib.bgnn_profile_enable(1)
        g, o = step(); torch.cuda.synchronize()
        n = lib.bgnn_profile_collect(0, None, None)
        slots = (C.c_int32 * n)(); tms = (C.c_float * n)()
        lib.bgnn_profile_collect(n, slots, tms)
        lib.bgnn_profile_enable(0)
        kern = {}
        for s_, m_ in zip(slots, tms):
            k = SLOT_NAMES.get(int(s_), str(int(s_)))
            kern[k] = round(kern.get(k, 0.0) + float(m_), 4)
        N, E = g.num_nodes, g.num_edges
        outs[prec] = {k: o[k].float().cpu() for k in ("class_logits", "confidence", "correction", "predicted_class")}
        # algorithmic bytes per launch: projections N*(in + out + 8H), aggregations N*(h + out + 8H) + E*12 + 4*cells + 4N,
        # element size 2 (bf16 storage) or 4; the last layer's output stays fp32
        peak, _ = peaks()
        esz = 2 if prec == "bf16" else 4
        roof = {}
        for l in range(4):
            last = l == 3
            H_l, HC = (1 if last else 4), (128 if last else 512)
            in_l = 128 if l == 0 else 512
            b_proj = N * (esz * in_l + esz * HC + 8 * H_l)
            b_agg = N * (esz * HC + (4 if last else esz) * HC + 8 * H_l + 4) + E * 12 + 4 * T * T
            for name, b in ((f"proj_l{l}", b_proj), (f"agg_l{l}", b_agg)):
                if name in kern and kern[name] > 0:
                    gbs = b / (kern[name] * 1e-3) / 1e9
                    roof[name] = {"alg_bytes": int(b), "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        res[prec] = {"ms_per_step": round(ms, 3), "cells_per_s": world * T * T / (ms * 1e-3), "kernels_ms": kern, "roofline": roof}
        del g, o
    t = torch.tensor([res["fp32"]["ms_per_step"], res["bf16"]["ms_per_step"]], device=dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.destroy_process_group()
    diff = {k: float((outs["bf16"][k] - outs["fp32"][k]).abs().max()) for k in ("class_logits", "confidence", "correction")}
    diff["label_disagreement_frac"] = float((outs["bf16"]["predicted_class"] != outs["fp32"]["predicted_class"]).float().mean())
    diff["logit_scale"] = float(outs["fp32"]["class_logits"].abs().max())
    if rank == 0:
        line = ({"metric": METRIC + ", wide variant", "value": world * T * T / (float(t[1]) * 1e-3), "unit": UNIT,
                          "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(t[1]),
                          "higher_is_better": True, "scaling": "weak", "dtype": "bf16 storage / f32 accumulate", "data": "synthetic",
                          "config": {"workload": f"configs[4]: one {T}x{T} tile per GPU, 24-connected (5x5), F=7, GAT hidden 128 x 4 heads "
                                                 "x 4 layers, graph build + predict; value = bf16 mode, fp32 mode beside it",
                                     "nodes": N, "edges": E, "l2": "activations (GBs per layer) far exceed the 126 MB L2"},
                          "fp32": {"ms_per_step": float(t[0]), "cells_per_s": world * T * T / (float(t[0]) * 1e-3), "kernels_ms": res["fp32"]["kernels_ms"],
                                   "roofline": res["fp32"]["roofline"]},
                          "bf16": {"ms_per_step": float(t[1]), "kernels_ms": res["bf16"]["kernels_ms"], "roofline": res["bf16"]["roofline"]},
                          "bf16_vs_fp32_max_abs_diff": diff})
        if embed:
            return line
        print(json.dumps(line))


def cpu_oracle_sample(depth_np, mask_np, cfg, sd, crop=512, threads=None):
    """Oracle (numpy/scipy graph build + torch-CPU GAT forward) on a crop x crop corner of the tile."""
    from oracle import graph_oracle as go
    from oracle import model_oracle as mo
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    d = np.ascontiguousarray(depth_np[:crop, :crop])
    m = np.ascontiguousarray(mask_np[:crop, :crop])
    t0 = time.perf_counter()
    g = go.build_graph(d, m, None, RES)
    t1 = time.perf_counter()
    mo.predict(sd, cfg, g["x"], g["edge_index"], g["edge_attr"])
    t2 = time.perf_counter()
    return {"value": crop * crop / (t2 - t0), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{crop}x{crop} corner of the bench tile, one pass; build {t1 - t0:.2f}s (numpy/scipy, 1 thread, "
                      f"per-edge Python loop of the reference vectorised) + forward {t2 - t1:.2f}s (torch CPU fp32, {threads} threads)",
            "seconds": t2 - t0}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  The Python reference cannot
    travel to the GPU box (/root/reference is absent there), so this times the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    depth_np, mask_np = make_tile(0)
    cfg, sd = make_weights(0)
    crop = args.cpu_crop_ref
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_sample(depth_np, mask_np, cfg, sd, crop)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_oracle_sample(depth_np, mask_np, cfg, sd, crop)
    dt = time.perf_counter() - t0
    value = args.steps * crop * crop / dt
    last["value"] = value
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"configs[1] tile, bounded sample: {crop}x{crop} corner per step (CPU oracle port)"},
                      "cpu_baseline": last,
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="timed loop only (used under ncu): no e2e / per-kernel / cpu legs")
    ap.add_argument("--workload", default="tile", choices=["tile", "survey", "ragged", "wide"],
                    help="tile = BASELINE configs[1] (default, the headline line); survey = configs[2], ragged = configs[3], "
                         "wide = configs[4] fp32 + bf16 (secondary)")
    ap.add_argument("--grids", type=int, default=100000)
    ap.add_argument("--api-grids", type=int, default=20000, help="ragged workload: grids pushed through the reference's add_to_batch / flush_batch API")
    ap.add_argument("--no-api", action="store_true", help="ragged workload without the reference-API leg")
    ap.add_argument("--chunk-cells", type=int, default=8000000)
    ap.add_argument("--survey-size", type=int, default=60000, help="side of the configs[2] survey (the `survey` object of the default line)")
    ap.add_argument("--no-survey", action="store_true", help="default line without the configs[2] survey leg")
    ap.add_argument("--no-extra", action="store_true", help="default line without the configs[3] / configs[4] objects")
    ap.add_argument("--tile", type=int, default=2048)
    ap.add_argument("--overlap", type=int, default=128)
    ap.add_argument("--cpu-crop", type=int, default=1024, help="cpu_baseline sample: crop x crop corner (about 10-15 s of CPU work)")
    ap.add_argument("--cpu-crop-ref", type=int, default=512, help="--impl reference: crop x crop corner per step (about 3 s per step)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        if args.workload == "survey":
            run_survey(args)
        elif args.workload == "ragged":
            run_ragged(args)
        elif args.workload == "wide":
            run_wide(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
