"""Where a survey tile's 14.2 ms go: device time per library slot (events inside the library) vs wall time per tile."""
import ctypes as C, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bathymetric_gnn_b200 as bg
from bathymetric_gnn_b200 import survey as sv
from bathymetric_gnn_b200._lib import lib
import bench
from tests import synth

S = int(sys.argv[1]) if len(sys.argv) > 1 else 12000
dev = torch.device("cuda", 0)
cfg_m, sd = bench.make_weights(0)
model = bg.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.0, 3)
model.load_state_dict(sd); model = model.to(dev).eval()
cfg = bg.Config(); cfg.tile.tile_size, cfg.tile.overlap, cfg.tile.min_valid_ratio = 2048, 128, 0.1
pipe = bg.BathymetricPipeline(cfg); pipe.model = model
tabs = synth.survey_tables(S)
depth = synth.survey_block_torch(S, 0, S, 0, S, dev, tables=tabs).cpu().pin_memory()
ntr, _, specs = pipe.tile_manager.compute_tile_grid((S, S))
plan = sv.plan_bands(specs, ntr, 1, (S, S))[0]
out = sv.pinned_outputs(S, S)
sv.process_band(pipe, plan, depth, None, synth.SURVEY_NODATA, (0.5, 0.5), out, None)
for prof in (0, 1):
    lib.bgnn_profile_enable(prof)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = sv.process_band(pipe, plan, depth, None, synth.SURVEY_NODATA, (0.5, 0.5), out, None)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    tiles = res["stats"]["tiles"]
    print(f"profile={prof}: {dt*1e3:.1f} ms total, {tiles} tiles, {dt*1e3/tiles:.3f} ms per tile")
n = lib.bgnn_profile_collect(0, None, None)
slots = (C.c_int32 * n)(); ms = (C.c_float * n)()
lib.bgnn_profile_collect(n, slots, ms)
lib.bgnn_profile_enable(0)
agg = {}
for s_, m_ in zip(slots, ms):
    k = bench.SLOT_NAMES.get(int(s_), str(int(s_)))
    agg[k] = agg.get(k, 0.0) + float(m_)
tot = sum(agg.values())
print("device ms per tile by slot:", {k: round(v / tiles, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])})
print("device sum per tile %.3f ms" % (tot / tiles))
