"""Throughput of the calls the UNMODIFIED scripts/inference_native.py makes through its in-script NativeVRProcessor
(per-grid GraphBuilder.build_graph + Batch.from_data_list + predict + graph_to_grid), on the overlay."""
import os, sys, time, importlib.util, logging
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bathymetric_gnn_b200 import dropin
dropin.install()
import bathymetric_gnn_b200 as bg
import bench
from tests import synth
spec = importlib.util.spec_from_file_location("fixture_native", os.path.join(ROOT, "tests/fixtures/ref_layout/scripts/inference_native.py"))
mod = importlib.util.module_from_spec(spec); mod.logger = logging.getLogger("x"); spec.loader.exec_module(mod)
G = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cfg_m, sd = bench.make_weights(0)
model = bg.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.0, 3); model.load_state_dict(sd); model = model.cuda().eval()
grids = synth.ragged_grids(G, seed=7)
valid = sum(int(((d != 1.0e6) & np.isfinite(d)).sum()) for d, _, _ in grids)
for budget in (50000, 2000000):
    proc = mod.ScriptSideProcessor(model, bg.GraphBuilder("8-connected"), torch.device("cuda"))
    proc.BATCH_NODE_BUDGET = budget
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); acc = 0.0
        for d, u, res in grids:
            proc.add_to_batch(d, u, res, 1.0e6)
            if proc.batch_ready:
                for cls, conf, corr in proc.flush_batch(): acc += float(conf.flat[0])
        for cls, conf, corr in proc.flush_batch(): acc += float(conf.flat[0])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"script-side class, budget {budget}: {G} grids, {valid} valid cells in {dt:.3f} s = {valid/dt:.3e} cells/s, {dt/G*1e6:.0f} us per grid")
