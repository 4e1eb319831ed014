"""Dev aid: per-phase cycle accounting of gat_aggregate_tile (library built with -DBGNN_AGG_TIMING)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["BGNN_LIB"] = os.path.join(ROOT, "tools", "libbgnn_timing.so")
sys.path.insert(0, ROOT)
import torch
import bench
import bathymetric_gnn_b200 as bg
from bathymetric_gnn_b200._lib import lib
lib.bgnn_debug_agg_timing.argtypes = [C.c_void_p, C.c_int]
depth, mask = bench.make_tile(0)
cfg, sd = bench.make_weights(0)
model = bg.BathymetricGNN(cfg.in_channels, cfg.hidden, cfg.layers, "GAT", cfg.heads, cfg.num_classes, cfg.predict_correction, 0.0, cfg.edge_dim)
model.load_state_dict(sd); model = model.cuda().eval()
g = bg.GraphBuilder("8-connected").build_graph(torch.from_numpy(depth).cuda(), torch.from_numpy(mask).cuda(), None, (0.5, 0.5))
for _ in range(2):
    model.predict(g)
torch.cuda.synchronize()
lib.bgnn_debug_agg_timing(None, 1)
model.predict(g); torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
lib.bgnn_debug_agg_timing(out, 0)
names = ["ids+asrc", "eattr setup", "issue+scores (x heads)", "wait copies", "aggregate", "end sync"]
tot = sum(out[i] for i in range(6))
for i, n in enumerate(names):
    print(f"{n:26s} {out[i]/1e6:10.1f} Mcycles  {100*out[i]/tot:5.1f}%")
print("total (thread 0 of every CTA, 4 layers):", tot/1e6, "Mcycles")
