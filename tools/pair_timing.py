"""Dev aid: role cycle counters of proj_pair_kernel (cta_group::2) during one forward of the bench tile (needs tools/libbgnn_timing.so
built with -DBGNN_AGG_TIMING '-DBGNN_UT_FILTER=false' so that only the pair kernel's own counters tick)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("BGNN_LIB", os.path.join(ROOT, "tools", "libbgnn_timing.so"))
os.environ["BGNN_NO_CG2"] = "1"; os.environ["PAIR_MLP0_ONLY"] = "1"
sys.path.insert(0, ROOT)
import torch
import bench
import bathymetric_gnn_b200 as bg
from bathymetric_gnn_b200._lib import lib
lib.bgnn_debug_umma_timing.argtypes = [C.c_void_p, C.c_int]
depth, mask = bench.make_tile(0)
cfg, sd = bench.make_weights(0)
model = bg.BathymetricGNN(cfg.in_channels, cfg.hidden, cfg.layers, "GAT", cfg.heads, cfg.num_classes, cfg.predict_correction, 0.0, cfg.edge_dim)
model.load_state_dict(sd); model = model.cuda().eval()
g = bg.GraphBuilder("8-connected").build_graph(torch.from_numpy(depth).cuda(), torch.from_numpy(mask).cuda(), None, (0.5, 0.5))
for _ in range(2):
    model.predict(g)
torch.cuda.synchronize()
lib.bgnn_debug_umma_timing(None, 1)
model.predict(g); torch.cuda.synchronize()
out = (C.c_ulonglong * 12)()
lib.bgnn_debug_umma_timing(out, 0)
names = ["mma: wait TEMPTY (epilogues)", "mma: wait FULL pass 1", "mma: issue", "epi w0: wait TFULL", "epi w0: drain + store",
         "loader: wait EMPTY", "loader: issue", "mma: wait FULL pass 2", "mma thread: total cycles", "mma thread: total ns"]
n = 74 * (1 if os.environ.get("PAIR_MLP0_ONLY") else 2)        # leader CTAs x launches
for i, nm in enumerate(names):
    print(f"   {nm:34s} {out[i]/n/1e3:10.1f} kcycles per leader CTA and launch")
print(f"   SM clock {out[8]/max(out[9],1):.3f} GHz; kernel time {out[9]/n/1e6:.3f} ms")
