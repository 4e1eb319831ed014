"""Dev aid: where the MMA issuer and one epilogue thread of proj_umma_kernel spend their cycles (needs -DBGNN_AGG_TIMING)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("BGNN_LIB", os.path.join(ROOT, "tools", "libbgnn_timing.so"))
sys.path.insert(0, ROOT)
import numpy as np, torch
from bathymetric_gnn_b200._lib import lib, check, ptr
lib.bgnn_debug_umma_timing.argtypes = [C.c_void_p, C.c_int]
fn = lib.bgnn_test_proj_umma_presplit if os.environ.get('PRESPLIT') else lib.bgnn_test_proj_umma
M = 4099615
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for (K, N, heads) in [(256, 256, 4), (256, 64, 1)]:
    A = torch.relu(torch.randn(M, K, device="cuda"))
    W = torch.randn(N, K) / np.sqrt(K)
    att = torch.randn(N, device="cuda")
    Cout = torch.empty((M, N), dtype=torch.float32, device="cuda")
    alpha = torch.empty((M, 2 * heads), dtype=torch.float32, device="cuda")
    for rep in range(3):
        lib.bgnn_debug_umma_timing(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        check(fn(ptr(A), M, K, C.c_void_p(W.data_ptr()), N, ptr(Cout), ptr(att), ptr(att), ptr(alpha), heads, st))
    out = (C.c_ulonglong * 12)()
    lib.bgnn_debug_umma_timing(out, 0)
    names = ["mma: wait TEMPTY (epilogue)", "mma: wait FULL (operands)", "mma: issue", "epi: wait TFULL (MMA)", "epi: drain + store",
             "prod g0: wait RFULL (TMA A)", "prod g0: wait EMPTY (MMA frees stage)", "prod g0: split + STS", "mma thread: total cycles", "mma thread: total ns (globaltimer)"]
    e0.record(); check(fn(ptr(A), M, K, C.c_void_p(W.data_ptr()), N, ptr(Cout), ptr(att), ptr(att), ptr(alpha), heads, st)); e1.record(); torch.cuda.synchronize()
    print(f"K={K} N={N}: per-CTA average cycles (148 CTAs); call incl. setup {e0.elapsed_time(e1):.3f} ms; WHOLE={os.environ.get('BGNN_N256_WHOLE')}")
    for i, n in enumerate(names):
        print(f"   {n:30s} {out[i]/148/1e3:10.1f} kcycles")
    print(f"   SM clock during the kernel = {out[8]/max(out[9],1):.3f} GHz; kernel time per call = {out[9]/148/2/1e6:.3f} ms")
    del A, Cout, alpha
