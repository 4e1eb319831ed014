"""dev aid: bf16 tcgen05 projection vs fp64 -- where do the errors sit?"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from bathymetric_gnn_b200._lib import lib, check, ptr, stream_ptr

def bits(t): return (t.contiguous().view(torch.int32) >> 16).to(torch.int16).contiguous()

for N, K, heads, M in [(512, 512, 4, 1000), (256, 64, 4, 77), (64, 256, 1, 2049)]:
    g = torch.Generator(device="cpu").manual_seed(1)
    A = torch.randn(M, K, generator=g).bfloat16().float()
    W = (torch.randn(N, K, generator=g) / np.sqrt(K)).bfloat16().float()
    att_s, att_d = torch.randn(N, generator=g), torch.randn(N, generator=g)
    Ad, Wd = bits(A).cuda(), bits(W).cuda()
    CH = N // heads
    Cd = torch.zeros(heads * M * CH, dtype=torch.int16, device="cuda")
    al = torch.zeros(M, 2 * heads, dtype=torch.float32, device="cuda")
    as_d, ad_d = att_s.cuda(), att_d.cuda()      # kept alive across the launch
    check(lib.bgnn_test_proj_bf16(ptr(Ad), M, K, ptr(Wd), N, ptr(Cd), ptr(as_d), ptr(ad_d), ptr(al), heads, stream_ptr()))
    torch.cuda.synchronize()
    got = (Cd.cpu().to(torch.int32) << 16).view(torch.float32).view(heads, M, CH).permute(1, 0, 2).reshape(M, N).double().numpy()
    ref = (A.double() @ W.double().t()).numpy()
    half = 2.0 ** (np.floor(np.log2(np.maximum(np.abs(ref), 1e-30))) - 8)
    err = np.abs(got - ref)
    bad = err > 1.02 * half + 5e-6
    print(f"N={N} K={K} M={M}: bad {bad.sum()} of {bad.size}; max err/half {np.max(err / (half + 1e-12)):.3f}; max abs err {err.max():.3e}")
    if bad.any():
        r, c = np.nonzero(bad)
        print("  bad rows (first 10)", np.unique(r)[:10], "n rows", len(np.unique(r)), "bad cols (first 10)", np.unique(c)[:10], "n cols", len(np.unique(c)))
        i = np.argmax(err); print("  worst", np.unravel_index(i, err.shape), got.flat[i], ref.flat[i])
    ra = np.concatenate([(ref.reshape(M, heads, CH) * att_s.double().numpy().reshape(heads, CH)).sum(-1),
                         (ref.reshape(M, heads, CH) * att_d.double().numpy().reshape(heads, CH)).sum(-1)], axis=1)
    print("  alpha max err", np.abs(al.cpu().double().numpy() - ra).max(), "scale", np.abs(ra).max())
