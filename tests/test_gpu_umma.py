"""GPU: the tcgen05 (bf16x3 split, fp32 accumulate in TMEM) projection GEMM vs an fp64 reference.
fp32-grade accuracy is required: the error must be of the order of the fp32 FMA GEMM's own error."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def run_umma(A, W, att_s=None, att_d=None, heads=0, presplit=False):
    from bathymetric_gnn_b200._lib import lib, check, ptr
    M, K = A.shape
    N = W.shape[0]
    Cout = torch.empty((M, N), dtype=torch.float32, device="cuda")
    alpha = torch.empty((M, 2 * heads), dtype=torch.float32, device="cuda") if heads else None
    Wh = W.cpu().contiguous()
    fn = lib.bgnn_test_proj_umma_presplit if presplit else lib.bgnn_test_proj_umma
    check(fn(ptr(A), M, K, C.c_void_p(Wh.data_ptr()), N, ptr(Cout), ptr(att_s), ptr(att_d), ptr(alpha),
                                  heads, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return Cout, alpha


@pytest.mark.parametrize("M,K,N,heads", [(256, 64, 256, 4), (1000, 256, 256, 4), (70000, 256, 256, 4), (4099, 256, 64, 1),
                                         (513, 64, 64, 1), (300, 32, 32, 1), (777, 128, 128, 4), (50000, 64, 256, 4)])
def test_umma_matches_fp64(M, K, N, heads):
    g = torch.Generator(device="cpu").manual_seed(M + K + N)
    A = (torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, 1, generator=g))).cuda()      # rows of varied scale
    W = (torch.randn(N, K, generator=g) / np.sqrt(K)).cuda()
    att_s, att_d = torch.randn(N, generator=g).cuda(), torch.randn(N, generator=g).cuda()
    Cout, alpha = run_umma(A, W, att_s, att_d, heads)
    ref = A.double() @ W.double().t()
    fp32 = (A @ W.t()).double()                       # cuBLAS fp32 (may itself use TF32 off: default highest precision)
    scale = (A.double().abs() @ W.double().abs().t())              # sum |a||b| : the natural error scale
    err = ((Cout.double() - ref).abs() / scale).max().item()
    err32 = ((fp32 - ref).abs() / scale).max().item()
    assert err < 1e-6 and err < 4 * err32 + 2e-7, (err, err32)   # fp32-grade: same order as the cuBLAS fp32 GEMM
    CH = N // heads
    a_ref = torch.stack([(ref[:, h * CH:(h + 1) * CH] * att_s.double()[h * CH:(h + 1) * CH]).sum(1) for h in range(heads)]
                        + [(ref[:, h * CH:(h + 1) * CH] * att_d.double()[h * CH:(h + 1) * CH]).sum(1) for h in range(heads)], 1)
    a_scale = torch.stack([(scale[:, h * CH:(h + 1) * CH] * att_s.double().abs()[h * CH:(h + 1) * CH]).sum(1) for h in range(heads)]
                          + [(scale[:, h * CH:(h + 1) * CH] * att_d.double().abs()[h * CH:(h + 1) * CH]).sum(1) for h in range(heads)], 1)
    assert ((alpha.double() - a_ref).abs() / a_scale).max().item() < 1e-6
    print(f"M={M} K={K} N={N}: err/sum|a||b| umma={err:.2e} fp32={err32:.2e}")


def test_umma_toggle_equivalence_in_forward():
    """Whole forward with tensor-core projections vs fp32 FMA projections."""
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200._lib import lib
    from oracle import model_oracle as mo
    from tests import synth
    d = synth.noisy_tile(160, 200, seed=3)
    cfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(cfg, seed=5)
    m = bg.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.0, 3)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = bg.GraphBuilder().build_graph(d, None, None, (0.5, 0.5))
    a = m.set_option("tensor_cores", 1)(g)
    b = m.set_option("tensor_cores", 0)(g)           # a switch of THIS model's handle, not of the process
    m2 = bg.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.0, 3)
    m2.load_state_dict(sd)
    c = m2.cuda().eval()(g)                           # another handle is unaffected: tensor cores on
    assert torch.equal(a["class_logits"], c["class_logits"])
    np.testing.assert_allclose(a["class_logits"].cpu().numpy(), b["class_logits"].cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert torch.equal(a["predicted_class"], b["predicted_class"]) or (a["predicted_class"] != b["predicted_class"]).float().mean() < 1e-3


@pytest.mark.parametrize("M,K,N,heads", [(1, 256, 256, 4), (129, 256, 256, 4), (257, 64, 256, 4), (383, 192, 256, 4),
                                         (256, 256, 256, 4), (1000, 256, 256, 4), (70001, 256, 256, 4), (4099, 256, 64, 1),
                                         (513, 64, 64, 1), (50000, 64, 256, 4), (333, 128, 256, 8)])
def test_umma_presplit_is_bit_identical(M, K, N, heads):
    """A handed over as fp16 hi / lo planes (what the tiled aggregation writes) vs the in-kernel split of fp32 A.  N = 64: the
    same operands reach the tensor cores in the same order, so C and alpha match bit for bit.  N = 256 (K % 64 == 0) runs the
    cta_group::2 pair kernel, which folds the correction products into the ONE accumulator (D = A*B + D * 2^-11) instead of
    adding a second accumulator in the epilogue: the same three fp16 products, another fp32 summation order -- agreement to
    the accuracy both hold against fp64 (3.3e-7 * sum|a||w|, test_umma_matches_fp32_reference), rows past M untouched."""
    g = torch.Generator(device="cpu").manual_seed(M + K + N)
    A = (torch.randn(M, K, generator=g) * torch.exp(torch.randn(M, 1, generator=g))).cuda()
    W = (torch.randn(N, K, generator=g) / np.sqrt(K)).cuda()
    att_s, att_d = torch.randn(N, generator=g).cuda(), torch.randn(N, generator=g).cuda()
    C0, a0 = run_umma(A, W, att_s, att_d, heads)
    C1, a1 = run_umma(A, W, att_s, att_d, heads, presplit=True)
    if N == 256 and K % 64 == 0:
        bound = A.abs().double() @ W.abs().double().T                       # sum_k |a||w| per output
        assert bool(((C0.double() - C1.double()).abs() <= 4e-7 * bound + 1e-30).all())
        ref = A.double() @ W.double().T
        assert bool(((C1.double() - ref).abs() <= 4e-7 * bound + 1e-30).all())
        CH = N // heads
        for att, off in ((att_s, 0), (att_d, heads)):
            ab = (bound.view(M, heads, CH) * att.abs().double().view(1, heads, CH)).sum(2)
            d = (a0[:, off:off + heads].double() - a1[:, off:off + heads].double()).abs()
            assert bool((d <= 1e-6 * ab + 1e-30).all())
    else:
        assert torch.equal(C0, C1) and torch.equal(a0, a1)


def test_presplit_handover_in_forward():
    """Whole forward with the fp16 hi / lo hand-over between aggregation and projection vs fp32 intermediates: the last
    projection (N = 64) consumes bit-identical operands either way; the N = 256 projections run the cta_group::2 pair kernel
    on the hand-over path (other fp32 summation order, see above) -> agreement far inside the fp32 tolerance of the path."""
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200._lib import lib
    from oracle import model_oracle as mo
    from tests import synth
    d = synth.noisy_tile(300, 333, seed=9)
    mask = synth.blob_mask(300, 333, seed=9, invalid_frac=0.05)
    cfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(cfg, seed=7, randomize_bn=True)
    m = bg.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.0, 3)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = bg.GraphBuilder().build_graph(d, mask, None, (0.5, 0.5))
    a = m.set_option("presplit", 1)._run(g, 0.85, 0.6, True, return_hidden=True)
    b = m.set_option("presplit", 0)._run(g, 0.85, 0.6, True, return_hidden=True)
    for k in ("hidden", "class_logits", "confidence", "correction"):
        np.testing.assert_allclose(a[k].cpu().numpy(), b[k].cpu().numpy(), rtol=2e-6, atol=2e-6, err_msg=k)
    for k in ("predicted_class", "action"):
        assert (a[k] != b[k]).float().mean().item() < 1e-3, k


def test_adversarial_checkpoint_operand_scaling():
    """the fp16 hi / lo operand split scales A by an ANALYTIC bound (bn_mul * max|h| + bn_add), not by the measured
    maximum.  Adversarial checkpoint: one BatchNorm channel per layer with running_var ~ 0 (1 / sqrt(var + eps) = 316), one
    projection-weight outlier (x 50) and input spikes of 1000 x the typical depth range, so that ordinary activations sit
    ~2^-15 below the bound.  The forward must still hold the fp32 tolerance against the fp64 oracle, scaled by the oracle's own
    fp32 envelope on this input (no silent degradation)."""
    import bathymetric_gnn_b200 as bg
    from oracle import graph_oracle as go
    from oracle import model_oracle as mo
    from tests import synth
    d = synth.noisy_tile(120, 150, seed=4)
    rng = np.random.default_rng(0)
    idx = rng.integers(0, d.size, 12)
    d.ravel()[idx] += rng.choice([-1.0, 1.0], 12).astype(np.float32) * 3000.0        # spikes
    cfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(cfg, seed=9, randomize_bn=True)
    for l in range(cfg.layers):
        sd[f"gnn.norms.{l}.module.running_var"][3 + l] = 1e-12
        sd[f"gnn.convs.{l}.lin.weight"][5, 7] *= 50.0
    m = bg.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.0, 3)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    g = bg.GraphBuilder().build_graph(d, None, None, (0.5, 0.5))
    ref_g = go.build_graph(d, None, None, (0.5, 0.5))
    x = g.x.cpu().numpy()
    r32 = mo.forward(sd, cfg, x, ref_g["edge_index"], ref_g["edge_attr"])
    r64 = mo.forward(sd, cfg, x, ref_g["edge_index"], ref_g["edge_attr"], dtype=torch.float64)
    out_tc = m.set_option("tensor_cores", 1)(g)
    out_fma = m.set_option("tensor_cores", 0)(g)
    for k in ("class_logits", "confidence", "correction"):
        ref = r64[k].numpy()
        env = float(np.abs(r32[k].numpy().astype(np.float64) - ref).max())
        for name, o in (("tensor cores", out_tc), ("fma", out_fma)):
            err = np.abs(o[k].cpu().numpy().astype(np.float64) - ref)
            assert np.all(err <= max(2e-5, 3.0 * env) + 1e-5 * np.abs(ref)), (k, name, float(err.max()), env)
