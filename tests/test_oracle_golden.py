"""CPU: the oracle restatements vs the fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from oracle import model_oracle as mo
from oracle import stitch_oracle as so
from tests import golden_io as gio
from tests import synth

GZ = gio.load("graph_cases.npz")


@pytest.mark.parametrize("name", gio.graph_case_names(GZ))
def test_graph_oracle_bit_exact_vs_reference(name):
    c, ref = gio.graph_case(GZ, name)
    out = go.build_graph(c["depth"], c["mask"], c["unc"], c["res"], c["conn"])
    for k in ("x", "edge_index", "edge_attr", "pos", "local_std"):
        assert out[k].dtype == ref[k].dtype, k
        assert out[k].shape == ref[k].shape, k
        assert np.array_equal(out[k], ref[k], equal_nan=True), k
    if "valid_rows" in ref:
        assert np.array_equal(out["valid_rows"], ref["valid_rows"])
        assert np.array_equal(out["valid_cols"], ref["valid_cols"])
        assert out["num_valid_cells"] == int(ref["num_valid_cells"])
        assert tuple(out["grid_shape"]) == tuple(ref["grid_shape"])
    else:
        assert out["grid_shape"] is None            # reference's empty graph has no grid_shape


def test_graph_oracle_c1_digests():
    """BASELINE config 1 (512^2, reference SyntheticNoiseGenerator seed 42)."""
    z = gio.load("c1_tile.npz")
    d = z["depth"]
    out = go.build_graph(d, np.ones_like(d, bool), None, tuple(z["res"]))
    assert out["x"].shape[0] == int(z["N"]) and out["edge_index"].shape[1] == int(z["E"])
    for k in ("x", "edge_index", "edge_attr", "pos", "local_std", "valid_rows", "valid_cols"):
        digest = hashlib.sha256(np.ascontiguousarray(out[k]).tobytes()).hexdigest()
        assert digest == str(z["sha_" + k]), k


def test_edge_count_formula():
    # E = 4(n-1)(2n-1) for a full n x n 8-connected tile
    d = synth.noisy_tile(20, 20)
    out = go.build_graph(d, None, None)
    assert out["edge_index"].shape[1] == 4 * 19 * 39


def test_graph_to_grid_errors():
    with pytest.raises(ValueError):
        go.graph_to_grid((2, 2), np.array([0]), np.array([0]), np.zeros((1, 2)))


MZ = gio.load("model_cases.npz")


def _model_case(key):
    F, seed, hidden, heads = (int(v) for v in MZ[key + "/meta"])
    c, ref = gio.graph_case(GZ, str(MZ[key + "/graph"]))
    cfg = mo.ModelCfg(in_channels=F, hidden=hidden, heads=heads)
    sd = mo.init_state_dict(cfg, seed=seed, randomize_bn=True, head_gain=8.0)
    sd["classification_head.mlp.3.bias"] = torch.from_numpy(MZ[key + "/cls_bias"])
    sd["confidence_head.mlp.3.bias"] = torch.from_numpy(MZ[key + "/conf_bias"])
    return cfg, sd, ref


@pytest.mark.parametrize("key", gio.model_case_keys(MZ))
def test_model_oracle_vs_reference_module(key):
    cfg, sd, g = _model_case(key)
    assert mo.param_count(sd) == int(MZ[key + "/param_count"])
    out = mo.predict(sd, cfg, g["x"], g["edge_index"], g["edge_attr"])
    # same torch ops in the same order as the reference module on the restated PyG classes
    for k in ("class_logits", "class_probs", "confidence", "correction"):
        np.testing.assert_allclose(out[k].numpy(), MZ[f"{key}/{k}"], rtol=1e-6, atol=1e-6, err_msg=k)
    for k in ("predicted_class", "action", "needs_review", "auto_correct"):
        assert np.array_equal(out[k].numpy(), MZ[f"{key}/{k}"]), k
    assert len(np.unique(MZ[f"{key}/predicted_class"])) == 3     # all three classes occur


def test_structural_pins():
    # docs/QUICK_REFERENCE.md:185 "182K parameters"
    assert int(MZ["param_count_F8"]) == 182469 and int(MZ["param_count_F7"]) == 182405
    for F in (7, 8):
        sd = mo.init_state_dict(mo.ModelCfg(in_channels=F), seed=0)
        assert mo.param_count(sd) == int(MZ[f"param_count_F{F}"])
    sd = mo.init_state_dict(mo.ModelCfg(in_channels=7), seed=0)
    assert sorted(sd.keys()) == [str(k) for k in MZ["state_keys"]]


def test_model_oracle_fp64_envelope():
    """fp32 oracle vs its own fp64 evaluation: the envelope the CUDA tolerance is
    stated against (~2.5e-6 on logits)."""
    key = gio.model_case_keys(MZ)[0]
    cfg, sd, g = _model_case(key)
    o32 = mo.forward(sd, cfg, g["x"], g["edge_index"], g["edge_attr"], torch.float32)
    o64 = mo.forward(sd, cfg, g["x"], g["edge_index"], g["edge_attr"], torch.float64)
    err = (o32["class_logits"].double() - o64["class_logits"]).abs().max().item()
    assert err < 2e-4, err


SZ = gio.load("stitch_cases.npz")


def test_tile_specs_vs_reference():
    i = 0
    while f"specs{i}/cfg" in SZ.files:
        H, W, ts, ov = (int(v) for v in SZ[f"specs{i}/cfg"])
        ntr, ntc, specs = so.tile_specs((H, W), ts, ov)
        assert [ntr, ntc] == list(SZ[f"specs{i}/n"])
        assert np.array_equal(np.array(specs, np.int64), SZ[f"specs{i}/specs"])
        i += 1
    assert i >= 5
    with pytest.raises(ValueError):
        so.tile_specs((10, 10), 16, 16)


def test_blend_tables_vs_reference():
    for k in SZ.files:
        if k.startswith("blend1d/"):
            size, ov = (int(v) for v in k.split("/")[1].split("_"))
            assert np.array_equal(so.blend_1d(size, ov), SZ[k]), k
    assert np.array_equal(so.blend_2d((50, 64), 16), SZ["blend2d/50x64_16"])


def test_merge_vs_reference():
    H, W, ts, ov = (int(v) for v in SZ["merge/cfg"])
    _, _, specs = so.tile_specs((H, W), ts, ov)
    m = so.Merger((H, W), ["cleaned_depth", "classification", "confidence", "correction"], ov)
    for i in SZ["merge/tile_ids"]:
        td = {k: SZ[f"merge/tile{i}/{k}"] for k in ("cleaned_depth", "classification", "confidence", "correction")}
        m.add_tile(specs[int(i)], td)
    res = m.finalize()
    for k, v in res.items():
        assert np.array_equal(v, SZ[f"merge/ref/{k}"], equal_nan=True), k
    # quirk: outer ring of the survey keeps weight 0 -> stays 0.0
    assert np.all(res["confidence"][0, :40] == 0.0)
    with pytest.raises(ValueError):
        m2 = so.Merger((H, W), ["confidence"], ov)
        m2.add_tile(specs[0], {"bogus": np.zeros((ts, ts), np.float32)})


def test_pipeline_oracle_vs_reference():
    """graph oracle + model oracle + stitch oracle == reference BathymetricPipeline.process."""
    z = gio.load("pipeline_case.npz")
    H, W, ts, ov = (int(v) for v in z["cfg"])
    depth, nodata = z["depth"], float(z["nodata"])
    valid = np.isfinite(depth) & (depth != nodata)
    cfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(cfg, seed=int(z["seed"]), randomize_bn=True, head_gain=8.0)
    sd["classification_head.mlp.3.bias"] = torch.from_numpy(z["cls_bias"])
    sd["confidence_head.mlp.3.bias"] = torch.from_numpy(z["conf_bias"])
    _, _, specs = so.tile_specs((H, W), ts, ov)
    chans = ["cleaned_depth", "classification", "confidence", "correction"]
    m = so.Merger((H, W), chans, ov)
    for s in specs:
        r0, c0, r1, c1 = s[:4]
        td, tm = depth[r0:r1, c0:c1], valid[r0:r1, c0:c1]
        if tm.sum() / tm.size < float(z["min_valid_ratio"]):
            continue
        g = go.build_graph(td, tm, None, tuple(z["res"]))
        o = mo.predict(sd, cfg, g["x"], g["edge_index"], g["edge_attr"])
        cls, conf, corr = so.tile_outputs_to_grids(td.shape, g["valid_rows"], g["valid_cols"],
                                                   o["predicted_class"].numpy(), o["confidence"].numpy(),
                                                   o["correction"].numpy(), g["local_std"])
        m.add_tile(s, dict(cleaned_depth=td, classification=cls, confidence=conf, correction=corr))
    res = so.finish_survey(depth, valid, m.finalize())
    assert np.array_equal(res["classification"], z["ref/classification"], equal_nan=True)
    np.testing.assert_allclose(res["confidence"], z["ref/confidence"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(res["correction"], z["ref/correction"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res["cleaned_depth"], z["ref/cleaned_depth"], rtol=1e-6, atol=1e-5)
    assert (res["cleaned_depth"] != depth).sum() > 100


@pytest.mark.parametrize("key", ["g64_holes_s0_h64x4", "g33_k24_s4_h128x4"])
def test_model_oracle_bf16_storage_vs_golden(key):
    """The bf16-storage restatement (the checker of the CUDA bf16 mode) reproduces its committed vectors, and stays within
    the stated bf16 bound of the UNMODIFIED reference module's fp32 outputs stored beside them (logit bounds x 8: these
    fixtures use classification head gain 8).  Across machines torch's CPU matmul may sum in a different order, which can
    move an activation over a bf16 rounding boundary: hence 5e-3, not 0."""
    z = gio.load("model_bf16_cases.npz")
    F, seed, hidden, heads = (int(v) for v in z[key + "/meta"])
    c, _ = gio.graph_case(GZ, str(z[key + "/graph"]))
    cfg = mo.ModelCfg(in_channels=F, hidden=hidden, heads=heads)
    sd = mo.init_state_dict(cfg, seed=seed, randomize_bn=True, head_gain=8.0)
    sd["classification_head.mlp.3.bias"] = torch.from_numpy(z[key + "/cls_bias"])
    sd["confidence_head.mlp.3.bias"] = torch.from_numpy(z[key + "/conf_bias"])
    g = go.build_graph(c["depth"], c["mask"], c["unc"], c["res"], connectivity=c["conn"])
    out = mo.forward(sd, cfg, g["x"], g["edge_index"], g["edge_attr"], torch.float32, bf16_storage=True)
    for k, gain in (("class_logits", 8.0), ("confidence", 1.0), ("correction", 1.0)):
        np.testing.assert_allclose(out[k].numpy(), z[f"{key}/{k}"], rtol=0, atol=5e-3 * gain)
        np.testing.assert_allclose(out[k].numpy(), z[f"{key}/ref32/{k}"], rtol=0, atol=(5e-3 if k == "confidence" else 2e-2 * gain))
    # fp32 evaluation of the same oracle == the reference module's fp32 vectors (the restatement itself is unchanged)
    o32 = mo.forward(sd, cfg, g["x"], g["edge_index"], g["edge_attr"], torch.float32)
    np.testing.assert_allclose(o32["class_logits"].numpy(), z[key + "/ref32/class_logits"], rtol=1e-5, atol=2e-5)


class _Grid:
    """Attribute bag with the RefinementGrid surface (data/vr_bag.py:66-97) the sidecar code reads."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    @property
    def valid_mask(self):
        return (self.depth != 1.0e6) & np.isfinite(self.depth)


def test_sidecar_oracle_bit_exact_vs_reference():
    """oracle/vr_post_oracle.py::SidecarOracle == the UNMODIFIED reference SidecarBuilder.add_refinement_results on the
    committed scene (overlapping grids, raster-edge clipping, NaN / nodata cells): all four planes bit for bit."""
    from oracle import vr_post_oracle as vo
    z = gio.load("sidecar_case.npz")
    geo, grids, results = synth.sidecar_scene()
    sc = vo.SidecarOracle(**geo)
    for g, (cls, conf, corr) in zip(grids, results):
        sc.add_refinement_results(_Grid(**g), cls, conf, corr)
    for k in ("classification", "confidence", "correction", "valid_mask"):
        assert np.array_equal(getattr(sc, k), z[k], equal_nan=True), k
    assert np.isnan(z["classification"]).any() and (z["valid_mask"] == 1).any()


def test_vr_apply_results_oracle_properties():
    """apply_results restatement (inference_native.py:480-503): untouched outside the apply mask, `>=` threshold, fp32 ops."""
    from oracle import vr_post_oracle as vo
    _, grids, results = synth.sidecar_scene()
    n_apply = 0
    for g, (cls, conf, corr) in zip(grids, results):
        cd, cu, st = vo.apply_results(g["depth"], g["uncertainty"], cls, conf, corr, 0.85)
        v = vo.valid_mask(g["depth"])
        m = (cls == 2) & v & (conf >= np.float32(0.85))
        assert st["cells_corrected"] == int(m.sum()) and st["cells_processed"] == int(v.sum())
        assert np.array_equal(cd[~m], g["depth"][~m], equal_nan=True) and np.array_equal(cu[~m], g["uncertainty"][~m])
        assert np.array_equal(cd[m], g["depth"][m] - corr[m]) and np.array_equal(cu[m], g["uncertainty"][m] * (np.float32(2.0) - conf[m]))
        assert cd.dtype == np.float32 and cu.dtype == np.float32
        n_apply += int(m.sum())
        assert ((conf == np.float32(0.85)) & (cls == 2) & v).sum() == ((conf == np.float32(0.85)) & m).sum()
    assert n_apply > 50


def test_gat_layer_vs_independent_dense_formulation():
    """Second opinion on the restated GATConv (PyG is absent, so the attention arithmetic cannot be pinned to it): the same
    layer written INDEPENDENTLY as dense [N, N] attention in fp64 -- Velickovic et al. 2018, eq. (1)-(4), plus PyG's
    documented edge term a_e^T (W_e e_ij) and `add_self_loops(fill_value='mean')` -- shares no code with
    oracle/pyg_restated.py (no scatter / index_add / segment softmax).  Agreement to fp64 round-off on a graph with holes,
    an isolated node and asymmetric edge attributes."""
    torch.manual_seed(3)
    N, Fin, H, C, ED = 37, 10, 3, 8, 3
    x = torch.randn(N, Fin, dtype=torch.float64)
    dense = (torch.rand(N, N) < 0.15)
    dense.fill_diagonal_(False)
    dense[:, 5] = False                                          # node 5 has no out-edges ...
    dense[5, :] = False                                          # ... and no in-edges: only its self loop, loop attribute 0
    src, dst = dense.nonzero(as_tuple=True)                      # edge j -> i stored as dense[j, i]
    ei = torch.stack([src, dst])
    ea = torch.randn(ei.shape[1], ED, dtype=torch.float64)
    sd = {}
    p = "gnn.convs.0."
    sd[p + "lin.weight"] = torch.randn(H * C, Fin, dtype=torch.float64) * 0.3
    sd[p + "att_src"] = torch.randn(1, H, C, dtype=torch.float64)
    sd[p + "att_dst"] = torch.randn(1, H, C, dtype=torch.float64)
    sd[p + "lin_edge.weight"] = torch.randn(H * C, ED, dtype=torch.float64) * 0.3
    sd[p + "att_edge"] = torch.randn(1, H, C, dtype=torch.float64)
    sd[p + "bias"] = torch.randn(H * C, dtype=torch.float64)
    got = mo.gat_layer(sd, 0, x, ei, ea, H, C, True)
    # ---- dense formulation
    h = (x @ sd[p + "lin.weight"].T).view(N, H, C)
    a_s = torch.einsum("nhc,hc->nh", h, sd[p + "att_src"][0])
    a_d = torch.einsum("nhc,hc->nh", h, sd[p + "att_dst"][0])
    E = torch.zeros(N, N, ED, dtype=torch.float64)
    E[src, dst] = ea
    deg = dense.sum(0)                                           # in-degree of i
    loop_attr = torch.where(deg[:, None] > 0, E.sum(0) / deg.clamp(min=1)[:, None], torch.zeros(N, ED, dtype=torch.float64))
    A = dense.clone()
    idx = torch.arange(N)
    A[idx, idx] = True
    E[idx, idx] = loop_attr
    We = sd[p + "lin_edge.weight"].view(H, C, ED)
    m_e = torch.einsum("hcd,hc->hd", We, sd[p + "att_edge"][0])               # a_e^T W_e as one [H, ED] map
    a_e = torch.einsum("jid,hd->jih", E, m_e)
    score = torch.nn.functional.leaky_relu(a_s[:, None, :] + a_d[None, :, :] + a_e, 0.2)      # [j, i, h]
    score = score.masked_fill(~A[:, :, None], float("-inf"))
    att = torch.softmax(score, dim=0)                                          # over sources j of target i
    want = torch.einsum("jih,jhc->ihc", att, h).reshape(N, H * C) + sd[p + "bias"]
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-12, atol=1e-12)
    # mean over heads (the last layer's concat=False)
    got_m = mo.gat_layer({**sd, p + "bias": sd[p + "bias"][:C]}, 0, x, ei, ea, H, C, False)
    want_m = torch.einsum("jih,jhc->ihc", att, h).mean(1) + sd[p + "bias"][:C]
    np.testing.assert_allclose(got_m.numpy(), want_m.numpy(), rtol=1e-12, atol=1e-12)


def test_multiscale_oracle_bit_exact_vs_reference():
    """graph_oracle.build_multiscale_graph == the UNMODIFIED reference MultiScaleGraphBuilder (scales 1 / 2 / 4, cropped
    block means with NaN holes, majority mask, scaled resolution): every Data field bit for bit."""
    import warnings
    z = gio.load("multiscale_case.npz")
    d, m, u = synth.multiscale_case()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)          # nanmean of an all-NaN block, as in the reference
        graphs = go.build_multiscale_graph(d, m, u, (0.5, 1.0), scales=(4, 1, 2))
    assert sorted(graphs) == [1, 2, 4]
    for s, g in graphs.items():
        for k in ("x", "edge_index", "edge_attr", "pos", "valid_rows", "valid_cols", "local_std"):
            assert np.array_equal(g[k], z[f"s{s}/{k}"], equal_nan=True), (s, k)
        assert tuple(g["grid_shape"]) == tuple(z[f"s{s}/grid_shape"]) and int(g["num_valid_cells"]) == int(z[f"s{s}/num_valid_cells"])
