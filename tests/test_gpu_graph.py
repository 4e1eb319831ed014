"""GPU parity: device graph build (through the C ABI) vs the CPU oracle and the
reference golden fixtures.  Integer outputs bit-exact; float columns bit-exact
except local_mean / local_std (and what depends on the local-mean fill), which
are held to the tolerances stated here (SciPy's running-sum
box filter cannot be reproduced bit-for-bit by a direct sum)."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from tests import golden_io as gio
from tests import synth

pytestmark = pytest.mark.gpu

GZ = gio.load("graph_cases.npz")


def ulp32(a):
    return np.spacing(np.abs(a).astype(np.float32)).astype(np.float64)


def build(c):
    import bathymetric_gnn_b200 as bg
    conn = c["conn"]
    b = bg.GraphBuilder(conn)
    return b, b.build_graph(c["depth"], c["mask"], c["unc"], c["res"])


def check_graph(g, ref, depth, has_holes=True):
    N = ref["x"].shape[0]
    assert g.num_nodes == N
    x = g.x.cpu().numpy()
    assert x.dtype == np.float32 and x.shape == ref["x"].shape
    scale = float(np.nanmax(np.abs(np.where(np.isfinite(ref["x"][:, 0]), ref["x"][:, 0], 0)))) if N else 1.0
    # integer paths: bit exact
    ei = g.edge_index.cpu().numpy()
    assert ei.dtype == np.int64
    assert np.array_equal(ei, ref["edge_index"])
    assert np.array_equal(g.valid_rows.cpu().numpy(), ref["valid_rows"])
    assert np.array_equal(g.valid_cols.cpu().numpy(), ref["valid_cols"])
    assert np.array_equal(g.pos.cpu().numpy(), ref["pos"])
    assert g.num_valid_cells == int(ref["num_valid_cells"]) and tuple(g.grid_shape) == tuple(ref["grid_shape"])
    # depth column bit exact
    assert np.array_equal(x[:, 0], ref["x"][:, 0], equal_nan=True)
    # local_mean: <= 1 ulp32 ; local_std: <= 1e-7 * max|depth| (tolerances of SURVEY.md 8(c))
    assert np.all(np.abs(x[:, 1].astype(np.float64) - ref["x"][:, 1]) <= ulp32(ref["x"][:, 1]))
    assert np.all(np.abs(x[:, 2].astype(np.float64) - ref["x"][:, 2]) <= 1e-7 * max(scale, 1.0) + 1e-12)
    ls = g.local_std.cpu().numpy()
    assert np.all(np.abs(ls.astype(np.float64) - ref["local_std"]) <= 1e-7 * max(scale, 1.0) + 1e-12)
    # gradients / magnitude / curvature: exact unless a local-mean-filled neighbour is involved
    tol = 4 * float(np.spacing(np.float32(max(scale, 1.0))))
    for col in (3, 4, 5, 6):
        diff = np.abs(x[:, col].astype(np.float64) - ref["x"][:, col])
        assert np.all(diff <= tol), (col, diff.max())
        if not has_holes:
            assert np.array_equal(x[:, col], ref["x"][:, col]), col
    if x.shape[1] > 7:
        assert np.array_equal(x[:, 7], ref["x"][:, 7])
    # edge features: distance and depth_difference bit exact, slope <= 1 ulp32
    ea = g.edge_attr.cpu().numpy()
    assert ea.shape == ref["edge_attr"].shape
    assert np.array_equal(ea[:, 0], ref["edge_attr"][:, 0])
    assert np.array_equal(ea[:, 1], ref["edge_attr"][:, 1])
    assert np.all(np.abs(ea[:, 2].astype(np.float64) - ref["edge_attr"][:, 2]) <= ulp32(ref["edge_attr"][:, 2]))
    # the CSR the forward pass consumes == stable sort of the reference COO by target
    dg = g._dev
    rowptr, src, eattr, _ = go.coo_to_in_csr(ref["edge_index"], ref["edge_attr"], N)
    assert np.array_equal(dg.rowptr.cpu().numpy().astype(np.int64), rowptr)
    E = ref["edge_index"].shape[1]
    assert np.array_equal(dg.csr_src.cpu().numpy()[:E].astype(np.int64), src)
    ce = dg.csr_eattr.cpu().numpy()[:E]
    assert np.array_equal(ce[:, :2], eattr[:, :2])
    assert np.all(np.abs(ce[:, 2].astype(np.float64) - eattr[:, 2]) <= ulp32(eattr[:, 2]))
    # the dense per-cell layout the tiled aggregation reads: slot k of cell i == CSR entry of the in-edge from i - off_k
    # (bit-identical to the CSR payload, zeros where the edge is absent)
    ge = getattr(dg, "grid_eattr", None)
    if ge is not None and N:
        ge = ge.cpu().numpy()                                   # [cells, K, ED]
        K = ge.shape[1]
        cell = dg.node_cell.cpu().numpy().astype(np.int64)
        nid = dg.node_id.cpu().numpy().reshape(ref["grid_shape"])
        H, W = nid.shape
        rr, cc = cell // W, cell % W
        present = np.zeros((N, K), bool)
        for k in range(K):
            dr, dc = int(dg.cfg.off_dr[k]), int(dg.cfg.off_dc[k])
            jr, jc = rr - dr, cc - dc
            ok = (jr >= 0) & (jr < H) & (jc >= 0) & (jc < W)
            present[ok, k] = nid[jr[ok], jc[ok]] >= 0
        gn = ge[cell]                                           # [N, K, ED] in node order
        assert np.array_equal(gn[present], ce)                  # row-major over (node, k) == CSR order
        assert not gn[~present].any()
    return (x[:, 3:7] != ref["x"][:, 3:7]).mean(), (ea[:, 2] != ref["edge_attr"][:, 2]).mean()


@pytest.mark.parametrize("name", [n for n in gio.graph_case_names(GZ) if n != "g16_empty"])
def test_graph_vs_reference_golden(name):
    c, ref = gio.graph_case(GZ, name)
    has_holes = c["mask"] is None or not c["mask"].all()
    _, g = build(c)
    check_graph(g, ref, c["depth"], has_holes)


def test_empty_graph():
    c, ref = gio.graph_case(GZ, "g16_empty")
    _, g = build(c)
    assert g.num_nodes == 0 and g.num_edges == 0
    assert not hasattr(g, "grid_shape")                     # reference quirk (:460-469)
    assert tuple(g.x.shape) == (0, 7) and tuple(g.edge_index.shape) == (2, 0)
    assert tuple(g.edge_attr.shape) == (0, 3) and g.local_std.numel() == 0


def test_c1_tile_vs_oracle():
    """BASELINE config 1: 512x512 reference-noise tile (oracle digests pinned in test_oracle_golden)."""
    import bathymetric_gnn_b200 as bg
    z = gio.load("c1_tile.npz")
    d = z["depth"]
    ref = go.build_graph(d, np.ones_like(d, bool), None, tuple(z["res"]))
    g = bg.GraphBuilder("8-connected").build_graph(d, np.ones_like(d, bool), None, tuple(z["res"]))
    fx, fs = check_graph(g, ref, d, has_holes=False)
    assert g.num_nodes == int(z["N"]) and g.num_edges == int(z["E"])
    assert fs < 1e-3           # fraction of slopes that differ (by 1 ulp) from NumPy's fp64 arctan


def test_strided_view_and_device_input():
    """A tile taken as a strided CUDA view of a survey gives the same graph as the copied tile."""
    import bathymetric_gnn_b200 as bg
    survey = synth.noisy_tile(96, 130, seed=3)
    mask = synth.blob_mask(96, 130, seed=3, invalid_frac=0.05)
    survey[~mask] = 1.0e6
    dev = torch.device("cuda")
    sd = torch.from_numpy(survey).to(dev)
    r0, c0, h, w = 10, 17, 64, 80
    b = bg.GraphBuilder()
    g1 = b.build_graph(sd[r0:r0 + h, c0:c0 + w], None, None, (0.5, 0.5), nodata=1.0e6)
    tile = survey[r0:r0 + h, c0:c0 + w].copy()
    ref = go.build_graph(tile, mask[r0:r0 + h, c0:c0 + w].copy(), None, (0.5, 0.5))
    check_graph(g1, ref, tile)


def test_ragged_batch_equals_per_grid():
    """One ragged device build == reference-style per-grid builds concatenated (Batch.from_data_list)."""
    import bathymetric_gnn_b200 as bg
    grids = synth.ragged_grids(25, seed=5, lo=2, hi=20, all_invalid_every=6)
    b = bg.GraphBuilder()
    g = b.build_graph_batch(grids, nodata=1.0e6, use_uncertainty=False)
    xs, eis, eas, lss, rows, cols, off = [], [], [], [], [], [], 0
    for d, u, res in grids:
        o = go.build_graph(d, (d != 1.0e6) & np.isfinite(d), None, res)
        if o["x"].shape[0] == 0:
            continue
        xs.append(o["x"]); eis.append(o["edge_index"] + off); eas.append(o["edge_attr"]); lss.append(o["local_std"])
        rows.append(o["valid_rows"]); cols.append(o["valid_cols"])
        off += o["x"].shape[0]
    X, EI, EA = np.concatenate(xs), np.concatenate(eis, 1), np.concatenate(eas)
    assert g.num_nodes == X.shape[0] and g.num_edges == EI.shape[1]
    x = g.x.cpu().numpy()
    assert np.array_equal(x[:, 0], X[:, 0])
    np.testing.assert_allclose(x, X, rtol=0, atol=2e-5)
    assert np.array_equal(g.valid_rows.cpu().numpy(), np.concatenate(rows))
    assert np.array_equal(g.valid_cols.cpu().numpy(), np.concatenate(cols))
    # CSR equals the in-edge CSR of the concatenated reference COO
    rowptr, src, eattr, _ = go.coo_to_in_csr(EI, EA, X.shape[0])
    dg = g._dev
    assert np.array_equal(dg.rowptr.cpu().numpy().astype(np.int64), rowptr)
    assert np.array_equal(dg.csr_src.cpu().numpy()[:len(src)].astype(np.int64), src)
    assert np.array_equal(dg.csr_eattr.cpu().numpy()[:len(src), :2], eattr[:, :2])


def test_too_small_grid_raises_like_reference():
    import bathymetric_gnn_b200 as bg
    with pytest.raises(ValueError):
        bg.GraphBuilder().build_graph(np.ones((1, 5), np.float32))
    with pytest.raises(ValueError):
        bg.GraphBuilder("16-connected")


def test_graph_to_grid():
    import bathymetric_gnn_b200 as bg
    c, ref = gio.graph_case(GZ, "g64_holes")
    b, g = build(c)
    vals = torch.arange(g.num_nodes, dtype=torch.float32)
    out = b.graph_to_grid(g.cpu(), vals, fill_value=0.0)
    exp = go.graph_to_grid(ref["grid_shape"], ref["valid_rows"], ref["valid_cols"], vals.numpy(), 0.0)
    assert out.dtype == np.float32 and np.array_equal(out, exp)
    out = b.graph_to_grid(g, vals.cuda())
    exp = go.graph_to_grid(ref["grid_shape"], ref["valid_rows"], ref["valid_cols"], vals.numpy())
    assert np.array_equal(out, exp, equal_nan=True)
    with pytest.raises(ValueError):
        b.graph_to_grid(g, torch.zeros(g.num_nodes, 2))
    with pytest.raises(ValueError):
        b.graph_to_grid(bg.Data(x=g.x), vals)


def test_determinism():
    c, _ = gio.graph_case(GZ, "g64_holes_unc")
    _, g1 = build(c)
    _, g2 = build(c)
    for k in ("x", "edge_index", "edge_attr", "local_std"):
        assert torch.equal(getattr(g1, k), getattr(g2, k))


def test_packed_ragged_entry_equals_batch():
    """build_graph_packed (device-resident packed buffer + descriptor table) == build_graph_batch (list of host grids)."""
    import bathymetric_gnn_b200 as bg
    rng = np.random.default_rng(11)
    grids = []
    for _ in range(40):
        r, c = int(rng.integers(3, 20)), int(rng.integers(3, 20))
        d = (30 + rng.normal(0, 1, (r, c))).astype(np.float32)
        d[rng.random((r, c)) < 0.1] = 1.0e6
        grids.append((d, None, (float(rng.choice([0.5, 1.0, 2.0])),) * 2))
    b = bg.GraphBuilder("8-connected")
    ga = b.build_graph_batch(grids, nodata=1.0e6)
    rows = np.array([g[0].shape[0] for g in grids], np.int32); cols = np.array([g[0].shape[1] for g in grids], np.int32)
    off = np.zeros(len(grids) + 1, np.int64); np.cumsum(rows.astype(np.int64) * cols, out=off[1:])
    depth = torch.from_numpy(np.concatenate([g[0].ravel() for g in grids])).cuda()
    res = np.array([g[2] for g in grids], np.float64)
    gb = b.build_graph_packed(depth, off, rows, cols, res, nodata=1.0e6)
    assert ga.num_nodes == gb.num_nodes and ga.num_edges == gb.num_edges
    for k in ("x", "edge_index", "edge_attr", "local_std", "valid_rows", "valid_cols"):
        assert torch.equal(getattr(ga, k), getattr(gb, k)), k


def test_multiscale_graph_builder_vs_reference_golden():
    """MultiScaleGraphBuilder drop-in (graph_construction.py:508-607): host block reduction with the reference's NumPy calls,
    device graph build per scale -- against the golden Data fields of the unmodified reference (same tolerances as a tile)."""
    import warnings
    import bathymetric_gnn_b200 as bg
    z = gio.load("multiscale_case.npz")
    d, m, u = synth.multiscale_case()
    b = bg.MultiScaleGraphBuilder(scales=[4, 1, 2], connectivity="8-connected")
    assert b.scales == [1, 2, 4]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        graphs = b.build_multiscale_graph(d, m, u, (0.5, 1.0))
    assert sorted(graphs) == [1, 2, 4]
    for s, g in graphs.items():
        ref = {k: z[f"s{s}/{k}"] for k in ("x", "edge_index", "edge_attr", "pos", "valid_rows", "valid_cols", "local_std")}
        ref["grid_shape"], ref["num_valid_cells"] = tuple(z[f"s{s}/grid_shape"]), z[f"s{s}/num_valid_cells"]
        check_graph(g, ref, None, has_holes=True)


def test_lazy_arrays_refuse_a_survey_edited_after_the_build():
    """build_graph consumes a CUDA depth plane in place and emits edge_index / edge_attr / the CSR payload lazily from it: an
    in-place edit in between must be refused instead of returning features inconsistent with x / rowptr."""
    import bathymetric_gnn_b200 as bg
    d = torch.from_numpy(synth.noisy_tile(40, 50, seed=2)).cuda()
    g = bg.GraphBuilder().build_graph(d, None, None, (1.0, 1.0))
    ea0 = bg.GraphBuilder().build_graph(d.clone(), None, None, (1.0, 1.0)).edge_attr.clone()
    d[3, 4] += 100.0
    with pytest.raises(RuntimeError, match="modified in place"):
        g.edge_attr
    with pytest.raises(RuntimeError, match="modified in place"):
        g._dev.csr_src
    g2 = bg.GraphBuilder().build_graph(d, None, None, (1.0, 1.0))
    assert torch.equal(g2.edge_index, bg.GraphBuilder().build_graph(d.clone(), None, None, (1.0, 1.0)).edge_index)
    assert not torch.equal(g2.edge_attr, ea0)


@pytest.mark.parametrize("seed", range(24))
def test_random_small_grids_vs_oracle(seed):
    """Seeded sweep over the shapes / masks / values the golden cases do not enumerate: 2 ... 40-sided grids, scattered and
    blocky holes, masked-out NaN / +-inf / nodata cells, anisotropic resolutions, 4- / 8- / 24-offset stencils, with and
    without the uncertainty column -- device graph (eager AND deferred build) vs the oracle (pinned bit-for-bit to the
    reference by tests/test_oracle_golden.py)."""
    import bathymetric_gnn_b200 as bg
    rng = np.random.default_rng(1000 + seed)
    H, W = int(rng.integers(2, 41)), int(rng.integers(2, 41))
    d = (20.0 + rng.normal(0, 1.0, (H, W)) + 0.05 * np.arange(W)[None, :]).astype(np.float32)
    mask = rng.random((H, W)) > rng.choice([0.0, 0.05, 0.3, 0.7])
    if rng.random() < 0.5 and H > 4 and W > 4:
        r, c = int(rng.integers(0, H - 2)), int(rng.integers(0, W - 2))
        mask[r:r + int(rng.integers(1, 4)), c:c + int(rng.integers(1, 4))] = False
    if not mask.any():
        mask[0, 0] = True
    bad = ~mask
    d[bad] = rng.choice([np.nan, np.inf, -np.inf, 1.0e6], size=int(bad.sum())).astype(np.float32)   # never leaks: masked out
    conn = ["8-connected", "4-connected", "24-connected"][seed % 3]
    res = [(1.0, 1.0), (0.5, 0.5), (2.0, 0.5), (4.0, 4.0)][seed % 4]
    unc = (0.1 + 0.05 * rng.random((H, W))).astype(np.float32) if seed % 2 else None
    ref = go.build_graph(d, mask, unc, res, connectivity=conn)
    b = bg.GraphBuilder(conn)
    g_def = b.build_graph(d, mask, unc, res)                                   # NumPy input, <= 4096 cells: deferred build
    assert isinstance(g_def, bg.HostGridGraph) and g_def.num_nodes == int(mask.sum())
    check_graph(g_def, ref, d, has_holes=not mask.all())
    g_now = b.build_graph(torch.from_numpy(d).cuda(), torch.from_numpy(mask).cuda(), None if unc is None else torch.from_numpy(unc).cuda(), res)
    assert not isinstance(g_now, bg.HostGridGraph)
    check_graph(g_now, ref, d, has_holes=not mask.all())
    for k in ("x", "edge_index", "edge_attr", "local_std"):
        assert torch.equal(getattr(g_def, k), getattr(g_now, k)), k


def test_edge_total_guard_refuses_a_cell_space_beyond_int32():
    """node / edge ids are int32; a cell space whose edge count reaches 2^31 (268 M valid cells at K = 8,
    89 M at K = 24) must be REFUSED by bgnn_graph_count (ValueError) instead of returning a wrapped total that would size the
    CSR arrays wrongly.  9 600 x 9 600 valid cells with the 24-offset stencil: 2.2e9 edges."""
    import bathymetric_gnn_b200 as bg
    side = 9600
    d = torch.zeros((side, side), dtype=torch.float32, device="cuda")
    with pytest.raises(ValueError, match="does not fit int32"):
        bg.GraphBuilder("24-connected").build_graph(d, None, None, (1.0, 1.0))
    g = bg.GraphBuilder("8-connected").build_graph(d[:3000, :3000], None, None, (1.0, 1.0))      # the same plane, in range: fine
    assert g.num_nodes == 9_000_000 and g.num_edges == 4 * 2999 * (2 * 3000 - 1)
