"""VR / SR BAG containers -> ragged descriptor.  CPU: the oracle restatement and the product's
host logic (bathymetric_gnn_b200.vr_bag) against the fixture produced by the UNMODIFIED reference handler / writer on
the h5py stand-in (tests/golden/vr_bag_case.npz); GPU: the descriptor + packed buffers feed build_graph_packed."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import vr_bag_oracle as vo
from tests import fake_h5py
from tests import golden_io as gio

Z = gio.load("vr_bag_case.npz")
CASES = ["shuffled", "ordered"]


def _kw(case):
    return ast.literal_eval(str(Z[f"{case}/kw"]))


def _corrections(grids, depth_of, unc_of, valid_of, shape_of):
    """The seeded corrections make_golden.py pushed through the reference writer."""
    rng = np.random.default_rng(17)
    out = []
    for g in grids:
        cd, cu = depth_of(g).copy(), unc_of(g).copy()
        m = valid_of(g) & (rng.random(shape_of(g)) < 0.3)
        cd[m] -= rng.normal(0, 0.5, int(m.sum())).astype(np.float32)
        cu[m] *= np.float32(1.25)
        out.append((cd, cu))
    return out


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("mvr", [0.0, 0.5])
def test_oracle_iterate_bit_exact_vs_reference(case, mvr):
    grids = list(vo.iterate_refinements(Z[f"{case}/meta"], Z[f"{case}/ref"], mvr))
    tag = f"{case}/mvr{mvr}"
    desc = np.array([[g["base_row"], g["base_col"], g["dimensions"][0], g["dimensions"][1], g["start_index"]] for g in grids], np.int64)
    assert np.array_equal(desc, Z[tag + "/desc"])
    assert np.array_equal(np.array([g["resolution"] for g in grids]), Z[tag + "/res"])
    assert np.array_equal(np.array([g["sw_corner"] for g in grids]), Z[tag + "/sw"])
    assert np.array_equal(np.concatenate([g["depth"].ravel() for g in grids]), Z[tag + "/depth"], equal_nan=True)
    assert np.array_equal(np.concatenate([g["uncertainty"].ravel() for g in grids]), Z[tag + "/unc"])
    assert np.array_equal(np.array([g["num_valid"] for g in grids]), Z[tag + "/num_valid"])


@pytest.mark.parametrize("case", CASES)
def test_oracle_write_back_bit_exact_vs_reference(case):
    ref = Z[f"{case}/ref"].copy()
    grids = list(vo.iterate_refinements(Z[f"{case}/meta"], ref, 0.0))
    corr = _corrections(grids, lambda g: g["depth"], lambda g: g["uncertainty"], lambda g: vo.valid_mask(g["depth"]),
                        lambda g: g["depth"].shape)
    n = sum(vo.update_refinement_batch(ref, g, cd, cu) for g, (cd, cu) in zip(grids, corr))
    assert n == int(Z[f"{case}/corrections_applied"])
    assert ref.tobytes() == Z[f"{case}/ref_after"].tobytes()


@pytest.fixture()
def h5(tmp_path):
    import sys
    saved = sys.modules.get("h5py")
    fake_h5py.install()
    yield tmp_path
    if saved is None:
        sys.modules.pop("h5py", None)
    else:
        sys.modules["h5py"] = saved


@pytest.mark.parametrize("case", CASES)
def test_product_handler_table_and_iteration_vs_reference(case, h5):
    from bathymetric_gnn_b200 import vr_bag as vb
    src = os.path.join(h5, case + ".bag")
    meta, ref = fake_h5py.write_vr_bag(src, **_kw(case))
    assert meta.tobytes() == Z[f"{case}/meta"].tobytes() and ref.tobytes() == Z[f"{case}/ref"].tobytes()
    assert vb.detect_bag_type(src) == str(Z[f"{case}/type"]) == "VR"
    h = vb.VRBagHandler(src)
    info = h.get_refinement_info()
    assert [info["num_refined_cells"], info["total_refinement_nodes"]] == Z[f"{case}/info"].tolist()
    assert np.array_equal(np.array(info["unique_resolutions"], np.float64), Z[f"{case}/unique_resolutions"])
    assert h.num_refinement_cells == info["num_refined_cells"] and h.geotransform is None
    for mvr in (0.0, 0.5):
        tag = f"{case}/mvr{mvr}"
        tab = h.refinement_table(mvr)
        desc = np.stack([tab.base_row, tab.base_col, tab.rows, tab.cols, tab.start_index], axis=1).astype(np.int64)
        assert np.array_equal(desc, Z[tag + "/desc"])
        assert np.array_equal(tab.res, Z[tag + "/res"]) and np.array_equal(tab.sw_corner, Z[tag + "/sw"])
        assert tab.depth.dtype == np.float32 and np.array_equal(tab.depth, Z[tag + "/depth"], equal_nan=True)
        assert np.array_equal(tab.uncertainty, Z[tag + "/unc"])
        assert np.array_equal(np.diff(tab.cell_off), tab.rows.astype(np.int64) * tab.cols)
        # the generator yields the reference's RefinementGrid objects in the reference's order
        grids = list(h.iterate_refinements(mvr))
        assert len(grids) == len(tab)
        assert np.array_equal(np.array([g.num_valid for g in grids]), Z[tag + "/num_valid"])
        for i, g in enumerate(grids):
            assert (g.base_row, g.base_col, g.dimensions, g.start_index) == (desc[i, 0], desc[i, 1], (desc[i, 2], desc[i, 3]), desc[i, 4])
            assert g.shape == g.dimensions and g.depth.flags.owndata or g.depth.base is not tab.depth
        # chunking by cell budget covers every grid exactly once
        parts = list(tab.chunks(400))
        assert sum(len(p) for p in parts) == len(tab)
        assert np.array_equal(np.concatenate([p.depth for p in parts]), tab.depth, equal_nan=True)
        assert all(p.n_cells <= 400 or len(p) == 1 for p in parts)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("bulk", [False, True])
def test_product_writer_bit_exact_vs_reference(case, bulk, h5):
    from bathymetric_gnn_b200 import vr_bag as vb
    src, dst = os.path.join(h5, case + ".bag"), os.path.join(h5, case + "_out.bag")
    fake_h5py.write_vr_bag(src, **_kw(case))
    h = vb.VRBagHandler(src)
    w = h.copy_and_open_for_writing(dst)
    grids = list(h.iterate_refinements(0.0))
    corr = _corrections(grids, lambda g: g.depth, lambda g: g.uncertainty, lambda g: g.valid_mask, lambda g: g.shape)
    if bulk:        # one read-modify-write for the whole table (two chunks, to exercise partial ranges)
        tab = h.refinement_table(0.0)
        cd = np.concatenate([c[0].ravel() for c in corr])
        cu = np.concatenate([c[1].ravel() for c in corr])
        half = len(tab) // 2
        for g0, g1 in ((0, half), (half, len(tab))):
            part = tab.slice(g0, g1)
            a, b = int(tab.cell_off[g0]), int(tab.cell_off[g1])
            w.update_refinements_packed(part, torch.from_numpy(cd[a:b]), cu[a:b])
        with pytest.raises(ValueError):
            w.update_refinements_packed(tab, cd[:-1])
    else:
        for g, (cd, cu) in zip(grids, corr):
            w.update_refinement_batch(g, cd, cu)
        with pytest.raises(ValueError):
            w.update_refinement_batch(grids[0], np.zeros((1, 1), np.float32))
    assert w._corrections_applied == int(Z[f"{case}/corrections_applied"])
    w.close()
    with fake_h5py.File(dst, "r") as f:
        assert f["BAG_root"]["varres_refinements"][:].tobytes() == Z[f"{case}/ref_after"].tobytes()
    with fake_h5py.File(src, "r") as f:        # the source container is untouched
        assert f["BAG_root"]["varres_refinements"][:].tobytes() == Z[f"{case}/ref"].tobytes()


def test_product_sr_handler_and_writer_vs_reference(h5):
    from bathymetric_gnn_b200 import vr_bag as vb
    src, dst = os.path.join(h5, "sr.bag"), os.path.join(h5, "sr_out.bag")
    fake_h5py.write_sr_bag(src)
    assert vb.detect_bag_type(src) == str(Z["sr/type"]) == "SR"
    h = vb.SRBagHandler(src)
    g = next(iter(h.iterate_refinements(0.0)))
    assert np.array_equal(g.depth, Z["sr/depth"]) and np.array_equal(g.uncertainty, Z["sr/unc"])
    assert np.array_equal(np.array(g.resolution), Z["sr/res"])
    assert h.get_refinement_info()["total_refinement_nodes"] == int(Z["sr/info"][0])
    assert len(list(h.iterate_refinements(0.95))) == int(Z["sr/n_grids_mvr0.95"]) == 0
    tab = h.refinement_table(0.0)
    assert len(tab) == 1 and np.array_equal(tab.depth.reshape(g.shape), g.depth) and len(h.refinement_table(0.95)) == 0
    w = h.copy_and_open_for_writing(dst)
    w.update_refinements_packed(tab, (g.depth + np.float32(1.0)).ravel(), (g.uncertainty * np.float32(2.0)).ravel())
    w.close()
    with fake_h5py.File(dst, "r") as f:
        assert np.array_equal(f["BAG_root"]["elevation"][:], Z["sr/elev_after"])
        assert np.array_equal(f["BAG_root"]["uncertainty"][:], Z["sr/unc_after"])


def test_not_a_bag_and_missing_backend(h5, monkeypatch):
    from bathymetric_gnn_b200 import vr_bag as vb
    p = os.path.join(h5, "x.bag")
    f = fake_h5py.File(p, "w"); f.create_dataset("other/thing", np.zeros(3)); f.close()
    with pytest.raises(ValueError):
        vb.VRBagHandler(p)
    src = os.path.join(h5, "sr.bag")
    fake_h5py.write_sr_bag(src)
    with pytest.raises(ValueError):               # an SR BAG is not a VR BAG (vr_bag.py:201-213)
        vb.VRBagHandler(src)
    import sys
    monkeypatch.setitem(sys.modules, "h5py", None)      # import h5py -> ImportError
    with pytest.raises(ImportError):
        vb.detect_bag_type(src)


@pytest.mark.gpu
def test_table_feeds_packed_graph_build(h5):
    """descriptor + packed buffers -> build_graph_packed == per-grid build_graph of the yielded grids (bit-exact)."""
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200 import vr_bag as vb
    src = os.path.join(h5, "vr.bag")
    fake_h5py.write_vr_bag(src, base_shape=(9, 11), seed=11)
    h = vb.VRBagHandler(src)
    tab = h.refinement_table(0.05)
    b = bg.GraphBuilder("8-connected")
    d, u = tab.to_device()
    g = b.build_graph_packed(d, tab.cell_off, tab.rows, tab.cols, tab.res, uncertainty=u, nodata=vb.NODATA)
    xs, ls, ne = [], [], 0
    for grid in h.iterate_refinements(0.05):
        if grid.num_valid == 0:
            continue
        gi = b.build_graph(grid.depth, grid.valid_mask, grid.uncertainty, grid.resolution)
        xs.append(gi.x); ls.append(gi.local_std); ne += gi.num_edges
    assert g.num_edges == ne
    assert torch.equal(g.x, torch.cat(xs)) and torch.equal(g.local_std, torch.cat(ls))
