"""GPU parity of the native VR-BAG post-processing through the C ABI: bgnn_vr_apply_results vs the
restated `apply_results` closure (scripts/inference_native.py:480-503) and bgnn_sidecar_add vs the golden planes painted by
the UNMODIFIED reference SidecarBuilder.add_refinement_results (data/vr_bag.py:695-778).  Everything here is bit-exact
(fp32 elementwise arithmetic compiled -fmad=false; integer placement computed on the host like the reference)."""
import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from oracle import vr_post_oracle as vo
from tests import golden_io as gio
from tests import synth
from tests.test_oracle_golden import _Grid

pytestmark = pytest.mark.gpu


def _packed(grids, results):
    off = np.zeros(len(grids) + 1, np.int64)
    np.cumsum([g["depth"].size for g in grids], out=off[1:])
    cat = lambda arrs: torch.from_numpy(np.concatenate([np.asarray(a, np.float32).ravel() for a in arrs])).cuda()
    return (off, cat([g["depth"] for g in grids]), cat([g["uncertainty"] for g in grids]),
            cat([r[0] for r in results]), cat([r[1] for r in results]), cat([r[2] for r in results]))


def test_sidecar_batch_bit_exact_vs_reference_golden():
    import bathymetric_gnn_b200 as bg
    z = gio.load("sidecar_case.npz")
    geo, grids, results = synth.sidecar_scene()
    off, d, u, cl, cf, cr = _packed(grids, results)
    # one batch, and the same scene in three batches + one single-grid call: "later grids overwrite" must hold across calls
    for cuts in ([0, len(grids)], [0, 7, 8, 31, len(grids)]):
        sb = bg.SidecarBuilder(**geo)
        for a, b in zip(cuts[:-1], cuts[1:]):
            gs = [_Grid(**g) for g in grids[a:b]]
            if b - a == 1:
                sb.add_refinement_results(gs[0], *results[a])
            else:
                sl = slice(int(off[a]), int(off[b]))
                sb.add_batch(gs, off[a:b + 1] - off[a], d[sl], cl[sl], cf[sl], cr[sl])
        for k in ("classification", "confidence", "correction", "valid_mask"):
            assert np.array_equal(getattr(sb, k), z[k], equal_nan=True), (k, cuts)


def test_sidecar_errors_and_empty():
    import bathymetric_gnn_b200 as bg
    geo, grids, results = synth.sidecar_scene()
    with pytest.raises(RuntimeError):
        bg.SidecarBuilder()                                     # no georeferencing (the GDAL lookup is not built)
    sb = bg.SidecarBuilder(**geo)
    sb.add_batch([], np.zeros(1, np.int64), torch.zeros(0).cuda(), torch.zeros(0).cuda(), torch.zeros(0).cuda(), torch.zeros(0).cuda())
    assert np.isnan(sb.classification).all() and (sb.valid_mask == 0).all()
    g = _Grid(**grids[0])
    with pytest.raises(ValueError):
        sb.add_batch([g], np.array([0, 3], np.int64), torch.zeros(3).cuda(), torch.zeros(3).cuda(), torch.zeros(3).cuda(), torch.zeros(3).cuda())


def test_vr_apply_results_bit_exact_vs_oracle():
    import bathymetric_gnn_b200 as bg
    _, grids, results = synth.sidecar_scene(seed=9, n=80)
    off, d, u, cl, cf, cr = _packed(grids, results)
    od, ou, st = bg.apply_results_packed(d, u, cl, cf, cr, 0.85)
    od, ou = od.cpu().numpy(), ou.cpu().numpy()
    tot = dict(cells_corrected=0, cells_classified_noise=0, cells_processed=0, total_confidence=0.0)
    for i, (g, (c0, c1, c2)) in enumerate(zip(grids, results)):
        rd, ru, rs = vo.apply_results(g["depth"], g["uncertainty"], c0, c1, c2, 0.85)
        sl = slice(int(off[i]), int(off[i + 1]))
        assert np.array_equal(od[sl].reshape(rd.shape), rd, equal_nan=True)
        assert np.array_equal(ou[sl].reshape(ru.shape), ru)
        for k in tot:
            tot[k] += rs[k]
    assert st["cells_corrected"] == tot["cells_corrected"] > 100
    assert st["cells_classified_noise"] == tot["cells_classified_noise"] and st["cells_processed"] == tot["cells_processed"]
    assert abs(st["total_confidence"] - tot["total_confidence"]) <= 1e-6 * tot["total_confidence"]
    # the host (per-grid NumPy) variant shipped for the HDF5 write-back path agrees too
    g, (c0, c1, c2) = grids[3], results[3]
    hd, hu, n = bg.apply_results(g["depth"], g["uncertainty"], vo.valid_mask(g["depth"]), c0, c1, c2, 0.85)
    rd, ru, rs = vo.apply_results(g["depth"], g["uncertainty"], c0, c1, c2, 0.85)
    assert np.array_equal(hd, rd, equal_nan=True) and np.array_equal(hu, ru) and n == rs["cells_corrected"]
    # depth only (no uncertainty plane)
    od2, ou2, _ = bg.apply_results_packed(d, None, cl, cf, cr, 0.85)
    assert ou2 is None and np.array_equal(od2.cpu().numpy(), od, equal_nan=True)


def test_native_processor_packed_flush_matches_per_grid_results():
    """flush_batch_packed (results stay on the device) == flush_batch (per-grid host arrays), then apply + sidecar on the
    packed planes == the per-grid oracle post-processing of those results."""
    import bathymetric_gnn_b200 as bg
    from tests.test_gpu_gnn import make_model
    geo, grids, _ = synth.sidecar_scene(seed=4, n=30)
    cfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(cfg, seed=2, randomize_bn=True, head_gain=8.0)
    model = make_model(cfg, sd)
    procs = [bg.NativeVRProcessor(model, bg.GraphBuilder("8-connected"), None, 0.85) for _ in range(2)]
    kept = []
    for i, g in enumerate(grids):
        r0 = procs[0].add_to_batch(g["depth"], g["uncertainty"], g["resolution"], 1.0e6)
        r1 = procs[1].add_to_batch(g["depth"], g["uncertainty"], g["resolution"], 1.0e6)
        assert (r0 is None) == (r1 is None)
        if r0 is None:
            kept.append(i)
    host = procs[0].flush_batch()
    pk = procs[1].flush_batch_packed()
    assert len(host) == len(kept) == len(pk["shapes"])
    sb, so = bg.SidecarBuilder(**geo), vo.SidecarOracle(**geo)
    od, ou, st = bg.apply_results_packed(pk["depth"], pk["uncertainty"], pk["classification"], pk["confidence"], pk["correction"], 0.85)
    sb.add_batch([_Grid(**grids[i]) for i in kept], pk["cell_off"], pk["depth"], pk["classification"], pk["confidence"], pk["correction"])
    od, ou = od.cpu().numpy(), ou.cpu().numpy()
    n_corr = 0
    for j, i in enumerate(kept):
        sl = slice(int(pk["cell_off"][j]), int(pk["cell_off"][j + 1]))
        for k, name in enumerate(("classification", "confidence", "correction")):
            assert np.array_equal(pk[name][sl].cpu().numpy().reshape(host[j][k].shape), host[j][k]), (name, i)
        g = grids[i]
        rd, ru, rs = vo.apply_results(g["depth"], g["uncertainty"], *host[j], 0.85)
        assert np.array_equal(od[sl].reshape(rd.shape), rd, equal_nan=True) and np.array_equal(ou[sl].reshape(ru.shape), ru)
        n_corr += rs["cells_corrected"]
        so.add_refinement_results(_Grid(**g), *host[j])
    assert st["cells_corrected"] == n_corr
    for k in ("classification", "confidence", "correction", "valid_mask"):
        assert np.array_equal(getattr(sb, k), getattr(so, k), equal_nan=True), k
