"""The drop-in boundary: reference-layout entry scripts run on the B200 packages.

The reference scripts prepend THEIR checkout root to sys.path before importing `config` / `data` / `models`
(scripts/inference.py:22-25, scripts/inference_native.py:23-36, scripts/test_pipeline.py:30), so the overlay is
installed through sys.modules (`bathymetric_gnn_b200.dropin.install`).  CPU tests: the overlay wins over a checkout's
own packages, for the fixture layout (decoy packages that raise) and -- in the build container -- for the module level
of the UNMODIFIED reference scripts.  GPU tests: the scripts run end to end and reproduce the outputs of the unmodified
reference (pipeline_case.npz = reference BathymetricPipeline.process; native_script_case.npz = the unmodified
scripts/inference_native.py run as __main__ on the reference's own packages)."""
import ast
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import model_oracle as mo
from tests import fake_h5py
from tests import golden_io as gio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYOUT = os.path.join(ROOT, "tests", "fixtures", "ref_layout")
REFERENCE = "/root/reference"


def _probe(script, layout_root):
    """Run the MODULE LEVEL of an entry script (its sys.path.insert + imports) through the overlay in a fresh
    interpreter and report where the names it imported come from."""
    code = f"""
import json, sys
sys.path.insert(0, {ROOT!r})
from bathymetric_gnn_b200 import dropin
g = dropin.run_script({os.path.join(layout_root, 'scripts', script)!r}, run_name="module_level_only")
names = ["Config", "BathymetricPipeline", "BathymetricGNN", "GraphBuilder", "VRBagHandler", "VRBagWriter", "SRBagHandler",
         "SRBagWriter", "SidecarBuilder", "detect_bag_type", "PyGBatch"]
out = {{n: getattr(g[n], "__module__", None) for n in names if n in g}}
out["PYG_BATCH_AVAILABLE"] = g.get("PYG_BATCH_AVAILABLE")
out["CORRECTION_NORM_FLOOR"] = g.get("CORRECTION_NORM_FLOOR")
out["path0"] = sys.path[0]
out["pkg_files"] = {{n: getattr(sys.modules.get(n), "__file__", None) for n in ("config", "data", "models", "torch_geometric")}}
print("PROBE" + json.dumps(out))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("PROBE")][-1][5:])


def _check_probe(p, layout_root):
    assert os.path.samefile(p["path0"], layout_root)          # the script DID put its own checkout first
    for pkg, f in p["pkg_files"].items():
        if f is not None:
            assert os.path.join("bathymetric-gnn_b200", "dropin", pkg) in f, (pkg, f)
    for name, mod in p.items():
        if name in ("path0", "pkg_files", "PYG_BATCH_AVAILABLE", "CORRECTION_NORM_FLOOR"):
            continue
        assert mod is not None and mod.startswith("bathymetric_gnn_b200"), (name, mod)


@pytest.mark.parametrize("script", ["inference.py", "inference_native.py"])
def test_overlay_wins_over_checkout_packages_fixture_layout(script):
    p = _probe(script, LAYOUT)
    _check_probe(p, LAYOUT)
    if script == "inference_native.py":
        assert p["PYG_BATCH_AVAILABLE"] is True and p["CORRECTION_NORM_FLOOR"] == 0.01
        assert {"VRBagHandler", "SidecarBuilder", "detect_bag_type", "PyGBatch"} <= set(p)


def test_without_overlay_the_checkout_packages_are_imported():
    """Control: the same script WITHOUT install() picks the checkout's own packages up (here: the decoys raise)."""
    r = subprocess.run([sys.executable, os.path.join(LAYOUT, "scripts", "inference.py"), "--help"], capture_output=True, text=True,
                       env={**os.environ, "PYTHONPATH": os.path.join(ROOT, "bathymetric-gnn_b200", "dropin")}, timeout=600)
    assert r.returncode != 0 and "did not take effect" in r.stderr


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "scripts")), reason="build container only (/root/reference)")
@pytest.mark.parametrize("script", ["inference.py", "inference_native.py", "test_pipeline.py"])
def test_overlay_wins_for_the_unmodified_reference_scripts(script):
    p = _probe(script, REFERENCE)
    _check_probe(p, REFERENCE)
    if script == "inference_native.py":
        assert p["PYG_BATCH_AVAILABLE"] is True
        assert {"BathymetricGNN", "GraphBuilder", "VRBagHandler", "VRBagWriter", "SRBagHandler", "SRBagWriter",
                "SidecarBuilder", "detect_bag_type"} <= set(p)


def test_dropin_exports_every_name_of_the_reference_packages():
    """config/__init__.py:1-22, data/__init__.py:1-33, models/__init__.py:1-19 export lists (fresh interpreter)."""
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
from bathymetric_gnn_b200 import dropin
dropin.install()
import config, data, models
want = dict(config=["Config", "TileConfig", "GraphConfig", "ModelConfig", "TrainingConfig", "SyntheticNoiseConfig", "InferenceConfig",
                    "CORRECTION_NORM_FLOOR", "CORRECTION_NORM_CAP"],
            data=["BathymetricGrid", "BathymetricLoader", "BathymetricWriter", "Tile", "TileSpec", "TileManager", "TileMerger",
                  "GraphBuilder", "MultiScaleGraphBuilder", "VRBagHandler", "VRBagWriter", "SRBagHandler", "SRBagWriter",
                  "RefinementGrid", "SidecarBuilder", "process_vr_bag_native", "detect_bag_type"],
            models=["BathymetricGNN", "LocalFeatureExtractor", "GNNBackbone", "ClassificationHead", "ConfidenceHead",
                    "CorrectionHead", "BathymetricPipeline"])
for pkg, names in want.items():
    for n in names:
        assert hasattr(sys.modules[pkg], n), (pkg, n)
    assert set(names) <= set(sys.modules[pkg].__all__), pkg
try:
    data.SyntheticNoiseGenerator
    raise SystemExit("expected ImportError without a reference checkout")
except ImportError as e:
    assert "BGNN_REFERENCE_ROOT" in str(e)
import torch_geometric
from torch_geometric.data import Data, Batch
from torch_geometric.nn import GATConv, GCNConv, SAGEConv, GINConv, BatchNorm, global_mean_pool
assert dropin.installed()
dropin.uninstall()
assert "config" not in sys.modules and not dropin.installed()
print("OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env={k: v for k, v in os.environ.items() if k != "BGNN_REFERENCE_ROOT"})
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "data")), reason="build container only (/root/reference)")
def test_noise_generator_passes_through_from_the_checkout():
    code = f"""
import sys
sys.path.insert(0, {ROOT!r})
from bathymetric_gnn_b200 import dropin
dropin.install(reference_root={REFERENCE!r})
from data import SyntheticNoiseGenerator
import numpy as np
lab = SyntheticNoiseGenerator(seed=42).generate(np.full((32, 32), 20.0, np.float32), np.ones((32, 32), bool), 1.0)
assert lab.noisy_depth.shape == (32, 32) and "synthetic_noise.py" in SyntheticNoiseGenerator.generate.__code__.co_filename
print("OK")
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


# ------------------------------------------------------------------------------------------------ GPU: end to end
@pytest.fixture()
def overlay():
    """Run a script in-process on the overlay, then restore sys.modules / sys.path (other tests import the oracle's own
    `torch_geometric` restatement)."""
    saved_mods = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("config", "data", "models", "torch_geometric", "h5py")}
    saved_path = list(sys.path)
    from bathymetric_gnn_b200 import dropin
    yield dropin
    dropin.uninstall()
    for k in [k for k in sys.modules if k.split(".")[0] in ("config", "data", "models", "torch_geometric", "h5py")]:
        del sys.modules[k]
    sys.modules.update(saved_mods)
    sys.path[:] = saved_path


def _state_dict(seed, cls_bias, conf_bias):
    cfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(cfg, seed=seed, randomize_bn=True, head_gain=8.0)
    sd["classification_head.mlp.3.bias"] = torch.from_numpy(cls_bias)
    sd["confidence_head.mlp.3.bias"] = torch.from_numpy(conf_bias)
    return sd


@pytest.mark.gpu
def test_inference_script_end_to_end_vs_reference_pipeline(tmp_path, overlay):
    """scripts/inference.py layout -> BathymetricPipeline(config, vr_bag_mode).load_model().process(), config.yaml next to
    the model, CLI overrides; result grids vs the unmodified reference pipeline (pipeline_case.npz)."""
    import bathymetric_gnn_b200 as bg
    z = gio.load("pipeline_case.npz")
    H, W, ts, ov = (int(v) for v in z["cfg"])
    cfg = bg.Config()
    cfg.tile.overlap = ov
    cfg.save(tmp_path / "config.yaml")
    torch.save({"model_state_dict": _state_dict(int(z["seed"]), z["cls_bias"], z["conf_bias"]), "in_channels": 7, "edge_dim": 3},
               tmp_path / "model.pt")
    np.savez(tmp_path / "in.npz", depth=z["depth"], nodata=z["nodata"], resolution=z["res"])
    g = overlay.run_script(os.path.join(LAYOUT, "scripts", "inference.py"),
                           ["--input", str(tmp_path / "in.npz"), "--model", str(tmp_path / "model.pt"),
                            "--output", str(tmp_path / "o" / "out.npz"), "--tile-size", str(ts),
                            "--min-valid-ratio", str(float(z["min_valid_ratio"])), "--log-level", "ERROR"])
    assert g["BathymetricPipeline"] is bg.BathymetricPipeline and g["Config"] is bg.Config
    res = g["RESULTS"]
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref/")}
    same = (res["classification"] == ref["classification"]) | (np.isnan(res["classification"]) & np.isnan(ref["classification"]))
    assert same.mean() > 0.998
    np.testing.assert_allclose(res["confidence"], ref["confidence"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(res["correction"][same], ref["correction"][same], rtol=1e-4, atol=2e-4)
    ok = same & (np.abs(ref["confidence"] - 0.85) > 1e-4)
    np.testing.assert_allclose(res["cleaned_depth"][ok], ref["cleaned_depth"][ok], rtol=1e-6, atol=2e-4)
    out = np.load(tmp_path / "o" / "out.npz")
    assert np.array_equal(out["depth"], res["cleaned_depth"]) and "classification" in out.files


@pytest.mark.gpu
def test_inference_native_script_end_to_end_vs_unmodified_reference_script(tmp_path, overlay):
    """scripts/inference_native.py layout on a VR container: per-grid build_graph + Batch.from_data_list + predict +
    graph_to_grid + handler / writer / sidecar, all from the overlay; the output container's refinement records vs the
    UNMODIFIED reference script run as __main__ on the reference's own packages (native_script_case.npz)."""
    import bathymetric_gnn_b200 as bg
    z = gio.load("native_script_case.npz")
    fake_h5py.install()
    src, dst = tmp_path / "in.bag", tmp_path / "o" / "out.bag"
    meta, before = fake_h5py.write_vr_bag(src, **ast.literal_eval(str(z["kw"])))
    assert before.tobytes() == z["ref_before"].tobytes()
    torch.save({"model_state_dict": _state_dict(int(z["model_seed"]), z["cls_bias"], z["conf_bias"]), "config": bg.Config(),
                "in_channels": 7, "edge_dim": 3}, tmp_path / "model.pt")
    flags = [a for a in z["argv"].tolist() if a not in ("--device", "cpu", "--no-export-sidecar")]
    g = overlay.run_script(os.path.join(LAYOUT, "scripts", "inference_native.py"),
                           ["--input", str(src), "--model", str(tmp_path / "model.pt"), "--output", str(dst), *flags])
    assert g["BathymetricGNN"] is bg.BathymetricGNN and g["PyGBatch"] is bg.Batch and g["PYG_BATCH_AVAILABLE"]
    from bathymetric_gnn_b200 import vr_bag as vb
    assert g["VRBagHandler"] is vb.VRBagHandler and g["SidecarBuilder"] is bg.SidecarBuilder
    stats, sidecar = g["STATS"], g["SIDECAR"]
    with fake_h5py.File(dst, "r") as f:
        after = f["BAG_root"]["varres_refinements"][:]
    ref = z["ref_after"]
    # a cell is corrected when class == noise and conf >= threshold: decisions can only differ on fp32 near-ties
    nan_kept = np.isnan(before["depth"])                 # a NaN record stays NaN (NaN != NaN is not a correction)
    dec, dec_ref = (after["depth"] != before["depth"]) & ~nan_kept, (ref["depth"] != before["depth"]) & ~nan_kept
    assert np.isnan(after["depth"][nan_kept]).all()
    flips = int((dec != dec_ref).sum())
    assert dec_ref.sum() > 500 and flips <= max(2, int(2e-3 * dec_ref.sum())), (flips, int(dec_ref.sum()))
    both = dec & dec_ref
    np.testing.assert_allclose(after["depth"][both], ref["depth"][both], rtol=1e-6, atol=2e-4)
    np.testing.assert_allclose(after["depth_uncrt"][both], ref["depth_uncrt"][both], rtol=1e-5, atol=1e-5)
    untouched = ~dec & ~dec_ref
    assert np.array_equal(after["depth"][untouched], before["depth"][untouched], equal_nan=True)
    assert stats["cells_corrected"] == int(dec.sum()) and stats["grids_processed"] > 20
    # the sidecar was built (GDAL-free georeferencing) and saved next to the output
    assert sidecar is not None and np.isfinite(sidecar.classification).sum() > 0
    assert (tmp_path / "o" / "out_gnn_outputs.npz").exists() or (tmp_path / "o" / "out_gnn_outputs.tif").exists()


@pytest.mark.gpu
def test_test_pipeline_script_bodies_on_c1_tile(overlay):
    """BASELINE configs[0]'s driver (scripts/test_pipeline.py): tiling -> graph construction -> random-init forward on
    the 512^2 reference tile; the graph is bit-exact vs the reference build (c1_tile.npz digests)."""
    import hashlib
    import bathymetric_gnn_b200 as bg
    z = gio.load("c1_tile.npz")
    g = overlay.run_script(os.path.join(LAYOUT, "scripts", "test_pipeline.py"), run_name="module_level_only")
    assert g["check_imports"]().endswith("bgnn.b200")
    grid = bg.BathymetricGrid(depth=z["depth"], uncertainty=None, nodata_value=1.0e6, resolution=tuple(float(v) for v in z["res"]))
    tile, n_specs = g["test_tiling"](grid)
    assert n_specs == 1 and tile.shape == (512, 512)
    graph, info = g["test_graph_construction"](tile, grid.resolution)
    assert info["nodes"] == int(z["N"]) and info["edges"] == int(z["E"]) and info["x"] == (int(z["N"]), 7)
    for k in ("edge_index", "valid_rows", "valid_cols", "pos"):
        assert hashlib.sha256(np.ascontiguousarray(getattr(graph, k).cpu().numpy()).tobytes()).hexdigest() == str(z["sha_" + k]), k
    model, out, finfo = g["test_model_forward"](graph)
    assert type(model) is bg.BathymetricGNN and finfo["params"] == 182405
    assert out["class_logits"].shape == (int(z["N"]), 3) and out["predicted_class"].dtype == torch.int64
    assert sum(finfo["counts"]) == int(z["N"]) and 0.0 < finfo["mean_confidence"] < 1.0
    assert bool(torch.isfinite(out["class_logits"]).all())
