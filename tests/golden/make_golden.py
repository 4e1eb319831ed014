#!/usr/bin/env python3
"""Generate the committed golden fixtures by running the UNMODIFIED reference
(/root/reference) in the build container.

    python tests/golden/make_golden.py          # writes tests/golden/*.npz

The reference has no tests or golden vectors of its own, so
its own outputs on seeded inputs are the pins.  Third-party modules missing from
this image are shimmed: `torch_geometric` by the restatement in
oracle/pyg_restated.py (=> the attention arithmetic itself stays "parity
unpinned"), `osgeo.gdal` by a two-function stub.  /root/reference does not exist
on the GPU box; only the .npz files travel.
"""
import hashlib
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyg_restated  # noqa: E402
from oracle import model_oracle as mo  # noqa: E402
from tests import synth  # noqa: E402

pyg_restated.install()
REF = os.environ.get("BGNN_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import config as ref_config  # noqa: E402
import data.graph_construction as ref_gc  # noqa: E402
import data.tiling as ref_tiling  # noqa: E402
import data.loaders as ref_loaders  # noqa: E402
import data.synthetic_noise as ref_noise  # noqa: E402
import models.gnn as ref_gnn  # noqa: E402
import models.pipeline as ref_pipeline  # noqa: E402

logging.disable(logging.CRITICAL)


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def graph_cases():
    """name -> dict(depth, mask|None, unc|None, res, conn)"""
    cases = {}
    d = synth.noisy_tile(64, 64, seed=1)
    cases["g64_full"] = dict(depth=d, mask=np.ones_like(d, bool), unc=None, res=(0.5, 0.5), conn="8-connected")
    d = synth.noisy_tile(64, 64, seed=2)
    m = synth.scattered_mask(64, 64, seed=2, invalid_frac=0.05) & synth.blob_mask(64, 64, seed=2, corner=8)
    d2 = d.copy(); d2[~m] = 1.0e6
    cases["g64_holes_unc"] = dict(depth=d2, mask=m, unc=synth.uncertainty_plane(64, 64, 2), res=(1.0, 2.0), conn="8-connected")
    cases["g64_holes"] = dict(depth=d2, mask=m, unc=None, res=(1.0, 2.0), conn="8-connected")
    d = synth.noisy_tile(48, 40, seed=3)
    m = synth.blob_mask(48, 40, seed=3, invalid_frac=0.08)
    cases["g48_4conn"] = dict(depth=d, mask=m, unc=None, res=(2.0, 2.0), conn="4-connected")
    d = synth.noisy_tile(40, 56, seed=4)
    d[~synth.scattered_mask(40, 56, seed=4, invalid_frac=0.1)] = np.nan
    cases["g40_nan_nomask"] = dict(depth=d, mask=None, unc=None, res=(1.0, 1.0), conn="8-connected")
    d = synth.noisy_tile(33, 37, seed=5)
    m = synth.scattered_mask(33, 37, seed=5, invalid_frac=0.07)
    cases["g33_k24"] = dict(depth=d, mask=m, unc=None, res=(0.5, 0.5), conn="24-connected")
    cases["g33_k24_aniso"] = dict(depth=d, mask=m, unc=None, res=(0.5, 1.5), conn="24-connected")
    d = synth.noisy_tile(32, 32, seed=6)
    m = synth.scattered_mask(32, 32, seed=6, invalid_frac=0.7)
    cases["g32_sparse"] = dict(depth=d, mask=m, unc=None, res=(1.0, 1.0), conn="8-connected")
    for name, (h, w) in dict(t2x2=(2, 2), t3x3=(3, 3), t2x5=(2, 5), t7x3=(7, 3)).items():
        d = synth.noisy_tile(h, w, seed=7 + h * w)
        m = np.ones((h, w), bool)
        if h * w > 4:
            m[0, 1] = False
        cases[name] = dict(depth=d, mask=m, unc=None, res=(1.0, 1.0), conn="8-connected")
    d = synth.noisy_tile(16, 16, seed=8)
    cases["g16_empty"] = dict(depth=d, mask=np.zeros((16, 16), bool), unc=None, res=(1.0, 1.0), conn="8-connected")
    d = (synth.noisy_tile(24, 24, seed=9) * 0 + 12.5).astype(np.float32)     # perfectly flat: std == 0
    cases["g24_flat"] = dict(depth=d, mask=np.ones((24, 24), bool), unc=None, res=(1.0, 1.0), conn="8-connected")
    return cases


def ref_builder(conn):
    from oracle.graph_oracle import OFFSETS_24
    if conn == "24-connected":
        b = ref_gc.GraphBuilder("8-connected")
        b.neighbor_offsets = list(OFFSETS_24)       # plain list attribute (graph_construction.py:78-87)
        return b
    return ref_gc.GraphBuilder(conn)


def run_ref_graph(c):
    b = ref_builder(c["conn"])
    g = b.build_graph(c["depth"], None if c["mask"] is None else c["mask"], c["unc"], c["res"])
    out = dict(x=g.x.numpy(), edge_index=g.edge_index.numpy(), edge_attr=g.edge_attr.numpy(),
               pos=g.pos.numpy(), local_std=g.local_std.numpy())
    if hasattr(g, "valid_rows"):
        out.update(valid_rows=g.valid_rows.numpy(), valid_cols=g.valid_cols.numpy(),
                   num_valid_cells=np.int64(g.num_valid_cells), grid_shape=np.array(g.grid_shape))
    return g, out


def make_graph_fixtures():
    blob = {}
    for name, c in graph_cases().items():
        _, out = run_ref_graph(c)
        blob[f"{name}/depth"] = c["depth"]
        blob[f"{name}/mask"] = c["mask"] if c["mask"] is not None else np.zeros(0, bool)
        blob[f"{name}/unc"] = c["unc"] if c["unc"] is not None else np.zeros(0, np.float32)
        blob[f"{name}/res"] = np.array(c["res"], np.float64)
        blob[f"{name}/conn"] = np.array(c["conn"])
        for k, v in out.items():
            blob[f"{name}/ref/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "graph_cases.npz"), **blob)
    print("graph_cases.npz:", len(graph_cases()), "cases")


def make_c1_fixture():
    """BASELINE config 1: 512x512 tile, reference SyntheticNoiseGenerator(seed=42),
    intensity 1.0 on the analytic surface.  Outputs are stored as SHA-256 digests."""
    clean = synth.clean_surface(512, 512)
    mask = np.ones_like(clean, bool)
    lab = ref_noise.SyntheticNoiseGenerator(seed=42).generate(clean, mask, 1.0)
    noisy = lab.noisy_depth.astype(np.float32)
    b = ref_gc.GraphBuilder("8-connected", edge_features=["distance", "depth_difference", "slope"])
    import time
    t0 = time.time()
    g = b.build_graph(noisy, mask, None, (0.5, 0.5))
    dt = time.time() - t0
    blob = dict(depth=noisy, res=np.array([0.5, 0.5]), ref_build_seconds=np.float64(dt),
                N=np.int64(g.num_nodes), E=np.int64(g.num_edges))
    for k in ("x", "edge_index", "edge_attr", "pos", "local_std", "valid_rows", "valid_cols"):
        blob["sha_" + k] = np.array(sha(getattr(g, k).numpy()))
    np.savez_compressed(os.path.join(HERE, "c1_tile.npz"), **blob)
    print(f"c1_tile.npz: N={g.num_nodes} E={g.num_edges} reference build {dt:.1f}s")


def ref_model(cfg: mo.ModelCfg, sd):
    m = ref_gnn.BathymetricGNN(in_channels=cfg.in_channels, hidden_channels=cfg.hidden,
                               num_gnn_layers=cfg.layers, gnn_type=cfg.gnn_type, heads=cfg.heads,
                               num_classes=cfg.num_classes, predict_correction=cfg.predict_correction,
                               dropout=0.0, edge_dim=cfg.edge_dim)
    missing = m.load_state_dict(sd, strict=True)
    m.eval()
    return m


def make_model_fixtures():
    cases = graph_cases()
    blob = {}
    specs = [("g64_holes", 7, 0, 64, 4), ("g64_holes", 7, 1, 64, 4), ("g64_holes_unc", 8, 2, 64, 4),
             ("g32_sparse", 7, 3, 64, 4), ("g33_k24", 7, 4, 128, 4), ("g48_4conn", 7, 5, 32, 2)]
    for gname, F, seed, hidden, heads in specs:
        g, _ = run_ref_graph(cases[gname])
        cfg = mo.ModelCfg(in_channels=F, hidden=hidden, heads=heads)
        sd = mo.init_state_dict(cfg, seed=seed, randomize_bn=True, head_gain=8.0)
        cls_b, conf_b = mo.calibrate_heads(sd, cfg, g.x, g.edge_index, g.edge_attr)
        m = ref_model(cfg, sd)
        assert set(m.state_dict().keys()) == set(sd.keys())
        out = m.predict(g, 0.85, 0.6)
        key = f"{gname}_s{seed}_h{hidden}x{heads}"
        blob[key + "/cls_bias"], blob[key + "/conf_bias"] = cls_b.numpy(), conf_b.numpy()
        blob[key + "/meta"] = np.array([F, seed, hidden, heads], np.int64)
        blob[key + "/graph"] = np.array(gname)
        blob[key + "/param_count"] = np.int64(sum(p.numel() for p in m.parameters()))
        for k, v in out.items():
            blob[f"{key}/{k}"] = v.numpy()
        print(key, "class hist", np.bincount(out["predicted_class"].numpy(), minlength=3),
              "params", int(blob[key + "/param_count"]))
    # structural pins
    for F in (7, 8):
        cfg = mo.ModelCfg(in_channels=F)
        m = ref_gnn.BathymetricGNN(F, 64, 4, "GAT", 4, 3, True, 0.1, 3)
        blob[f"param_count_F{F}"] = np.int64(sum(p.numel() for p in m.parameters()))
    blob["state_keys"] = np.array(sorted(ref_gnn.BathymetricGNN(7, 64, 4, "GAT", 4, 3, True, 0.1, 3).state_dict().keys()))
    np.savez_compressed(os.path.join(HERE, "model_cases.npz"), **blob)
    print("model_cases.npz: params F7/F8 =", blob["param_count_F7"], blob["param_count_F8"])


def make_model_types_fixtures():
    """GCN / GraphSAGE / GIN backbones (models/gnn.py:120-143): the UNMODIFIED reference module on the restated PyG layers."""
    cases = graph_cases()
    blob = {}
    for t, gname, F, seed in (("GCN", "g64_holes", 7, 11), ("GraphSAGE", "g64_holes_unc", 8, 12), ("GIN", "g32_sparse", 7, 13),
                              ("GCN", "g48_4conn", 7, 14), ("GraphSAGE", "g33_k24", 7, 15), ("GIN", "g64_holes", 7, 16)):
        g, _ = run_ref_graph(cases[gname])
        cfg = mo.ModelCfg(in_channels=F, gnn_type=t)
        sd = mo.init_state_dict(cfg, seed=seed, randomize_bn=True, head_gain=8.0)
        cls_b, conf_b = mo.calibrate_heads(sd, cfg, g.x, g.edge_index, g.edge_attr)
        m = ref_model(cfg, sd)
        assert set(m.state_dict().keys()) == set(sd.keys()), set(m.state_dict().keys()) ^ set(sd.keys())
        out = m.predict(g, 0.85, 0.6)
        key = f"{t}_{gname}_s{seed}"
        blob[key + "/cls_bias"], blob[key + "/conf_bias"] = cls_b.numpy(), conf_b.numpy()
        blob[key + "/meta"] = np.array([F, seed], np.int64)
        blob[key + "/graph"], blob[key + "/type"] = np.array(gname), np.array(t)
        blob[key + "/param_count"] = np.int64(sum(p.numel() for p in m.parameters()))
        blob[key + "/state_keys"] = np.array(sorted(m.state_dict().keys()))
        for k, v in out.items():
            blob[f"{key}/{k}"] = v.numpy()
        print(key, "class hist", np.bincount(out["predicted_class"].numpy(), minlength=3), "params", int(blob[key + "/param_count"]))
    np.savez_compressed(os.path.join(HERE, "model_types_cases.npz"), **blob)


def make_stitch_fixtures():
    blob = {}
    # tile grids for assorted shapes
    shapes = [((150, 170), 64, 16), ((1000, 1000), 256, 32), ((60, 60), 64, 16), ((4096, 3968), 2048, 128),
              ((129, 64), 64, 16), ((300, 90), 100, 40)]
    for i, (shape, ts, ov) in enumerate(shapes):
        tm = ref_tiling.TileManager(ts, ov, 0.1)
        ntr, ntc, specs = tm.compute_tile_grid(shape)
        blob[f"specs{i}/cfg"] = np.array([shape[0], shape[1], ts, ov], np.int64)
        blob[f"specs{i}/n"] = np.array([ntr, ntc], np.int64)
        blob[f"specs{i}/specs"] = np.array([[s.row_start, s.col_start, s.row_end, s.col_end, s.tile_row, s.tile_col]
                                            for s in specs], np.int64)
    # 1-D blend tables
    for size, ov in [(64, 16), (2048, 128), (20, 16), (3, 16), (100, 40), (1024, 128)]:
        tm = ref_tiling.TileManager(max(size, ov + 1), ov, 0.1)
        blob[f"blend1d/{size}_{ov}"] = tm._create_1d_blend(size)
    tm = ref_tiling.TileManager(64, 16, 0.1)
    blob["blend2d/50x64_16"] = tm._create_blend_weights((50, 64))
    # merge of random tile outputs (class / conf / corr) in reference tile order
    shape, ts, ov = (150, 170), 64, 16
    rng = np.random.default_rng(11)
    tm = ref_tiling.TileManager(ts, ov, 0.1)
    merger = ref_tiling.TileMerger(tm)
    chans = ["cleaned_depth", "classification", "confidence", "correction"]
    merger.initialize(shape, chans)
    _, _, specs = tm.compute_tile_grid(shape)
    skip = {3, 7}
    tiles = []
    for i, s in enumerate(specs):
        if i in skip:
            continue
        h, w = s.row_end - s.row_start, s.col_end - s.col_start
        td = dict(cleaned_depth=rng.normal(30, 1, (h, w)).astype(np.float32),
                  classification=rng.integers(0, 3, (h, w)).astype(np.float32),
                  confidence=(np.round(rng.random((h, w)) * 8) / 8).astype(np.float32),   # ties on purpose
                  correction=rng.normal(0, 0.5, (h, w)).astype(np.float32))
        tiles.append((i, td))
        merger.add_tile(s, td)
    res = merger.finalize()
    blob["merge/cfg"] = np.array([shape[0], shape[1], ts, ov], np.int64)
    blob["merge/tile_ids"] = np.array([i for i, _ in tiles], np.int64)
    for i, td in tiles:
        for k, v in td.items():
            blob[f"merge/tile{i}/{k}"] = v
    for k, v in res.items():
        blob[f"merge/ref/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "stitch_cases.npz"), **blob)
    print("stitch_cases.npz")


def make_pipeline_fixture():
    """Reference BathymetricPipeline.process end to end (loader/writer patched out)."""
    H, W, ts, ov = 150, 170, 64, 16
    depth = synth.noisy_tile(H, W, seed=21, sigma=0.3, spike_p=0.02)
    mask = synth.blob_mask(H, W, seed=21, invalid_frac=0.06, corner=0)
    mask[:70, 110:] = False        # makes one tile fall below min_valid_ratio
    nodata = 1.0e6
    depth = depth.copy(); depth[~mask] = nodata
    cfg = ref_config.Config()
    cfg.tile.tile_size, cfg.tile.overlap, cfg.tile.min_valid_ratio = ts, ov, 0.3
    cfg.device = "cpu"
    pipe = ref_pipeline.BathymetricPipeline(cfg)
    mcfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(mcfg, seed=1, randomize_bn=True, head_gain=8.0)
    gfull = ref_gc.GraphBuilder("8-connected").build_graph(depth, mask, None, (0.5, 0.5))
    cls_b, conf_b = mo.calibrate_heads(sd, mcfg, gfull.x, gfull.edge_index, gfull.edge_attr, conf_target=0.88)
    pipe.model = ref_model(mcfg, sd)
    grid = ref_loaders.BathymetricGrid(depth=depth, uncertainty=None, nodata_value=nodata,
                                       transform=(0, 1, 0, 0, 0, -1), crs=None, resolution=(0.5, 0.5),
                                       bounds=(0, 0, W, H), source_path=None)
    pipe.loader = types.SimpleNamespace(load=lambda p: grid)
    pipe.writer = types.SimpleNamespace(save=lambda *a, **k: None)
    res = pipe.process("in.tif", "out.tif")
    blob = dict(depth=depth, cfg=np.array([H, W, ts, ov], np.int64), min_valid_ratio=np.float64(0.3),
                nodata=np.float64(nodata), res=np.array([0.5, 0.5]), seed=np.int64(1),
                cls_bias=cls_b.numpy(), conf_bias=conf_b.numpy())
    for k, v in res.items():
        blob["ref/" + k] = v
    np.savez_compressed(os.path.join(HERE, "pipeline_case.npz"), **blob)
    cl = res["classification"]
    print("pipeline_case.npz: class hist", [int((cl == k).sum()) for k in range(3)],
          "nan", int(np.isnan(cl).sum()), "corrected", int((res["cleaned_depth"] != depth).sum()))


def make_native_fixture():
    """Reference NativeVRProcessor batching (scripts/inference_native.py:117-342)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_inference_native", os.path.join(REF, "scripts", "inference_native.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.logger = logging.getLogger("ref_native")   # the shipped ctor uses an undefined module-level `logger`
    grids = synth.ragged_grids(40, seed=7, lo=2, hi=24, all_invalid_every=9)
    mcfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(mcfg, seed=2, randomize_bn=True, head_gain=8.0)
    big = max(grids, key=lambda g: ((g[0] != 1.0e6).sum()))
    gbig = ref_gc.GraphBuilder("8-connected").build_graph(big[0], big[0] != 1.0e6, None, big[2])
    cls_b, conf_b = mo.calibrate_heads(sd, mcfg, gbig.x, gbig.edge_index, gbig.edge_attr, conf_target=0.8)
    model = ref_model(mcfg, sd)
    proc = mod.NativeVRProcessor(model, ref_gc.GraphBuilder("8-connected"), torch.device("cpu"), 0.85)
    proc.BATCH_NODE_BUDGET = 3000
    results, pending = [None] * len(grids), []
    for i, (d, u, res) in enumerate(grids):
        r = proc.add_to_batch(d, u, res, 1.0e6)
        if r is not None:
            results[i] = r
        else:
            pending.append(i)
        if proc.batch_ready:
            for j, rr in zip(pending, proc.flush_batch()):
                results[j] = rr
            pending = []
    for j, rr in zip(pending, proc.flush_batch()):
        results[j] = rr
    blob = dict(n=np.int64(len(grids)), seed=np.int64(7), model_seed=np.int64(2),
                cls_bias=cls_b.numpy(), conf_bias=conf_b.numpy())
    for i, r in enumerate(results):
        blob[f"g{i}/classification"], blob[f"g{i}/confidence"], blob[f"g{i}/correction"] = r
    # unbatched per-grid results of the same processor (process_grid) for the first 10 grids
    for i in range(10):
        d, u, res = grids[i]
        r = proc.process_grid(d, u, res, 1.0e6)
        blob[f"single{i}/classification"], blob[f"single{i}/confidence"], blob[f"single{i}/correction"] = r
    np.savez_compressed(os.path.join(HERE, "native_case.npz"), **blob)
    print("native_case.npz:", len(grids), "grids")


def make_model_bf16_fixtures():
    """bf16 mode (BASELINE configs[4]): the reference has no reduced-precision path, so these vectors come from the ORACLE's
    bf16-storage restatement (oracle/model_oracle.py::forward(bf16_storage=True)) on the same graphs / weights as two of the
    fp32 model cases; the unmodified reference module's fp32 outputs are stored beside them, so the fixture also pins how far
    bf16 storage moves the reference's results."""
    cases = graph_cases()
    blob = {}
    for gname, F, seed, hidden, heads in [("g64_holes", 7, 0, 64, 4), ("g33_k24", 7, 4, 128, 4)]:
        g, _ = run_ref_graph(cases[gname])
        cfg = mo.ModelCfg(in_channels=F, hidden=hidden, heads=heads)
        sd = mo.init_state_dict(cfg, seed=seed, randomize_bn=True, head_gain=8.0)
        cls_b, conf_b = mo.calibrate_heads(sd, cfg, g.x, g.edge_index, g.edge_attr)
        ref32 = ref_model(cfg, sd).predict(g, 0.85, 0.6)
        out = mo.forward(sd, cfg, g.x.numpy(), g.edge_index.numpy(), g.edge_attr.numpy(), torch.float32, bf16_storage=True)
        key = f"{gname}_s{seed}_h{hidden}x{heads}"
        blob[key + "/cls_bias"], blob[key + "/conf_bias"] = cls_b.numpy(), conf_b.numpy()
        blob[key + "/meta"] = np.array([F, seed, hidden, heads], np.int64)
        blob[key + "/graph"] = np.array(gname)
        for k in ("class_logits", "confidence", "correction", "predicted_class"):
            blob[f"{key}/{k}"] = out[k].numpy()
            blob[f"{key}/ref32/{k}"] = ref32[k].numpy()
        print(key, "max |bf16-storage - reference fp32| logits",
              float((out["class_logits"] - ref32["class_logits"]).abs().max()), "logit scale", float(ref32["class_logits"].abs().max()),
              "label flips", int((out["predicted_class"] != ref32["predicted_class"]).sum()), "of", len(ref32["predicted_class"]))
    np.savez_compressed(os.path.join(HERE, "model_bf16_cases.npz"), **blob)


def make_multiscale_fixture():
    """Reference MultiScaleGraphBuilder (data/graph_construction.py:508-607), UNMODIFIED: scales 1 / 2 / 4 of a 37 x 50 grid
    with holes (NaN-marked in the depth, so nanmean has work to do) and an uncertainty plane."""
    d, m, u = synth.multiscale_case()
    b = ref_gc.MultiScaleGraphBuilder(scales=[4, 1, 2], connectivity="8-connected")
    graphs = b.build_multiscale_graph(d, m, u, (0.5, 1.0))
    blob = {}
    for s, g in graphs.items():
        for k in ("x", "edge_index", "edge_attr", "pos", "valid_rows", "valid_cols", "local_std"):
            blob[f"s{s}/{k}"] = getattr(g, k).numpy()
        blob[f"s{s}/grid_shape"] = np.array(g.grid_shape, np.int64)
        blob[f"s{s}/num_valid_cells"] = np.int64(g.num_valid_cells)
        print("scale", s, "nodes", g.num_nodes, "edges", g.edge_index.shape[1], "grid", g.grid_shape)
    np.savez_compressed(os.path.join(HERE, "multiscale_case.npz"), **blob)


def make_sidecar_fixture():
    """Reference SidecarBuilder.add_refinement_results (data/vr_bag.py:695-778), UNMODIFIED, on a synthetic layout.  The object
    is created without __init__ (its GDAL georeferencing step is file I/O); the state __init__ would set is set here."""
    import importlib
    vb = importlib.import_module("data.vr_bag")
    geo, grids, results = synth.sidecar_scene()
    sb = object.__new__(vb.SidecarBuilder)

    class H:
        base_shape = geo["base_shape"]
    sb.handler = H()
    sb.shape, sb.geotransform, sb.bounds = geo["shape"], geo["geotransform"], geo["bounds"]
    sb.resolution = abs(sb.geotransform[1])
    sb.classification = np.full(sb.shape, np.nan, dtype=np.float32)
    sb.confidence = np.full(sb.shape, np.nan, dtype=np.float32)
    sb.correction = np.full(sb.shape, np.nan, dtype=np.float32)
    sb.valid_mask = np.zeros(sb.shape, dtype=np.float32)
    for g, (cls, conf, corr) in zip(grids, results):
        sb.add_refinement_results(vb.RefinementGrid(**g), cls, conf, corr)
    blob = dict(classification=sb.classification, confidence=sb.confidence, correction=sb.correction, valid_mask=sb.valid_mask,
                seed=np.int64(5), n=np.int64(len(grids)))
    np.savez_compressed(os.path.join(HERE, "sidecar_case.npz"), **blob)
    print("sidecar_case.npz: painted pixels", int(np.isfinite(sb.classification).sum()), "of", sb.classification.size,
          "valid", int(sb.valid_mask.sum()))


def make_vr_bag_fixture():
    """Reference VRBagHandler.iterate_refinements (data/vr_bag.py:243-298), VRBagWriter.update_refinement_batch (:550-588),
    detect_bag_type (:29-63) and the SR handler / writer (:315-475), UNMODIFIED, on synthetic containers served by the
    h5py stand-in of tests/fake_h5py.py (h5py itself is absent from this image)."""
    import importlib
    import tempfile
    from tests import fake_h5py
    fake_h5py.install()
    vb = importlib.import_module("data.vr_bag")
    vb = importlib.reload(vb)                      # picks the stand-in up (H5PY_AVAILABLE is latched at import)
    assert vb.H5PY_AVAILABLE
    blob = {}
    with tempfile.TemporaryDirectory() as tmp:
        for case, kw in dict(shuffled=dict(seed=3, shuffle_storage=True), ordered=dict(seed=5, shuffle_storage=False, base_shape=(4, 6))).items():
            src, dst = os.path.join(tmp, case + ".bag"), os.path.join(tmp, case + "_out.bag")
            meta, ref = fake_h5py.write_vr_bag(src, **kw)
            blob[f"{case}/kw"] = np.array(repr(kw))
            blob[f"{case}/meta"], blob[f"{case}/ref"] = meta, ref
            blob[f"{case}/type"] = np.array(vb.detect_bag_type(src))
            h = vb.VRBagHandler(src)
            info = h.get_refinement_info()
            blob[f"{case}/info"] = np.array([info["num_refined_cells"], info["total_refinement_nodes"]], np.int64)
            blob[f"{case}/unique_resolutions"] = np.array(info["unique_resolutions"], np.float64)
            for mvr in (0.0, 0.5):
                grids = list(h.iterate_refinements(mvr))
                tag = f"{case}/mvr{mvr}"
                blob[tag + "/desc"] = np.array([[g.base_row, g.base_col, g.dimensions[0], g.dimensions[1], g.start_index] for g in grids], np.int64)
                blob[tag + "/res"] = np.array([g.resolution for g in grids], np.float64)
                blob[tag + "/sw"] = np.array([g.sw_corner for g in grids], np.float64)
                blob[tag + "/depth"] = np.concatenate([g.depth.ravel() for g in grids])
                blob[tag + "/unc"] = np.concatenate([g.uncertainty.ravel() for g in grids])
                blob[tag + "/num_valid"] = np.array([g.num_valid for g in grids], np.int64)
            # write-back of seeded "corrections" through the reference writer, grid by grid
            w = h.copy_and_open_for_writing(dst)
            rng = np.random.default_rng(17)
            for g in h.iterate_refinements(0.0):
                cd = g.depth.copy()
                cu = g.uncertainty.copy()
                m = g.valid_mask & (rng.random(g.shape) < 0.3)
                cd[m] -= rng.normal(0, 0.5, int(m.sum())).astype(np.float32)
                cu[m] *= np.float32(1.25)
                w.update_refinement_batch(g, cd, cu)
            blob[f"{case}/corrections_applied"] = np.int64(w._corrections_applied)
            w.close()
            with fake_h5py.File(dst, "r") as f:
                blob[f"{case}/ref_after"] = f["BAG_root"]["varres_refinements"][:]
        # SR container
        src, dst = os.path.join(tmp, "sr.bag"), os.path.join(tmp, "sr_out.bag")
        d = fake_h5py.write_sr_bag(src)
        blob["sr/type"] = np.array(vb.detect_bag_type(src))
        h = vb.SRBagHandler(src)
        g = next(iter(h.iterate_refinements(0.0)))
        blob["sr/depth"], blob["sr/unc"], blob["sr/res"] = g.depth, g.uncertainty, np.array(g.resolution)
        blob["sr/info"] = np.array([h.get_refinement_info()["total_refinement_nodes"]], np.int64)
        blob["sr/n_grids_mvr0.95"] = np.int64(len(list(h.iterate_refinements(0.95))))
        w = h.copy_and_open_for_writing(dst)
        w.update_refinement_batch(g, g.depth + np.float32(1.0), g.uncertainty * np.float32(2.0))
        w.close()
        with fake_h5py.File(dst, "r") as f:
            blob["sr/elev_after"], blob["sr/unc_after"] = f["BAG_root"]["elevation"][:], f["BAG_root"]["uncertainty"][:]
    np.savez_compressed(os.path.join(HERE, "vr_bag_case.npz"), **blob)
    print("vr_bag_case.npz:", {k: blob[k].tolist() for k in blob if k.endswith("/info")})


def make_native_script_fixture():
    """The UNMODIFIED scripts/inference_native.py run as __main__ (runpy) on the reference's own packages, CPU: VR container
    served by the h5py stand-in, checkpoint with the pickled reference Config, `--no-export-sidecar` (its georeferencing
    needs GDAL).  A module-level `logger` is pre-seeded: the shipped script uses one it never defines (:148 vs :348).
    The output container's refinement records are the fixture."""
    import importlib
    import runpy
    import tempfile
    from tests import fake_h5py
    fake_h5py.install()
    importlib.reload(importlib.import_module("data.vr_bag"))
    import data as ref_data
    importlib.reload(ref_data)                   # re-exports the reloaded handler classes
    kw = dict(base_shape=(6, 8), seed=23, lo=3, hi=20, nodata_frac=0.08)
    mcfg = mo.ModelCfg(in_channels=7)
    sd = mo.init_state_dict(mcfg, seed=4, randomize_bn=True, head_gain=8.0)
    with tempfile.TemporaryDirectory() as tmp:
        src, dst, ckpt = os.path.join(tmp, "in.bag"), os.path.join(tmp, "out", "out.bag"), os.path.join(tmp, "model.pt")
        meta, ref = fake_h5py.write_vr_bag(src, **kw)
        # calibrate the heads on the biggest grid so that noise / confident cells exist
        big = max(importlib.import_module("data.vr_bag").VRBagHandler(src).iterate_refinements(0.0), key=lambda g: g.num_valid)
        gbig = ref_gc.GraphBuilder("8-connected").build_graph(big.depth, big.valid_mask, None, big.resolution)
        cls_b, conf_b = mo.calibrate_heads(sd, mcfg, gbig.x, gbig.edge_index, gbig.edge_attr, conf_target=0.8)
        torch.save({"model_state_dict": sd, "config": ref_config.Config(), "in_channels": 7, "edge_dim": 3}, ckpt)
        argv = ["--input", src, "--model", ckpt, "--output", dst, "--device", "cpu", "--no-export-sidecar",
                "--min-valid-ratio", "0.05", "--auto-correct-threshold", "0.7", "--log-level", "ERROR"]
        old = sys.argv
        sys.argv = [os.path.join(REF, "scripts", "inference_native.py")] + argv
        try:
            g = runpy.run_path(sys.argv[0], init_globals={"logger": logging.getLogger("ref_native_script")}, run_name="__main__")
        finally:
            sys.argv = old
        g["NativeVRProcessor"].BATCH_NODE_BUDGET = 50000
        with fake_h5py.File(dst, "r") as f:
            after = f["BAG_root"]["varres_refinements"][:]
    blob = dict(kw=np.array(repr(kw)), model_seed=np.int64(4), cls_bias=cls_b.numpy(), conf_bias=conf_b.numpy(),
                argv=np.array(argv[6:]), ref_before=ref, ref_after=after)
    np.savez_compressed(os.path.join(HERE, "native_script_case.npz"), **blob)
    changed = int((after["depth"] != ref["depth"]).sum())
    print("native_script_case.npz: records", ref.shape[1], "depth changed", changed,
          "uncertainty changed", int((after["depth_uncrt"] != ref["depth_uncrt"]).sum()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["native_script", "graph", "c1", "model", "model_bf16", "model_types", "stitch", "pipeline", "native", "sidecar", "multiscale", "vr_bag"]
    if "vr_bag" in which: make_vr_bag_fixture()
    if "native_script" in which: make_native_script_fixture()
    if "graph" in which: make_graph_fixtures()
    if "model" in which: make_model_fixtures()
    if "model_bf16" in which: make_model_bf16_fixtures()
    if "model_types" in which: make_model_types_fixtures()
    if "sidecar" in which: make_sidecar_fixture()
    if "multiscale" in which: make_multiscale_fixture()
    if "stitch" in which: make_stitch_fixtures()
    if "pipeline" in which: make_pipeline_fixture()
    if "native" in which: make_native_fixture()
    if "c1" in which: make_c1_fixture()
