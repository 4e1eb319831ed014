"""CPU: host-side logic that needs no GPU -- the C-ABI library loads and exports every
symbol include/bgnn.h declares, tile geometry / blend tables vs the reference goldens,
config round trip, band partition + seam merge (world_size-2 gloo)."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import stitch_oracle as so
from tests import golden_io as gio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import bathymetric_gnn_b200 as bg
    hdr = open(os.path.join(ROOT, "include", "bgnn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(bgnn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(bg._lib.SYMBOLS.keys())
    for name in declared:
        assert hasattr(bg._lib.lib, name), name
    assert bg._lib.lib.bgnn_version() == 100


def test_no_cpu_fallback():
    import bathymetric_gnn_b200 as bg
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bg.GraphBuilder().build_graph(np.zeros((4, 4), np.float32))
    m = bg.BathymetricGNN(7, edge_dim=3).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(bg.Data(x=torch.zeros(3, 7), edge_index=torch.zeros(2, 0, dtype=torch.long)))
    with pytest.raises(RuntimeError):
        bg.BathymetricPipeline(bg.Config())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bathymetric-gnn_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/", ""), fn


def test_tile_geometry_and_blend_vs_reference_golden():
    import bathymetric_gnn_b200 as bg
    z = gio.load("stitch_cases.npz")
    i = 0
    while f"specs{i}/cfg" in z.files:
        H, W, ts, ov = (int(v) for v in z[f"specs{i}/cfg"])
        ntr, ntc, specs = bg.TileManager(ts, ov, 0.1).compute_tile_grid((H, W))
        assert [ntr, ntc] == list(z[f"specs{i}/n"])
        got = np.array([[s.row_start, s.col_start, s.row_end, s.col_end, s.tile_row, s.tile_col] for s in specs], np.int64)
        assert np.array_equal(got, z[f"specs{i}/specs"])
        i += 1
    for k in z.files:
        if k.startswith("blend1d/"):
            size, ov = (int(v) for v in k.split("/")[1].split("_"))
            assert np.array_equal(bg.TileManager(max(size, ov + 1), ov)._create_1d_blend(size), z[k])
    assert np.array_equal(bg.TileManager(64, 16)._create_blend_weights((50, 64)), z["blend2d/50x64_16"])
    with pytest.raises(ValueError):
        bg.TileManager(16, 16)


def test_config_roundtrip(tmp_path):
    import bathymetric_gnn_b200 as bg
    c = bg.Config()
    assert (c.tile.tile_size, c.tile.overlap, c.tile.min_valid_ratio) == (1024, 128, 0.1)
    assert c.graph.connectivity == "8-connected" and c.model.gnn_heads == 4 and c.inference.auto_correct_threshold == 0.85
    c.tile.tile_size = 512
    c.noise.blob_size_range = (4, 9)
    c.save(tmp_path / "c.yaml")
    d = bg.Config.load(tmp_path / "c.yaml")
    assert d.tile.tile_size == 512 and d.noise.blob_size_range == (4, 9) and d.graph.edge_features == c.graph.edge_features
    with pytest.raises(AssertionError):
        bg.Config(tile=bg.TileConfig(tile_size=100, overlap=80))
    assert bg.CORRECTION_NORM_FLOOR == 0.01 and bg.CORRECTION_NORM_CAP == 50.0


def test_band_partition():
    from bathymetric_gnn_b200 import parallel as par
    assert par.band_partition(32, 8) == [(4 * i, 4 * i + 4) for i in range(8)]
    assert par.band_partition(5, 2) == [(0, 3), (3, 5)]
    assert par.band_partition(1, 2) == [(0, 1), (1, 1)]


def _band_worker(rank, world, port, q):
    """One rank: stitch its band of tiles with the CPU oracle merger (stands in for the device
    TileMerger), then gather the unfinalised band planes on rank 0 over gloo."""
    import torch.distributed as dist
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200 import parallel as par
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    z = gio.load("stitch_cases.npz")
    H, W, ts, ov = (int(v) for v in z["merge/cfg"])
    ntr, ntc, specs = bg.TileManager(ts, ov, 0.1).compute_tile_grid((H, W))
    ids = set(int(i) for i in z["merge/tile_ids"])
    band = par.band_partition(ntr, world)[rank]
    mine = par.band_specs(specs, band)
    r0, r1 = par.band_rows(mine)
    chans = ["classification", "confidence", "correction"]
    m = so.Merger((r1 - r0, W), chans, ov)
    for s, s_local in zip(mine, par.shift_specs(mine, r0)):
        i = specs.index(s)
        if i in ids:
            m.add_tile((s_local.row_start, s_local.col_start, s_local.row_end, s_local.col_end),
                       {k: z[f"merge/tile{i}/{k}"] for k in chans})
    part = {"classification": torch.from_numpy(m.out["classification"]), "tracker": torch.from_numpy(m.tracker)}
    for c in ("confidence", "correction"):
        part["acc_" + c], part["wsum_" + c] = torch.from_numpy(m.out[c]), torch.from_numpy(m.w[c])
    gathered = [None] * world
    dist.gather_object((part, (r0, r1)), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        full = par.merge_band_partials((H, W), [g[0] for g in gathered], [g[1] for g in gathered])
        res = par.finalize_merged(full)
        q.put({k: v.numpy() for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stitch_two_ranks_gloo():
    """world_size 2 over gloo: banded stitch + seam merge == the single-process reference result
    (classification bit-exact; continuous channels within 4 eps32 on seam cells, exact off the seams)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_band_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = gio.load("stitch_cases.npz")
    assert np.array_equal(res["classification"], z["merge/ref/classification"], equal_nan=True)
    for c in ("confidence", "correction"):
        ref = z[f"merge/ref/{c}"]
        assert np.array_equal(np.isnan(res[c]), np.isnan(ref))
        ok = ~np.isnan(ref)
        diff = np.abs(res[c][ok].astype(np.float64) - ref[ok])
        # seam cells: (A+B)+(C+D) vs ((A+B)+C)+D, then acc/wsum -> a few ulp32 of the term magnitude
        assert np.all(diff <= 4 * np.finfo(np.float32).eps * np.maximum(1.0, np.abs(ref[ok]))), diff.max()
        assert (diff == 0).mean() > 0.9


def test_plan_bands_partitions_the_rows():
    """survey.plan_bands: owned row ranges partition [0, H); seams are exactly the rows two neighbouring bands touch;
    thin bands (one tile-row each, shifted last tile-row) may have overlapping seams."""
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200 import survey as sv
    for (H, W, ts, ov) in ((60000, 60000, 2048, 128), (5000, 3000, 2048, 128), (180, 200, 96, 16), (300, 90, 96, 16)):
        ntr, ntc, specs = bg.TileManager(ts, ov, 0.1).compute_tile_grid((H, W))
        for world in (1, 2, 3, 8, 40):
            plans = sv.plan_bands(specs, ntr, world, (H, W))
            live = [p for p in plans if not p.empty]
            assert len(plans) == world and sum(len(p.specs) for p in plans) == len(specs)
            assert [s for p in plans for s in p.specs] == list(specs)          # reference order kept
            cover = 0
            for i, p in enumerate(live):
                assert p.own_rows[0] == cover and p.own_rows[1] >= p.own_rows[0]
                cover = p.own_rows[1]
                if i + 1 < len(live):
                    assert p.next_rank == live[i + 1].rank and p.bottom0 == live[i + 1].r0 and live[i + 1].top_end == p.r1
                else:
                    assert p.next_rank is None and p.bottom0 == p.r1 == H
            assert cover == H
    ntr, _, specs = bg.TileManager(96, 16, 0.1).compute_tile_grid((180, 200))
    mid = sv.plan_bands(specs, ntr, 3, (180, 200))[1]
    assert mid.top_end > mid.bottom0                                           # the thin-band case exists


def _seam_worker(rank, world, port, q):
    """One rank of the EXACT seam protocol of survey.py on the CPU: the band loop of `process_band` with the oracle
    merger standing in for the device planes -- park the strips that lie in the previous band's seam, hand the bottom seam
    down through `DistLink` (gloo), copy the received planes in, replay the parked strips in tile order, emit the owned rows."""
    import torch.distributed as dist
    import bathymetric_gnn_b200 as bg
    from bathymetric_gnn_b200 import survey as sv
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    z = gio.load("stitch_cases.npz")
    H, W, ts, ov = (int(v) for v in z["merge/cfg"])
    ntr, ntc, specs = bg.TileManager(ts, ov, 0.1).compute_tile_grid((H, W))
    ids = set(int(i) for i in z["merge/tile_ids"])
    plan = sv.plan_bands(specs, ntr, world, (H, W))[rank]
    chans = ["classification", "confidence", "correction"]
    own = {}
    if not plan.empty:
        bh = plan.r1 - plan.r0
        m = so.Merger((bh, W), chans, ov)

        def add_rows(spec, data, i0, i1):            # Merger.add_tile restricted to tile rows [i0, i1)
            if i1 <= i0:
                return
            r0, c0, r1, c1 = spec.row_start - plan.r0, spec.col_start, spec.row_end - plan.r0, spec.col_end
            wts = so.blend_2d((r1 - r0, c1 - c0), ov)[i0:i1]
            conf = data["confidence"][i0:i1]
            for ch in chans:
                d = data[ch][i0:i1]
                region = m.out[ch][r0 + i0:r0 + i1, c0:c1]
                if ch == "classification":
                    creg = m.tracker[r0 + i0:r0 + i1, c0:c1]
                    upd = np.isfinite(d) & ((conf > creg) | np.isnan(region))
                    region[upd] = d[upd]
                    creg[upd] = conf[upd]
                else:
                    wreg = m.w[ch][r0 + i0:r0 + i1, c0:c1]
                    valid = np.isfinite(d)
                    region[np.isnan(region) & valid] = 0.0
                    wreg[valid] += wts[valid]
                    region[valid] += (d * wts)[valid]
        top, bottom = plan.top_end - plan.r0, plan.bottom0 - plan.r0
        parked = []
        for s in plan.specs:
            i = specs.index(s)
            if i not in ids:
                continue
            data = {k: z[f"merge/tile{i}/{k}"] for k in chans}
            k = max(0, min(top, s.row_end - plan.r0) - (s.row_start - plan.r0))
            if k:
                parked.append((s, data, k))
            add_rows(s, data, k, s.row_end - s.row_start)
        views = lambda a, b: [torch.from_numpy(p[a:b]) for p in (m.out["confidence"], m.out["correction"], m.w["confidence"],
                                                                 m.w["correction"], m.out["classification"], m.tracker)]
        send = (plan.next_rank, views(bottom, bh)) if plan.next_rank is not None and bh > bottom else None
        recv = (plan.prev_rank, views(0, top)) if plan.prev_rank is not None and top > 0 else None
        link = sv.DistLink(dist)

        def replay():
            for s, data, k in parked:
                add_rows(s, data, 0, k)
        if send and recv and top > bottom:
            link.exchange(None, recv); replay(); link.exchange(send, None)
        else:
            link.exchange(send, recv)
            if recv:
                replay()
        res = m.finalize()
        own = {k: v[:bottom].copy() for k, v in res.items()}
    gathered = [None] * world
    dist.gather_object((plan.own_rows, own), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        full = {k: np.concatenate([g[1][k] for g in gathered if g[1]], axis=0) for k in chans}
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exact_seam_protocol_gloo(world):
    """world_size 2 and 3 over gloo: bands + parked seam strips + point-to-point hand-over reproduce the UNMODIFIED
    reference TileMerger result bit for bit (stitch_cases.npz merge/ref/*), seams included -- no tolerance."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_seam_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = gio.load("stitch_cases.npz")
    for c in ("classification", "confidence", "correction"):
        assert np.array_equal(res[c], z[f"merge/ref/{c}"], equal_nan=True), c


def test_dropin_packages_resolve_reference_imports():
    """The import lines of scripts/inference.py:24-25, models/pipeline.py:21-31 and
    scripts/inference_native.py:32-36 resolve against the drop-in directory."""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from config import Config\nfrom config.constants import CORRECTION_NORM_FLOOR\n"
            "from data import BathymetricGrid, BathymetricLoader, BathymetricWriter, TileManager, TileMerger, GraphBuilder\n"
            "from models import BathymetricPipeline\nfrom models.gnn import BathymetricGNN\n"
            "assert Config().tile.tile_size == 1024 and CORRECTION_NORM_FLOOR == 0.01\n"
            "assert BathymetricGNN.CLASS_NOISE == 2\nprint('ok')\n") % os.path.join(ROOT, "bathymetric-gnn_b200", "dropin")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_every_reference_gnn_type_constructs_with_the_reference_state_dict_keys():
    """models/gnn.py:120-146: GAT / GCN / GraphSAGE / GIN module trees load the oracle's (= the reference module's) state
    dicts with strict=True; an unknown type raises the reference's ValueError."""
    import bathymetric_gnn_b200 as bg
    from oracle import model_oracle as mo
    z = gio.load("model_types_cases.npz")
    for t in ("GCN", "GraphSAGE", "GIN"):
        cfg = mo.ModelCfg(in_channels=7, gnn_type=t)
        sd = mo.init_state_dict(cfg, seed=1)
        m = bg.BathymetricGNN(7, 64, 4, t, 4, 3, True, 0.1, 3)
        m.load_state_dict(sd, strict=True)
        key = [k for k in z.files if k.startswith(t + "_") and k.endswith("/state_keys")][0]
        assert sorted(m.state_dict().keys()) == z[key].tolist()
    with pytest.raises(ValueError, match="Unknown GNN type"):
        bg.BathymetricGNN(7, 64, 4, "Transformer")


def test_set_precision_validates_and_needs_no_gpu():
    """BathymetricGNN.set_precision (additive; BASELINE configs[4] bf16 mode): host-side state only, bad names raise
    ValueError like the reference's other string options (`Unknown connectivity`, graph_construction.py:88-89)."""
    import bathymetric_gnn_b200 as bg
    m = bg.BathymetricGNN(7)
    assert m.precision == "fp32"
    assert m.set_precision("bf16") is m and m.precision == "bf16"
    with pytest.raises(ValueError):
        m.set_precision("fp8")
    assert m.set_precision("fp32").precision == "fp32"
