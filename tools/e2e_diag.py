"""Dev aid: where the e2e leg of bench.py loses time against the resident loop (variants of the streaming loop)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import bathymetric_gnn_b200 as bg

depth_np, mask_np = bench.make_tile(0)
cfg, sd = bench.make_weights(0)
model = bg.BathymetricGNN(cfg.in_channels, cfg.hidden, cfg.layers, "GAT", cfg.heads, cfg.num_classes, cfg.predict_correction, 0.0, cfg.edge_dim)
model.load_state_dict(sd); model = model.cuda().eval()
builder = bg.GraphBuilder("8-connected")
dev = torch.device("cuda", 0)
depth_dev, mask_dev = torch.from_numpy(depth_np).to(dev), torch.from_numpy(mask_np).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
depth_pin, mask_pin = torch.from_numpy(depth_np).pin_memory(), torch.from_numpy(mask_np).pin_memory()
g = builder.build_graph(depth_dev, mask_dev, None, bench.RES); N = g.num_nodes; del g
host_out = [(torch.empty(N, dtype=torch.int64).pin_memory(), torch.empty(N, dtype=torch.float32).pin_memory(),
             torch.empty(N, dtype=torch.float32).pin_memory()) for _ in range(2)]
dev_in = [(torch.empty_like(depth_dev), torch.empty_like(mask_dev)) for _ in range(2)]
main_s, copy_s = torch.cuda.current_stream(), torch.cuda.Stream()


def loop(steps, h2d, d2h, keep_alive=True, do_flush=True):
    ev_in = [torch.cuda.Event() for _ in range(2)]; ev_cmp = [torch.cuda.Event() for _ in range(2)]; ev_out = [torch.cuda.Event() for _ in range(2)]
    keep = [None, None]

    def upload(t):
        b = t % 2
        with torch.cuda.stream(copy_s):
            if t >= 2:
                copy_s.wait_event(ev_cmp[b])
            if h2d:
                dev_in[b][0].copy_(depth_pin, non_blocking=True); dev_in[b][1].copy_(mask_pin, non_blocking=True)
            ev_in[b].record(copy_s)
    upload(0)
    for t in range(steps):
        b = t % 2
        if t + 1 < steps:
            upload(t + 1)
        main_s.wait_event(ev_in[b])
        if do_flush:
            flush.fill_(1)
        g = builder.build_graph(dev_in[b][0], dev_in[b][1], None, bench.RES)
        out = model.predict(g)
        ev_cmp[b].record(main_s)
        if keep_alive:
            keep[b] = (g, out)
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(ev_cmp[b])
            if d2h == "full":
                host_out[b][0].copy_(out["predicted_class"], non_blocking=True)
                host_out[b][1].copy_(out["confidence"], non_blocking=True)
                host_out[b][2].copy_(out["correction"], non_blocking=True)
            elif d2h == "tiny":
                host_out[b][1][:1].copy_(out["confidence"][:1], non_blocking=True)
            ev_out[b].record(copy_s)
        if t >= 1:
            ev_out[1 - b].synchronize(); keep[1 - b] = None
    ev_out[(steps - 1) % 2].synchronize()
    main_s.wait_event(ev_out[(steps - 1) % 2])


def timed(name, **kw):
    dev_in[0][0].copy_(depth_dev); dev_in[0][1].copy_(mask_dev); dev_in[1][0].copy_(depth_dev); dev_in[1][1].copy_(mask_dev)
    loop(3, **kw); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop(10, **kw); e1.record(); torch.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/10:8.3f} ms/step")


timed("no copies at all", h2d=False, d2h="none")
timed("H2D only", h2d=True, d2h="none")
timed("D2H tiny (4 bytes)", h2d=False, d2h="tiny")
timed("D2H full (65 MB)", h2d=False, d2h="full")
timed("H2D + D2H full (= bench e2e)", h2d=True, d2h="full")
timed("no copies, results not kept alive", h2d=False, d2h="none", keep_alive=False)
timed("no copies, no L2 flush", h2d=False, d2h="none", do_flush=False)
