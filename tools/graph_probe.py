#!/usr/bin/env python3
"""Deferred cfg-2 frames through the lane graphs: device time per frame over several batches (how stable is the pipeline?),
with and without lone (non-deferred, stage-timed) steps in between, and what vp_info says about captures / node updates."""
import ctypes as C, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes"); abi = vp.abi
lib = vp.load_product()
sc, cam, fpd = scenes.config(2)
W, H = fpd["width"], fpd["height"]
dev = torch.device("cuda", 0)
acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev); rgb = torch.zeros((H, W), dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream(dev); cc = cam.to_c()
with vp.Session(lib) as s:
    s.set_scene(sc)
    def run(a, n, lone_every=0):
        for f in range(a, a + n):
            lone = lone_every and (f - a) % lone_every == 0
            if lone:
                lib.check(lib.c_render_fence(s.ctx, stream.cuda_stream, 0), "fence")
            fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=f, flags=abi.VP_FLAG_TIME_STAGES if lone else abi.VP_FLAG_DEFER)
            st = abi.VpStats()
            lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), stream.cuda_stream, C.byref(st) if lone else None), "rd")
        lib.check(lib.c_render_fence(s.ctx, stream.cuda_stream, 0), "fence")
    run(0, 8); torch.cuda.synchronize()
    for lone_every in (0, 20):
        out = []
        for b in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); run(8 + 60 * b, 60, lone_every); e1.record(stream); torch.cuda.synchronize()
            out.append(round(e0.elapsed_time(e1) / 60, 3))
        print("lone_every", lone_every, "ms/frame per batch of 60:", out)
    i = s.info()
    print({k: i[k] for k in ("graph_replays", "graph_node_updates", "graph_captures", "graph_enabled", "hbm_in_use")})
