"""How much do independent frames overlap on one GPU?  S contexts (own streams and path state) each
render K frame-shares; compare the wall time per frame with S = 1.  Upper bound for frames-in-flight
pipelining inside one context."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import numpy as np, torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
lib = vp.load_product(); abi = vp.abi
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc, cam, fpd = scenes.config(cfg)
W, H = fpd["width"], fpd["height"]
cc = cam.to_c()
K = 24
for world in [int(x) for x in os.environ.get("PROBE_WORLDS", "1,8").split(",")]:
    for S in [int(x) for x in os.environ.get("PROBE_CTX", "1,3").split(",")]:
        sess = [vp.Session(lib, 0) for _ in range(S)]
        for s in sess: s.set_scene(sc)
        streams = [torch.cuda.Stream() for _ in range(S)]
        acc = [torch.zeros((H, W, 4), device="cuda") for _ in range(S)]
        rgb = [torch.zeros((H, W), dtype=torch.int32, device="cuda") for _ in range(S)]
        def frame(j, i):
            fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0,
                                 row_tile=8 if world > 1 else 0, row_tile_stride=world, row_tile_phase=0)
            lib.check(lib.c_render_device(sess[j].ctx, C.byref(cc), C.byref(fp), acc[j].data_ptr(), rgb[j].data_ptr(), streams[j].cuda_stream, None), "render")
        for i in range(3 * S): frame(i % S, i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream())
        for s_ in streams: s_.wait_stream(torch.cuda.current_stream())
        h0 = time.perf_counter()
        for i in range(K): frame(i % S, i)
        h1 = time.perf_counter()
        for s_ in streams: torch.cuda.current_stream().wait_stream(s_)
        b.record(torch.cuda.current_stream()); torch.cuda.synchronize()
        print(f"cfg{cfg} shard 1/{world} contexts {S}: {a.elapsed_time(b) / K:.3f} ms per frame-share, host enqueue {1e3 * (h1 - h0) / K:.3f} ms per frame", flush=True)
        for s in sess: s.close()
        del sess, acc, rgb
        torch.cuda.empty_cache()
