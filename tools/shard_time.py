"""Per-rank device time of one sharded frame on ONE GPU (what a rank of an N-GPU run does), to
separate fixed per-frame overhead from work that scales with the number of rows."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
lib = vp.load_product(); abi = vp.abi
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
sc, cam, fpd = scenes.config(cfg)
W, H = fpd["width"], fpd["height"]
s = vp.Session(lib, 0); s.set_scene(sc); cc = cam.to_c()
acc = torch.zeros((H, W, 4), device="cuda"); rgb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
for world in (1, 2, 4, 8, 16):
    ts = []
    for i in range(13):
        fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0,
                             row_tile=8 if world > 1 else 0, row_tile_stride=world, row_tile_phase=0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), st.cuda_stream, None), "render")
        b.record(st); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    print(f"cfg{cfg} shard 1/{world}: {np.mean(ts):.3f} ms per frame-share (ideal {np.mean(ts) if world==1 else 0:.3f})")
# the same shares with frames in flight (VP_FLAG_DEFER): throughput of a rank in steady state
K = 30
for world in (1, 2, 4, 8, 16):
    def frame(i):
        fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0,
                             row_tile=8 if world > 1 else 0, row_tile_stride=world, row_tile_phase=0, flags=abi.VP_FLAG_DEFER)
        lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), st.cuda_stream, None), "render")
    for i in range(6): frame(i)
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence"); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(K): frame(i)
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence")
    b.record(st); torch.cuda.synchronize()
    print(f"cfg{cfg} shard 1/{world}, frames in flight: {a.elapsed_time(b) / K:.3f} ms per frame-share")
