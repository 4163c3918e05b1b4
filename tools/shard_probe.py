"""Steady-state device time per frame-share (1/world of the rows as interleaved 8-row tiles, frames in flight) on ONE
GPU: what one rank of an N-GPU run does.  usage: shard_probe.py [cfg] [world] [frames]; knobs through VP_* variables."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
lib = vp.load_product(); abi = vp.abi
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
K = int(sys.argv[3]) if len(sys.argv) > 3 else 120
sc, cam, fpd = scenes.config(cfg)
W, H = fpd["width"], fpd["height"]
s = vp.Session(lib, 0); s.set_scene(sc); cc = cam.to_c()
acc = torch.zeros((H, W, 4), device="cuda"); rgb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
def frame(i):
    fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0,
                         row_tile=8 if world > 1 else 0, row_tile_stride=world, row_tile_phase=i % world, flags=abi.VP_FLAG_DEFER)
    lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), st.cuda_stream, None), "render")
for i in range(12): frame(i)
lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence"); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(K): frame(i)
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence")
    b.record(st); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / K)
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("VP_"))
print(f"cfg{cfg} 1/{world}: {best:.4f} ms per frame-share   [{knobs}]")
