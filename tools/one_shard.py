"""Render a few frame-shares (1/world of the rows, interleaved tiles) on one GPU; meant to be run under ncu."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
lib = vp.load_product()
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sc, cam, fpd = scenes.config(cfg)
W, H = fpd["width"], fpd["height"]
s = vp.Session(lib, 0); s.set_scene(sc); cc = cam.to_c()
acc = torch.zeros((H, W, 4), device="cuda"); rgb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
for i in range(frames):
    fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0,
                         row_tile=8 if world > 1 else 0, row_tile_stride=world, row_tile_phase=0)
    lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), st.cuda_stream, None), "render")
    torch.cuda.synchronize()
