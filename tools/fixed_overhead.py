"""Times a tiny shard (1/64 of the rows) of cfg2: essentially the fixed per-frame overhead of each pipeline."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
lib = vp.load_product(); abi = vp.abi
sc, cam, fpd = scenes.config(2)
W, H = fpd["width"], fpd["height"]
s = vp.Session(lib, 0); s.set_scene(sc); cc = cam.to_c()
acc = torch.zeros((H, W, 4), device="cuda"); rgb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream()
for name, flags in (("staged", 0),):
    ts = []
    for i in range(13):
        fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0,
                             row_tile=8, row_tile_stride=64, row_tile_phase=0, flags=flags)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), st.cuda_stream, None), "render")
        b.record(st); torch.cuda.synchronize()
        if i >= 3: ts.append(a.elapsed_time(b))
    print(f"{name}: 1/64 shard {np.mean(ts):.3f} ms")
