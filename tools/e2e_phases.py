"""Where an end-to-end vp_render call (the synchronous Update()) spends its time: wall time of the call for a few band
settings (environment is read at context creation).  usage: e2e_phases.py [cfg]"""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
abi = vp.abi
lib = vp.load_product()
sc, cam, fpd = scenes.config(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
W, H = fpd["width"], fpd["height"]
cc = cam.to_c()
h_acc = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
h_rgb = torch.zeros((H, W), dtype=torch.int32).pin_memory()
KEYS = ("VP_BANDS", "VP_BAND_RAMP", "VP_BAND_STAGGER")
for env in ({"VP_BANDS": "1"}, {"VP_BANDS": "4"}, {"VP_BANDS": "5"}, {"VP_BANDS": "6"}, {"VP_BANDS": "8"}, {"VP_BANDS": "6", "VP_BAND_STAGGER": "1"},
            {"VP_BANDS": "6", "VP_BAND_STAGGER": "-1"}, {"VP_BANDS": "8", "VP_BAND_STAGGER": "1"}, {"VP_BANDS": "12"}):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    s = vp.Session(lib, 0); s.set_scene(sc)
    wall = []
    for i in range(16):
        fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0)
        t0 = time.perf_counter()
        lib.check(lib.c_render(s.ctx, C.byref(cc), C.byref(fp), h_acc.data_ptr(), h_rgb.data_ptr(), None), "render")
        t1 = time.perf_counter()
        if i >= 6:
            wall.append(1e3 * (t1 - t0))
    print(env, f"vp_render wall {np.mean(wall):.3f} ms (min {np.min(wall):.3f})", flush=True)
    s.close()
