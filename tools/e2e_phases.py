"""Where an end-to-end vp_render call spends its time: wall time of the call against the device time between its first
launch and its last resolve (VpStats.device_ms), for a few band settings (environment is read at context creation)."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
abi = vp.abi
lib = vp.load_product()
sc, cam, fpd = scenes.config(2)
W, H = fpd["width"], fpd["height"]
cc = cam.to_c()
h_acc = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
h_rgb = torch.zeros((H, W), dtype=torch.int32).pin_memory()
for env in ({"VP_BANDS": "1"}, {"VP_BANDS": "4"}, {"VP_BANDS": "4", "VP_BAND_RAMP": "0"}, {"VP_BANDS": "2"}):
    for k in ("VP_BANDS", "VP_BAND_RAMP"):
        os.environ.pop(k, None)
    os.environ.update(env)
    s = vp.Session(lib, 0); s.set_scene(sc)
    wall, dev, wall_ns = [], [], []
    for i in range(14):
        fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0)
        st = abi.VpStats()
        t0 = time.perf_counter()
        lib.check(lib.c_render(s.ctx, C.byref(cc), C.byref(fp), h_acc.data_ptr(), h_rgb.data_ptr(), C.byref(st)), "render")
        t1 = time.perf_counter()
        lib.check(lib.c_render(s.ctx, C.byref(cc), C.byref(fp), h_acc.data_ptr(), h_rgb.data_ptr(), None), "render")
        t2 = time.perf_counter()
        if i >= 4:
            wall.append(1e3 * (t1 - t0)); dev.append(st.device_ms); wall_ns.append(1e3 * (t2 - t1))
    print(env, f"wall {np.mean(wall):.3f} ms  device(first launch..last resolve) {np.mean(dev):.3f} ms  wall without stats {np.mean(wall_ns):.3f} ms", flush=True)
    s.close()
