"""Throughput of every SURVEY.md 8(d) config at its full size on one GPU (frames in flight, 1 spp per frame) beside the
reference's CPU path on the host cores (one frame each).  One JSON line per config."""
import ctypes as C, importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
abi = vp.abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = vp.load_product()
ref = os.path.join(ROOT, "oracle", "_ref", "libvpref.so")
clib, kind = (abi.Lib(ref, "vpref_"), "reference") if os.path.exists(ref) else (abi.Lib(os.path.join(ROOT, "oracle", "libvporacle.so"), "vpo_"), "port")
which = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4, 5]
for cfg in which:
    sc, cam, fpd = scenes.config(cfg)
    W, H = fpd["width"], fpd["height"]
    K = 24 if cfg != 5 else 8
    s = vp.Session(lib, 0); s.set_scene(sc); cc = cam.to_c()
    acc = torch.zeros((H, W, 4), device="cuda"); rgb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream()
    def frame(i):
        fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, seed=scenes.SEED, aa_strength=1.0, flags=abi.VP_FLAG_DEFER)
        lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), st.cuda_stream, None), "render")
    for i in range(4): frame(i)
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence"); torch.cuda.synchronize()
    lib.check(lib.c_stats_begin(s.ctx), "b")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for i in range(K): frame(i)
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence")
    b.record(st); torch.cuda.synchronize()
    tot = abi.VpStats(); lib.check(lib.c_stats_end(s.ctx, C.byref(tot)), "e")
    ms = a.elapsed_time(b) / K
    free, total = torch.cuda.mem_get_info()
    s.close(); del acc, rgb; torch.cuda.empty_cache()
    with vp.Session(clib) as cs:
        cs.set_scene(sc)
        t0 = time.perf_counter()
        _, _, cst = cs.render(cam, vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=0, seed=scenes.SEED, aa_strength=1.0), want_rgb8=False)
        cpu = time.perf_counter() - t0
    print(json.dumps({"workload": scenes.WORKLOAD_NAMES[cfg], "gpu_ms_per_frame": round(ms, 3), "gpu_mrays_per_s": round(tot.rays / K / (ms * 1e-3) / 1e6, 1),
                      "rays_per_frame": int(tot.rays / K), "cpu_kind": kind, "cpu_ms_per_frame": round(1e3 * cpu, 1), "cpu_mrays_per_s": round(cst.rays / cpu / 1e6, 1),
                      "speedup": round(1e3 * cpu / ms, 1), "hbm_in_use_gb": round((total - free) / 1e9, 1)}), flush=True)
