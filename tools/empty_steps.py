"""Where do the DDA steps of the extend stage go?  (instrumented pass: steps inside empty 16^3 regions / empty 4^3 bricks)"""
import ctypes as C, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
lib = vp.load_product(); abi = vp.abi
for cfg in (2, 4):
    sc, cam, fpd = scenes.config(cfg)
    W, H = fpd["width"], fpd["height"]
    s = vp.Session(lib, 0); s.set_scene(sc); cc = cam.to_c()
    acc = torch.zeros((H, W, 4), device="cuda"); rgb = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    lib.check(lib.c_stats_begin(s.ctx), "b")
    fp = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=0, seed=scenes.SEED, aa_strength=1.0, flags=abi.VP_FLAG_COUNT_STEPS)
    st = abi.VpStats()
    lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), torch.cuda.current_stream().cuda_stream, C.byref(st)), "r")
    b = C.create_string_buffer(4096); lib.c_info(s.ctx, b, 4096); m = json.loads(b.value.decode())
    print(cfg, "extend steps", m["extend_steps"], "in empty region %.3f" % (m["dbg_extend_steps_in_empty_region"] / m["extend_steps"]),
          "in empty brick %.3f" % (m["dbg_extend_steps_in_empty_brick"] / m["extend_steps"]), "all steps", st.dda_steps, "rays", st.rays)
    s.close()
