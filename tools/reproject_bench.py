"""Reprojection frame mode (vp_render_reproject_device) at a config's full size: ms/frame and Mrays/s on the GPU with the
history ping-pong resident in HBM, the reference's CPU path beside it, and a bit comparison of the last frame.
Prints one JSON line."""
import ctypes as C, importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
host = importlib.import_module("raytracer-voxpopuli_b200.host")
abi = vp.abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lib = vp.load_product()
sc, cam, fpd = scenes.config(cfg)
W, H = fpd["width"], fpd["height"]
pos = np.array(cam.cam_pos, np.float32)
tl, tr, bl = (np.array(v, np.float32) for v in (cam.top_left, cam.top_right, cam.bottom_left))
target = tl + (tr - tl) * np.float32(0.5) + (bl - tl) * np.float32(0.5)
step = np.array([0.002, 0.001, -0.0015], np.float32)
cams = [host.look_at(pos + step * f, target, W, H) for f in range(frames + 1)]
prevs = [host.prev_camera(pos + step * max(f - 1, 0), target, W, H) for f in range(frames + 1)]

def fp_of(f):
    return vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=f, seed=scenes.SEED, aa_strength=1.0)

dev = torch.device("cuda", 0)
hist = [torch.zeros((H, W, 3), dtype=torch.float32, device=dev) for _ in range(2)]
rgb = torch.zeros((H, W), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream(dev)
s = vp.Session(lib, 0); s.set_scene(sc)
def gpu_frame(f):
    cc = cams[f].to_c(); fp = fp_of(f)
    lib.check(lib.c_render_reproject_device(s.ctx, C.byref(cc), C.byref(prevs[f]), C.byref(fp), hist[f % 2].data_ptr(), hist[(f + 1) % 2].data_ptr(),
                                            rgb.data_ptr(), st.cuda_stream, None), "reproject")
for f in range(3): gpu_frame(f)
torch.cuda.synchronize()
for h in hist: h.zero_()
lib.check(lib.c_stats_begin(s.ctx), "b")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st)
for f in range(frames): gpu_frame(f)
b.record(st); torch.cuda.synchronize()
ms = a.elapsed_time(b) / frames
tot = abi.VpStats(); lib.check(lib.c_stats_end(s.ctx, C.byref(tot)), "e")
gpu_hist = hist[frames % 2].cpu().numpy(); gpu_rgb = rgb.cpu().numpy().view(np.uint32)

# CPU reference (or port) on the same frames: time the last two, compare the last one
ref = os.path.join(ROOT, "oracle", "_ref", "libvpref.so")
clib, kind = (abi.Lib(ref, "vpref_"), "reference") if os.path.exists(ref) else (abi.Lib(os.path.join(ROOT, "oracle", "libvporacle.so"), "vpo_"), "port")
cpu_frames = min(frames, 4)
# a fair, bounded check: replay the whole GPU sequence on the CPU only when it is short
full_check = frames <= 6
with vp.Session(clib) as cs:
    cs.set_scene(sc)
    h = None; times = []; rays = []
    for f in range(frames if full_check else cpu_frames):
        t0 = time.perf_counter()
        h, crgb, _, _, cst = cs.render_reproject(cams[f], prevs[f], fp_of(f), h, want_buffers=False)
        times.append(time.perf_counter() - t0); rays.append(cst.rays)
line = {"mode": "reprojection (Renderer::Tick static-camera branch)", "workload": scenes.WORKLOAD_NAMES[cfg], "frames": frames,
        "gpu_ms_per_frame": ms, "gpu_mrays_per_s": tot.rays / frames / (ms * 1e-3) / 1e6, "rays_per_frame": tot.rays / frames,
        "cpu_kind": kind, "cpu_ms_per_frame": 1e3 * min(times[1:] or times), "cpu_mrays_per_s": rays[-1] / min(times[1:] or times) / 1e6}
if full_check:
    line["last_frame_history_bit_identical"] = bool(np.array_equal(gpu_hist.view(np.uint32), h.view(np.uint32)))
    line["last_frame_rgb8_identical"] = bool(np.array_equal(gpu_rgb, crgb))
print(json.dumps(line))
