"""Multi-device context on every GPU of the box: steady-state ms per frame of vp_render_async (host buffers, one process)
and of deferred vp_render_device (image on the first GPU, peer stores), with and without a host thread per device."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes"); abi = vp.abi
lib = vp.load_product()
n = torch.cuda.device_count()
sc, cam, fpd = scenes.config(2)
W, H = fpd["width"], fpd["height"]
cc = cam.to_c()
def fp(i, flags=0): return vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, flags=flags)
with vp.Session(lib, list(range(n))) as s:
    s.set_scene(sc)
    acc = [torch.zeros((H, W, 4)).pin_memory() for _ in range(2)]; rgb = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    tk = []
    def issue(i):
        t = C.c_uint64(); f = fp(i)
        lib.check(lib.c_render_async(s.ctx, C.byref(cc), C.byref(f), acc[i & 1].data_ptr(), acc[(i + 1) & 1].data_ptr(), rgb[i & 1].data_ptr(), C.byref(t)), "async")
        tk.append(t.value)
        if len(tk) > 3: lib.check(lib.c_render_wait(s.ctx, tk[-4]), "wait")
    for i in range(10): issue(i)
    lib.check(lib.c_render_wait(s.ctx, tk[-1]), "wait")
    t0 = time.perf_counter()
    for i in range(10, 70): issue(i)
    t_issue = time.perf_counter() - t0
    lib.check(lib.c_render_wait(s.ctx, tk[-1]), "wait")
    dt = time.perf_counter() - t0
    print(f"{n} GPUs, vp_render_async: {1e3 * dt / 60:.3f} ms/frame (host issue {1e3 * t_issue / 60:.3f} ms/frame)  threads={os.environ.get('VP_GROUP_THREADS', '1')}")
    dev = torch.device("cuda", 0)
    dacc = torch.zeros((H, W, 4), device=dev); drgb = torch.zeros((H, W), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)
    def dframe(i):
        f = fp(i, abi.VP_FLAG_DEFER)
        lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(f), dacc.data_ptr(), drgb.data_ptr(), st.cuda_stream, None), "rd")
    for i in range(10): dframe(i)
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence"); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(10, 70): dframe(i)
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence"); e1.record(st); torch.cuda.synchronize(dev)
    print(f"{n} GPUs, vp_render_device (image on GPU 0, peer stores): {e0.elapsed_time(e1) / 60:.3f} ms/frame")
