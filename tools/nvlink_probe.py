"""Two-device context, a few deferred frames into an image on the first GPU (peer stores from the second GPU's resolve
stage): meant to run under  ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum -k regex:k_wf_resolve."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
vp = importlib.import_module("raytracer-voxpopuli_b200"); scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes"); abi = vp.abi
lib = vp.load_product()
sc, cam, fpd = scenes.config(2)
W, H = fpd["width"], fpd["height"]
cc = cam.to_c()
with vp.Session(lib, [0, 1]) as s:
    s.set_scene(sc)
    dev = torch.device("cuda", 0)
    dacc = torch.zeros((H, W, 4), device=dev); drgb = torch.zeros((H, W), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev)
    for i in range(4):
        f = vp.frame_params(W, H, max_bounces=fpd["max_bounces"], spp=1, frame_index=i, flags=abi.VP_FLAG_DEFER)
        lib.check(lib.c_render_device(s.ctx, C.byref(cc), C.byref(f), dacc.data_ptr(), drgb.data_ptr(), st.cuda_stream, None), "rd")
    lib.check(lib.c_render_fence(s.ctx, st.cuda_stream, 0), "fence"); torch.cuda.synchronize(dev)
    print("ok", int(drgb.sum().item()) != 0)
