"""CUDA library (through the C ABI) against the oracle on the same seeded inputs, against
the committed goldens, and - at BASELINE.json's full sizes - through checksums and
size-independent properties.  Bars (BASELINE.json north_star): primary hit
{volume, cell, t, normal, material, face} bit-exact; shaded RGB |d| <= 1e-3 per channel;
accumulated frame relative L2 <= 1e-4.  The implementation folds colours in the reference's
operation order, so in practice the accumulators are bit-identical; the tests assert the
stated bars and additionally that (almost) all pixels are bit-identical."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest
from _util import GOLDEN, abi, assert_hits_equal, bits, params, scenes, small_config, vp

pytestmark = pytest.mark.gpu

RGB_ABS_TOL = 1e-3      # north_star: shaded RGB within 1e-3 absolute per channel
ACCUM_REL_L2_TOL = 1e-4  # north_star: relative L2 of accumulated frames


def check_accum(a, b, what):
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    rel = np.linalg.norm((a - b).astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30)
    nbit = int(np.count_nonzero((bits(a) != bits(b)).any(axis=-1)))
    npx = a.shape[0] * a.shape[1]
    print(f"{what}: max|d|={d.max():.3g} relL2={rel:.3g} pixels not bit-identical={nbit}/{npx}")
    assert d.max() <= RGB_ABS_TOL, f"{what}: max abs diff {d.max()}"
    assert rel <= ACCUM_REL_L2_TOL, f"{what}: rel L2 {rel}"
    assert nbit <= max(1, npx // 1000), f"{what}: {nbit} pixels differ bitwise"


def hit_digest(hits):
    h = hashlib.sha256()
    for f in ("volume", "cell", "t", "normal", "material", "face"):
        h.update(np.ascontiguousarray(hits[f]).tobytes())
    return h.hexdigest()


CASES = [
    (1, 160, 90, {}),
    (1, 160, 90, dict(max_bounces=4, spp=2)),
    (2, 160, 90, dict(spp=2)),
    (3, 128, 72, dict(spp=3)),
    (4, 160, 90, dict(spp=2)),
    (5, 96, 54, dict(spp=1)),
]


@pytest.mark.parametrize("pipeline", ["staged", "megakernel"])
@pytest.mark.parametrize("idx,w,h,over", CASES)
def test_config_vs_oracle(product, oracle, idx, w, h, over, pipeline):
    """Both formulations of the path against the oracle: wavefront stages as kernels (default), single-kernel walk
    (the single persistent-kernel formulation of round 1 was retired: it lost its A/B at every frame size)."""
    sc, cam, fpd = small_config(idx, w, h, **over)
    flags = abi.VP_FLAG_COUNT_STEPS | {"staged": 0, "megakernel": abi.VP_FLAG_MEGAKERNEL}[pipeline]
    res = []
    for lib in (oracle, product):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            hits = s.trace_primary(cam, params(fpd))
            acc, rgb8, st = s.render(cam, params(fpd, flags=flags))
            res.append((hits, acc, rgb8, st))
    assert_hits_equal(res[0][0], res[1][0], f"cfg{idx}")
    check_accum(res[1][1], res[0][1], f"cfg{idx} accum")
    assert np.count_nonzero(res[0][2] != res[1][2]) <= max(1, w * h // 1000)
    a, b = res[0][3], res[1][3]
    for k in ("primary_rays", "secondary_rays", "shadow_rays", "exit_rays", "dda_steps", "sky_samples"):
        assert getattr(a, k) == getattr(b, k), (k, getattr(a, k), getattr(b, k))
    assert b.kernel_launches >= 1


@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5])
def test_config_vs_golden(product, idx):
    g = np.load(os.path.join(GOLDEN, f"cfg{idx}_64x36.npz"))
    sc, cam, fpd = small_config(idx, 64, 36)
    with vp.Session(product) as s:
        s.set_scene(sc)
        hits = s.trace_primary(cam, params(fpd))
        acc, rgb8, st = s.render(cam, params(fpd, spp=int(g["spp"])))
    assert_hits_equal(g["hits"], hits, f"cfg{idx} golden")
    check_accum(acc, g["accum"], f"cfg{idx} golden accum")
    assert [st.primary_rays, st.secondary_rays, st.shadow_rays, st.exit_rays] == list(g["stats"][:4])


@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5])
def test_fullres_primary_hits_digest(product, idx):
    """Bit-exact primary hits at BASELINE.json's full resolutions (checksum of the dump made
    by the reference build)."""
    want = json.load(open(os.path.join(GOLDEN, "fullres_hit_digests.json")))[f"cfg{idx}"]
    sc, cam, fpd = scenes.config(idx)
    with vp.Session(product) as s:
        s.set_scene(sc)
        hits = s.trace_primary(cam, params(fpd))
    assert int((hits["volume"] >= 0).sum()) == want["hit_pixels"]
    assert hit_digest(hits) == want["sha256"]


@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5])
def test_fullres_frames_digest(product, idx):
    """Bit-exact FRAMES at BASELINE.json's full sizes: the accumulator and RGB8 image after two accumulated samples, the
    ray / step counters, and one reprojection frame on top (history + RGB8), against SHA-256 digests of what the
    reference build produced (tools/make_golden.py)."""
    from _util import host
    want = json.load(open(os.path.join(GOLDEN, "fullres_frame_digests.json")))[f"cfg{idx}"]
    sc, cam, fpd = scenes.config(idx)
    W, H = fpd["width"], fpd["height"]
    assert (W, H) == (want["width"], want["height"])
    with vp.Session(product) as s:
        s.set_scene(sc)
        acc, rgb8, st = s.render(cam, params(fpd, spp=2, flags=abi.VP_FLAG_COUNT_STEPS))
        prev = host.prev_camera(np.array(cam.cam_pos, np.float32) + np.array([0.01, 0, 0], np.float32), scenes.CAMERA_TARGETS[idx], W, H)
        hist, rrgb, _, _, rst = s.render_reproject(cam, prev, params(fpd, spp=1, frame_index=3), np.ascontiguousarray(acc[:, :, :3]),
                                                   want_buffers=False)
    assert [st.primary_rays, st.secondary_rays, st.shadow_rays, st.exit_rays, st.dda_steps] == want["stats"]
    assert hashlib.sha256(acc.tobytes()).hexdigest() == want["accum_sha256"]
    assert hashlib.sha256(rgb8.tobytes()).hexdigest() == want["rgb8_sha256"]
    assert [rst.primary_rays, rst.secondary_rays, rst.shadow_rays, rst.exit_rays] == want["reproject_rays"]
    assert hashlib.sha256(hist.tobytes()).hexdigest() == want["reproject_history_sha256"]
    assert hashlib.sha256(rrgb.tobytes()).hexdigest() == want["reproject_rgb8_sha256"]


def test_batch_seams_and_prims(product, oracle):
    sc, cam, fpd = small_config(4, 64, 36)
    sc.spheres = [((0.2, 0.6, -0.5), 0.3, 5), ((1.0, 0.4, 0.2), 0.25, 1), ((-0.8, 0.5, 0.1), 0.2, 20)]
    sc.triangles = [((-0.5, 0, 0), (0, 0.6, 0), (0.5, 0, 0), (0.4, 0.1, -0.9), 7), ((-1, 0, 0), (0, 1, 0), (1, 0, 0), (0.0, 0.2, 3.5), 30)]
    rng = np.random.default_rng(3)
    n = 50000
    o = rng.uniform(-2.5, 3.0, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays = vp.make_rays(o, d)
    rays["t"][::3] = rng.uniform(0.2, 3.0, rays["t"][::3].shape).astype(np.float32)
    seeds = rng.integers(1, 2 ** 32 - 1, n, dtype=np.uint64).astype(np.uint32)
    res = []
    for lib in (oracle, product):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            res.append((s.find_nearest(rays), s.is_occluded(rays), s.trace(rays, seeds, 6)))
    assert_hits_equal(res[0][0], res[1][0], "find_nearest")
    assert np.array_equal(res[0][1], res[1][1])
    assert np.abs(res[0][2] - res[1][2]).max() <= RGB_ABS_TOL
    assert np.count_nonzero((bits(res[0][2]) != bits(res[1][2])).any(axis=1)) <= n // 1000


def test_edge_cases(product, oracle):
    """Empty volume list, empty grid, 1^3 volumes (the reference's scaled single-voxel
    platforms, renderer.cpp:482-530), ragged resolution, rays starting inside solids,
    axis-aligned rays (NaN tmax path), zero spp."""
    full = np.zeros(1, np.uint8)  # NON_METAL_WHITE
    M, iM, a, b = __import__("_util").host.make_transform((0, 0, 1), (0, 0.4, 0), (3.0, 0.5, 2.0))
    one = vp.Volume(full, 1, a, b, M, iM)
    M2, iM2, a2, b2 = __import__("_util").host.make_transform((-1, 0, 0))
    empty = vp.Volume(np.full(8 ** 3, 255, np.uint8), 8, a2, b2, M2, iM2)
    g7 = np.full(7 ** 3, 255, np.uint8)
    g7[::5] = 17
    M3, iM3, a3, b3 = __import__("_util").host.make_transform((1, 0, 0))
    odd = vp.Volume(g7, 7, a3, b3, M3, iM3)
    mats = __import__("_util").host.default_materials()
    for vols in ([], [empty], [one], [one, empty, odd]):
        sc = vp.SceneDesc(volumes=vols, materials=mats, point_lights=[((0.5, 3.0, 0.0), (5, 5, 5))])
        cam = __import__("_util").host.look_at((0.5, 1.2, -2.0), (0.5, 0.2, 1.2), 61, 37)
        fp = vp.frame_params(61, 37, max_bounces=5, spp=2)
        res = []
        for lib in (oracle, product):
            with vp.Session(lib) as s:
                s.set_scene(sc)
                rays = vp.make_rays([[0.5, 0.5, 1.5], [0.5, 2.0, 1.5], [0.25, 0.5, -1.0], [1.2, 0.3, 0.3]],
                                    [[0, -1, 0], [0, -1, 0], [0, 0, 1], [1, 0, 0]])
                res.append((s.trace_primary(cam, fp), s.render(cam, fp)[0], s.find_nearest(rays), s.is_occluded(rays)))
        assert_hits_equal(res[0][0], res[1][0], f"{len(vols)} volumes")
        check_accum(res[1][1], res[0][1], f"{len(vols)} volumes")
        assert_hits_equal(res[0][2], res[1][2], "axis rays")
        assert np.array_equal(res[0][3], res[1][3])
    with vp.Session(product) as s:
        s.set_scene(vp.SceneDesc(volumes=[one], materials=mats))
        acc = np.full((37, 61, 4), 0.5, np.float32)
        out, _, st = s.render(cam, vp.frame_params(61, 37, spp=0), accum=acc.copy())
        assert np.array_equal(out, acc) and st.primary_rays == 0


def test_properties_fullsize(product):
    """Size-independent properties at full resolution (cfg2, 1920x1080):
    - tile sharding: rendering interleaved row tiles in separate calls == one call (image independent of #GPUs);
    - spp splitting: 2 passes in one call == 2 calls with frame_index advanced (running average is sequential);
    - host-buffer call == device-resident call; trace_primary is idempotent."""
    sc, cam, fpd = scenes.config(2)
    W, H = fpd["width"], fpd["height"]
    with vp.Session(product) as s:
        s.set_scene(sc)
        whole, rgb_whole, st = s.render(cam, params(fpd, spp=2))
        parts = np.zeros_like(whole)
        for phase in range(4):
            parts, _, _ = s.render(cam, params(fpd, spp=2, row_tile=8, row_tile_stride=4, row_tile_phase=phase), accum=parts)
        assert np.array_equal(bits(whole), bits(parts))
        seq, _, _ = s.render(cam, params(fpd, spp=1))
        seq, rgb_seq, _ = s.render(cam, params(fpd, spp=1, frame_index=1), accum=seq)
        assert np.array_equal(bits(whole), bits(seq)) and np.array_equal(rgb_whole, rgb_seq)
        L = product
        L.check(L.c_accum_reset(s.ctx, W, H), "accum_reset")
        cc, fp = cam.to_c(), params(fpd, spp=2)
        st2 = abi.VpStats()
        L.check(L.c_render_resident(s.ctx, C.byref(cc), C.byref(fp), C.byref(st2)), "render_resident")
        res = np.zeros_like(whole)
        rgb = np.zeros((H, W), np.uint32)
        L.check(L.c_accum_download(s.ctx, res.ctypes.data, rgb.ctypes.data), "accum_download")
        assert np.array_equal(bits(whole), bits(res)) and np.array_equal(rgb, rgb_whole)
        assert st2.rays == st.rays and st.rays > W * H
        h1 = s.trace_primary(cam, params(fpd))
        h2 = s.trace_primary(cam, params(fpd))
        assert_hits_equal(h1, h2, "idempotence")
        assert np.isfinite(whole).all() and (whole[..., :3] >= 0).all()


def test_dirty_updates(product, oracle):
    """vp_update_volume_grid / vp_update_volume_transform (Scene::Set + SetTransform between frames)."""
    from _util import host
    sc, cam, fpd = small_config(4, 96, 54)
    fp = params(fpd, spp=1)
    with vp.Session(product) as s:
        s.set_scene(sc)
        v = sc.volumes[4]
        g = v.grid.copy()
        g[1000:5000] = 3
        product.check(product.c_update_volume_grid(s.ctx, 4, 1000, 4000, g[1000:5000].ctypes.data), "update_volume_grid")
        M, iM, a, b = host.make_transform((-0.75, 0.25, 0.5), (0.1, 0.7, 0.0))
        product.check(product.c_update_volume_transform(s.ctx, 6, M.ctypes.data, iM.ctypes.data, a.ctypes.data, b.ctypes.data),
                      "update_volume_transform")
        got_hits, got = s.trace_primary(cam, fp), s.render(cam, fp)[0]
    sc.volumes[4].grid = g
    sc.volumes[6].matrix, sc.volumes[6].inv_matrix, sc.volumes[6].cube_min, sc.volumes[6].cube_max = M, iM, a, b
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        want_hits, want = s.trace_primary(cam, fp), s.render(cam, fp)[0]
    assert_hits_equal(want_hits, got_hits, "after dirty update")
    check_accum(got, want, "after dirty update")


def test_error_paths(product):
    with vp.Session(product) as s:
        sc, cam, fpd = small_config(1, 32, 18)
        cc = cam.to_c()
        fp = params(fpd)
        acc = np.zeros((18, 32, 4), np.float32)
        assert product.c_render(s.ctx, C.byref(cc), C.byref(fp), acc.ctypes.data, None, None) == abi.VP_ESTATE
        assert b"vp_set_volumes" in product.c_last_error()
        s.set_scene(sc)
        bad = params(fpd, max_bounces=40)
        assert product.c_render(s.ctx, C.byref(cc), C.byref(bad), acc.ctypes.data, None, None) == abi.VP_EINVAL
        assert product.c_render(s.ctx, C.byref(cc), C.byref(fp), None, None, None) == abi.VP_EINVAL
        info = C.create_string_buffer(1024)
        product.check(product.c_info(s.ctx, info, 1024), "info")
        meta = json.loads(info.value.decode())
        assert meta["sm"] == "10.0" and meta["k_render_regs"] > 0


def test_cpp_renderer_mirror(product, oracle):
    """The C++ host mirror (voxpop::Renderer / Scene / Camera over the C ABI) driven like reference
    host code: Scene + SetTransform, lights, camera.LookAt, UploadScene, Update() x frames."""
    import ctypes as C
    from _util import host
    so = os.path.join(os.path.dirname(vp.PRODUCT_SO), "libvoxpop_renderer.so")
    dll = C.CDLL(so)
    dll.vphr_demo_render.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.POINTER(C.c_ulonglong)]
    grid, n, _ = scenes.load_grid("teapot")
    W, H, frames, mb = 96, 54, 3, 5
    acc = np.zeros((H, W, 4), np.float32)
    rgb = np.zeros((H, W), np.uint32)
    rays = C.c_ulonglong()
    assert dll.vphr_demo_render(None, grid.ctypes.data, n, W, H, frames, mb, acc.ctypes.data, rgb.ctypes.data, C.byref(rays)) == 0
    M, iM, a, b = host.make_transform((0, 0, 0), (0, 0.3, 0))
    sc = vp.SceneDesc(volumes=[vp.Volume(grid, n, a, b, M, iM)], materials=host.default_materials(),
                      point_lights=[((0.5, 1.6, -0.6), (4, 4, 4))])
    cam = host.look_at((0.5, 0.55, -0.75), (0.5, 0.3, 0.5), W, H)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        want, want_rgb, st = s.render(cam, vp.frame_params(W, H, max_bounces=mb, spp=frames))
    check_accum(acc, want, "C++ mirror")
    assert np.count_nonzero(rgb != want_rgb) == 0 and rays.value == st.rays


def test_frames_in_flight(product, oracle):
    """vp_render_device with VP_FLAG_DEFER: several frames overlap on the library's stream pairs; they must resolve
    in call order, i.e. give the accumulator of the same frames rendered one after the other (oracle: one call
    with spp = number of frames).  Mixes deferred single-sample calls, a multi-sample call (more samples than
    lanes), a non-deferred call with stats in the middle and a scene edit (which has to wait for the frames)."""
    import torch
    sc, cam, fpd = small_config(2, 160, 90)
    W, H = fpd["width"], fpd["height"]
    plan = [(1, True), (1, True), (1, True), (1, True), (5, True), (1, False), (1, True), (2, True), (1, True)]
    total = sum(n for n, _ in plan)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        want, want_rgb, st_want = s.render(cam, params(fpd, spp=total))
    dev = torch.device("cuda", 0)
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    rgb = torch.zeros((H, W), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)
    cc = cam.to_c()
    with vp.Session(product) as s:
        s.set_scene(sc)
        product.check(product.c_stats_begin(s.ctx), "stats_begin")
        frame = 0
        for n, defer in plan:
            fp = params(fpd, spp=n, frame_index=frame, flags=abi.VP_FLAG_DEFER if defer else 0)
            st = None if defer else abi.VpStats()
            product.check(product.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), stream.cuda_stream,
                                                  C.byref(st) if st is not None else None), "render_device")
            frame += n
        product.check(product.c_render_fence(s.ctx, stream.cuda_stream, 0), "render_fence")
        got = acc.cpu().numpy()
        got_rgb = rgb.cpu().numpy().view(np.uint32)
        tot = abi.VpStats()
        product.check(product.c_stats_end(s.ctx, C.byref(tot)), "stats_end")
        assert tot.rays == st_want.rays
        # a scene edit while frames are in flight waits for them, and later frames see the edit
        for i in range(3):
            fp = params(fpd, spp=1, frame_index=100 + i, flags=abi.VP_FLAG_DEFER)
            product.check(product.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), stream.cuda_stream, None),
                          "render_device")
        s.set_scene(sc)
        assert product.c_render_fence(s.ctx, stream.cuda_stream, 8) != 0  # older than the ring of completion events
        product.check(product.c_render_fence(s.ctx, stream.cuda_stream, 7), "render_fence")
        torch.cuda.synchronize(dev)
    assert np.array_equal(bits(got), bits(want)), "frames in flight changed the accumulator"
    assert np.array_equal(got_rgb, want_rgb)


@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5])
def test_reprojection_vs_oracle(product, oracle, idx):
    """Second frame mode (vp_render_reproject): pass-1 buffers, the reprojected history and the RGB8 image of three
    drifting-camera frames against the oracle - bit-identical, counters included - and against the committed goldens
    generated from the reference build."""
    from _util import reproject_sequence
    want = reproject_sequence(oracle, idx, 96, 54)
    got = reproject_sequence(product, idx, 96, 54)
    for f, (w, g) in enumerate(zip(want, got)):
        for k, name in ((2, "albedo"), (3, "illumination"), (0, "history")):
            d = np.abs(w[k].astype(np.float64) - g[k].astype(np.float64))
            nbit = int(np.count_nonzero((bits(w[k]) != bits(g[k])).any(axis=-1)))
            print(f"cfg{idx} frame {f} {name}: max|d|={d.max():.3g} pixels not bit-identical={nbit}")
            assert d.max() <= RGB_ABS_TOL and nbit == 0, f"cfg{idx} frame {f}: {name}"
        assert np.array_equal(w[1], g[1]), f"cfg{idx} frame {f}: rgb8"
        assert w[4][:4] == g[4][:4], f"cfg{idx} frame {f}: ray counters {w[4]} vs {g[4]}"
    if idx in (2, 3, 4):
        gold = np.load(os.path.join(GOLDEN, f"reproject_cfg{idx}_64x36.npz"))
        out = reproject_sequence(product, idx, 64, 36)
        for f, o in enumerate(out):
            assert np.array_equal(bits(gold[f"history{f}"]), bits(o[0])) and np.array_equal(gold[f"rgb8_{f}"], o[1])


def test_reprojection_device_and_errors(product, oracle):
    """vp_render_reproject_device on caller-owned history ping-pong buffers (torch tensors), interleaved with ordinary
    accumulating frames on the same context; plus the argument checks."""
    import torch
    from _util import host
    sc, cam, fpd = small_config(2, 128, 72)
    W, H = fpd["width"], fpd["height"]
    pos = np.array(cam.cam_pos, np.float32)
    tl, tr, bl = (np.array(v, np.float32) for v in (cam.top_left, cam.top_right, cam.bottom_left))
    target = tl + (tr - tl) * np.float32(0.5) + (bl - tl) * np.float32(0.5)
    prev = host.prev_camera(pos, target, W, H)
    camf = host.look_at(pos + np.array([0.02, 0, 0], np.float32), target, W, H)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        h1, _, _, _, _ = s.render_reproject(camf, prev, params(fpd, frame_index=0))
        h2, rgb2, _, _, _ = s.render_reproject(camf, prev, params(fpd, frame_index=1), h1)
    dev = torch.device("cuda", 0)
    hist = [torch.zeros((H, W, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    rgb = torch.zeros((H, W), dtype=torch.int32, device=dev)
    acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)
    cc = camf.to_c()
    with vp.Session(product) as s:
        s.set_scene(sc)
        for f in range(2):
            fp = params(fpd, frame_index=f)
            product.check(product.c_render_reproject_device(s.ctx, C.byref(cc), C.byref(prev), C.byref(fp), hist[f % 2].data_ptr(),
                                                            hist[(f + 1) % 2].data_ptr(), rgb.data_ptr(), stream.cuda_stream, None), "reproject_device")
            # an accumulating frame in between (deferred: it overlaps the next reprojection frame's start)
            fpa = params(fpd, frame_index=f, flags=abi.VP_FLAG_DEFER)
            product.check(product.c_render_device(s.ctx, C.byref(cc), C.byref(fpa), acc.data_ptr(), None, stream.cuda_stream, None), "render_device")
        product.check(product.c_render_fence(s.ctx, stream.cuda_stream, 0), "fence")
        torch.cuda.synchronize(dev)
        assert np.array_equal(bits(hist[0].cpu().numpy()), bits(h2))
        assert np.array_equal(rgb.cpu().numpy().view(np.uint32), rgb2)
        # argument checks
        fp = params(fpd, frame_index=0)
        assert product.c_render_reproject_device(s.ctx, C.byref(cc), C.byref(prev), C.byref(fp), hist[0].data_ptr(), hist[0].data_ptr(),
                                                 rgb.data_ptr(), stream.cuda_stream, None) != 0  # aliasing history
        fp = params(fpd, frame_index=0, max_bounces=-1)
        assert product.c_render_reproject_device(s.ctx, C.byref(cc), C.byref(prev), C.byref(fp), hist[0].data_ptr(), hist[1].data_ptr(),
                                                 rgb.data_ptr(), stream.cuda_stream, None) != 0
    with vp.Session(oracle) as s:  # the accumulating frames were not disturbed
        s.set_scene(sc)
        want_acc, _, _ = s.render(camf, params(fpd, spp=2))
    assert np.array_equal(bits(acc.cpu().numpy()), bits(want_acc))


def test_cpp_renderer_mirror_reproject(product, oracle):
    """voxpop::Renderer::UpdateReproject / CopyToPrevCamera (the C++ mirror of the static-camera branch of Tick)."""
    from _util import host
    so = os.path.join(os.path.dirname(vp.PRODUCT_SO), "libvoxpop_renderer.so")
    dll = C.CDLL(so)
    dll.vphr_demo_reproject.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    grid, n, _ = scenes.load_grid("teapot")
    W, H, frames, mb = 96, 54, 3, 5
    hist = np.zeros((H, W, 3), np.float32)
    rgb = np.zeros((H, W), np.uint32)
    assert dll.vphr_demo_reproject(grid.ctypes.data, n, W, H, frames, mb, hist.ctypes.data, rgb.ctypes.data) == 0
    M, iM, a, b = host.make_transform((0, 0, 0), (0, 0.3, 0))
    sc = vp.SceneDesc(volumes=[vp.Volume(grid, n, a, b, M, iM)], materials=host.default_materials(),
                      point_lights=[((0.5, 1.6, -0.6), (4, 4, 4))])
    tgt = (0.5, 0.3, 0.5)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        h = None
        for f in range(frames):
            pos = (np.float32(0.5) + np.float32(0.01) * np.float32(f), 0.55, -0.75)
            if f == 0:
                prev = host.prev_camera(pos, tgt, W, H)
            h, want_rgb, _, _, _ = s.render_reproject(host.look_at(pos, tgt, W, H), prev, vp.frame_params(W, H, max_bounces=mb, spp=1, frame_index=f), h)
            prev = host.prev_camera(pos, tgt, W, H)
    assert np.array_equal(bits(hist), bits(h)) and np.array_equal(rgb, want_rgb)


def test_gameplay_side_channels(product, oracle):
    """vp_in_light, vp_find_nearest_player, vp_focus_distance against the oracle (which is pinned to the reference's
    inLight / FindNearestPlayer / focus ray by tests/test_oracle_vs_ref.py)."""
    from _util import side_channels
    a, b = side_channels(oracle), side_channels(product)
    for k in ("idle", "after_render", "after_read", "after_reproject"):
        assert a[k] == b[k], k
    assert_hits_equal(a["player0"], b["player0"], "FindNearestPlayer")
    assert_hits_equal(a["player2"], b["player2"], "FindNearestPlayer chunk 2")
    assert np.array_equal(bits(a["focus"]), bits(b["focus"]))
    assert np.array_equal(bits(a["accum"]), bits(b["accum"]))


@pytest.mark.parametrize("world", [2, 3])
def test_reprojection_row_sharded(product, oracle, world):
    """Row-sharded reprojection (what each rank of a multi-GPU run does): every shard traces its rows plus their direct
    neighbours in pass 1, writes only its own rows in pass 2; the merged shards equal the whole frame bit for bit, also
    on the second frame, which reprojects into the merged (= gathered) history."""
    from _util import host
    sharding = __import__("importlib").import_module("raytracer-voxpopuli_b200.sharding")
    sc, cam, fpd = small_config(4, 128, 72)
    W, H = fpd["width"], fpd["height"]
    pos = np.array(cam.cam_pos, np.float32)
    tl, tr, bl = (np.array(v, np.float32) for v in (cam.top_left, cam.top_right, cam.bottom_left))
    target = tl + (tr - tl) * np.float32(0.5) + (bl - tl) * np.float32(0.5)
    cams = [host.look_at(pos + np.array([0.015 * f, 0, 0], np.float32), target, W, H) for f in range(2)]
    prevs = [host.prev_camera(pos, target, W, H), host.prev_camera(pos, target, W, H)]
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        want, hist = [], None
        for f in range(2):
            hist, rgb, _, _, st = s.render_reproject(cams[f], prevs[f], params(fpd, frame_index=f), hist)
            want.append((hist, rgb, st.rays))
    with vp.Session(product) as s:
        s.set_scene(sc)
        hist = None
        for f in range(2):
            merged_h = np.zeros((H, W, 3), np.float32)
            merged_rgb = np.zeros((H, W), np.uint32)
            for r in range(world):
                h, rgb, _, _, st = s.render_reproject(cams[f], prevs[f], params(fpd, frame_index=f, **sharding.shard_params(world, r)), hist)
                for t in sharding.tiles_of_rank(H, world, r):
                    rows = slice(t * sharding.ROW_TILE, (t + 1) * sharding.ROW_TILE)
                    merged_h[rows], merged_rgb[rows] = h[rows], rgb[rows]
            assert np.array_equal(bits(merged_h), bits(want[f][0])), f"frame {f}: merged history differs"
            assert np.array_equal(merged_rgb, want[f][1]), f"frame {f}: merged rgb8 differs"
            hist = merged_h


def test_render_row_bands(product, oracle, monkeypatch):
    """vp_render cuts a call into row bands that run as deferred passes on the library's lanes, the rows of a band
    travelling back while later bands are traced (VP_BANDS / VP_BAND_MIN_PIXELS; large frames only by default).
    Forced on for a small frame here: accumulator, RGB8 and ray counters must equal the oracle's un-banded frame,
    for one and for several samples per call, with an accumulator carried over between calls and with a row range."""
    sc, cam, fpd = small_config(3, 128, 72)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        w1, w1_rgb, st1 = s.render(cam, params(fpd, spp=1))
        w3, w3_rgb, st3 = s.render(cam, params(fpd, spp=3, frame_index=1), accum=w1.copy())
        wr, wr_rgb, str_ = s.render(cam, params(fpd, spp=2, frame_index=7, row_begin=10, row_end=61))
    for bands, stagger in (("6", "-1"), ("4", "0"), ("3", "1"), ("1", "-1")):  # the default cut, round 2's, a late start, none
        monkeypatch.setenv("VP_BANDS", bands)
        monkeypatch.setenv("VP_BAND_STAGGER", stagger)
        monkeypatch.setenv("VP_BAND_MIN_PIXELS", "0")
        with vp.Session(product) as s:  # the knobs are read when the context is created
            s.set_scene(sc)
            g1, g1_rgb, gs1 = s.render(cam, params(fpd, spp=1))
            g3, g3_rgb, gs3 = s.render(cam, params(fpd, spp=3, frame_index=1), accum=g1.copy())
            gr, gr_rgb, gsr = s.render(cam, params(fpd, spp=2, frame_index=7, row_begin=10, row_end=61))
        for a, b, what in ((w1, g1, "1 spp"), (w3, g3, "3 spp carried"), (wr, gr, "row range")):
            assert np.count_nonzero(bits(a) != bits(b)) == 0, f"bands={bands} {what}: accumulator differs"
        for a, b in ((w1_rgb, g1_rgb), (w3_rgb, g3_rgb), (wr_rgb, gr_rgb)):
            assert np.count_nonzero(a != b) == 0
        assert (gs1.rays, gs3.rays, gsr.rays) == (st1.rays, st3.rays, str_.rays)
        if bands != "1":
            assert gs1.kernel_launches > 2 * (3 * (fpd["max_bounces"] + 1) + 2)  # really several passes


@pytest.mark.parametrize("n", [20, 50, 100])
def test_grid_sizes_off_the_brick_lattice(product, oracle, n):
    """Grid sizes that are neither a power of two nor a multiple of the 4-cell brick: the brick walk runs in the
    padded power-of-two domain (mirrored coordinates) and walks the empty padding before it leaves.  Rotated, scaled
    volume with a floor and random voxels; primary hits, accumulator, batched nearest / any-hit queries against the
    oracle, bit for bit."""
    from _util import host
    rng = np.random.default_rng(n)
    g = np.full((n, n, n), 255, np.uint8)  # [z, y, x]
    g[:, 0, :] = 3
    fill = rng.uniform(size=(n, n, n)) < 0.04
    g[fill] = rng.integers(0, 8, int(fill.sum())).astype(np.uint8)
    M, iM, a, b = host.make_transform((0.25, 0.0, 0.5), (0.0, 0.6, 0.0), (1.5, 1.0, 1.25))
    sc = vp.SceneDesc(volumes=[vp.Volume(np.ascontiguousarray(g.reshape(-1)), n, a, b, M, iM)], materials=host.default_materials(),
                      point_lights=[((0.5, 2.5, -0.5), (6, 6, 6)), ((2.0, 1.5, 2.5), (3, 3, 4))])
    cam = host.look_at((0.2, 1.4, -1.6), (1.0, 0.3, 1.2), 96, 54)
    fp = vp.frame_params(96, 54, max_bounces=6, spp=2)
    o = rng.uniform(-0.5, 2.5, (4096, 3)).astype(np.float32)
    d = rng.normal(size=(4096, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[::7, 1] = 0.0  # some rays parallel to an axis plane of the world
    rays = vp.make_rays(o, d)
    res = []
    for lib in (oracle, product):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            res.append((s.trace_primary(cam, fp), s.render(cam, fp), s.find_nearest(rays), s.is_occluded(rays)))
    assert_hits_equal(res[0][0], res[1][0], f"n={n} primary")
    assert np.count_nonzero(bits(res[0][1][0]) != bits(res[1][1][0])) == 0
    assert res[0][1][2].rays == res[1][1][2].rays
    assert_hits_equal(res[0][2], res[1][2], f"n={n} nearest")
    assert np.array_equal(res[0][3], res[1][3])
    assert np.count_nonzero(res[1][0]["volume"] >= 0) > 500  # the camera really sees the volume


@pytest.mark.parametrize("env", [{"VP_GRAPH": "0"}, {"VP_BRICKWALK": "0"}, {"VP_GRAPH": "0", "VP_BRICKWALK": "0", "VP_BANDS": "1"}])
def test_fallback_paths(product, oracle, monkeypatch, env):
    """The A/B switches keep working: kernels launched from the host instead of the per-lane CUDA graph, the cell walk
    instead of the brick walk, no row bands.  Multi-volume config with area lights, several samples (lanes rotate)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sc, cam, fpd = small_config(4, 128, 72)
    res = []
    for lib in (oracle, product):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            res.append((s.trace_primary(cam, params(fpd)), s.render(cam, params(fpd, spp=5))))
    assert_hits_equal(res[0][0], res[1][0], str(env))
    assert np.count_nonzero(bits(res[0][1][0]) != bits(res[1][1][0])) == 0
    assert np.count_nonzero(res[0][1][1] != res[1][1][1]) == 0
    assert res[0][1][2].rays == res[1][1][2].rays


def test_graph_recapture(product, oracle):
    """Deferred frames go through a per-lane CUDA graph whose topology depends on the bounce count and whose nodes point
    into the lane's buffers: alternate bounce counts, then a larger frame (the lanes reallocate), then a scene with
    another volume count, all with frames in flight, and compare every accumulator with the oracle."""
    import torch
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(dev)

    def run_gpu(s, cam, fpd, plan, W, H):
        acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        rgb = torch.zeros((H, W), dtype=torch.int32, device=dev)
        cc = cam.to_c()
        for i, mb in enumerate(plan):
            fp = params(fpd, spp=1, frame_index=i, max_bounces=mb, flags=abi.VP_FLAG_DEFER)
            product.check(product.c_render_device(s.ctx, C.byref(cc), C.byref(fp), acc.data_ptr(), rgb.data_ptr(), stream.cuda_stream, None),
                          "render_device")
        product.check(product.c_render_fence(s.ctx, stream.cuda_stream, 0), "render_fence")
        return acc.cpu().numpy()

    def run_cpu(sc, cam, fpd, plan, W, H):
        acc = np.zeros((H, W, 4), np.float32)
        with vp.Session(oracle) as s:
            s.set_scene(sc)
            for i, mb in enumerate(plan):
                acc, _, _ = s.render(cam, params(fpd, spp=1, frame_index=i, max_bounces=mb), accum=acc)
        return acc

    plan = [3, 3, 5, 3, 5, 5, 3, 0, 5, 3, 3]
    with vp.Session(product) as s:
        for idx, (W, H) in ((2, (96, 54)), (2, (160, 90)), (4, (128, 72)), (2, (64, 36))):
            sc, cam, fpd = small_config(idx, W, H)
            s.set_scene(sc)
            got = run_gpu(s, cam, fpd, plan, W, H)
            want = run_cpu(sc, cam, fpd, plan, W, H)
            assert np.count_nonzero(bits(got) != bits(want)) == 0, f"cfg {idx} {W}x{H}"
