"""Shared test helpers: library loading (product + checkers) and comparison utilities.
The checkers (oracle/libvporacle.so, oracle/_ref/libvpref.so) are test infrastructure and are
loaded ONLY from here, from __graft_entry__.smoke() and from bench.py's baseline legs."""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
vp = importlib.import_module("raytracer-voxpopuli_b200")
abi = vp.abi
scenes = importlib.import_module("raytracer-voxpopuli_b200.scenes")
host = importlib.import_module("raytracer-voxpopuli_b200.host")

ORACLE_SO = os.path.join(ROOT, "oracle", "libvporacle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libvpref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def oracle_lib():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "vp_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libvporacle.so"])
    lib = abi.Lib(ORACLE_SO, "vpo_")
    _pin_protos(lib, "vpo_")
    return lib


def ref_lib():
    """The reference itself (built by oracle/ref/build_ref.py); None when it is not available."""
    if not os.path.exists(REF_SO):
        return None
    lib = abi.Lib(REF_SO, "vpref_")
    _pin_protos(lib, "vpref_")
    d = lib.dll
    d.vpref_load_vox.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
    d.vpref_make_transform.argtypes = [C.c_void_p] * 3 + [C.c_uint32] + [C.c_void_p] * 4
    d.vpref_default_materials.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    return lib


def _pin_protos(lib, pre):
    d = lib.dll
    getattr(d, pre + "offset_ray").argtypes = [C.c_void_p] * 3
    getattr(d, pre + "sample_sky").argtypes = [C.c_void_p] * 3
    getattr(d, pre + "wang_hash").argtypes = [C.c_uint32]
    getattr(d, pre + "wang_hash").restype = C.c_uint32
    getattr(d, pre + "rng_floats").argtypes = [C.c_uint32, C.c_int, C.c_void_p]
    getattr(d, pre + "sampling").argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    getattr(d, pre + "tonemap").argtypes = [C.c_void_p] * 4
    getattr(d, pre + "threads").restype = C.c_int


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def hits_equal(a, b):
    """Field-wise bit comparison of two VpHit arrays; returns {field: mismatch count}."""
    out = {}
    for f in a.dtype.names:
        if f == "pad":
            continue
        x, y = a[f], b[f]
        if x.dtype == np.float32:
            x, y = x.view(np.uint32), y.view(np.uint32)
        out[f] = int(np.count_nonzero(x != y))
    return out


def assert_hits_equal(a, b, what=""):
    mm = hits_equal(a, b)
    assert all(v == 0 for v in mm.values()), f"{what} hit mismatches: {mm}"


def small_config(idx, width, height, **over):
    sc, cam, fpd = scenes.config(idx, width, height)
    fpd.update(over)
    return sc, cam, fpd


def params(fpd, **over):
    d = dict(fpd)
    d.update(over)
    return vp.frame_params(d["width"], d["height"], max_bounces=d["max_bounces"], spp=d["spp"],
                           frame_index=d.get("frame_index", 0), seed=d.get("seed", scenes.SEED),
                           aa_strength=d["aa_strength"], row_begin=d.get("row_begin", 0), row_end=d.get("row_end", 0),
                           row_tile=d.get("row_tile", 0), row_tile_stride=d.get("row_tile_stride", 1),
                           row_tile_phase=d.get("row_tile_phase", 0), flags=d.get("flags", 0))


def reproject_sequence(lib, idx, width, height, frames=3, max_bounces=None):
    """`frames` frames of the reprojection mode (Renderer::Tick static-camera branch) with a camera that drifts between
    frames, each frame reprojecting into the previous camera and history.  Returns per frame
    (history, rgb8, albedo, illumination, [primary, secondary, shadow, exit, steps])."""
    sc, cam, fpd = small_config(idx, width, height)
    if max_bounces is not None:
        fpd["max_bounces"] = max_bounces
    pos = np.array(cam.cam_pos, np.float32)
    tl, tr, bl = (np.array(v, np.float32) for v in (cam.top_left, cam.top_right, cam.bottom_left))
    target = tl + (tr - tl) * np.float32(0.5) + (bl - tl) * np.float32(0.5)  # centre of the image plane
    out, hist = [], None
    with vp.Session(lib) as s:
        s.set_scene(sc)
        prev = host.prev_camera(pos, target, width, height)
        for f in range(frames):
            camf = host.look_at(pos, target, width, height)
            h, rgb, alb, ill, st = s.render_reproject(camf, prev, params(fpd, spp=1, frame_index=f), hist)
            out.append((h, rgb, alb, ill, [st.primary_rays, st.secondary_rays, st.shadow_rays, st.exit_rays, st.dda_steps]))
            hist = h
            prev = host.prev_camera(pos, target, width, height)
            pos = pos + np.array([0.01, 0.004, -0.006], np.float32)
    return out


def side_channels(lib, width=96, height=54):
    """The gameplay side channels on scenes.game_scene(): inLight after an accumulating frame and after a reprojection
    frame (and that the read clears it), FindNearestPlayer over a fan of rays with and without the chunk rule, the
    autofocus distance for a few lens seeds.  Returns a dict of arrays / flags."""
    sc, cam, fpd = scenes.game_scene(width, height)
    out = {}
    with vp.Session(lib) as s:
        s.set_scene(sc)
        out["idle"] = s.in_light()
        acc, rgb, st = s.render(cam, params(fpd, spp=2))
        out["after_render"] = s.in_light()
        out["after_read"] = s.in_light()
        pos = np.array(cam.cam_pos, np.float32)
        prev = host.prev_camera(pos, (0.75, 0.4, 0.5), width, height)
        s.render_reproject(cam, prev, params(fpd, frame_index=5))
        out["after_reproject"] = s.in_light()
        rng = np.random.default_rng(11)
        n = 512
        o = np.tile(np.array([[0.75, 0.6, -1.2]], np.float32), (n, 1)) + rng.uniform(-0.05, 0.05, (n, 3)).astype(np.float32)
        d = np.stack([rng.uniform(-0.8, 0.8, n), rng.uniform(-0.35, 0.25, n), np.ones(n)], axis=1).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        rays = vp.make_rays(o, d)
        out["player0"] = s.find_nearest_player(rays, 0)
        out["player2"] = s.find_nearest_player(rays, 2)
        out["nearest"] = s.find_nearest(rays)
        out["focus"] = np.array([s.focus_distance(cam, width, height, sd) for sd in (1, 0x12345678, 0xDEADBEEF, 77)], np.float32)
        out["accum"] = acc
    return out


# ---- round-2 parity scenes (tools/make_golden_r2.py writes their reference-build digests) ---------------------------
def prims_scene(glass=False, width=160, height=90):
    """cfg 4 (city, area + point lights) with analytic primitives through the whole render path (reference
    renderer.cpp:994-1016, src/BVH/Shapes.h): three spheres and two triangles; with `glass`, two GLASS spheres in front of
    them - rays refracted into a sphere carry isInsideGlass = 1 until a primitive hit overwrites it (renderer.cpp:1012)."""
    sc, cam, fpd = small_config(4, width, height, spp=2, max_bounces=6)
    sc.spheres = [((0.2, 0.6, -0.5), 0.3, 5), ((1.0, 0.4, 0.2), 0.25, 1), ((-0.8, 0.5, 0.1), 0.2, 20)]
    sc.triangles = [((-0.5, 0, 0), (0, 0.6, 0), (0.5, 0, 0), (0.4, 0.1, -0.9), 7), ((-1, 0, 0), (0, 1, 0), (1, 0, 0), (0.0, 0.2, 3.5), 30)]
    if glass:
        sc.spheres = sc.spheres + [((0.3, 0.75, -1.3), 0.22, 8), ((0.75, 0.5, -1.0), 0.2, 8)]
    return sc, cam, fpd


def light_scene(kind, width=160, height=90):
    """Light evaluation variants no BASELINE config reaches: 'dir' = only the directional light, coloured
    (DirectionalLightEvaluate, reference renderer.cpp:315-338); 'area1' / 'area5' = cfg 4's area lights with
    numCheckShadowsAreaLight 1 / 5 (AreaLightEvaluation, :163-207) next to the coloured directional light."""
    sc, cam, fpd = small_config(4, width, height, spp=2, max_bounces=6)
    sc.dir_light = (tuple(float(x) for x in host.normalize((-0.3, -0.8, 0.5))), (2.0, 1.8, 1.5))
    if kind == "dir":
        sc.point_lights, sc.area_lights, sc.spot_lights = [], [], []
    else:
        sc.area_shadow_samples = {"area1": 1, "area5": 5}[kind]
    return sc, cam, fpd


def frame_digest(acc, rgb8, st):
    import hashlib
    return dict(accum_sha256=hashlib.sha256(np.ascontiguousarray(acc).tobytes()).hexdigest(),
                rgb8_sha256=hashlib.sha256(np.ascontiguousarray(rgb8).tobytes()).hexdigest(),
                stats=[int(st.primary_rays), int(st.secondary_rays), int(st.shadow_rays), int(st.exit_rays)])


R2_SMALL = {"prims": lambda: prims_scene(False), "prims_glass": lambda: prims_scene(True), "dir": lambda: light_scene("dir"),
            "area1": lambda: light_scene("area1"), "area5": lambda: light_scene("area5")}
