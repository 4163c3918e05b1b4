#!/usr/bin/env python3
"""Generate the committed golden vectors under tests/golden/ from the REFERENCE ITSELF
(oracle/_ref/libvpref.so = the reference's own renderer.cpp/scene.cpp/... built headless by
oracle/ref/build_ref.py).  Runs in the build container only.

Per config at 64x36:  primary-hit records (volume, cell, t bits, normal bits,
material, face), the float4 accumulator after `spp` Update() passes, the RGB8 image and the
ray counters; plus SHA-256 digests of the full-resolution primary-hit dumps.  Also the small
pins: WangHash / xorshift floats, sampling helpers, OffsetRay, tonemap, SampleSky,
Scene::SetTransform matrices and the MaterialSetUp table."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _util import GOLDEN, abi, host, params, ref_lib, scenes, small_config, vp  # noqa: E402

SMALL = (64, 36)
SMALL_SPP = {1: 2, 2: 2, 3: 3, 4: 2, 5: 2}


def hit_digest(hits):
    h = hashlib.sha256()
    for f in ("volume", "cell", "t", "normal", "material", "face"):
        h.update(np.ascontiguousarray(hits[f]).tobytes())
    return h.hexdigest()


def main():
    ref = ref_lib()
    assert ref is not None, "build oracle/_ref first (python oracle/ref/build_ref.py)"
    os.makedirs(GOLDEN, exist_ok=True)
    digests = {}
    for idx in (1, 2, 3, 4, 5):
        sc, cam, fpd = small_config(idx, *SMALL)
        s = vp.Session(ref)
        s.set_scene(sc)
        hits = s.trace_primary(cam, params(fpd))
        acc, rgb8, st = s.render(cam, params(fpd, spp=SMALL_SPP[idx]))
        np.savez_compressed(os.path.join(GOLDEN, f"cfg{idx}_64x36.npz"), hits=hits, accum=acc, rgb8=rgb8,
                            stats=np.array([st.primary_rays, st.secondary_rays, st.shadow_rays, st.exit_rays, st.dda_steps], np.uint64),
                            spp=np.int32(SMALL_SPP[idx]))
        # full resolution primary-hit digest (bit-exact gate at BASELINE.json's sizes)
        sc, cam, fpd = scenes.config(idx)
        s.set_scene(sc)
        hits = s.trace_primary(cam, params(fpd))
        digests[f"cfg{idx}"] = dict(width=fpd["width"], height=fpd["height"], sha256=hit_digest(hits),
                                    hit_pixels=int((hits["volume"] >= 0).sum()))
        print(idx, digests[f"cfg{idx}"], flush=True)
        s.close()
    with open(os.path.join(GOLDEN, "fullres_hit_digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)

    # ---- full-size frames: digests of the accumulator / RGB8 after 2 accumulated samples, and of one reprojection frame ----
    frames = {}
    for idx in (1, 2, 3, 4, 5):
        sc, cam, fpd = scenes.config(idx)
        W, H = fpd["width"], fpd["height"]
        with vp.Session(ref) as s:
            s.set_scene(sc)
            acc, rgb8, st = s.render(cam, params(fpd, spp=2))
            prev = host.prev_camera(np.array(cam.cam_pos, np.float32) + np.array([0.01, 0, 0], np.float32), scenes.CAMERA_TARGETS[idx], W, H)
            hist, rrgb, _, _, rst = s.render_reproject(cam, prev, params(fpd, spp=1, frame_index=3),
                                                       np.ascontiguousarray(acc[:, :, :3]), want_buffers=False)
        frames[f"cfg{idx}"] = dict(width=W, height=H, accum_sha256=hashlib.sha256(acc.tobytes()).hexdigest(),
                                   rgb8_sha256=hashlib.sha256(rgb8.tobytes()).hexdigest(),
                                   stats=[int(st.primary_rays), int(st.secondary_rays), int(st.shadow_rays), int(st.exit_rays), int(st.dda_steps)],
                                   reproject_history_sha256=hashlib.sha256(hist.tobytes()).hexdigest(),
                                   reproject_rgb8_sha256=hashlib.sha256(rrgb.tobytes()).hexdigest(),
                                   reproject_rays=[int(rst.primary_rays), int(rst.secondary_rays), int(rst.shadow_rays), int(rst.exit_rays)])
        print("fullres frame", idx, frames[f"cfg{idx}"]["stats"], flush=True)
    with open(os.path.join(GOLDEN, "fullres_frame_digests.json"), "w") as f:
        json.dump(frames, f, indent=1, sort_keys=True)

    # ---- reprojection frame mode: three drifting-camera frames per config -------------------
    from _util import reproject_sequence
    for idx in (2, 3, 4):
        out = reproject_sequence(ref, idx, *SMALL)
        blob = {}
        for f, o in enumerate(out):
            blob[f"history{f}"] = o[0]; blob[f"rgb8_{f}"] = o[1]; blob[f"stats{f}"] = np.array(o[4], np.uint64)
        blob["albedo2"] = out[2][2]; blob["illumination2"] = out[2][3]
        np.savez_compressed(os.path.join(GOLDEN, f"reproject_cfg{idx}_64x36.npz"), **blob)
        print("reproject", idx, out[2][4], flush=True)

    # ---- small pins -----------------------------------------------------------------------
    d = ref.dll
    rng = np.random.default_rng(7)
    seeds = rng.integers(1, 2 ** 32 - 1, size=64, dtype=np.uint64).astype(np.uint32)
    wang = np.array([d.vpref_wang_hash(int(s)) for s in seeds], np.uint32)
    floats = np.zeros((64, 8), np.float32)
    samp = np.zeros((64, 15), np.float32)
    after = np.zeros(64, np.uint32)
    nrm = rng.normal(size=(64, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    for i, s in enumerate(seeds):
        d.vpref_rng_floats(int(s), 8, floats[i].ctypes.data)
        a = C.c_uint32()
        d.vpref_sampling(int(s), nrm[i].ctypes.data, samp[i].ctypes.data, C.byref(a))
        after[i] = a.value
    pts = (rng.normal(size=(256, 3)) * np.array([3.0, 0.02, 1.0])).astype(np.float32)
    pn = rng.normal(size=(256, 3)).astype(np.float32)
    pn /= np.linalg.norm(pn, axis=1, keepdims=True).astype(np.float32)
    off = np.zeros((256, 3), np.float32)
    for i in range(256):
        d.vpref_offset_ray(pts[i].ctypes.data, pn[i].ctypes.data, off[i].ctypes.data)
    ctx = ref.c_create(0)
    col = np.abs(rng.normal(size=(128, 4)) * 2).astype(np.float32)
    tm = np.zeros((128, 3), np.float32)
    tm8 = np.zeros(128, np.uint32)
    for i in range(128):
        a = C.c_uint32()
        d.vpref_tonemap(ctx, col[i].ctypes.data, tm[i].ctypes.data, C.byref(a))
        tm8[i] = a.value
    sky = scenes.synthetic_sky(256, 128)
    ref.c_set_sky(ctx, sky.ctypes.data, 256, 128, 1.5, 1)
    sd = rng.normal(size=(512, 3)).astype(np.float32)
    sd /= np.linalg.norm(sd, axis=1, keepdims=True).astype(np.float32)
    so = np.zeros((512, 3), np.float32)
    for i in range(512):
        d.vpref_sample_sky(ctx, sd[i].ctypes.data, so[i].ctypes.data)
    mats = (abi.VpMaterial * 256)()
    d.vpref_default_materials(ctx, mats, 256)
    mtab = np.array([[x.albedo[0], x.albedo[1], x.albedo[2], x.roughness, x.emissive, x.ior] for x in mats], np.float32)
    tr_in, tr_out = [], []
    for i in range(64):
        pos = np.round(rng.uniform(-5, 5, 3) * 4).astype(np.float32) / np.float32(4)
        rot = rng.uniform(-3, 3, 3).astype(np.float32) if i % 3 else np.zeros(3, np.float32)
        scl = rng.uniform(0.5, 10, 3).astype(np.float32) if i % 4 == 0 else np.ones(3, np.float32)
        M, iM, a, b = (np.zeros(16, np.float32), np.zeros(16, np.float32), np.zeros(3, np.float32), np.zeros(3, np.float32))
        d.vpref_make_transform(pos.ctypes.data, rot.ctypes.data, scl.ctypes.data, 8, M.ctypes.data, iM.ctypes.data,
                               a.ctypes.data, b.ctypes.data)
        tr_in.append(np.concatenate([pos, rot, scl]))
        tr_out.append(np.concatenate([M, iM, a, b]))
    ref.c_destroy(ctx)
    np.savez_compressed(os.path.join(GOLDEN, "pins.npz"), seeds=seeds, wang=wang, floats=floats, normals=nrm, sampling=samp,
                        seed_after=after, offset_p=pts, offset_n=pn, offset_out=off, tonemap_in=col, tonemap_out=tm,
                        tonemap_rgb8=tm8, sky_dirs=sd, sky_out=so, default_materials=mtab,
                        transform_in=np.array(tr_in, np.float32), transform_out=np.array(tr_out, np.float32))
    print("golden written to", GOLDEN)


if __name__ == "__main__":
    main()
