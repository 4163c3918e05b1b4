"""The oracle against the committed golden vectors (generated from the reference build by
tools/make_golden.py).  Runs everywhere, no reference needed: this is what pins the oracle
on the GPU box and in CI."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest
from _util import GOLDEN, assert_hits_equal, bits, host, params, scenes, small_config, vp


@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5])
def test_config_small(oracle, idx):
    g = np.load(os.path.join(GOLDEN, f"cfg{idx}_64x36.npz"))
    sc, cam, fpd = small_config(idx, 64, 36)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        hits = s.trace_primary(cam, params(fpd))
        acc, rgb8, st = s.render(cam, params(fpd, spp=int(g["spp"])))
    assert_hits_equal(g["hits"], hits, f"cfg{idx}")
    assert np.array_equal(bits(g["accum"]), bits(acc))
    assert np.array_equal(g["rgb8"], rgb8)
    assert list(g["stats"]) == [st.primary_rays, st.secondary_rays, st.shadow_rays, st.exit_rays, st.dda_steps]


def hit_digest(hits):
    h = hashlib.sha256()
    for f in ("volume", "cell", "t", "normal", "material", "face"):
        h.update(np.ascontiguousarray(hits[f]).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("idx", [1, 4])
def test_fullres_hit_digest(oracle, idx):
    """Primary hits at BASELINE.json's full resolution: checksum of the whole dump."""
    want = json.load(open(os.path.join(GOLDEN, "fullres_hit_digests.json")))[f"cfg{idx}"]
    sc, cam, fpd = scenes.config(idx)
    assert (fpd["width"], fpd["height"]) == (want["width"], want["height"])
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        hits = s.trace_primary(cam, params(fpd))
    assert int((hits["volume"] >= 0).sum()) == want["hit_pixels"]
    assert hit_digest(hits) == want["sha256"]


def test_pins(oracle):
    g = np.load(os.path.join(GOLDEN, "pins.npz"))
    d = oracle.dll
    for i, s in enumerate(g["seeds"]):
        assert d.vpo_wang_hash(int(s)) == int(g["wang"][i])
        f = np.zeros(8, np.float32)
        d.vpo_rng_floats(int(s), 8, f.ctypes.data)
        assert np.array_equal(bits(f), bits(g["floats"][i]))
        out, after = np.zeros(15, np.float32), C.c_uint32()
        n = np.ascontiguousarray(g["normals"][i])
        d.vpo_sampling(int(s), n.ctypes.data, out.ctypes.data, C.byref(after))
        assert np.array_equal(bits(out), bits(g["sampling"][i])) and after.value == int(g["seed_after"][i])
    for i in range(len(g["offset_p"])):
        o = np.zeros(3, np.float32)
        p, n = np.ascontiguousarray(g["offset_p"][i]), np.ascontiguousarray(g["offset_n"][i])
        d.vpo_offset_ray(p.ctypes.data, n.ctypes.data, o.ctypes.data)
        assert np.array_equal(bits(o), bits(g["offset_out"][i]))
    ctx = oracle.c_create(0)
    for i in range(len(g["tonemap_in"])):
        o, r8 = np.zeros(3, np.float32), C.c_uint32()
        c = np.ascontiguousarray(g["tonemap_in"][i])
        d.vpo_tonemap(ctx, c.ctypes.data, o.ctypes.data, C.byref(r8))
        assert np.array_equal(bits(o), bits(g["tonemap_out"][i])) and r8.value == int(g["tonemap_rgb8"][i])
    sky = scenes.synthetic_sky(256, 128)
    oracle.c_set_sky(ctx, sky.ctypes.data, 256, 128, 1.5, 1)
    for i in range(len(g["sky_dirs"])):
        o = np.zeros(3, np.float32)
        dd = np.ascontiguousarray(g["sky_dirs"][i])
        d.vpo_sample_sky(ctx, dd.ctypes.data, o.ctypes.data)
        assert np.array_equal(bits(o), bits(g["sky_out"][i]))
    oracle.c_destroy(ctx)


def test_host_authoring_matches_reference():
    """libvoxpop_host.so (C++ restatement of Scene::SetTransform / MaterialSetUp) against the
    reference's own outputs stored in the goldens."""
    g = np.load(os.path.join(GOLDEN, "pins.npz"))
    assert np.array_equal(bits(host.default_materials()), bits(g["default_materials"]))
    for tin, tout in zip(g["transform_in"], g["transform_out"]):
        M, iM, a, b = host.make_transform(tin[0:3], tin[3:6], tin[6:9])
        assert np.array_equal(bits(np.concatenate([M, iM, a, b])), bits(tout))


@pytest.mark.skipif(not os.path.isdir("/root/reference/assets"), reason="reference assets not present on this box")
def test_vox_loader_matches_fixtures():
    """The C++ .vox reader + LoadModel restatement reproduces the committed grids (which the
    reference's own ogt_vox-based loader produced)."""
    for name in ["teapot", "room", "roomGlass", "SmallBuilding01", "SmallBuilding02", "TallBuilding01", "monu1", "monu2",
                 "monu3", "player"]:
        grid, n, mats = scenes.load_grid(name)
        g2, m2 = host.load_vox(f"/root/reference/assets/{name}.vox", n)
        assert np.array_equal(grid, g2), name
        assert np.array_equal(bits(mats), bits(m2)), name


@pytest.mark.parametrize("idx", [2, 3, 4])
def test_reprojection_golden(oracle, idx):
    """Reprojection frame mode against goldens generated from the reference build (tools/make_golden.py)."""
    from _util import reproject_sequence
    g = np.load(os.path.join(GOLDEN, f"reproject_cfg{idx}_64x36.npz"))
    out = reproject_sequence(oracle, idx, 64, 36)
    for f, o in enumerate(out):
        assert np.array_equal(bits(g[f"history{f}"]), bits(o[0])), f"frame {f} history"
        assert np.array_equal(g[f"rgb8_{f}"], o[1]), f"frame {f} rgb8"
        assert list(g[f"stats{f}"]) == o[4]
    assert np.array_equal(bits(g["albedo2"]), bits(out[2][2])) and np.array_equal(bits(g["illumination2"]), bits(out[2][3]))


@pytest.mark.parametrize("idx", [1, 4])
def test_fullres_frame_digest(oracle, idx):
    """The oracle at full size against the reference build's frame digests (two accumulated samples, counters)."""
    want = json.load(open(os.path.join(GOLDEN, "fullres_frame_digests.json")))[f"cfg{idx}"]
    sc, cam, fpd = scenes.config(idx)
    with vp.Session(oracle) as s:
        s.set_scene(sc)
        acc, rgb8, st = s.render(cam, params(fpd, spp=2))
    assert [st.primary_rays, st.secondary_rays, st.shadow_rays, st.exit_rays, st.dda_steps] == want["stats"]
    assert hashlib.sha256(acc.tobytes()).hexdigest() == want["accum_sha256"]
    assert hashlib.sha256(rgb8.tobytes()).hexdigest() == want["rgb8_sha256"]
