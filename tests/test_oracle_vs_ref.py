"""Pins the plain-C restatement (oracle/vp_oracle.c) against the REFERENCE ITSELF
(oracle/_ref/libvpref.so) on identical scene bytes, cameras and seeds: bit-exact.
Skipped only where the reference build is unavailable (the committed goldens in
test_golden.py carry the same pin everywhere)."""
import ctypes as C

import numpy as np
import pytest
from _util import abi, assert_hits_equal, bits, params, small_config, vp

CASES = [  # (config, width, height, overrides)
    (1, 96, 54, {}),
    (1, 96, 54, dict(max_bounces=3, spp=2)),
    (2, 96, 54, dict(spp=2)),
    (3, 80, 45, dict(spp=2)),
    (4, 96, 54, dict(spp=2)),
    (5, 64, 36, dict(spp=1)),
]


@pytest.mark.parametrize("idx,w,h,over", CASES)
def test_render_and_hits_bit_exact(ref, oracle, idx, w, h, over):
    sc, cam, fpd = small_config(idx, w, h, **over)
    out = []
    for lib in (ref, oracle):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            hits = s.trace_primary(cam, params(fpd))
            acc, rgb8, st = s.render(cam, params(fpd))
            out.append((hits, acc, rgb8, st))
    assert_hits_equal(out[0][0], out[1][0], f"cfg{idx}")
    assert np.array_equal(bits(out[0][1]), bits(out[1][1])), "accumulator differs"
    assert np.array_equal(out[0][2], out[1][2]), "RGB8 differs"
    a, b = out[0][3], out[1][3]
    for k in ("primary_rays", "secondary_rays", "shadow_rays", "exit_rays", "dda_steps"):
        assert getattr(a, k) == getattr(b, k), k


def test_defocus_and_second_frame(ref, oracle):
    """Depth of field on (sin/cos pin exercised) and accumulation continuing from frame 3."""
    sc, cam, fpd = small_config(3, 64, 36, spp=2)
    cam.defocus_jitter = 2.0
    cam.focal_distance = 0.6
    res = []
    for lib in (ref, oracle):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            acc = np.full((36, 64, 4), 0.25, np.float32)
            acc, rgb8, _ = s.render(cam, params(fpd, frame_index=3), accum=acc)
            res.append((acc, rgb8))
    assert np.array_equal(bits(res[0][0]), bits(res[1][0]))
    assert np.array_equal(res[0][1], res[1][1])


def test_prims_and_batch_seams(ref, oracle):
    sc, cam, fpd = small_config(4, 64, 36)
    sc.spheres = [((0.2, 0.6, -0.5), 0.3, 5), ((1.0, 0.4, 0.2), 0.25, 1), ((-0.8, 0.5, 0.1), 0.2, 20)]
    sc.triangles = [((-0.5, 0, 0), (0, 0.6, 0), (0.5, 0, 0), (0.4, 0.1, -0.9), 7), ((-1, 0, 0), (0, 1, 0), (1, 0, 0), (0.0, 0.2, 3.5), 30)]
    rng = np.random.default_rng(3)
    n = 4000
    o = rng.uniform(-2.5, 3.0, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    rays = vp.make_rays(o, d)
    rays["t"][::3] = rng.uniform(0.2, 3.0, rays["t"][::3].shape).astype(np.float32)
    seeds = rng.integers(1, 2 ** 32 - 1, n, dtype=np.uint64).astype(np.uint32)
    res = []
    for lib in (ref, oracle):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            res.append((s.find_nearest(rays), s.is_occluded(rays), s.trace(rays, seeds, 6),
                        s.render(cam, params(fpd, spp=2))[0]))
    assert_hits_equal(res[0][0], res[1][0], "find_nearest")
    assert np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(bits(res[0][2]), bits(res[1][2]))
    assert np.array_equal(bits(res[0][3]), bits(res[1][3]))


def test_small_pins(ref, oracle):
    rng = np.random.default_rng(11)
    for s in rng.integers(1, 2 ** 32 - 1, 32, dtype=np.uint64):
        assert ref.dll.vpref_wang_hash(int(s)) == oracle.dll.vpo_wang_hash(int(s))
        a, b = np.zeros(16, np.float32), np.zeros(16, np.float32)
        ref.dll.vpref_rng_floats(int(s), 16, a.ctypes.data)
        oracle.dll.vpo_rng_floats(int(s), 16, b.ctypes.data)
        assert np.array_equal(bits(a), bits(b))
        nrm = rng.normal(size=3).astype(np.float32)
        nrm /= np.float32(np.linalg.norm(nrm))
        x, y = np.zeros(15, np.float32), np.zeros(15, np.float32)
        sa, sb = C.c_uint32(), C.c_uint32()
        ref.dll.vpref_sampling(int(s), nrm.ctypes.data, x.ctypes.data, C.byref(sa))
        oracle.dll.vpo_sampling(int(s), nrm.ctypes.data, y.ctypes.data, C.byref(sb))
        assert np.array_equal(bits(x), bits(y)) and sa.value == sb.value


@pytest.mark.parametrize("idx", [1, 2, 3, 4, 5])
def test_reprojection_frames_bit_exact(ref, oracle, idx):
    """Second frame mode (TraceReproject + history reprojection, renderer.cpp:1330-1586,1997-2101): three frames with a
    drifting camera; history, albedo, illumination, RGB8 and the ray / step counters must be identical."""
    from _util import reproject_sequence
    a = reproject_sequence(ref, idx, 80, 45)
    b = reproject_sequence(oracle, idx, 80, 45)
    for f, (x, y) in enumerate(zip(a, b)):
        for k, name in ((0, "history"), (2, "albedo"), (3, "illumination")):
            assert np.array_equal(bits(x[k]), bits(y[k])), f"cfg{idx} frame {f}: {name} differs"
        assert np.array_equal(x[1], y[1]), f"cfg{idx} frame {f}: rgb8 differs"
        assert x[4] == y[4], f"cfg{idx} frame {f}: counters {x[4]} vs {y[4]}"
    # the history is really used: the last frame differs from its own illumination buffer
    assert np.count_nonzero(bits(a[-1][0]) != bits(a[-1][3])) > 0


def test_gameplay_side_channels(ref, oracle):
    """inLight, FindNearestPlayer / FindNearestExcept and the autofocus ray (renderer.cpp:1228-1240,1020-1071,1987-1991)."""
    from _util import side_channels
    a, b = side_channels(ref), side_channels(oracle)
    for k in ("idle", "after_render", "after_read", "after_reproject"):
        assert a[k] == b[k], k
    assert a["after_render"] and not a["after_read"] and not a["idle"]  # the scene really lights the player; the read clears the flag
    assert_hits_equal(a["player0"], b["player0"], "FindNearestPlayer")
    assert_hits_equal(a["player2"], b["player2"], "FindNearestPlayer chunk 2")
    assert np.array_equal(bits(a["focus"]), bits(b["focus"]))
    assert np.array_equal(bits(a["accum"]), bits(b["accum"]))
    # the side channel differs from the ordinary nearest hit (player volume skipped, smoke transparent) and the chunk rule bites
    assert np.count_nonzero(a["player0"]["volume"] != a["nearest"]["volume"]) > 0
    assert np.count_nonzero(a["player0"]["volume"] != a["player2"]["volume"]) > 0


@pytest.mark.parametrize("n", [44, 100, 128])
def test_axis_parallel_rays_on_cell_boundaries(ref, oracle, n):
    """The reference's own walk on rays with exactly zero direction components that start exactly on cell boundaries
    (tmax of such an axis is NaN, +inf or - on grids whose size is not a power of two - -inf): the plain-C restatement
    must reproduce whatever the reference does there, bit for bit (NaN normals compared as NaN).  The GPU twin of this
    test is tests/test_gpu_round2.py::test_axis_parallel_rays_on_cell_boundaries."""
    from _util import host
    rng = np.random.default_rng(1000 + n)
    g = np.full((n, n, n), 255, np.uint8)  # [z, y, x]
    g[:, : max(1, n // 10), :] = 3
    fill = rng.uniform(size=(n, n, n)) < 0.02
    g[fill] = rng.integers(0, 8, int(fill.sum())).astype(np.uint8)
    M, iM, a, b = host.make_transform((0.0, 0.0, 0.0))
    sc = vp.SceneDesc(volumes=[vp.Volume(np.ascontiguousarray(g.reshape(-1)), n, a, b, M, iM)], materials=host.default_materials(),
                      point_lights=[((0.5, 2.5, -0.5), (6, 6, 6))])
    count = 30000
    lo, hi = np.asarray(a, np.float32), np.asarray(b, np.float32)
    o = rng.uniform(0.02, 0.9, (count, 3)).astype(np.float32)
    o = (np.round(o * n) / n).astype(np.float32) * (hi - lo) + lo
    d = rng.normal(size=(count, 3)).astype(np.float32)
    d[rng.uniform(size=(count, 3)) < 0.3] = 0.0
    d[np.all(d == 0, axis=1)] = (1, 0, 0)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[count // 2:][d[count // 2:] == 0] = -0.0
    # ... and components so small that tdelta is huge or overflows to inf (1 / denormal)
    tiny = np.array([1e-30, -1e-30, 1e-38, -1e-38, 1e-39, -1e-39, 1e-44, -1e-44], np.float32)
    sel = rng.uniform(size=(count // 4, 3)) < 0.3
    d[: count // 4][sel] = rng.choice(tiny, int(sel.sum()))
    t = np.where(rng.uniform(size=count) < 0.5, 1e34, rng.uniform(0, 1.0, count)).astype(np.float32)
    rays = vp.make_rays(o, d.astype(np.float32), t=t)
    res = []
    for lib in (ref, oracle):
        with vp.Session(lib) as s:
            s.set_scene(sc)
            res.append((s.find_nearest(rays), s.is_occluded(rays)))
    want, got = res[0][0], res[1][0]
    for k in ("volume", "cell", "material", "face", "inside"):
        assert np.array_equal(want[k], got[k]), f"n={n} nearest: {k}"
    assert np.array_equal(bits(want["t"]), bits(got["t"])), f"n={n} nearest: t"
    wn, gn = want["normal"], got["normal"]
    assert np.array_equal(np.isnan(wn), np.isnan(gn)) and np.array_equal(bits(np.nan_to_num(wn)), bits(np.nan_to_num(gn)))
    assert np.array_equal(res[0][1], res[1][1])
    if n == 100:
        assert np.count_nonzero(np.isneginf(want["t"])) > 0  # the case this test exists for does occur
