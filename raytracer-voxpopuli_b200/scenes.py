"""The benchmark / parity configurations of BASELINE.json as explicit inputs.

Every config is a (SceneDesc, Camera, frame-parameter dict).  Grids come from the committed
fixtures under data/scenes/ (decoded from the reference's .vox assets by the reference's own
Scene::LoadModel, see tools/make_scene_fixtures.py); transforms, materials, lights and
cameras are built here with the host-side restatement of the reference's authoring code
(host.py -> libvoxpop_host.so).  Nothing here touches /root/reference or oracle/.
"""
import json
import os

import numpy as np

from . import host
from .session import SceneDesc, Volume

DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "scenes")
SEED = 0x12345678


def load_grid(name):
    d = np.load(os.path.join(DATA_DIR, f"{name}.npz"))
    return d["grid"].copy(), int(d["n"]), d["materials"].copy()


def synthetic_sky(w=2048, h=1024):
    """Deterministic HDR environment standing in for the missing assets/sky_19.hdr
    (reference .MISSING_LARGE_BLOBS): vertical gradient + a sun disc, fp32 RGB, pure
    integer/rational arithmetic so every platform generates identical bytes."""
    v = (np.arange(h, dtype=np.float32) + np.float32(0.5)) / np.float32(h)
    u = (np.arange(w, dtype=np.float32) + np.float32(0.5)) / np.float32(w)
    horizon = np.float32(1.0) - np.abs(v - np.float32(0.5)) * np.float32(2.0)
    base = np.stack([np.float32(0.25) + np.float32(0.55) * horizon,
                     np.float32(0.45) + np.float32(0.45) * horizon,
                     np.float32(0.85) + np.float32(0.10) * horizon], axis=-1).astype(np.float32)
    sky = np.broadcast_to(base[:, None, :], (h, w, 3)).copy()
    du = u[None, :] - np.float32(0.3)
    dv = v[:, None] - np.float32(0.2)
    sun = (du * du * np.float32(4.0) + dv * dv) < np.float32(0.0004)
    sky[sun] = np.array([40.0, 36.0, 28.0], dtype=np.float32)
    return np.ascontiguousarray(sky, dtype=np.float32)


def _volume(name, position, rotation=(0.0, 0.0, 0.0), scale=(1.0, 1.0, 1.0), remap=None):
    grid, n, mats = load_grid(name)
    if remap:
        for a, b in remap.items():
            grid[grid == a] = b
    M, iM, bmin, bmax = host.make_transform(position, rotation, scale)
    return Volume(grid, n, bmin, bmax, M, iM, (name, tuple(sorted((remap or {}).items())))), mats


def _merge_materials(tables):
    """LoadModel edits the shared materials table per loaded model (scene.cpp:516-520): later loads win."""
    base = host.default_materials()
    out = base.copy()
    for t in tables:
        changed = np.any(t != base, axis=1)
        out[changed] = t[changed]
    return out


def config(idx, width=None, height=None):
    """Returns (SceneDesc, Camera, dict(width, height, max_bounces, spp, aa_strength)) for configs 1..5; 6 is config 2 with
    room.vox's native material mapping (palette 8 = GLASS for 99 % of its voxels, reference template/scene.cpp:508-522):
    the "room.vox ... with glass" wording of the north_star target."""
    if idx == 6:
        sc, cam, fpd = config(2, width, height)
        vol, mats = _volume("room", (0, 0, 0))
        sc.volumes, sc.materials = [vol], _merge_materials([mats])
        return sc, cam, fpd
    if idx == 1:
        # teapot.vox single volume, 1280x720, primary rays + 1 point light shadow, 1 spp
        vol, mats = _volume("teapot", (0, 0, 0), remap={8: 249})  # all-glass -> diffuse palette id so a shadow ray exists
        sc = SceneDesc(volumes=[vol], materials=_merge_materials([mats]),
                       point_lights=[((0.5, 1.6, -0.6), (4.0, 4.0, 4.0))])
        W, H = width or 1280, height or 720
        cam = host.look_at((0.5, 0.55, -0.75), (0.5, 0.3, 0.5), W, H)
        return sc, cam, dict(width=W, height=H, max_bounces=0, spp=1, aa_strength=1.0)
    if idx in (2, 3):
        W, H = width or 1920, height or 1080
        # the model is a closed room (floor y~0.01, ceiling y~0.37, x 0..0.92, z 0..0.69): camera and lights sit inside
        lights = dict(point_lights=[((0.5, 0.30, 0.35), (0.35, 0.32, 0.28)), ((0.2, 0.25, 0.55), (0.15, 0.18, 0.25))],
                      spot_lights=[((0.7, 0.34, 0.3), host.normalize((-0.3, -0.9, 0.1)), (0.6, 0.6, 0.5), host.cos_degrees(50.0)),
                                   ((0.3, 0.34, 0.2), host.normalize((0.2, -0.9, 0.2)), (0.4, 0.5, 0.6), host.cos_degrees(45.0))])
        cam = host.look_at((0.85, 0.22, 0.62), (0.2, 0.1, 0.15), W, H)
        if idx == 2:
            # room.vox Whitted-style materials: 8 -> NON_METAL_WHITE, 246 -> METAL_LOW ("diffuse+mirror")
            vol, mats = _volume("room", (0, 0, 0), remap={8: 0, 246: 7})
            sc = SceneDesc(volumes=[vol], materials=_merge_materials([mats]), **lights)
            return sc, cam, dict(width=W, height=H, max_bounces=14, spp=1, aa_strength=1.0)
        vol, mats = _volume("roomGlass", (0, 0, 0))
        sc = SceneDesc(volumes=[vol], materials=_merge_materials([mats]), sky=synthetic_sky(), sky_active=True, **lights)
        return sc, cam, dict(width=W, height=H, max_bounces=8, spp=16, aa_strength=1.0)
    if idx in (4, 5):
        names = ["SmallBuilding01", "SmallBuilding02", "TallBuilding01", "monu1", "monu2", "monu3", "player"]
        reps = 1 if idx == 4 else 3
        vols, tabs = [], []
        for r in range(reps):
            for k, nm in enumerate(names):
                pos = (float(k % 4) * 1.25 - 2.0, 0.0, float(k // 4) * 1.5 + float(r) * 3.0)  # multiples of 1/4: integral in grid units for n = 16 and 64
                rot = (0.0, float(np.float32(np.pi) / np.float32(4.0)), 0.0) if k in (1, 4) else (0.0, 0.0, 0.0)
                v, m = _volume(nm, pos, rot)
                vols.append(v)
                tabs.append(m)
        mats = _merge_materials(tabs)
        if idx == 4:
            W, H = width or 2560, height or 1440
            sc = SceneDesc(volumes=vols, materials=mats,
                           area_lights=[((-1.5, 2.0, 0.5), (1.0, 0.9, 0.8), 1.2, 0.5), ((0.5, 2.3, 1.0), (0.8, 0.9, 1.0), 1.2, 0.5),
                                        ((2.0, 2.0, 0.2), (1.0, 1.0, 1.0), 1.2, 0.5), ((0.0, 1.8, 2.5), (0.9, 0.8, 1.0), 1.2, 0.5)],
                           point_lights=[((-1.0, 1.2, -0.8), (2.0, 2.0, 2.0)), ((1.5, 1.2, 2.5), (2.0, 1.7, 1.4))],
                           area_shadow_samples=3)
            cam = host.look_at((0.45, 1.1, -2.4), (0.45, 0.35, 0.9), W, H)
            return sc, cam, dict(width=W, height=H, max_bounces=14, spp=1, aa_strength=1.0)
        W, H = width or 3840, height or 2160
        rng = host.Xorshift32(1)
        pl, sl, al = [], [], []
        for _ in range(32):
            pl.append(((rng.uniform(-2.5, 3.5), rng.uniform(1.2, 3.0), rng.uniform(-1, 9)), (rng.uniform(0.5, 3), rng.uniform(0.5, 3), rng.uniform(0.5, 3))))
        for _ in range(16):
            p = (rng.uniform(-2.5, 3.5), rng.uniform(2.0, 3.5), rng.uniform(-1, 9))
            d = host.normalize((rng.uniform(-0.5, 0.5), -1.0, rng.uniform(-0.5, 0.5)))
            sl.append((p, d, (rng.uniform(1, 4), rng.uniform(1, 4), rng.uniform(1, 4)), host.cos_degrees(35.0)))
        for _ in range(16):
            al.append(((rng.uniform(-2.5, 3.5), rng.uniform(2.0, 3.5), rng.uniform(-1, 9)), (rng.uniform(0.5, 1), rng.uniform(0.5, 1), rng.uniform(0.5, 1)), 1.2, rng.uniform(0.4, 1.2)))
        sc = SceneDesc(volumes=vols, materials=mats, point_lights=pl, spot_lights=sl, area_lights=al, area_shadow_samples=3)
        cam = host.look_at((0.45, 2.0, -3.2), (0.45, 0.3, 3.5), W, H)
        return sc, cam, dict(width=W, height=H, max_bounces=14, spp=64, aa_strength=1.0)
    raise ValueError(f"unknown config {idx}")


def game_scene(width=96, height=54):
    """A game-like arrangement for the gameplay side channels: the player volume FIRST (the reference
    keeps it at voxelVolumes[0]), buildings behind it, one monument with smoke cells, and a strong point light right above
    the player so that Renderer::inLight fires."""
    vols, tabs = [], []
    for nm, pos, rot in (("player", (0.25, 0.0, 0.0), (0.0, 0.0, 0.0)), ("SmallBuilding01", (-1.0, 0.0, 1.5), (0.0, 0.0, 0.0)),
                         ("monu1", (1.0, 0.0, 1.25), (0.0, float(np.float32(np.pi) / np.float32(4.0)), 0.0)),
                         ("TallBuilding01", (0.0, 0.0, 2.75), (0.0, 0.0, 0.0)), ("monu2", (2.25, 0.0, 2.5), (0.0, 0.0, 0.0))):
        v, m = _volume(nm, pos, rot)
        vols.append(v)
        tabs.append(m)
    sc = SceneDesc(volumes=vols, materials=_merge_materials(tabs),
                   point_lights=[((0.75, 1.4, 0.4), (60.0, 60.0, 60.0)), ((0.0, 2.0, 2.0), (3.0, 3.0, 3.0))])
    cam = host.look_at((0.75, 0.8, -1.6), (0.75, 0.4, 0.5), width, height, focal_distance=1.0, defocus_jitter=2.0)
    return sc, cam, dict(width=width, height=height, max_bounces=6, spp=1, aa_strength=1.0)


# look-at targets of config()'s cameras (for building a previous-frame camera next to them)
CAMERA_TARGETS = {1: (0.5, 0.3, 0.5), 2: (0.2, 0.1, 0.15), 3: (0.2, 0.1, 0.15), 4: (0.45, 0.35, 0.9), 5: (0.45, 0.3, 3.5)}

WORKLOAD_NAMES = {
    1: "cfg1 teapot.vox 128^3 1280x720 1spp maxBounces=0 1 point light",
    2: "cfg2 room.vox 128^3 1920x1080 1spp maxBounces=14 2 point + 2 spot lights, diffuse+mirror remap",
    3: "cfg3 roomGlass.vox 128^3 1920x1080 16spp maxBounces=8 glass + synthetic HDR sky, 2 point + 2 spot lights",
    4: "cfg4 city 6x64^3+16^3 2560x1440 1spp maxBounces=14 4 area + 2 point lights",
    5: "cfg5 city x3 (21 volumes) 3840x2160 64spp maxBounces=14 64 lights",
    6: "cfg2g room.vox 128^3 1920x1080 1spp maxBounces=14 2 point + 2 spot lights, native materials (palette 8 = GLASS)",
}

BENCH_DIR = os.path.join(os.path.dirname(DATA_DIR), "bench")


def save_inputs(idx, path):
    """Everything config(idx) computes with the host mirror (transforms, material table, lights, camera), as plain arrays:
    bench.py's reference arm reads this file instead, so that no library of this repository is mapped into that process
    (tools/make_bench_inputs.py writes it, tests/test_bench_inputs.py keeps it in step with config())."""
    sc, cam, fpd = config(idx)
    f = lambda rows, k: np.array([np.concatenate([np.ravel(np.asarray(x, np.float32)) for x in r]) for r in rows], np.float32).reshape(len(rows), k)
    np.savez_compressed(
        path, names=np.array([v.source[0] for v in sc.volumes]), remap=np.array([json.dumps(v.source[1]) for v in sc.volumes]),
        n=np.array([v.n for v in sc.volumes], np.int32), cube_min=np.array([v.cube_min for v in sc.volumes], np.float32),
        cube_max=np.array([v.cube_max for v in sc.volumes], np.float32), matrix=np.array([v.matrix for v in sc.volumes], np.float32),
        inv_matrix=np.array([v.inv_matrix for v in sc.volumes], np.float32), materials=np.asarray(sc.materials, np.float32),
        point=f(sc.point_lights, 6), area=f(sc.area_lights, 8), spot=f(sc.spot_lights, 10),
        dir=np.concatenate([np.asarray(sc.dir_light[0], np.float32), np.asarray(sc.dir_light[1], np.float32)]),
        area_samples=np.int32(sc.area_shadow_samples), sky=np.int32(1 if sc.sky is not None else 0), sky_contribution=np.float32(sc.sky_contribution),
        sky_active=np.int32(sc.sky_active),
        camera=np.concatenate([np.asarray(getattr(cam, k), np.float32) for k in ("cam_pos", "top_left", "top_right", "bottom_left", "right", "up")] +
                              [np.array([cam.focal_distance, cam.defocus_jitter], np.float32)]),
        frame=np.array([fpd["width"], fpd["height"], fpd["max_bounces"], fpd["spp"]], np.int32), aa=np.float32(fpd["aa_strength"]))


def load_inputs(path):
    """(SceneDesc, Camera, frame dict) from a save_inputs file; numpy only."""
    from .session import Camera
    d = np.load(path)
    vols = []
    for i, nm in enumerate(d["names"]):
        grid, n, _ = load_grid(str(nm))
        for a, b in json.loads(str(d["remap"][i])):
            grid[grid == a] = b
        assert n == int(d["n"][i])
        vols.append(Volume(grid, n, d["cube_min"][i], d["cube_max"][i], d["matrix"][i], d["inv_matrix"][i], (str(nm), ())))
    sc = SceneDesc(volumes=vols, materials=d["materials"],
                   point_lights=[(r[:3], r[3:6]) for r in d["point"]], area_lights=[(r[:3], r[3:6], float(r[6]), float(r[7])) for r in d["area"]],
                   spot_lights=[(r[:3], r[3:6], r[6:9], float(r[9])) for r in d["spot"]], dir_light=(tuple(d["dir"][:3]), tuple(d["dir"][3:])),
                   area_shadow_samples=int(d["area_samples"]), sky=synthetic_sky() if int(d["sky"]) else None,
                   sky_contribution=float(d["sky_contribution"]), sky_active=bool(int(d["sky_active"])))
    c = d["camera"]
    cam = Camera(c[0:3], c[3:6], c[6:9], c[9:12], c[12:15], c[15:18], float(c[18]), float(c[19]))
    fr = d["frame"]
    return sc, cam, dict(width=int(fr[0]), height=int(fr[1]), max_bounces=int(fr[2]), spp=int(fr[3]), aa_strength=float(d["aa"]))
