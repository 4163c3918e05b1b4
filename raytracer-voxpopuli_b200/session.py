"""Scene description + a typed session over any library that speaks include/voxpop.h.

`SceneDesc` is the exact set of bytes the reference's `Renderer` holds for the render
path (volumes, materials, lights, primitives, sky); `Session` hands those bytes to a
library.  The same SceneDesc is given to the CUDA product and (in tests / bench baseline
only) to the checkers, so that all of them see identical inputs.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import abi


@dataclass
class Volume:
    grid: np.ndarray          # uint8 [n*n*n], index x + y*n + z*n*n (reference template/scene.h:241-248)
    n: int
    cube_min: np.ndarray
    cube_max: np.ndarray
    matrix: np.ndarray        # float32[16], row-major mat4 cells
    inv_matrix: np.ndarray
    source: tuple = None      # (fixture name, ((from, to), ...) palette remap) when the grid comes from data/scenes/


@dataclass
class SceneDesc:
    volumes: list = field(default_factory=list)
    materials: np.ndarray = None          # float32 [256, 6] albedo rgb, roughness, emissive, ior
    point_lights: list = field(default_factory=list)   # (pos3, color3)
    area_lights: list = field(default_factory=list)    # (pos3, color3, multiplier, radius)
    spot_lights: list = field(default_factory=list)    # (pos3, dir3, color3, cos_angle)
    dir_light: tuple = ((1.0, 0.0, 0.0), (0.0, 0.0, 0.0))  # reference DirectionalLight.h:12
    area_shadow_samples: int = 3                        # reference renderer.h:205
    spheres: list = field(default_factory=list)         # (center3, radius, material)
    triangles: list = field(default_factory=list)       # (v0, v1, v2, position, material)
    sky: np.ndarray = None                              # float32 [h, w, 3] or None
    sky_contribution: float = 1.0
    sky_active: bool = False

    @property
    def light_count(self):
        return len(self.point_lights) + len(self.area_lights) + len(self.spot_lights) + 1


@dataclass
class Camera:
    cam_pos: np.ndarray
    top_left: np.ndarray
    top_right: np.ndarray
    bottom_left: np.ndarray
    right: np.ndarray
    up: np.ndarray
    focal_distance: float = 1.0
    defocus_jitter: float = 0.0

    def to_c(self):
        return abi.VpCamera(abi.f3(self.cam_pos), abi.f3(self.top_left), abi.f3(self.top_right),
                            abi.f3(self.bottom_left), abi.f3(self.right), abi.f3(self.up),
                            float(self.focal_distance), float(self.defocus_jitter))


def frame_params(width, height, max_bounces=14, spp=1, frame_index=0, seed=0x12345678, aa_strength=1.0,
                 row_begin=0, row_end=0, row_tile=0, row_tile_stride=1, row_tile_phase=0, flags=0):
    return abi.VpFrameParams(width, height, max_bounces, spp, frame_index, seed, aa_strength, row_begin, row_end,
                             row_tile, row_tile_stride, row_tile_phase, flags)


class Session:
    """One renderer context on one library (the analogue of one reference `Renderer`)."""

    def __init__(self, lib: abi.Lib, device=-1):
        """device: a CUDA device index, or a list of indices for a multi-device context (vp_create_multi)."""
        self.lib = lib
        if isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*[int(d) for d in device])
            self.ctx = lib.c_create_multi(arr, len(device))
        else:
            self.ctx = lib.c_create(device)
        if not self.ctx:
            msg = lib.c_last_error().decode() if lib.has("last_error") else ""
            raise abi.VpError(f"{lib.prefix}create failed: {msg}")
        self._keep = []

    def close(self):
        if self.ctx:
            self.lib.c_destroy(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads -------------------------------------------------------------------
    def set_scene(self, s: SceneDesc):
        L = self.lib
        vols = []
        self._keep = []
        for v in s.volumes:
            g = np.ascontiguousarray(v.grid, dtype=np.uint8).reshape(-1)
            assert g.size == v.n ** 3
            self._keep.append(g)
            vols.append(abi.VpVolume(g.ctypes.data, v.n, abi.f3(v.cube_min), abi.f3(v.cube_max), abi.f16(v.matrix),
                                     abi.f16(v.inv_matrix)))
        L.check(L.c_set_volumes(self.ctx, abi._arr(abi.VpVolume, vols), len(vols)), "set_volumes")
        if s.materials is not None:
            m = np.ascontiguousarray(s.materials, dtype=np.float32).reshape(-1, 6)
            mats = [abi.VpMaterial(abi.f3(r[:3]), float(r[3]), float(r[4]), float(r[5])) for r in m]
            L.check(L.c_set_materials(self.ctx, abi._arr(abi.VpMaterial, mats), len(mats)), "set_materials")
        pl = [abi.VpPointLight(abi.f3(p), abi.f3(c)) for p, c in s.point_lights]
        al = [abi.VpAreaLight(abi.f3(p), abi.f3(c), float(m), float(r)) for p, c, m, r in s.area_lights]
        sl = [abi.VpSpotLight(abi.f3(p), abi.f3(d), abi.f3(c), float(a)) for p, d, c, a in s.spot_lights]
        dl = abi.VpDirLight(abi.f3(s.dir_light[0]), abi.f3(s.dir_light[1]))
        L.check(L.c_set_lights(self.ctx, abi._arr(abi.VpPointLight, pl), len(pl), abi._arr(abi.VpAreaLight, al), len(al),
                               abi._arr(abi.VpSpotLight, sl), len(sl), C.byref(dl), int(s.area_shadow_samples)),
                "set_lights")
        sp = [abi.VpSphere(abi.f3(c), float(r), int(m)) for c, r, m in s.spheres]
        tr = [abi.VpTriangle(abi.f3(a), abi.f3(b), abi.f3(c), abi.f3(p), int(m)) for a, b, c, p, m in s.triangles]
        L.check(L.c_set_prims(self.ctx, abi._arr(abi.VpSphere, sp), len(sp), abi._arr(abi.VpTriangle, tr), len(tr)),
                "set_prims")
        if s.sky is not None:
            sky = np.ascontiguousarray(s.sky, dtype=np.float32)
            h, w = sky.shape[0], sky.shape[1]
            L.check(L.c_set_sky(self.ctx, sky.ctypes.data, w, h, float(s.sky_contribution), int(s.sky_active)),
                    "set_sky")
        else:
            L.check(L.c_set_sky(self.ctx, None, 0, 0, float(s.sky_contribution), 0), "set_sky")

    # ---- Renderer::Update ------------------------------------------------------------
    def render(self, cam: Camera, fp, accum=None, want_rgb8=True):
        """`fp.spp` Update() passes.  accum: float32 [H, W, 4] in-out (zeros if None)."""
        W, H = fp.width, fp.height
        if accum is None:
            accum = np.zeros((H, W, 4), dtype=np.float32)
        assert accum.dtype == np.float32 and accum.flags.c_contiguous and accum.size == W * H * 4
        rgb8 = np.zeros((H, W), dtype=np.uint32) if want_rgb8 else None
        st = abi.VpStats()
        cc = cam.to_c()
        self.lib.check(self.lib.c_render(self.ctx, C.byref(cc), C.byref(fp), accum.ctypes.data,
                                         rgb8.ctypes.data if want_rgb8 else None, C.byref(st)), "render")
        return accum, rgb8, st

    def render_async(self, cam: Camera, fp, accum_in, accum_out, rgb8_out):
        """vp_render_async: returns the ticket; the arrays must stay alive and untouched until wait(ticket)."""
        t = C.c_uint64()
        cc = cam.to_c()
        self.lib.check(self.lib.c_render_async(self.ctx, C.byref(cc), C.byref(fp), accum_in.ctypes.data if accum_in is not None else None,
                                               accum_out.ctypes.data, rgb8_out.ctypes.data if rgb8_out is not None else None, C.byref(t)),
                       "render_async")
        return t.value

    def wait(self, ticket):
        self.lib.check(self.lib.c_render_wait(self.ctx, ticket), "render_wait")

    def info(self):
        import json
        buf = C.create_string_buffer(4096)
        self.lib.check(self.lib.c_info(self.ctx, buf, 4096), "info")
        return json.loads(buf.value.decode())

    def render_reproject(self, cam: Camera, prev, fp, history=None, want_buffers=True):
        """One frame of the reprojection mode (Renderer::Tick static-camera branch).  prev: abi.VpPrevCamera;
        history: float32 [H, W, 3] of the previous frame (zeros if None).  Returns (new history, rgb8, albedo,
        illumination, stats)."""
        W, H = fp.width, fp.height
        if history is None:
            history = np.zeros((H, W, 3), dtype=np.float32)
        assert history.dtype == np.float32 and history.flags.c_contiguous and history.size == W * H * 3
        out = np.zeros((H, W, 3), dtype=np.float32)
        rgb8 = np.zeros((H, W), dtype=np.uint32)
        alb = np.zeros((H, W, 3), dtype=np.float32) if want_buffers else None
        ill = np.zeros((H, W, 3), dtype=np.float32) if want_buffers else None
        st = abi.VpStats()
        cc = cam.to_c()
        self.lib.check(self.lib.c_render_reproject(self.ctx, C.byref(cc), C.byref(prev), C.byref(fp), history.ctypes.data, out.ctypes.data,
                                                   rgb8.ctypes.data, alb.ctypes.data if want_buffers else None,
                                                   ill.ctypes.data if want_buffers else None, C.byref(st)), "render_reproject")
        return out, rgb8, alb, ill, st

    def trace_primary(self, cam: Camera, fp):
        hits = np.zeros(fp.width * fp.height, dtype=abi.HIT_DTYPE)
        cc = cam.to_c()
        self.lib.check(self.lib.c_trace_primary(self.ctx, C.byref(cc), C.byref(fp), hits.ctypes.data), "trace_primary")
        return hits.reshape(fp.height, fp.width)

    # ---- batch seams -------------------------------------------------------------------
    def find_nearest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        hits = np.zeros(rays.size, dtype=abi.HIT_DTYPE)
        self.lib.check(self.lib.c_find_nearest(self.ctx, rays.ctypes.data, rays.size, hits.ctypes.data), "find_nearest")
        return hits

    # ---- gameplay side channels --------------------------------------------------------
    def in_light(self):
        """Renderer::inLight since the previous call (cleared by the call)."""
        v = C.c_int()
        self.lib.check(self.lib.c_in_light(self.ctx, C.byref(v)), "in_light")
        return bool(v.value)

    def find_nearest_player(self, rays, current_chunk=0):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        hits = np.zeros(rays.size, dtype=abi.HIT_DTYPE)
        self.lib.check(self.lib.c_find_nearest_player(self.ctx, rays.ctypes.data, rays.size, int(current_chunk), hits.ctypes.data), "find_nearest_player")
        return hits

    def focus_distance(self, cam: Camera, width, height, seed):
        v = C.c_float()
        cc = cam.to_c()
        self.lib.check(self.lib.c_focus_distance(self.ctx, C.byref(cc), int(width), int(height), int(seed) & 0xFFFFFFFF, C.byref(v)), "focus_distance")
        return np.float32(v.value)

    def is_occluded(self, rays):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        out = np.zeros(rays.size, dtype=np.uint8)
        self.lib.check(self.lib.c_is_occluded(self.ctx, rays.ctypes.data, rays.size, out.ctypes.data), "is_occluded")
        return out

    def trace(self, rays, seeds, depth):
        rays = np.ascontiguousarray(rays, dtype=abi.RAY_DTYPE)
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        assert seeds.size == rays.size
        rgb = np.zeros((rays.size, 3), dtype=np.float32)
        self.lib.check(self.lib.c_trace(self.ctx, rays.ctypes.data, seeds.ctypes.data, rays.size, int(depth),
                                        rgb.ctypes.data), "trace")
        return rgb


def make_rays(o, d, t=1e34, inside=0):
    o = np.asarray(o, dtype=np.float32).reshape(-1, 3)
    d = np.asarray(d, dtype=np.float32).reshape(-1, 3)
    r = np.zeros(o.shape[0], dtype=abi.RAY_DTYPE)
    r["o"], r["d"] = o, d
    r["t"] = t
    r["inside"] = inside
    return r
