"""ctypes mirror of include/voxpop.h.

One binding class serves the three libraries that share these PODs:
  prefix "vp_"    -> libvoxpop_b200.so       (CUDA, the product)
  prefix "vpo_"   -> oracle/libvporacle.so   (plain-C restatement; tests/bench-baseline only)
  prefix "vpref_" -> oracle/_ref/libvpref.so (the reference itself; tests/bench-baseline only)
This module only describes the ABI; which library gets loaded is the caller's business.
"""
import ctypes as C

import numpy as np

F3 = C.c_float * 3
F16 = C.c_float * 16


class VpVolume(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("n", C.c_uint32), ("cube_min", F3), ("cube_max", F3),
                ("matrix", F16), ("inv_matrix", F16)]


class VpMaterial(C.Structure):
    _fields_ = [("albedo", F3), ("roughness", C.c_float), ("emissive", C.c_float), ("ior", C.c_float)]


class VpPointLight(C.Structure):
    _fields_ = [("position", F3), ("color", F3)]


class VpAreaLight(C.Structure):
    _fields_ = [("position", F3), ("color", F3), ("color_multiplier", C.c_float), ("radius", C.c_float)]


class VpSpotLight(C.Structure):
    _fields_ = [("position", F3), ("direction", F3), ("color", F3), ("angle", C.c_float)]


class VpDirLight(C.Structure):
    _fields_ = [("direction", F3), ("color", F3)]


class VpSphere(C.Structure):
    _fields_ = [("center", F3), ("radius", C.c_float), ("material", C.c_uint32)]


class VpTriangle(C.Structure):
    _fields_ = [("vertex0", F3), ("vertex1", F3), ("vertex2", F3), ("position", F3), ("material", C.c_uint32)]


class VpCamera(C.Structure):
    _fields_ = [("cam_pos", F3), ("top_left", F3), ("top_right", F3), ("bottom_left", F3), ("right", F3),
                ("up", F3), ("focal_distance", C.c_float), ("defocus_jitter", C.c_float)]


class VpPrevCamera(C.Structure):
    _fields_ = [("cam_pos", C.c_float * 3), ("left_normal", C.c_float * 3), ("right_normal", C.c_float * 3),
                ("top_normal", C.c_float * 3), ("bottom_normal", C.c_float * 3)]


class VpFrameParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("max_bounces", C.c_int32), ("spp", C.c_int32),
                ("frame_index", C.c_uint32), ("seed", C.c_uint32), ("aa_strength", C.c_float),
                ("row_begin", C.c_int32), ("row_end", C.c_int32), ("row_tile", C.c_int32),
                ("row_tile_stride", C.c_int32), ("row_tile_phase", C.c_int32), ("flags", C.c_uint32)]


class VpShareInfo(C.Structure):
    _fields_ = [("handle", C.c_ubyte * 64), ("bytes", C.c_uint64), ("device", C.c_int32), ("pad", C.c_int32 * 13)]


class VpStats(C.Structure):
    _fields_ = [("primary_rays", C.c_uint64), ("secondary_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("exit_rays", C.c_uint64), ("dda_steps", C.c_uint64), ("sky_samples", C.c_uint64), ("device_ms", C.c_double),
                ("kernel_launches", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}

    @property
    def rays(self):
        return self.primary_rays + self.secondary_rays + self.shadow_rays


VP_OK, VP_EINVAL, VP_ENODEVICE, VP_ECUDA, VP_ENOMEM, VP_ESTATE = 0, -1, -2, -3, -4, -5
VP_FLAG_COUNT_STEPS = 1
VP_FLAG_MEGAKERNEL = 2
VP_FLAG_PERSISTENT = 4
VP_FLAG_TIME_STAGES = 8
VP_FLAG_DEFER = 16

HIT_DTYPE = np.dtype([("volume", "<i4"), ("cell", "<u4"), ("t", "<f4"), ("normal", "<f4", (3,)),
                      ("material", "u1"), ("face", "i1"), ("inside", "u1"), ("pad", "u1")])
RAY_DTYPE = np.dtype([("o", "<f4", (3,)), ("d", "<f4", (3,)), ("t", "<f4"), ("inside", "<u4")])
assert HIT_DTYPE.itemsize == 28 and RAY_DTYPE.itemsize == 32

# every symbol include/voxpop.h declares (checked by tests/test_abi.py against the header text)
ABI_SYMBOLS = [
    "create", "destroy", "last_error", "set_volumes", "update_volume_grid", "update_volume_transform",
    "set_materials", "set_lights", "set_prims", "set_sky", "render", "render_resident", "accum_reset",
    "accum_download", "render_device", "trace_primary", "find_nearest", "is_occluded", "trace", "info", "stats_begin", "stats_end", "stage_times", "check_sincos_pin",
    "render_fence", "render_reproject", "render_reproject_device", "in_light", "find_nearest_player", "focus_distance",
    "create_multi", "device_count", "render_async", "render_wait", "host_register", "host_unregister",
    "render_to_gl", "share_create", "share_open", "share_ptr", "share_signal", "share_wait", "share_error", "share_close",
]


def _arr(ctype, items):
    a = (ctype * max(1, len(items)))()
    for i, it in enumerate(items):
        a[i] = it
    return a


class VpError(RuntimeError):
    pass


class Lib:
    """Thin typed wrapper around one of the three shared libraries."""

    def __init__(self, path, prefix):
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        f = self._f
        f("create", C.c_void_p, [C.c_int])
        f("destroy", None, [C.c_void_p])
        f("set_volumes", C.c_int, [C.c_void_p, C.POINTER(VpVolume), C.c_int])
        f("set_materials", C.c_int, [C.c_void_p, C.POINTER(VpMaterial), C.c_int])
        f("set_lights", C.c_int, [C.c_void_p, C.POINTER(VpPointLight), C.c_int, C.POINTER(VpAreaLight), C.c_int,
                                  C.POINTER(VpSpotLight), C.c_int, C.POINTER(VpDirLight), C.c_int])
        f("set_prims", C.c_int, [C.c_void_p, C.POINTER(VpSphere), C.c_int, C.POINTER(VpTriangle), C.c_int])
        f("set_sky", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int])
        f("render", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.POINTER(VpFrameParams), C.c_void_p, C.c_void_p,
                              C.POINTER(VpStats)])
        f("render_reproject", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.POINTER(VpPrevCamera), C.POINTER(VpFrameParams), C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(VpStats)], optional=True)
        f("render_reproject_device", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.POINTER(VpPrevCamera), C.POINTER(VpFrameParams),
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(VpStats)], optional=True)
        f("in_light", C.c_int, [C.c_void_p, C.POINTER(C.c_int)], optional=True)
        f("find_nearest_player", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p], optional=True)
        f("focus_distance", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_float)], optional=True)
        f("trace_primary", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.POINTER(VpFrameParams), C.c_void_p])
        f("find_nearest", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p])
        f("is_occluded", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p])
        f("trace", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p])
        # product-only entry points (optional on the checker libraries)
        f("last_error", C.c_char_p, [], optional=True)
        f("update_volume_grid", C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p], optional=True)
        f("update_volume_transform", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
          optional=True)
        f("render_resident", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.POINTER(VpFrameParams), C.POINTER(VpStats)],
          optional=True)
        f("accum_reset", C.c_int, [C.c_void_p, C.c_int, C.c_int], optional=True)
        f("accum_download", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p], optional=True)
        f("render_device", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.POINTER(VpFrameParams), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.POINTER(VpStats)], optional=True)
        f("info", C.c_int, [C.c_void_p, C.c_char_p, C.c_int], optional=True)
        f("stats_begin", C.c_int, [C.c_void_p], optional=True)
        f("stats_end", C.c_int, [C.c_void_p, C.POINTER(VpStats)], optional=True)
        f("stage_times", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)], optional=True)
        f("check_sincos_pin", C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)], optional=True)
        f("render_fence", C.c_int, [C.c_void_p, C.c_void_p, C.c_int], optional=True)
        f("create_multi", C.c_void_p, [C.POINTER(C.c_int), C.c_int], optional=True)
        f("device_count", C.c_int, [C.c_void_p], optional=True)
        f("render_async", C.c_int, [C.c_void_p, C.POINTER(VpCamera), C.POINTER(VpFrameParams), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(C.c_uint64)], optional=True)
        f("render_wait", C.c_int, [C.c_void_p, C.c_uint64], optional=True)
        f("host_register", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64], optional=True)
        f("host_unregister", C.c_int, [C.c_void_p, C.c_void_p], optional=True)
        f("render_to_gl", C.c_int, [C.c_void_p, C.c_uint32], optional=True)
        f("share_create", C.c_void_p, [C.c_void_p, C.c_uint64, C.POINTER(VpShareInfo)], optional=True)
        f("share_open", C.c_void_p, [C.c_void_p, C.POINTER(VpShareInfo)], optional=True)
        f("share_ptr", C.c_void_p, [C.c_void_p], optional=True)
        f("share_signal", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32], optional=True)
        f("share_wait", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_double], optional=True)
        f("share_error", C.c_int, [C.c_void_p, C.POINTER(C.c_uint32)], optional=True)
        f("share_close", C.c_int, [C.c_void_p], optional=True)

    def _f(self, name, restype, argtypes, optional=False):
        try:
            fn = getattr(self.dll, self.prefix + name)
        except AttributeError:
            if optional:
                setattr(self, "c_" + name, None)
                return
            raise
        fn.restype, fn.argtypes = restype, argtypes
        setattr(self, "c_" + name, fn)

    def has(self, name):
        return getattr(self, "c_" + name, None) is not None

    def check(self, rc, what):
        if rc != 0:
            msg = self.c_last_error().decode() if self.has("last_error") else ""
            raise VpError(f"{self.prefix}{what} failed rc={rc}: {msg}")


def f3(v):
    return F3(*[float(x) for x in v])


def f16(m):
    return F16(*[float(x) for x in np.asarray(m, dtype=np.float32).reshape(16)])
