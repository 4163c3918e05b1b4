"""ctypes face of libvoxpop_host.so: the C++ restatement of the reference's scene-authoring
code (Scene::SetTransform, Scene::LoadModel + .vox reader, Renderer::MaterialSetUp, the
camera frame of Camera::HandleInput).  CPU only; used to build the inputs of every config."""
import ctypes as C
import os

import numpy as np

from . import abi
from .session import Camera

_SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvoxpop_host.so")
_dll = None


def dll():
    global _dll
    if _dll is None:
        if not os.path.exists(_SO):
            raise RuntimeError(f"{_SO} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
        _dll = C.CDLL(_SO)
        _dll.vph_make_transform.argtypes = [C.c_void_p] * 8
        _dll.vph_default_materials.argtypes = [C.c_void_p]
        _dll.vph_load_vox.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
        _dll.vph_load_vox.restype = C.c_int
        _dll.vph_camera_look_at.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(abi.VpCamera)]
        _dll.vph_cos_degrees.argtypes = [C.c_float]
        _dll.vph_cos_degrees.restype = C.c_float
        _dll.vph_normalize.argtypes = [C.c_void_p, C.c_void_p]
    return _dll


def _f3(v):
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(3))


def make_transform(cube_position, rotation=(0, 0, 0), scale=(1, 1, 1), member_position=(0, 0, 0)):
    """Scene(cube_position) + SetScale + SetTransform -> (matrix16, inv_matrix16, cube_min, cube_max)."""
    cp, mp, r, s = _f3(cube_position), _f3(member_position), _f3(rotation), _f3(scale)
    M, iM = np.zeros(16, np.float32), np.zeros(16, np.float32)
    bmin, bmax = np.zeros(3, np.float32), np.zeros(3, np.float32)
    dll().vph_make_transform(cp.ctypes.data, mp.ctypes.data, r.ctypes.data, s.ctypes.data, M.ctypes.data, iM.ctypes.data,
                             bmin.ctypes.data, bmax.ctypes.data)
    return M, iM, bmin, bmax


def default_materials():
    m = np.zeros((256, 6), np.float32)
    dll().vph_default_materials(m.ctypes.data)
    return m


def load_vox(path, n):
    """Scene::LoadModel into an n^3 grid -> (grid uint8[n^3], materials float32[256,6])."""
    grid = np.zeros(n ** 3, np.uint8)
    mats = np.zeros((256, 6), np.float32)
    rc = dll().vph_load_vox(os.fsencode(path), n, grid.ctypes.data, mats.ctypes.data)
    if rc != 0:
        raise IOError(f"cannot load MagicaVoxel model {path}")
    return grid, mats


def look_at(pos, target, width, height, focal_distance=1.0, defocus_jitter=0.0):
    out = abi.VpCamera()
    p, t = _f3(pos), _f3(target)
    dll().vph_camera_look_at(p.ctypes.data, t.ctypes.data, int(width), int(height), float(focal_distance),
                             float(defocus_jitter), C.byref(out))
    g = lambda a: np.array(list(a), np.float32)
    return Camera(g(out.cam_pos), g(out.top_left), g(out.top_right), g(out.bottom_left), g(out.right), g(out.up),
                  out.focal_distance, out.defocus_jitter)


def prev_camera(pos, target, width, height):
    """What Renderer::CopyToPrevCamera keeps of a camera looking from pos at target: position + frustum-plane normals."""
    out = abi.VpPrevCamera()
    p, t = _f3(pos), _f3(target)
    dll().vph_camera_prev(p.ctypes.data, t.ctypes.data, int(width), int(height), C.byref(out))
    return out


def cos_degrees(d):
    return float(dll().vph_cos_degrees(float(d)))


def normalize(v):
    a, o = _f3(v), np.zeros(3, np.float32)
    dll().vph_normalize(a.ctypes.data, o.ctypes.data)
    return tuple(float(x) for x in o)


class Xorshift32:
    """The reference's RNG (template/tmpl8math.cpp:119-133) for reproducible light placement."""

    def __init__(self, seed):
        self.s = np.uint32(seed)

    def next_float(self):
        s = int(self.s)
        s ^= (s << 13) & 0xFFFFFFFF
        s ^= s >> 17
        s ^= (s << 5) & 0xFFFFFFFF
        self.s = np.uint32(s)
        return float(np.float32(s) * np.float32(2.3283064365387e-10))

    def uniform(self, lo, hi):
        return float(np.float32(lo) + np.float32(self.next_float()) * np.float32(hi - lo))
