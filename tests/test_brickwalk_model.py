"""CPU model of the kernels' brick walk (tests/brickwalk_model.c) against the reference's cell walk:
hit / no hit, hit cell and the bits of t must be identical on every ray - random rays, rays with a
stopping distance (shadow rays), axis-parallel and grid-aligned rays (inf / NaN tmax), -0 directions and
exact diagonals (ties between axes).  Four formulations: explicit catch-up (mode 0), the unified step
(mode 1), the unified step with the forward-box rule, which is what the kernels use (mode 2), and that rule at the bit
level of the kernels' round-2 step body - Morton walk word, mirrored walk, axis carried as its lowest field bit, forward-box
table with the cell's one-hot mask, brick key in mirrored coordinates (mode 3: uw_step / walk_step under VP_STEP_V2)."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from _util import scenes

HERE = os.path.dirname(os.path.abspath(__file__))


class Stats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("brick_steps", "cell_steps", "to_cell", "to_brick", "rays", "ref_steps", "nan_rays")]


@pytest.fixture(scope="module")
def model():
    so = os.path.join(tempfile.mkdtemp(prefix="bw_"), "libbw.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "brickwalk_model.c"), "-lm"])
    lib = C.CDLL(so)
    lib.bw_compare.restype = C.c_int64
    return lib


def _ray_sets(n, count, rng):
    O = rng.uniform(0.02, 0.9, (count, 3)).astype(np.float32)
    D = rng.normal(size=(count, 3)).astype(np.float32)
    D /= np.linalg.norm(D, axis=1, keepdims=True)
    far = np.full(count, 1e34, np.float32)
    yield "inside", O, D, far
    yield "stopping distance", O, D, rng.uniform(0, 1.0, count).astype(np.float32)
    yield "outside", rng.uniform(-1.5, 2.5, (count, 3)).astype(np.float32), D, far
    D3 = D.copy()
    D3[rng.uniform(size=(count, 3)) < 0.3] = 0.0
    D3[np.all(D3 == 0, axis=1)] = (1, 0, 0)
    D3 /= np.linalg.norm(D3, axis=1, keepdims=True)
    O3 = (np.round(O * n) / n).astype(np.float32)
    yield "axis-parallel, grid-aligned", O3, D3, far
    D4 = D3.copy()
    D4[D4 == 0] = -0.0
    yield "negative zero", O3, D4, far
    D5 = np.sign(D).astype(np.float32)
    D5[rng.uniform(size=(count, 3)) < 0.2] *= 0.5
    D5 /= np.linalg.norm(D5, axis=1, keepdims=True)
    yield "diagonal ties", O3, D5.astype(np.float32), far
    # direction components so small that tdelta is huge, overflows to inf (1 / denormal) or tmax * 0 games start
    D6 = D.copy()
    tiny = np.array([1e-30, -1e-30, 1e-38, -1e-38, 1e-39, -1e-39, 1e-44, -1e-44], np.float32)
    sel = rng.uniform(size=(count, 3)) < 0.3
    D6[sel] = rng.choice(tiny, int(sel.sum()))
    yield "tiny components", O, D6.astype(np.float32), far
    yield "tiny components, grid-aligned", O3, D6.astype(np.float32), far


@pytest.mark.parametrize("name", ["room", "monu2", "player"])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_brick_walk_equals_cell_walk(model, name, mode):
    grid, n, _ = scenes.load_grid(name)
    grid = np.ascontiguousarray(grid.reshape(-1))
    if mode == 3:
        jr = 0
        while (16 << jr) < n:
            jr += 1
        fine = np.zeros(64 << (3 * jr), np.uint64)
        model.bw_build_fine_morton(grid.ctypes.data_as(C.c_void_p), n, jr, fine.ctypes.data_as(C.c_void_p))
    else:
        fine = np.zeros((n // 4) ** 3, np.uint64)
        model.bw_build_fine(grid.ctypes.data_as(C.c_void_p), n, fine.ctypes.data_as(C.c_void_p))
    rng = np.random.default_rng(5 + mode)
    for what, O, D, tl in _ray_sets(n, 60000, rng):
        O, D, tl = (np.ascontiguousarray(a, np.float32) for a in (O, D, tl))
        st, first = Stats(), C.c_int64()
        bad = model.bw_compare(grid.ctypes.data_as(C.c_void_p), fine.ctypes.data_as(C.c_void_p), n, O.ctypes.data_as(C.c_void_p),
                               D.ctypes.data_as(C.c_void_p), tl.ctypes.data_as(C.c_void_p), C.c_int64(len(O)), mode, C.byref(st),
                               C.byref(first), None)
        assert bad == 0, f"{name} / {what}: {bad} rays differ, first {first.value}"
        assert st.rays > 0 and st.ref_steps >= st.rays
        if name == "room" and what == "inside":
            assert st.brick_steps > 0 and st.brick_steps + st.cell_steps < st.ref_steps  # the walk really skips cells


@pytest.mark.parametrize("n", [20, 37, 44, 100, 256])
def test_bit_level_step_on_padded_grids(model, n):
    """Grid sizes that are not a power of two (and not a multiple of the brick size): the kernels walk the empty padding
    of the (16 << jr)^3 lattice until the masked add carries out; hits, cells and t must still equal the cell walk."""
    rng = np.random.default_rng(100 + n)
    grid = np.full((n, n, n), 255, np.uint8)
    solid = rng.uniform(size=grid.shape) < 0.02
    solid[:, :, : max(1, n // 10)] = True   # a floor, so that long walks end in hits as well
    grid[solid] = 1
    grid = np.ascontiguousarray(grid.transpose(2, 1, 0).reshape(-1))  # x fastest
    jr = 0
    while (16 << jr) < n:
        jr += 1
    fine = np.zeros(64 << (3 * jr), np.uint64)
    model.bw_build_fine_morton(grid.ctypes.data_as(C.c_void_p), n, jr, fine.ctypes.data_as(C.c_void_p))
    for what, O, D, tl in _ray_sets(n, 40000, rng):
        O, D, tl = (np.ascontiguousarray(a, np.float32) for a in (O, D, tl))
        st, first = Stats(), C.c_int64()
        bad = model.bw_compare(grid.ctypes.data_as(C.c_void_p), fine.ctypes.data_as(C.c_void_p), n, O.ctypes.data_as(C.c_void_p),
                               D.ctypes.data_as(C.c_void_p), tl.ctypes.data_as(C.c_void_p), C.c_int64(len(O)), 3, C.byref(st),
                               C.byref(first), None)
        assert bad == 0, f"n={n} / {what}: {bad} rays differ, first {first.value}"
        assert st.rays > 0
