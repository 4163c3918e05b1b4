"""One-off sweep of the bit-level CPU model of the brick walk (tests/brickwalk_model.c, mode 3) against the reference's cell walk:
every scene grid plus synthetic grids of 24..200 cells per axis, the eight adversarial ray classes of the model test, four seeds
(28.8 M rays, half a minute on one core).  No GPU needed.  usage: python tools/brickwalk_campaign.py"""
import ctypes as C, os, subprocess, tempfile, sys, time
import numpy as np
sys.path.insert(0,'/root/repo/tests')
import test_brickwalk_model as T
from _util import scenes
so = os.path.join(tempfile.mkdtemp(prefix="bw_"), "libbw.so")
subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", so, "/root/repo/tests/brickwalk_model.c", "-lm"])
lib = C.CDLL(so); lib.bw_compare.restype = C.c_int64
def grids():
    for name in ("room", "monu2", "player", "teapot") if hasattr(scenes, "load_grid") else ():
        try:
            g, n, _ = scenes.load_grid(name)
            yield name, np.ascontiguousarray(g.reshape(-1)), n
        except Exception as e:
            print("skip", name, e)
    for n, dens in ((24, 0.3), (60, 0.001), (100, 0.05), (128, 0.0005), (200, 0.01)):
        rng = np.random.default_rng(n)
        g = np.full((n, n, n), 255, np.uint8)
        g[rng.uniform(size=g.shape) < dens] = 1
        yield f"synthetic{n}", np.ascontiguousarray(g.reshape(-1)), n
tot = 0; t0 = time.time()
for name, grid, n in grids():
    jr = 0
    while (16 << jr) < n: jr += 1
    finem = np.zeros(64 << (3 * jr), np.uint64)
    lib.bw_build_fine_morton(grid.ctypes.data_as(C.c_void_p), n, jr, finem.ctypes.data_as(C.c_void_p))
    for seed in range(4):
        rng = np.random.default_rng(1000 * seed + n)
        for what, O, D, tl in T._ray_sets(n, 100000, rng):
            O, D, tl = (np.ascontiguousarray(a, np.float32) for a in (O, D, tl))
            st, first = T.Stats(), C.c_int64()
            bad = lib.bw_compare(grid.ctypes.data_as(C.c_void_p), finem.ctypes.data_as(C.c_void_p), n, O.ctypes.data_as(C.c_void_p), D.ctypes.data_as(C.c_void_p), tl.ctypes.data_as(C.c_void_p), C.c_int64(len(O)), 3, C.byref(st), C.byref(first), None)
            tot += len(O)
            if bad: print("MISMATCH", name, seed, what, bad, first.value)
print("rays", tot, "time", round(time.time() - t0, 1))
