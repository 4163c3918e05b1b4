#!/usr/bin/env python3
"""Decode the reference's MagicaVoxel assets into the uint8 grids the render path consumes
and commit them as small compressed fixtures (data/scenes/*.npz).

Runs in the build container only (needs /root/reference and oracle/_ref/libvpref.so):
the grids are produced by the REFERENCE's own Scene::LoadModel (template/scene.cpp:441-524,
ogt_vox decode, y/z swap, palette index == material id), so they are exactly the bytes the
reference would trace.  The GPU box never reads /root/reference; it reads these fixtures.
Each file holds: grid (n^3 uint8, 255 = empty), n, and the 256x6 material table after the
load (LoadModel writes palette colours into materials[index], scene.cpp:516-520).
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
vp = importlib.import_module("raytracer-voxpopuli_b200")
abi = vp.abi

ASSETS = {  # name -> grid size used by the configs
    "teapot": 128, "room": 128, "roomGlass": 128,
    "SmallBuilding01": 64, "SmallBuilding02": 64, "TallBuilding01": 64,
    "monu1": 64, "monu2": 64, "monu3": 64, "player": 16,
}


def main():
    ref = abi.Lib(os.path.join(ROOT, "oracle/_ref/libvpref.so"), "vpref_")
    ref.dll.vpref_load_vox.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p]
    out = os.path.join(ROOT, "data/scenes")
    os.makedirs(out, exist_ok=True)
    for name, n in ASSETS.items():
        ctx = ref.c_create(0)  # fresh Renderer -> fresh default material table per asset
        grid = np.zeros(n ** 3, dtype=np.uint8)
        mats = (abi.VpMaterial * 256)()
        path = f"/root/reference/assets/{name}.vox".encode()
        ref.dll.vpref_load_vox(ctx, path, n, grid.ctypes.data, mats)
        ref.c_destroy(ctx)
        m = np.array([[x.albedo[0], x.albedo[1], x.albedo[2], x.roughness, x.emissive, x.ior] for x in mats], np.float32)
        vals, cnt = np.unique(grid[grid != 255], return_counts=True)
        print(f"{name}: n={n} solid={int((grid != 255).sum())} materials={dict(zip(vals.tolist(), cnt.tolist()))}")
        np.savez_compressed(os.path.join(out, f"{name}.npz"), grid=grid, n=np.int32(n), materials=m)


if __name__ == "__main__":
    main()
