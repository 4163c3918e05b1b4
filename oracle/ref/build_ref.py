#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY - builds oracle/_ref/libvpref.so: the reference's OWN
render path (renderer.cpp, template/scene.cpp, template/tmpl8math.*, template/camera.h,
src/**) compiled headless on Linux, driven through oracle/ref/ref_harness.cpp.

Recipe (nothing of the reference is committed to this repo):
  1. read the reference sources where they lie under /root/reference,
  2. stream them through the portability patches below into a scratch directory
     oracle/_ref/build/ (git-ignored),
  3. compile ONE translation unit (ref_harness.cpp, which #includes the patched
     .cpp files after the shim precomp.h) with g++ into oracle/_ref/libvpref.so,
  4. delete the scratch directory (only the .so stays; it travels to the GPU box).

The reference's own build system (MSVC .vcxproj) is not used.  Each patch is either a
portability fix (MSVC-only syntax) or a determinism pin that selects the reference's
own alternative code path; none of them introduces new arithmetic:

  P1  `#include "precomp.h"` dropped from every .cpp (the harness includes the shim once).
  P2  `__declspec(align(n))`           -> `alignas(n)`            (tmpl8math.h:2569,2596)
  P3  anonymous `struct { float3 X; float dummyN; };` inside unions -> plain `float3 X;`
      (scene.h:96-138; g++ rejects non-POD members of anonymous structs).  Layout is
      unchanged: float3 occupies the first 12 bytes of the 16-byte __m128 union either way.
      aabb's unused `struct { float3 bmin3; };` members are dropped (tmpl8math.h:3252-3266).
  P4  `.m128_f32[i]`                    -> `[i]`                   (tmpl8math.cpp:359-390)
  P5  `_mm256_cos_ps/_mm256_sin_ps` (MSVC SVML) -> lane-wise cosf/sinf helper
      (renderer.cpp:1696-1697; AVX2 branch only - the harness drives the scalar branch).
  P6  SCRWIDTH/SCRHEIGHT macros -> runtime ints; `constexpr` users of them -> inline
      functions; the W*H `std::array` reprojection buffers -> `std::vector`
      (camera.h:4-8,174; renderer.h:240-244).
  P7  determinism pin: `FastReciprocal` (_mm_rcp_ps, vendor-specific bits) -> the
      reference's own `SlowReciprocal` branch (`#if 0` -> `#if 1`, renderer.cpp:966,1041).
  P8  `std::execution::par` (silently serial without TBB) is left as is: the harness
      re-states the scalar pixel loop of Update() (renderer.cpp:1851-1881) under OpenMP
      with a per-pixel reseed, calling the reference's Camera/Trace/Blend/tonemap.
  P9  `materials.resize(MaterialType::NONE)` -> 256 entries: LoadModel writes
      materials[255] for palette index 255 (scene.cpp:516-520 vs renderer.cpp:434-441),
      a heap overflow in the reference; entry 255 is never read on the path
      (255 == NONE == empty).
  P10 ray counters: one `++` at the top of Renderer::FindNearest / IsOccluded and of
      Scene::FindMaterialExit / FindSmokeExit, plus DDA-step counters (instrumentation).
      The hit cell index and pre-matrix face are also recorded (Scene::FindNearest hit block,
      Ray::GetNormalVoxel) for the bit-exact primary-hit dumps.
  P11 the stdout print inside Trace's smoke branch (renderer.cpp:1238) is silenced.

Two libraries are built from the same sources:
  libvpref.so       all patches - the parity checker (DDA-step counters, hit cell / face records);
  libvpref_fast.so  without P10e / P10f / P10g, i.e. nothing added inside the DDA loops (the per-step
                    `++g_vp_cnt.steps` is a thread-local read-modify-write in the reference's hottest loop) -
                    the build bench.py TIMES (--impl reference, cpu_baseline).  Its dda_steps counter reads 0
                    and vpref_trace_primary reports no cell / face; the per-ray counters (one increment per
                    Renderer::FindNearest / IsOccluded call) stay, they define the metric.
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.dirname(HERE)
REF = os.environ.get("VP_REFERENCE", "/root/reference")
OUT = os.path.join(ORACLE, "_ref")
BUILD = os.path.join(OUT, "build")

FILES = [
    "renderer.cpp", "renderer.h",
    "template/scene.cpp", "template/scene.h", "template/camera.h", "template/common.h",
    "template/tmpl8math.h", "template/tmpl8math.cpp",
    "src/BVH/BasicBVH.cpp", "src/BVH/BasicBVH.h", "src/BVH/Shapes.cpp", "src/BVH/Shapes.h",
    "src/Game/ModifyingProp.cpp", "src/Game/ModifyingProp.h",
    "src/Game/PlayerCharacter.cpp", "src/Game/PlayerCharacter.h",
    "src/Lighting/DirectionalLight.h", "src/Lighting/PointLight.h", "src/Lighting/SkyDome.h",
    "src/Lighting/SphereAreaLight.h", "src/Lighting/SpotLight.h",
    "src/Materials/Material.cpp", "src/Materials/Material.h",
    "lib/ogt_vox.h", "lib/stb_image.h",
]


def sub(text, pattern, repl, name, count=0, flags=0, expect=None):
    new, n = re.subn(pattern, repl, text, count=count, flags=flags)
    if (expect is not None and n != expect) or (expect is None and n == 0):
        raise SystemExit(f"build_ref: patch '{name}' matched {n} times (expected {expect or '>=1'})")
    return new


def patch(rel, text, timing=False):
    text = text.lstrip("﻿")
    if rel.endswith(".cpp"):
        text = re.sub(r'#include "precomp.h"', "", text)                                  # P1
    if rel == "template/tmpl8math.h":
        text = sub(text, r"__declspec\(align\((\d+)\)\)", r"alignas(\1)", "P2", expect=2)
        text = sub(text, r"struct\s*\{\s*float3 (bm(?:in|ax)3);\s*\};", r"/* \1 unused */", "P3-aabb", expect=2)
    if rel == "template/scene.h":
        text = sub(text, r"struct\s*\{\s*float3 (\w+);\s*float dummy\d;\s*\};", r"float3 \1;", "P3-ray", expect=4)
    if rel == "template/tmpl8math.cpp":
        text = sub(text, r"\.m128_f32\[", "[", "P4")
        text = sub(text, r"thread_local static uint32_t seed", "thread_local uint32_t seed", "seed-visible", expect=1)
    if rel == "template/camera.h":
        text = sub(text, r"#define SCRWIDTH\s+256", "#define SCRWIDTH g_vp_scrwidth", "P6-w", expect=1)
        text = sub(text, r"#define SCRHEIGHT\s+212", "#define SCRHEIGHT g_vp_scrheight", "P6-h", expect=1)
        text = sub(text, r"constexpr float HALF_PIXEL_W = ([^;]*);", r"#define HALF_PIXEL_W (\1)", "P6-hpw", expect=1)
        text = sub(text, r"constexpr float HALF_PIXEL_H = ([^;]*);", r"#define HALF_PIXEL_H (\1)", "P6-hph", expect=1)
        text = sub(text, r"static constexpr float ASPECT = ([^;]*);", "", "P6-aspect-decl", expect=1)
        text = text.replace("#define DOUBLESIZE",
                            "#define DOUBLESIZE\n#define ASPECT (static_cast<float>(SCRWIDTH) / static_cast<float>(SCRHEIGHT))")
    if rel == "renderer.h":
        text = sub(text, r"std::array<(\w+), SCRWIDTH \* SCRHEIGHT>", r"std::vector<\1>", "P6-buffers", expect=5)
    if rel == "renderer.cpp":
        text = sub(text, r"_mm256_cos_ps\(", "vp_mm256_cos_ps(", "P5-cos", expect=1)
        text = sub(text, r"_mm256_sin_ps\(", "vp_mm256_sin_ps(", "P5-sin", expect=1)
        text = sub(text, r"#if 0\s*\n(\s*__m128 rDSSE = SlowReciprocal)", r"#if 1\n\1", "P7", expect=2)
        text = sub(text, r"materials\.resize\(MaterialType::NONE\);", "materials.resize(256);", "P9a", expect=1)
        text = sub(text, r"for \(; i < MaterialType::NONE; \+\+i\)", "for (; i < 256; ++i)", "P9b", expect=1)
        text = sub(text, r"(int32_t Renderer::FindNearest\(Ray& ray\)\s*\{)", r"\1 ++g_vp_cnt.nearest;", "P10a", expect=1)
        text = sub(text, r"(bool Renderer::IsOccluded\(Ray& ray\) const\s*\{)", r"\1 ++g_vp_cnt.shadow;", "P10b", expect=1)
        text = sub(text, r'std::cout << "Player is in the light\\n";', "", "P11")
        # the reprojection-buffer copy at renderer.cpp:2100 stays valid for vectors.
    if rel == "template/scene.cpp":
        text = sub(text, r"(bool Scene::FindMaterialExit\(Ray& ray, MaterialType::MatType matType\) const\s*\{)",
                   r"\1 ++g_vp_cnt.exitray;", "P10c", expect=1)
        text = sub(text, r"(bool Scene::FindSmokeExit\(Ray& ray\) const\s*\{)", r"\1 ++g_vp_cnt.exitray;", "P10d", expect=1)
        if timing:
            return text
        # one DDA step == one grid load (scene.cpp:765,829,889,954,1017)
        text = sub(text, r"(const (?:MaterialType::MatType|auto) cell = grid\[GetVoxel\(s\.X, s\.Y, s\.Z\)\];)",
                   r"\1 ++g_vp_cnt.steps;", "P10e", expect=5)
        # hit-cell / face instrumentation for the bit-exact primary-hit dumps
        text = sub(text, r"(ray\.indexMaterial = cell;\s*return true;)",
                   r"g_vp_lastcell = s.X + s.Y * gridsize + s.Z * gridsize2; \1", "P10f", expect=4)
        text = sub(text, r"(?m)^(\s*float3 normal = float3\(mind == d\.x[^;]*;)", r"\1 g_vp_lastface = vp_face_code(normal);",
                   "P10g", expect=1)
    return text


def build_one(so_name, timing):
    shutil.rmtree(BUILD, ignore_errors=True)
    os.makedirs(BUILD, exist_ok=True)
    for rel in FILES:
        with open(os.path.join(REF, rel), encoding="utf-8", errors="surrogateescape") as f:
            text = f.read()
        dst = os.path.join(BUILD, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w", encoding="utf-8", errors="surrogateescape") as f:
            f.write(patch(rel, text, timing))
    so = os.path.join(OUT, so_name)
    cmd = [
        "g++", "-std=c++17", "-O2", "-g0", "-fPIC", "-shared", "-fopenmp",
        "-mavx2", "-mfma", "-mbmi2", "-ffp-contract=off", "-fno-fast-math",
        "-w", "-fpermissive",
        "-I", os.path.join(HERE, "shim"),
        "-I", BUILD, "-I", os.path.join(BUILD, "template"), "-I", os.path.join(BUILD, "src"),
        "-I", os.path.join(BUILD, "lib"),
        os.path.join(HERE, "ref_harness.cpp"),
        "-Wl,-Bsymbolic", "-Wl,--version-script=" + os.path.join(HERE, "vpref.map"), "-o", so, "-lm",
    ]
    print(" ".join(cmd))
    r = subprocess.run(cmd)
    if r.returncode == 0 and not os.environ.get("VP_KEEP_BUILD"):
        shutil.rmtree(BUILD, ignore_errors=True)
    return r.returncode


def main():
    if not os.path.isdir(REF):
        print(f"build_ref: {REF} not present - keeping any prebuilt oracle/_ref/libvpref*.so")
        return 0
    os.makedirs(OUT, exist_ok=True)
    jobs = [("libvpref.so", False), ("libvpref_fast.so", True)]
    if "--timing" in sys.argv:
        jobs = jobs[1:]
    if "--parity" in sys.argv:
        jobs = jobs[:1]
    for so_name, timing in jobs:
        rc = build_one(so_name, timing)
        if rc:
            return rc
    return 0


if __name__ == "__main__":
    sys.exit(main())
