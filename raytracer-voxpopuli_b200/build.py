"""In-tree build of the native libraries (no JIT cache, no pip install):
  libvoxpop_b200.so      CUDA sm_100a kernels + C ABI      (csrc/vp_api.cu)
  libvoxpop_host.so      C++ scene-authoring mirror        (host/voxpop_host.cpp)
  libvoxpop_renderer.so  C++ Renderer mirror over the ABI  (host/voxpop_renderer.cpp)
  libvoxpop_probe.so     L2 bandwidth probe for bench.py   (csrc/vp_probe.cu)
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_variant(name, defines):
    """Experiment builds: libvoxpop_b200_<name>.so with extra -D macros (select with VP_PRODUCT_SO=<path>)."""
    so = os.path.join(PKG, f"libvoxpop_b200_{name}.so")
    _run(["nvcc"] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", so, os.path.join(PKG, "csrc", "vp_api.cu")])
    return so


def build(force=False, verbose_ptxas=False):
    csrc = os.path.join(PKG, "csrc")
    hdr = os.path.join(os.path.dirname(PKG), "include", "voxpop.h")
    cu = [os.path.join(csrc, "vp_api.cu")] + sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")) + [hdr]
    so = os.path.join(PKG, "libvoxpop_b200.so")
    if force or _stale(so, cu):
        _run(["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-o", so, cu[0]])
    probe_src = os.path.join(csrc, "vp_probe.cu")
    pso = os.path.join(PKG, "libvoxpop_probe.so")
    if force or _stale(pso, [probe_src]):
        _run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
              "-cudart", "static", "-o", pso, probe_src])
    host_src = [os.path.join(PKG, "host", f) for f in ("voxpop_host.cpp", "voxpop_host.hpp")] + [hdr]
    hso = os.path.join(PKG, "libvoxpop_host.so")
    if force or _stale(hso, host_src):
        _run(["g++"] + CXX_FLAGS + ["-o", hso, host_src[0]])
    rsrc = [os.path.join(PKG, "host", f) for f in ("voxpop_renderer.cpp", "voxpop_host.hpp")] + [hdr]
    rso = os.path.join(PKG, "libvoxpop_renderer.so")
    if os.path.exists(rsrc[0]) and (force or _stale(rso, rsrc + [so, hso])):
        _run(["g++"] + CXX_FLAGS + ["-o", rso, rsrc[0], "-L" + PKG, "-lvoxpop_host", "-lvoxpop_b200", "-Wl,-rpath,$ORIGIN"])
    return so, hso


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
