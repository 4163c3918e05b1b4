"""The C-ABI library loads and exports every symbol include/voxpop.h declares; error paths
that need no GPU behave as documented."""
import ctypes as C
import os
import re
import subprocess

import pytest
from _util import ROOT, abi, vp

HEADER = os.path.join(ROOT, "include", "voxpop.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vp_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported():
    names = declared_symbols()
    assert set(names) == {"vp_" + s for s in abi.ABI_SYMBOLS}, "abi.ABI_SYMBOLS out of date with include/voxpop.h"
    out = subprocess.check_output(["nm", "-D", "--defined-only", vp.PRODUCT_SO], text=True)
    exported = set(re.findall(r" T (vp_[a-z_0-9]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, f"libvoxpop_b200.so lacks {missing}"


def test_library_is_sm100a_only_and_torch_free():
    sass = subprocess.check_output(["cuobjdump", "--list-elf", vp.PRODUCT_SO], text=True)
    archs = set(re.findall(r"sm_\d+a?", sass))
    assert archs == {"sm_100a"}, archs
    deps = subprocess.check_output(["ldd", vp.PRODUCT_SO], text=True)
    assert "torch" not in deps and "python" not in deps


def test_struct_layouts_match_header():
    assert C.sizeof(abi.VpVolume) == 8 + 4 + 12 + 12 + 64 + 64 + 4  # pointer + n + cube + 2 matrices + tail pad
    assert C.sizeof(abi.VpMaterial) == 24 and C.sizeof(abi.VpCamera) == 80
    assert C.sizeof(abi.VpFrameParams) == 52 and C.sizeof(abi.VpStats) == 64
    assert abi.HIT_DTYPE.itemsize == 28 and abi.RAY_DTYPE.itemsize == 32


def test_no_device_is_a_loud_error():
    """On a box without a GPU vp_create must fail with a message - never fall back to CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = vp.load_product()
    ctx = lib.c_create(0)
    assert not ctx
    assert b"no CUDA device" in lib.c_last_error() or b"CPU fallback" in lib.c_last_error()
    with pytest.raises(abi.VpError):
        vp.Session(lib)
