"""voxpop-b200: B200-native render hot path of Raytracer-VoxPopuli behind a C ABI.

The package holds only what the path needs: the CUDA library (csrc/ -> libvoxpop_b200.so),
the host-side mirror of the reference's Renderer/Scene/Camera/Material data model
(host/ -> libvoxpop_host.so, C++) and this thin ctypes layer used by tests and bench.py.
It never imports anything from oracle/.
"""
import os

from . import abi, session  # noqa: F401
from .session import Camera, SceneDesc, Session, Volume, frame_params, make_rays  # noqa: F401

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
PRODUCT_SO = os.environ.get("VP_PRODUCT_SO") or os.path.join(PKG_DIR, "libvoxpop_b200.so")  # the override selects an experiment build (build.build_variant)
HOST_SO = os.path.join(PKG_DIR, "libvoxpop_host.so")


def load_product():
    """Load the CUDA library.  Fails loudly when it has not been built: there is no fallback."""
    if not os.path.exists(PRODUCT_SO):
        raise RuntimeError(f"{PRODUCT_SO} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback for the render path")
    return abi.Lib(PRODUCT_SO, "vp_")
