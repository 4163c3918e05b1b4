import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from _util import oracle_lib
    return oracle_lib()


@pytest.fixture(scope="session")
def ref():
    from _util import ref_lib
    lib = ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libvpref.so not built (needs /root/reference; python oracle/ref/build_ref.py)")
    return lib


@pytest.fixture(scope="session")
def product():
    """The CUDA library through the C ABI.  No fallback: a missing .so or device is an error."""
    from _util import vp
    return vp.load_product()
