import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """The built libcwipc_util_cuda.so, loaded through the ctypes mirror."""
    from cwipc_util_b200 import build, util
    build.build_library()
    return util.cwipc_util_dll_load()


@pytest.fixture(scope="session")
def cw(lib):
    import cwipc_util_b200 as cw
    if cw.cuda_device_count() <= 0:
        pytest.fail("GPU test selected but libcwipc_util_cuda sees no CUDA device (there is no CPU fallback)")
    return cw


@pytest.fixture(scope="session")
def orc():
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle as orc_mod
    orc_mod.load()
    return orc_mod
