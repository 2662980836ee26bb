"""cwipc_util_b200 -- B200-native drop-in for the filter hot path of cwi-dis/cwipc_util.

The product is the C-ABI shared library lib/libcwipc_util_cuda.so (C++17 + CUDA, sm_100a, no
PyTorch); this package holds its sources (csrc/), the build recipe (build.py), a ctypes mirror of
the reference's python/cwipc/util.py for this path (util.py) and the synthetic workload
generator used by tests and bench (synthetic.py).
"""
from .util import *  # noqa: F401,F403
from . import util, synthetic  # noqa: F401
