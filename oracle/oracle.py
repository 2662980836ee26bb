"""ctypes loader of the CPU oracle (oracle/libcwipc_oracle.so).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never from the cwipc_util_b200 package (tests/test_abi_exports.py::test_product_does_not_use_oracle
greps for that).  Parity unpinned at the PCL boundary -- see cwipc_oracle.h.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Tuple

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcwipc_oracle.so")
LIB_ALT = os.path.join(HERE, "libcwipc_oracle_alt.so")   # both PCL-release-dependent switches flipped (see cwipc_oracle.c)

POINT_DTYPE = numpy.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("tile", "u1")])

_lib: Optional[ctypes.CDLL] = None
_lib_alt: Optional[ctypes.CDLL] = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "cwipc_oracle.c")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(HERE, "cwipc_oracle.h")), os.path.getmtime(os.path.join(HERE, "Makefile")))
    if force or any(not os.path.exists(l) or os.path.getmtime(l) < newest for l in (LIB, LIB_ALT)):
        subprocess.run(["make", "-C", HERE, "-B", "all"], check=True, capture_output=True)
    return LIB


def _declare(lib: ctypes.CDLL) -> ctypes.CDLL:
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    lib.orc_config.argtypes = []
    lib.orc_config.restype = ctypes.c_int
    lib.orc_tilefilter.argtypes = [vp, sz, ctypes.c_int, vp]
    lib.orc_tilefilter.restype = ctypes.c_long
    lib.orc_min_distance_to_first.argtypes = [vp, sz]
    lib.orc_min_distance_to_first.restype = ctypes.c_float
    lib.orc_downsample.argtypes = [vp, sz, ctypes.c_float, ctypes.c_float, vp, ctypes.POINTER(ctypes.c_float), vp, vp]
    lib.orc_downsample.restype = ctypes.c_long
    lib.orc_knn_mean_distances.argtypes = [vp, sz, ctypes.c_int, vp]
    lib.orc_knn_mean_distances.restype = ctypes.c_int
    lib.orc_knn_mean_distances_bruteforce.argtypes = [vp, sz, ctypes.c_int, vp]
    lib.orc_knn_mean_distances_bruteforce.restype = ctypes.c_int
    lib.orc_remove_outliers.argtypes = [vp, sz, ctypes.c_int, ctypes.c_float, ctypes.c_int, vp, ctypes.POINTER(ctypes.c_double)]
    lib.orc_remove_outliers.restype = ctypes.c_long
    return lib


def load(alt: bool = False) -> ctypes.CDLL:
    """The default oracle, or (alt=True) the build with ORC_LEAF_ORDER_DESCENDING and ORC_VOXEL_SORT_UNSTABLE set."""
    global _lib, _lib_alt
    if alt:
        if _lib_alt is None:
            build()
            _lib_alt = _declare(ctypes.CDLL(LIB_ALT))
        return _lib_alt
    if _lib is None:
        build()
        _lib = _declare(ctypes.CDLL(LIB))
    return _lib


def _pts(a: numpy.ndarray) -> numpy.ndarray:
    a = numpy.ascontiguousarray(a, dtype=POINT_DTYPE)
    return a


def tilefilter(pts: numpy.ndarray, tile: int) -> numpy.ndarray:
    pts = _pts(pts)
    out = numpy.zeros(max(len(pts), 1), POINT_DTYPE)
    m = load().orc_tilefilter(pts.ctypes.data, len(pts), tile, out.ctypes.data)
    return out[:m].copy()


def min_distance_to_first(pts: numpy.ndarray) -> float:
    pts = _pts(pts)
    return float(load().orc_min_distance_to_first(pts.ctypes.data, len(pts)))


def downsample(pts: numpy.ndarray, voxelsize: float, pc_cellsize: float = 0.0, want_keys: bool = False, alt: bool = False):
    """Returns (out_points or None, cellsize, point_keys (n,6) or None, counts)."""
    pts = _pts(pts)
    n = len(pts)
    out = numpy.zeros(max(n, 1), POINT_DTYPE)
    counts = numpy.zeros(max(n, 1), numpy.uint32)
    keys = numpy.zeros((max(n, 1), 6), numpy.int32) if want_keys else None
    cs = ctypes.c_float(0)
    m = load(alt).orc_downsample(pts.ctypes.data, n, voxelsize, pc_cellsize, out.ctypes.data, ctypes.byref(cs), keys.ctypes.data if want_keys else None, counts.ctypes.data)
    if m < 0:
        return None, cs.value, (keys[:n] if want_keys else None), None
    return out[:m].copy(), cs.value, (keys[:n] if want_keys else None), counts[:m].copy()


def knn_mean_distances(pts: numpy.ndarray, k: int, bruteforce: bool = False) -> numpy.ndarray:
    pts = _pts(pts)
    d = numpy.zeros(len(pts), numpy.float32)
    fn = load().orc_knn_mean_distances_bruteforce if bruteforce else load().orc_knn_mean_distances
    if fn(pts.ctypes.data, len(pts), k, d.ctypes.data) != 0:
        raise ValueError("oracle kNN needs n > k >= 1")
    return d


def remove_outliers(pts: numpy.ndarray, k: int, stddev_mul: float, per_tile: bool) -> Tuple[numpy.ndarray, float]:
    pts = _pts(pts)
    n = len(pts)
    out = numpy.zeros(max(2 * n, 1), POINT_DTYPE)
    thr = ctypes.c_double(float("nan"))
    m = load().orc_remove_outliers(pts.ctypes.data, n, k, stddev_mul, 1 if per_tile else 0, out.ctypes.data, ctypes.byref(thr))
    return out[:m].copy(), thr.value
