"""Synthetic workloads of BASELINE.json, as numpy record arrays of cwipc_point.

* synthetic_cloud(): the surface of revolution of the reference's fake camera, restated from
  src/cwipc_synthetic.cpp:182-222 (sqrt(N) x sqrt(N) grid over height x angle, tiles 1/2 by the
  sign of z) with the wall-clock colour phase replaced by a fixed `angle`.
* simulate_cameras(): hard tile assignment of python/cwipc/filters/simulatecams.py:23-28,45-59,70
  (camera c looks along (cos 2*pi*c/n, 0, sin 2*pi*c/n); tile = 1 << argmax dot).
* add_noise(): python/cwipc/filters/noise.py:44-50 (random direction, length distance*U(0,1)).
* add_outliers(): uniform points in the bounding box (our addition so that the outlier filter has
  something to remove).
"""
from __future__ import annotations

import numpy

from .util import cwipc_point_numpy_dtype


def synthetic_cloud(npoints: int = 160000, angle: float = 0.0) -> numpy.ndarray:
    side = int(numpy.sqrt(npoints))
    pi = numpy.float32(3.14159265358979)
    hi = numpy.arange(side, dtype=numpy.float32)
    h = hi * (numpy.float32(2.0) / numpy.float32(side))
    a = hi * (numpy.float32(2) * pi / numpy.float32(side))
    hh = (h * pi / numpy.float32(3) - pi / numpy.float32(6)).astype(numpy.float64)
    radius = (0.3 * numpy.power(numpy.cos(hh), 0.71)).astype(numpy.float32)
    H, A = numpy.meshgrid(h, a, indexing="ij")
    R = numpy.broadcast_to(radius[:, None], H.shape)
    x = (R.astype(numpy.float64) * numpy.sin(A.astype(numpy.float64))).astype(numpy.float32)
    z = (R.astype(numpy.float64) * numpy.cos(A.astype(numpy.float64))).astype(numpy.float32)
    pts = numpy.zeros(side * side, cwipc_point_numpy_dtype)
    pts["x"] = (-x).ravel()
    pts["y"] = H.ravel()
    pts["z"] = z.ravel()
    ang = numpy.float32(angle)
    for k, name in enumerate(("r", "g", "b")):
        phase = (numpy.float32(k + 2) * pi * H + ang + A).astype(numpy.float64)
        v = ((1 + numpy.sin(phase)) / 2).astype(numpy.float32)
        pts[name] = (v.astype(numpy.float64) * 255.0).astype(numpy.int32).astype(numpy.uint8).ravel()
    pts["tile"] = numpy.where(z.ravel() < 0, 1, 2).astype(numpy.uint8)
    return pts


def simulate_cameras(pts: numpy.ndarray, ncamera: int = 4) -> numpy.ndarray:
    out = pts.copy()
    ang = 2 * numpy.pi * numpy.arange(ncamera) / ncamera
    cams = numpy.stack([numpy.cos(ang), numpy.sin(ang)], axis=1)  # (x, z) components
    xz = numpy.stack([pts["x"].astype(numpy.float64), pts["z"].astype(numpy.float64)], axis=1)
    xz -= xz.mean(axis=0)
    best = numpy.argmax(xz @ cams.T, axis=1)
    out["tile"] = (1 << best).astype(numpy.uint8)
    return out


def add_noise(pts: numpy.ndarray, distance: float = 0.002, seed: int = 0) -> numpy.ndarray:
    rng = numpy.random.default_rng(seed)
    n = pts.shape[0]
    vec = rng.uniform(-1, 1, (n, 3))
    unif = rng.uniform(0, 1, n)
    norm = numpy.linalg.norm(vec, axis=1)
    norm[norm == 0] = 1.0
    vec = vec / (norm / unif)[:, None] * distance
    out = pts.copy()
    out["x"] = (pts["x"] + vec[:, 0]).astype(numpy.float32)
    out["y"] = (pts["y"] + vec[:, 1]).astype(numpy.float32)
    out["z"] = (pts["z"] + vec[:, 2]).astype(numpy.float32)
    return out


def add_outliers(pts: numpy.ndarray, fraction: float = 0.005, seed: int = 0) -> numpy.ndarray:
    """Overwrite a random `fraction` of the points with uniform samples of the bounding box."""
    rng = numpy.random.default_rng(seed + 7919)
    n = pts.shape[0]
    m = int(n * fraction)
    if m == 0:
        return pts.copy()
    out = pts.copy()
    which = rng.choice(n, m, replace=False)
    for ax in ("x", "y", "z"):
        lo, hi = float(pts[ax].min()), float(pts[ax].max())
        out[ax][which] = rng.uniform(lo, hi, m).astype(numpy.float32)
    return out


def camera_cloud(npoints: int, seed: int = 0, ncamera: int = 4, noise: float = 0.002, outliers: float = 0.005, angle: float = 0.0) -> numpy.ndarray:
    """BASELINE config 2/3/5 input: synthetic cloud, 4 simulated cameras (tile bits 1,2,4,8), jitter, sparse outliers."""
    pts = synthetic_cloud(npoints, angle)
    pts = simulate_cameras(pts, ncamera)
    if noise > 0:
        pts = add_noise(pts, noise, seed)
    if outliers > 0:
        pts = add_outliers(pts, outliers, seed)
    return pts


def cellsize_of(npoints: int) -> float:
    """cellsize metadata the synthetic source records: 2 / sqrt(N) (ref: src/cwipc_synthetic.cpp:131)."""
    return float(numpy.float32(2.0 / int(numpy.sqrt(npoints))))
