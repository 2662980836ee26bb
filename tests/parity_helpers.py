"""Comparisons shared by the GPU parity tests and __graft_entry__.smoke() (bars: BASELINE.json north_star)."""
import itertools

import numpy as np


def sor_group_check(grp, got, pos, d, mul):
    """`got[pos:]` must start with the inliers of `grp` as decided by the oracle's distances `d`, except that points
    within 1e-6 (relative) of the threshold may go either way.  Returns the position behind the group."""
    n = len(grp)
    dd = d.astype(np.float64)
    s, sq = dd.sum(), (d * d).astype(np.float64).sum()
    thr = s / n + mul * np.sqrt((sq - s * s / n) / (n - 1))
    keep = ~(dd > thr)
    ambiguous = np.flatnonzero(np.abs(dd - thr) <= 1e-6 * abs(thr))
    assert len(ambiguous) <= 12, f"{len(ambiguous)} points within 1e-6 of the threshold"
    for choice in itertools.product((False, True), repeat=len(ambiguous)):
        k2 = keep.copy()
        k2[ambiguous] = choice
        want = grp[k2]
        if pos + len(want) <= len(got) and np.array_equal(got[pos:pos + len(want)], want):
            return pos + len(want)
    raise AssertionError(f"group of {n} points at output position {pos}: kept points differ from the oracle's beyond threshold ties (thr={thr}, {len(ambiguous)} ties)")


def sor_set_check(grp, got, d, mul):
    """Order-free variant for partitioned results: `got` (any order) must contain every point of `grp` that is an
    inlier by more than 1e-6 of the threshold, and none that is an outlier by more than that."""
    n = len(grp)
    dd = d.astype(np.float64)
    s, sq = dd.sum(), (d * d).astype(np.float64).sum()
    thr = s / n + mul * np.sqrt((sq - s * s / n) / (n - 1))
    void = np.dtype((np.void, grp.dtype.itemsize))
    g = np.ascontiguousarray(grp).view(void).reshape(-1)
    o = np.ascontiguousarray(got).view(void).reshape(-1)
    present = np.isin(g, o)
    sure_keep, sure_drop = dd <= thr * (1 - 1e-6), dd > thr * (1 + 1e-6)
    assert present[sure_keep].all(), "an inlier was dropped"
    # a dropped point may still be `present` when an identical record survives elsewhere: count instead
    assert np.isin(o, g).all(), "output holds a point that is not in the input"
    assert len(got) <= int((~sure_drop).sum()) and len(got) >= int(sure_keep.sum())


def per_tile_check(orc, pts, got, k, mul):
    """ref: src/cwipc_filters.cpp:238-261 -- groups in first-appearance order of the tile value, tile 0 = the whole cloud"""
    _, first = np.unique(pts["tile"], return_index=True)
    tiles = pts["tile"][np.sort(first)]
    pos = 0
    for t in tiles:
        grp = pts if t == 0 else pts[pts["tile"] == t]
        if len(grp) <= k:
            assert np.array_equal(got[pos:pos + len(grp)], grp)
            pos += len(grp)
        else:
            pos = sor_group_check(grp, got, pos, orc.knn_mean_distances(grp, k), mul)
    assert pos == len(got)


def canonical_ranks(keys6):
    l = keys6[:, :3].astype(np.int64)
    m = np.zeros(len(l), np.int64)
    for b in range(21):
        m |= ((l[:, 0] >> b) & 1) << (3 * b + 2)
        m |= ((l[:, 1] >> b) & 1) << (3 * b + 1)
        m |= ((l[:, 2] >> b) & 1) << (3 * b)
    v = keys6[:, 3:].astype(np.int64)
    order = np.lexsort((v[:, 0], v[:, 1], v[:, 2], m))
    full = np.concatenate([m[:, None], v[:, ::-1]], axis=1)[order]
    heads = np.ones(len(full), bool)
    heads[1:] = np.any(full[1:] != full[:-1], axis=1)
    ranks = np.empty(len(full), np.int64)
    ranks[order] = np.cumsum(heads) - 1
    return ranks


def assert_points_close(got, want, cs):
    assert len(got) == len(want)
    for a in "xyz":
        scale = np.maximum(np.abs(want[a].astype(np.float64)), cs)
        err = np.abs(got[a].astype(np.float64) - want[a].astype(np.float64)) / scale
        assert err.max(initial=0.0) <= 1e-5, f"{a}: max rel err {err.max()}"
    for c in "rgb":
        assert np.abs(got[c].astype(np.int32) - want[c].astype(np.int32)).max(initial=0) <= 1
    assert np.array_equal(got["tile"], want["tile"])
