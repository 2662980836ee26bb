"""GPU parity on the BASELINE.json configurations themselves, at their full sizes, against the CPU oracle
(the C oracle needs seconds for these; ref: src/cwipc_filters.cpp:89-172, 222-278):

* configs[1]  1 000 000-point 4-camera cloud, cwipc_remove_outliers(30, 1.0, perTile=True), seeds 0/1/2;
* configs[3]  7 997 584-point cloud, cwipc_downsample at +0.002 / 0.005 / 0.01 / 0.02 / 0.05 (octree-split mode):
              voxel keys, counts, order and tile masks bit-exact, centroids 1e-5, colours +-1;
* configs[4]  the frames exactly as bench.py builds them: downsample(0.01) -> remove_outliers(30, 1.0, False).

Keep-masks are compared group by group: a point may differ from the oracle's decision only when its mean
neighbour distance lies within 1e-6 (relative) of the group's threshold.
"""
import os
import sys

import numpy as np
import pytest

from cwipc_util_b200 import synthetic
from parity_helpers import assert_exact_means, assert_points_close, canonical_ranks, per_tile_check, sor_group_check

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = synthetic.cwipc_point_numpy_dtype


def upload(cw, pts, timestamp=1, cellsize=None):
    pc = cw.cwipc_from_numpy_array(pts, timestamp)
    if cellsize is not None:
        pc._set_cellsize(cellsize)
    return pc


def download(pc):
    return pc.get_numpy_array().copy()


# ---- configs[1] ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_config2_remove_outliers_per_tile_1m(cw, orc, seed):
    n = 1000 * 1000
    pts = synthetic.camera_cloud(n, seed=seed)
    assert set(np.unique(pts["tile"]).tolist()) == {1, 2, 4, 8}
    out = cw.cwipc_remove_outliers(upload(cw, pts, timestamp=seed, cellsize=synthetic.cellsize_of(n)), 30, 1.0, True)
    got = download(out)
    assert 0 < len(got) < n and out.timestamp() == seed
    per_tile_check(orc, pts, got, 30, 1.0)


def test_per_tile_with_tile_zero_and_small_groups(cw, orc):
    pts = synthetic.camera_cloud(80000, seed=13)
    pts["tile"][5] = 0            # the whole cloud is processed again as the group of tile 0 (SURVEY.md finding 3)
    pts["tile"][100:110] = 77     # a group with fewer points than neighbours: kept as it is
    got = download(cw.cwipc_remove_outliers(upload(cw, pts), 30, 1.0, True))
    assert len(got) > len(pts)
    per_tile_check(orc, pts, got, 30, 1.0)


def test_per_tile_many_tiles_and_small_groups_in_one_pass(cw, orc):
    """No tile 0: every group goes through ONE search index (tile rank = a band of cells).  37 tile values of very
    different sizes, two groups with fewer points than neighbours, groups interleaved point by point."""
    pts = synthetic.camera_cloud(120000, seed=21)
    idx = np.arange(len(pts))
    pts["tile"] = 1 + (idx * 7919 % 37)          # 37 interleaved groups ...
    pts["tile"][idx % 5 == 0] = 200              # ... one much larger than the others
    pts["tile"][100:110] = 77                    # fewer points than neighbours: kept as they are
    pts["tile"][5000] = 78
    got = download(cw.cwipc_remove_outliers(upload(cw, pts), 30, 1.0, True))
    assert 0 < len(got) < len(pts)
    per_tile_check(orc, pts, got, 30, 1.0)
    # without a cellsize hint the pitch comes from the bounding box; other k / multiplier
    got = download(cw.cwipc_remove_outliers(upload(cw, pts, cellsize=0.0), 8, 2.0, True))
    per_tile_check(orc, pts, got, 8, 2.0)


def test_per_tile_one_pass_equals_group_after_group(cw):
    """CWIPC_CUDA_SOR_PER_TILE=sequential (tests only) runs the groups one after the other, each with an index of its
    own: same bytes."""
    import subprocess
    pts = synthetic.camera_cloud(300000, seed=22)
    got = download(cw.cwipc_remove_outliers(upload(cw, pts, cellsize=synthetic.cellsize_of(len(pts))), 30, 1.0, True))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from cwipc_util_b200 import synthetic, util as cw\n"
            "pts = synthetic.camera_cloud(300000, seed=22)\n"
            "pc = cw.cwipc_from_numpy_array(pts, 1); pc._set_cellsize(synthetic.cellsize_of(len(pts)))\n"
            "out = cw.cwipc_remove_outliers(pc, 30, 1.0, True).get_numpy_array()\n"
            "sys.stdout.buffer.write(out.tobytes())" % REPO)
    env = dict(os.environ, CWIPC_CUDA_SOR_PER_TILE="sequential")
    raw = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, check=True, timeout=600).stdout
    want = np.frombuffer(raw, DT)
    assert np.array_equal(got, want)


# ---- configs[3] ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cloud8m():
    n = 2828 * 2828
    return synthetic.camera_cloud(n, seed=4, noise=0.0005)


@pytest.mark.parametrize("voxel", [0.002, 0.005, 0.01, 0.02, 0.05])
def test_config4_downsample_8m_octree_mode(cw, orc, cloud8m, voxel):
    pts = cloud8m
    cellsize = synthetic.cellsize_of(len(pts))
    want, cs, keys6, counts = orc.downsample(pts, voxel, cellsize, want_keys=True)
    pc = upload(cw, pts, timestamp=4, cellsize=cellsize)
    out = cw.cwipc_downsample(pc, voxel)
    assert out.cellsize() == np.float32(cs) == np.float32(voxel)
    got = download(out)
    assert len(got) == len(want)
    assert_points_close(got, want, cs, counts)
    gkeys = cw.util.downsample_keys(pc, voxel)
    _, inverse = np.unique(gkeys, return_inverse=True)
    ranks = canonical_ranks(keys6)
    assert np.array_equal(inverse, ranks)
    assert np.array_equal(np.bincount(inverse, minlength=len(want)), counts)
    assert_exact_means(got, pts, ranks, counts, cs)


def test_config4_chain_8m(cw, orc, cloud8m):
    """downsample(0.005) -> remove_outliers(30, 1.0) on the 8 M-point cloud; the outlier stage is checked on the
    GPU's own centroids."""
    pts = cloud8m
    pc = upload(cw, pts, cellsize=synthetic.cellsize_of(len(pts)))
    d = cw.cwipc_downsample(pc, 0.005)
    o = cw.cwipc_remove_outliers(d, 30, 1.0, False)
    gd, go = download(d), download(o)
    assert 0 < len(go) < len(gd)
    assert sor_group_check(gd, go, 0, orc.knn_mean_distances(gd, 30), 1.0) == len(go)


# ---- configs[4]: the bench's own frames ---------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_config5_bench_frames(cw, orc, seed):
    sys.path.insert(0, REPO)
    import bench
    frame = bench.make_frames(seed, 1, 1)[0]
    assert len(frame) == bench.POINTS_PER_FRAME == 1000 * 1000
    cellsize = synthetic.cellsize_of(len(frame))
    pc = upload(cw, frame, timestamp=seed, cellsize=cellsize)
    d = cw.cwipc_downsample(pc, bench.VOXEL)
    o = cw.cwipc_remove_outliers(d, bench.K, bench.STDDEV, False)
    want, cs, keys6, counts = orc.downsample(frame, bench.VOXEL, cellsize, want_keys=True)
    gd = download(d)
    assert_points_close(gd, want, cs, counts)
    _, inverse = np.unique(cw.util.downsample_keys(pc, bench.VOXEL), return_inverse=True)
    ranks = canonical_ranks(keys6)
    assert np.array_equal(inverse, ranks)
    assert_exact_means(gd, frame, ranks, counts, cs)
    go = download(o)
    assert sor_group_check(gd, go, 0, orc.knn_mean_distances(gd, bench.K), bench.STDDEV) == len(go)
    assert o.cellsize() == np.float32(cs) and o.timestamp() == seed
