"""Partitioned clouds on the GPU: the protocol of cwipc_util_b200/slab.py with the library's CUDA kernels as
local steps, against the single-GPU cwipc_downsample / cwipc_remove_outliers on the concatenated cloud.

Two ranks sharing GPU 0 (gloo, host staging) run on any box; the NCCL variant (one GPU per rank, device buffers
straight into ncclSend/ncclRecv) needs two GPUs and is skipped otherwise."""
import numpy
import pytest

from cwipc_util_b200 import synthetic

import _slab_runner as runner
from test_slab_cpu import make_parts, sorted_records

pytestmark = pytest.mark.gpu


def check_against_single_gpu(cw, parts, dsn, sorn, chainn, voxelsize, k, mul, cellsize):
    whole = numpy.concatenate(parts)
    pc = cw.cwipc_from_numpy_array(whole, 7)
    pc._set_cellsize(cellsize)
    # downsample: same records (bit-exact: integer sums do not depend on which rank added what)
    ds = cw.cwipc_downsample(pc, voxelsize)
    got = numpy.concatenate(dsn)
    assert len(got) == ds.count()
    assert numpy.array_equal(sorted_records(got), sorted_records(ds.get_numpy_array()))
    assert sum(len(p) > 0 for p in dsn) == len(dsn), "every slab owns voxels"
    # outlier removal: identical except for points within 1e-6 of the threshold (the all-reduced sums are
    # added in another order than on one GPU)
    d = cw.util.knn_mean_distances(pc, k).astype(numpy.float64)
    n = len(d)
    s, sq = d.sum(), (d.astype(numpy.float32) ** 2).astype(numpy.float64).sum()
    thr = s / n + mul * numpy.sqrt((sq - s * s / n) / (n - 1))
    got = numpy.concatenate(sorn)
    sure_keep, sure_drop = d <= thr * (1 - 1e-6), d > thr * (1 + 1e-6)
    gi = 0
    for i in range(n):
        if gi < len(got) and got[gi] == whole[i] and not sure_drop[i]:
            gi += 1
        else:
            assert not sure_keep[i], f"point {i} (d={d[i]}, thr={thr}) was dropped"
    assert gi == len(got)
    single = cw.cwipc_remove_outliers(pc, k, mul, False)
    assert abs(single.count() - len(got)) <= 2
    # chain: outlier removal of the partitioned downsample result
    chain = cw.cwipc_remove_outliers(ds, k, mul, False)
    assert abs(chain.count() - sum(len(p) for p in chainn)) <= 2
    a, b = sorted_records(numpy.concatenate(chainn)), sorted_records(chain.get_numpy_array())
    assert len(numpy.setdiff1d(a, b)) + len(numpy.setdiff1d(b, a)) <= 2


@pytest.mark.parametrize("world,voxelsize", [(2, 0.01), (3, -0.02)])
def test_slabs_sharing_one_gpu(cw, world, voxelsize, tmp_path):
    parts = make_parts(40000, world, seed=20 + world)
    args = dict(voxelsize=voxelsize, k=30, mul=1.0, cellsize=0.002)
    dsn, sorn, chainn = runner.launch(world, "cuda-shared", parts, str(tmp_path), port=29661 + world, **args)
    check_against_single_gpu(cw, parts, dsn, sorn, chainn, **args)


def test_slabs_unordered_parts_and_tiny_halo(cw, tmp_path):
    """Parts that are NOT slabs (interleaved points): ownership still puts every voxel on one rank, the halo
    exchange degenerates to all-to-all, and the open-query merge keeps the statistics exact."""
    pts = synthetic.camera_cloud(30000, seed=5, outliers=0.01)
    parts = [pts[0::2].copy(), pts[1::2].copy()]
    args = dict(voxelsize=0.015, k=16, mul=1.5, cellsize=0.0)
    dsn, sorn, chainn = runner.launch(2, "cuda-shared", parts, str(tmp_path), port=29671, halo=1e-3, **args)
    whole = numpy.concatenate(parts)
    pc = cw.cwipc_from_numpy_array(whole, 7)
    ds = cw.cwipc_downsample(pc, 0.015)
    assert numpy.array_equal(sorted_records(numpy.concatenate(dsn)), sorted_records(ds.get_numpy_array()))
    single = cw.cwipc_remove_outliers(pc, 16, 1.5, False)
    assert abs(single.count() - sum(len(p) for p in sorn)) <= 2


def test_slabs_with_an_empty_part(cw, tmp_path):
    parts = make_parts(20000, 2, seed=77)
    parts = [parts[0], parts[1][:0], parts[1]]  # rank 1 holds nothing
    args = dict(voxelsize=0.02, k=10, mul=1.0, cellsize=0.0)
    dsn, sorn, chainn = runner.launch(3, "cuda-shared", parts, str(tmp_path), port=29691, **args)
    whole = numpy.concatenate(parts)
    pc = cw.cwipc_from_numpy_array(whole, 7)
    ds = cw.cwipc_downsample(pc, 0.02)
    assert len(dsn[1]) == 0 and len(sorn[1]) == 0
    assert numpy.array_equal(sorted_records(numpy.concatenate(dsn)), sorted_records(ds.get_numpy_array()))
    single = cw.cwipc_remove_outliers(pc, 10, 1.0, False)
    assert abs(single.count() - sum(len(p) for p in sorn)) <= 2


def test_slabs_nccl_one_gpu_per_rank(cw, tmp_path):
    if cw.cuda_device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    world = min(cw.cuda_device_count(), 4)
    parts = make_parts(200000, world, seed=31)
    args = dict(voxelsize=0.005, k=30, mul=1.0, cellsize=0.002)
    dsn, sorn, chainn = runner.launch(world, "cuda-nccl", parts, str(tmp_path), port=29681, **args)
    check_against_single_gpu(cw, parts, dsn, sorn, chainn, **args)
