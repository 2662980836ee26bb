"""The x-slab protocol (cwipc_util_b200/slab.py) over gloo with world sizes 2 and 3, local steps restated in numpy:
the partitioned result must equal the same steps run on the whole cloud by one rank."""
import numpy
import pytest

from cwipc_util_b200 import slab, synthetic

import _slab_runner as runner


def make_parts(n, world, seed, shuffle_inside=True):
    pts = synthetic.camera_cloud(n, seed=seed, outliers=0.01)
    order = numpy.argsort(pts["x"], kind="stable")
    pts = pts[order]
    cuts = [0] + [int(len(pts) * (i + 1) / world) for i in range(world)]
    rng = numpy.random.default_rng(seed)
    parts = []
    for r in range(world):
        p = pts[cuts[r]:cuts[r + 1]].copy()
        if shuffle_inside:
            rng.shuffle(p)
        parts.append(p)
    return parts


def sorted_records(a):
    return numpy.sort(numpy.ascontiguousarray(a).view("V16").reshape(-1))


def test_column_threshold_is_exact():
    inv = numpy.float32(1.0) / numpy.float32(0.013)
    for v in (-57.0, -1.0, 0.0, 1.0, 3.0, 1234.0):
        t = numpy.float32(slab.column_threshold(v, inv))
        assert numpy.floor(t * inv) >= v
        assert numpy.floor(numpy.nextafter(t, numpy.float32(-numpy.inf)) * inv) < v
    assert slab.column_threshold(float("inf"), inv) == float("inf")


@pytest.mark.parametrize("world", [2, 3])
def test_protocol_matches_single_rank(world, tmp_path):
    parts = make_parts(6000, world, seed=world)
    whole = numpy.concatenate(parts)
    args = dict(voxelsize=-0.02, k=12, mul=1.0, cellsize=0.004)
    ds1, sor1, chain1 = runner.launch(1, "numpy", [whole], str(tmp_path / "one"), port=29641, **args)
    dsn, sorn, chainn = runner.launch(world, "numpy", parts, str(tmp_path / "many"), port=29643 + world, **args)
    # downsample: every voxel reduced by exactly one rank, same records overall
    assert numpy.array_equal(sorted_records(numpy.concatenate(dsn)), sorted_records(ds1[0]))
    assert all(len(p) > 0 for p in dsn)
    # outlier removal keeps order inside a part, and the parts are the whole cloud in rank order
    assert numpy.array_equal(numpy.concatenate(sorn), sor1[0])
    assert 0 < len(sor1[0]) < len(whole)
    assert numpy.array_equal(sorted_records(numpy.concatenate(chainn)), sorted_records(chain1[0]))


def test_protocol_with_an_empty_part_and_a_tiny_halo(tmp_path):
    parts = make_parts(3000, 2, seed=9)
    parts = [parts[0], parts[1][:0], parts[1]]  # rank 1 holds nothing
    whole = numpy.concatenate(parts)
    args = dict(voxelsize=-0.03, k=8, mul=1.5, cellsize=0.0)
    ds1, sor1, _ = runner.launch(1, "numpy", [whole], str(tmp_path / "one"), port=29651, **args)
    dsn, sorn, _ = runner.launch(3, "numpy", parts, str(tmp_path / "many"), port=29653, halo=1e-4, **args)  # nearly every query goes through the merge
    assert numpy.array_equal(sorted_records(numpy.concatenate(dsn)), sorted_records(ds1[0]))
    assert numpy.array_equal(numpy.concatenate(sorn), sor1[0])
