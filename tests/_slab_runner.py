"""Helper of the partitioned-cloud tests: one process per rank (torch.multiprocessing.spawn), each runs the
protocol of cwipc_util_b200/slab.py on its part and saves what it ends up with.

backend "numpy": the local steps restated with numpy/scipy (test infrastructure; lets the PROTOCOL -- replay
pipeline, column ownership, boundary exchange, halo, open-query merge, all-reduced statistics -- run on CPU
over gloo).  backend "cuda": the library's C ABI on a GPU (gloo + host staging when the ranks share one
GPU, NCCL device-to-device when every rank has its own).
"""
import math
import os
import sys

import numpy

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

DT = numpy.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("tile", "u1")])


class NpCloud:
    def __init__(self, pts, cellsize=0.0):
        self.pts = numpy.ascontiguousarray(pts)
        self.cellsize = float(cellsize)


def _d2(q, c):
    """float32 ((dx*dx + dy*dy) + dz*dz), as the reference's L2_Simple<float>"""
    dx = (q["x"][:, None] - c["x"][None, :]).astype(numpy.float32)
    dy = (q["y"][:, None] - c["y"][None, :]).astype(numpy.float32)
    dz = (q["z"][:, None] - c["z"][None, :]).astype(numpy.float32)
    return ((dx * dx + dy * dy).astype(numpy.float32) + dz * dz).astype(numpy.float32)


def _lists(queries, cloud, k):
    from scipy.spatial import cKDTree
    kk = k + 1
    out = numpy.full((len(queries), kk), numpy.inf, numpy.float32)
    if len(cloud) == 0 or len(queries) == 0:
        return out
    xyz = numpy.stack([cloud["x"], cloud["y"], cloud["z"]], 1).astype(numpy.float64)
    qxyz = numpy.stack([queries["x"], queries["y"], queries["z"]], 1).astype(numpy.float64)
    take = min(len(cloud), kk + 8)  # a few spares: float32 ties may reorder the tail
    _, idx = cKDTree(xyz).query(qxyz, k=take)
    idx = idx.reshape(len(queries), take)
    for i in range(len(queries)):
        d = numpy.sort(_d2(queries[i:i + 1], cloud[idx[i]])[0])[:kk]
        out[i, :len(d)] = d
    return out


def _mean_from_lists(l, k):
    s = numpy.zeros(len(l), numpy.float64)
    for j in range(1, k + 1):
        s = s + numpy.sqrt(l[:, j].astype(numpy.float64))
    return (s / k).astype(numpy.float32)


class NumpyOps:
    def count(self, pc):
        return len(pc.pts)

    def cellsize(self, pc):
        return pc.cellsize

    def replay(self, pc, cellsize, state):
        st = numpy.array(state, numpy.float64)
        if len(pc.pts) == 0:
            return st, numpy.array([numpy.inf] * 3 + [-numpy.inf] * 3)
        st[7] = 1.0
        p = pc.pts
        return st, numpy.array([p["x"].min(), p["y"].min(), p["z"].min(), p["x"].max(), p["y"].max(), p["z"].max()], numpy.float64)

    def crop_x(self, pc, lo, hi):
        m = (pc.pts["x"] >= numpy.float32(lo)) & (pc.pts["x"] < numpy.float32(hi))
        return NpCloud(pc.pts[m], pc.cellsize)

    def join(self, pcs):
        return NpCloud(numpy.concatenate([p.pts for p in pcs]), min(p.cellsize for p in pcs))

    def downsample_planned(self, pc, voxelsize, state, bounds):
        assert voxelsize < 0, "the numpy stand-in restates the single-grid mode only"
        cs = numpy.float32(max(-voxelsize, pc.cellsize))
        inv = numpy.float32(1.0) / cs
        p = pc.pts
        if len(p) == 0:
            return NpCloud(p, float(cs))
        gmin = numpy.asarray(bounds[:3], numpy.float32)
        minb = numpy.floor(gmin * inv).astype(numpy.int64)
        ijk = numpy.stack([numpy.floor(p[a] * inv).astype(numpy.int64) - minb[i] for i, a in enumerate("xyz")], 1)
        key = ijk[:, 0] + (ijk[:, 1] << 21) + (ijk[:, 2] << 42)
        order = numpy.argsort(key, kind="stable")
        uk, start, cnt = numpy.unique(key[order], return_index=True, return_counts=True)
        out = numpy.zeros(len(uk), DT)
        for a in "xyz":
            out[a] = (numpy.add.reduceat(p[a][order].astype(numpy.float64), start) / cnt).astype(numpy.float32)
        for c in "rgb":
            out[c] = (numpy.add.reduceat(p[c][order].astype(numpy.int64), start) // cnt).astype(numpy.uint8)
        out["tile"] = numpy.bitwise_or.reduceat(p["tile"][order], start)
        return NpCloud(out, float(cs))

    def knn_open(self, pc, k, nquery, x_lo, x_hi):
        q = pc.pts[:nquery]
        if len(pc.pts) > k:
            l = _lists(q, pc.pts, k)
            mean, kth2 = _mean_from_lists(l, k), l[:, k]
        else:
            mean, kth2 = numpy.zeros(nquery, numpy.float32), numpy.full(nquery, numpy.inf, numpy.float32)
        x = q["x"].astype(numpy.float64)
        rk = numpy.sqrt(kth2.astype(numpy.float64)) * (1.0 + 1e-6)
        open_idx = numpy.nonzero(~((x - rk > x_lo) & (x + rk < x_hi)))[0].astype(numpy.uint32)
        return {"mean": mean.copy(), "idx": open_idx}, open_idx, q[open_idx], numpy.minimum(kth2[open_idx], numpy.float32(3.0e38))

    def keep_all(self, pc):
        return pc

    def patch(self, d, values):
        d["mean"][d["idx"]] = values

    def knn_lists(self, pc, queries, k, limits):
        l = _lists(queries, pc.pts, k)
        l[l > limits[:, None]] = numpy.inf   # only distances within the query's current bound are reported
        return l

    def merge_lists(self, lists, k):
        merged = numpy.sort(numpy.concatenate(list(lists), axis=1), axis=1)[:, :k + 1]
        return _mean_from_lists(merged, k)

    def distance_stats(self, dist):
        d = dist["mean"].astype(numpy.float32)
        return float(d.astype(numpy.float64).sum()), float((d * d).astype(numpy.float64).sum())

    def threshold(self, total, sq, n, mul):
        return total / n + mul * math.sqrt((sq - total * total / n) / (n - 1.0))

    def filter_by_distance(self, pc, dist, thr):
        return NpCloud(pc.pts[~(dist["mean"].astype(numpy.float64) > thr)], pc.cellsize)

    def to_wire(self, pc):
        import torch
        return torch.from_numpy(pc.pts.copy().view(numpy.uint8).reshape(-1)) if len(pc.pts) else None

    def from_wire(self, t, timestamp, cellsize):
        return NpCloud(t.numpy().view(DT).copy(), cellsize)

    def release_wire(self):
        pass


def run_rank_lib(rank, world, port, indir, outdir, voxelsize, k, mul, cellsize, halo):
    """The protocol INSIDE libcwipc_util_cuda (csrc/slab.cpp: NCCL through dlopen, one GPU per rank).  torch.distributed
    (gloo) only carries the 128-byte ncclUniqueId from rank 0 to the others."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import cwipc_util_b200 as cw
    from cwipc_util_b200 import util
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cw.cuda_set_device(rank)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(util.cuda_comm.unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(uid, src=0)
    comm = util.cuda_comm(bytes(uid.numpy().tobytes()), world, rank)
    part = numpy.load(os.path.join(indir, f"part{rank}.npy"))
    pc = cw.cwipc_from_numpy_array(part, 7)
    pc._set_cellsize(cellsize)
    ds = comm.downsample(pc, voxelsize)
    numpy.save(os.path.join(outdir, f"ds{rank}.npy"), ds.get_numpy_array().copy())
    kept = comm.remove_outliers(pc, k, mul, False, halo or 0.0)
    numpy.save(os.path.join(outdir, f"sor{rank}.npy"), kept.get_numpy_array().copy())
    chain = comm.remove_outliers(ds, k, mul, False, 0.0)
    numpy.save(os.path.join(outdir, f"chain{rank}.npy"), chain.get_numpy_array().copy())
    pertile = comm.remove_outliers(pc, k, mul, True, halo or 0.0)
    numpy.save(os.path.join(outdir, f"pertile{rank}.npy"), pertile.get_numpy_array().copy())
    tf, off, tot = comm.tilefilter(pc, 2)
    numpy.save(os.path.join(outdir, f"tf{rank}.npy"), tf.get_numpy_array().copy())
    numpy.save(os.path.join(outdir, f"tfmeta{rank}.npy"), numpy.array([off, tot], numpy.int64))
    comm.free()
    dist.barrier()
    dist.destroy_process_group()


def run_rank(rank, world, port, backend, indir, outdir, voxelsize, k, mul, cellsize, halo):
    if backend == "lib":
        return run_rank_lib(rank, world, port, indir, outdir, voxelsize, k, mul, cellsize, halo)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from cwipc_util_b200 import slab
    part = numpy.load(os.path.join(indir, f"part{rank}.npy"))
    if backend == "numpy":
        dist.init_process_group("gloo", rank=rank, world_size=world)
        comm, ops = slab.TorchComm("cpu"), NumpyOps()
        pc = NpCloud(part, cellsize)
        as_np = lambda c: c.pts  # noqa: E731
    else:
        import torch
        import cwipc_util_b200 as cw
        ndev = cw.cuda_device_count()
        own_gpu = backend == "cuda-nccl"
        dev = rank if own_gpu else 0
        assert dev < ndev
        if own_gpu:
            torch.cuda.set_device(dev)
            dist.init_process_group("nccl", rank=rank, world_size=world)
            comm = slab.TorchComm(f"cuda:{dev}")
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
            comm = slab.TorchComm("cpu")
        ops = slab.CudaOps(dev, host_wire=not own_gpu)
        pc = cw.cwipc_from_numpy_array(part, 7)
        pc._set_cellsize(cellsize)
        as_np = lambda c: c.get_numpy_array().copy()  # noqa: E731
    ds = slab.slab_downsample(pc, voxelsize, comm, ops)
    numpy.save(os.path.join(outdir, f"ds{rank}.npy"), as_np(ds))
    kept = slab.slab_remove_outliers(pc, k, mul, comm, ops, halo=halo)
    numpy.save(os.path.join(outdir, f"sor{rank}.npy"), as_np(kept))
    # the chain of BASELINE configs[3]: outlier removal of the partitioned downsample result
    chain = slab.slab_remove_outliers(ds, k, mul, comm, ops)
    numpy.save(os.path.join(outdir, f"chain{rank}.npy"), as_np(chain))
    dist.barrier()
    dist.destroy_process_group()


def launch(world, backend, parts, tmpdir, voxelsize, k, mul, cellsize, halo=None, port=29631):
    """Run the protocol on `parts` (one array per rank); returns (downsample parts, outlier parts, chain parts)."""
    import torch.multiprocessing as mp
    indir, outdir = os.path.join(tmpdir, "in"), os.path.join(tmpdir, "out")
    os.makedirs(indir, exist_ok=True)
    os.makedirs(outdir, exist_ok=True)
    for r, p in enumerate(parts):
        numpy.save(os.path.join(indir, f"part{r}.npy"), p)
    mp.spawn(run_rank, args=(world, port, backend, indir, outdir, voxelsize, k, mul, cellsize, halo), nprocs=world, join=True)
    load = lambda name: [numpy.load(os.path.join(outdir, f"{name}{r}.npy")) for r in range(world)]  # noqa: E731
    if backend == "lib":
        return load("ds"), load("sor"), load("chain"), load("pertile"), load("tf"), load("tfmeta")
    return load("ds"), load("sor"), load("chain")
