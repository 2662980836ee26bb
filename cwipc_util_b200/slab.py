"""slab.py -- one large cloud spread over several GPUs as x-slabs (BASELINE.json configs[3]).

The library calls are local to one GPU (include/cwipc_util_cuda.h, "partitioned clouds"); this module is
the protocol between the ranks, one process per GPU, `torch.distributed` as plumbing (NCCL over
NVLink on the GPUs: point exchanges are ncclSend/ncclRecv between device buffers; gloo in the CPU tests).

Semantics: the WHOLE cloud is the concatenation of the parts in rank order, and the results are those of
cwipc_downsample / cwipc_remove_outliers (ref: src/cwipc_filters.cpp:89-172, 181-278) on that cloud, left
partitioned (rank r holds the part of the result that belongs to its slab, in the reference's order).

slab_downsample
  1. octree box: PCL grows the octree's bounding box while points are inserted IN ORDER, so the box state
     travels rank 0 -> 1 -> ... (one small send/recv per hop) and the final state is broadcast.
  2. a voxel must be reduced by one rank: voxel columns (floorf(x / cellsize)) are assigned to ranks from the
     parts' own x minima, and every point that sits in a column owned by another rank is sent there
     (only the points of boundary voxels move when the parts are proper x-slabs).
  3. every rank runs the planned downsample on what it now holds.

slab_remove_outliers (whole cloud, perTile=False)
  1. halo: every rank receives the points within H of its x-extent from the other ranks.
  2. kNN statistics of the local points against local+halo; a query is final when its (k+1)-th distance
     stays inside the covered interval.
  3. the few queries that are not final (isolated points) are all-gathered with their current (k+1)-th distance;
     every rank whose x-extent that distance reaches answers with the k+1 smallest distances among its OWN points,
     the owner merges the lists: exact whatever H was.
  4. sum d, sum d*d, n are all-reduced; every rank thresholds its own points.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy


# ======================================================================================================
# communication: a thin adapter over torch.distributed
# ======================================================================================================
class TorchComm:
    """rank/size + the handful of collectives the protocol needs.  `device` is "cpu" (gloo) or "cuda:N" (nccl)."""

    def __init__(self, device: str = "cpu", group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.device = torch.device(device)
        self.rank = dist.get_rank(group)
        self.size = dist.get_world_size(group)

    # ---- small host arrays -----------------------------------------------------------------------------
    def _t(self, a: numpy.ndarray):
        # a private copy: collectives work in place and torch.from_numpy would alias the caller's array
        return self.torch.from_numpy(numpy.array(a, copy=True, order="C")).to(self.device)

    def allreduce(self, a: numpy.ndarray, op: str) -> numpy.ndarray:
        t = self._t(numpy.asarray(a, numpy.float64))
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op), group=self.group)
        return t.cpu().numpy()

    def gather_rows(self, row) -> numpy.ndarray:
        """Every rank contributes one row of floats; all get the size x len(row) matrix (one collective)."""
        row = numpy.asarray(row, numpy.float64)
        m = numpy.zeros((self.size, len(row)))
        m[self.rank] = row
        return self.allreduce(m.reshape(-1), "SUM").reshape(self.size, len(row))

    def broadcast(self, a: numpy.ndarray, src: int) -> numpy.ndarray:
        t = self._t(numpy.asarray(a, numpy.float64))
        self.dist.broadcast(t, src=src, group=self.group)
        return t.cpu().numpy()

    def send(self, a: numpy.ndarray, dst: int) -> None:
        self.dist.send(self._t(numpy.asarray(a, numpy.float64)), dst=dst, group=self.group)

    def recv(self, n: int, src: int) -> numpy.ndarray:
        t = self.torch.zeros(n, dtype=self.torch.float64, device=self.device)
        self.dist.recv(t, src=src, group=self.group)
        return t.cpu().numpy()

    def allgather_bytes(self, a: numpy.ndarray) -> List[numpy.ndarray]:
        """All-gather arrays of different lengths (same dtype and trailing shape): one array per rank."""
        a = numpy.ascontiguousarray(a)
        raw = a.view(numpy.uint8).reshape(-1)
        sizes = self.allreduce(numpy.eye(self.size)[self.rank] * raw.size, "SUM").astype(numpy.int64)
        width = int(sizes.max())
        mine = self.torch.zeros(max(width, 1), dtype=self.torch.uint8, device=self.device)
        if raw.size:
            mine[:raw.size] = self._t(raw)
        parts = [self.torch.zeros_like(mine) for _ in range(self.size)]
        self.dist.all_gather(parts, mine, group=self.group)
        out = []
        for q in range(self.size):
            b = parts[q][:int(sizes[q])].cpu().numpy()
            out.append(b.view(a.dtype).reshape((-1,) + a.shape[1:]))
        return out

    # ---- point exchange: every rank sends one (possibly empty) byte buffer to every other rank ---------
    def exchange(self, outgoing: Sequence[Optional[object]], counts: Sequence[int]) -> List[Optional[object]]:
        """outgoing[q]: uint8 tensor on self.device for rank q (None / empty allowed); counts[q] its length in bytes.
        Returns the tensors received from every rank (None where nothing came).  Grouped point-to-point: on NCCL
        this is one ncclGroupStart/End of ncclSend/ncclRecv pairs, device to device over NVLink."""
        torch, dist = self.torch, self.dist
        table = numpy.zeros((self.size, self.size))
        table[self.rank, :] = counts
        table = self.allreduce(table.reshape(-1), "SUM").reshape(self.size, self.size).astype(numpy.int64)
        ops, incoming = [], [None] * self.size
        for q in range(self.size):
            if q == self.rank:
                continue
            if table[self.rank, q] > 0:
                ops.append(dist.P2POp(dist.isend, outgoing[q], q, group=self.group))
            if table[q, self.rank] > 0:
                incoming[q] = torch.empty(int(table[q, self.rank]), dtype=torch.uint8, device=self.device)
                ops.append(dist.P2POp(dist.irecv, incoming[q], q, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
        return incoming


# ======================================================================================================
# local operations on the GPU: the C ABI of libcwipc_util_cuda
# ======================================================================================================
class _DeviceBytes:
    """Lets torch alias device memory owned by the library (zero copy) through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class CudaOps:
    """The local steps, on this rank's GPU.  Clouds are cwipc_pointcloud_wrapper objects (device resident)."""

    def __init__(self, device_index: int, host_wire: bool = False):
        """host_wire: stage exchanged points through host memory (gloo transport, e.g. several ranks sharing one GPU
        in the tests); otherwise the library's device buffers go straight into NCCL."""
        from . import util
        self.u = util
        self.device_index = device_index
        self.host_wire = host_wire
        util.cuda_set_device(device_index)
        self._keepalive = []

    def count(self, pc) -> int:
        return pc.count()

    def cellsize(self, pc) -> float:
        return pc.cellsize()

    def replay(self, pc, cellsize: float, state: numpy.ndarray):
        st = self.u.cwipc_cuda_octree_state.from_array(state)
        bounds = self.u.octree_replay(pc, cellsize, st)
        return st.to_array(), bounds.astype(numpy.float64)

    def crop_x(self, pc, lo: float, hi: float):
        inf = float("inf")
        return self.u.cwipc_crop(pc, [lo, hi, -inf, inf, -inf, inf])

    def join(self, pcs):
        return self.u.cwipc_join_multi(pcs)

    def downsample_planned(self, pc, voxelsize: float, state: numpy.ndarray, bounds: numpy.ndarray):
        return self.u.downsample_planned(pc, voxelsize, self.u.cwipc_cuda_octree_state.from_array(state), bounds)

    def knn_open(self, pc, k: int, nquery: int, x_lo: float, x_hi: float):
        """Distances of the first nquery points of `pc` (kept on the device) + indices and coordinates of the open queries."""
        d = self.u.cuda_distances(pc, k, nquery, x_lo, x_hi)
        idx, pts, kth2 = d.open_queries()
        return d, idx, pts, kth2

    def keep_all(self, pc):
        return self.u.cwipc_tilefilter(pc, 0)

    def knn_lists(self, pc, queries: numpy.ndarray, k: int, limits: numpy.ndarray) -> numpy.ndarray:
        return self.u.knn_lists(pc, queries, k, limits)

    def merge_lists(self, lists: numpy.ndarray, k: int) -> numpy.ndarray:
        return self.u.knn_merge_lists(lists, k)[0]

    def patch(self, d, values: numpy.ndarray) -> None:
        d.patch(values)

    def distance_stats(self, d):
        return d.stats()

    def threshold(self, total: float, sq: float, n: float, mul: float) -> float:
        return self.u.outlier_threshold(total, sq, n, mul)

    def filter_by_distance(self, pc, d, thr: float):
        out = d.filter(pc, thr)
        d.free()
        return out

    # ---- wire: device buffers straight into NCCL ---------------------------------------------------------
    def to_wire(self, pc):
        """uint8 tensor aliasing the cloud's device memory (valid while `pc` is alive)."""
        import torch
        n = pc.count()
        if n == 0:
            return None
        if self.host_wire:
            return torch.from_numpy(pc.get_numpy_array().copy().view(numpy.uint8).reshape(-1))
        self.u.cuda_synchronize()  # the library's stream produced the points; NCCL runs on torch's
        self._keepalive.append(pc)
        return torch.as_tensor(_DeviceBytes(self.u.pointcloud_device_ptr(pc), n * 16), device=f"cuda:{self.device_index}")

    def from_wire(self, t, timestamp: int, cellsize: float):
        if self.host_wire:
            pc = self.u.cwipc_from_numpy_array(t.numpy().view(self.u.cwipc_point_numpy_dtype), timestamp)
        else:
            pc = self.u.from_device_points(t.data_ptr(), t.numel() // 16, timestamp)
        pc._set_cellsize(cellsize)
        return pc

    def release_wire(self):
        self._keepalive = []


# ======================================================================================================
# the protocol
# ======================================================================================================
def _exchange_points(comm: TorchComm, ops, outgoing: Sequence[Optional[object]], timestamp: int, cellsize: float):
    """outgoing[q]: cloud for rank q or None.  Returns the clouds received, in rank order."""
    wires = [None if (pc is None or ops.count(pc) == 0) else ops.to_wire(pc) for pc in outgoing]
    counts = [0 if w is None else int(w.numel()) for w in wires]
    got = comm.exchange(wires, counts)
    ops.release_wire()
    return [ops.from_wire(t, timestamp, cellsize) for t in got if t is not None]


def column_threshold(v: float, inv: numpy.float32) -> float:
    """Smallest float32 x with floorf(x * inv) >= v (voxel columns are monotone in x), +-inf passed through."""
    if not math.isfinite(v):
        return v
    x = numpy.float32(v) / inv
    col = lambda y: numpy.floor(numpy.float32(y) * inv)  # noqa: E731  (float32 product, as the kernel computes it)
    for _ in range(64):
        if col(x) >= v:
            break
        x = numpy.nextafter(x, numpy.float32(numpy.inf))
    for _ in range(64):
        below = numpy.nextafter(x, numpy.float32(-numpy.inf))
        if col(below) < v:
            break
        x = below
    return float(x)


def slab_downsample(pc, voxelsize: float, comm: TorchComm, ops, timestamp: int = 0):
    """cwipc_downsample of the cloud whose parts are the ranks' `pc` (rank order); returns this rank's part of the result."""
    G, r = comm.size, comm.rank
    octree = not (voxelsize < 0)

    # 0. one collective for everything that is known locally: cellsize metadata, point count, bounding box
    _, b = ops.replay(pc, 1.0, numpy.zeros(9))  # bounding box of this part (the octree state of this call is not used)
    info = comm.gather_rows([float(ops.cellsize(pc)), float(ops.count(pc))] + list(b))
    have = info[:, 1] > 0
    cs = numpy.float32(max(abs(voxelsize), float(info[:, 0].max())))  # ref: src/cwipc_filters.cpp:103-107
    if not have.any():  # every part is empty
        return ops.downsample_planned(pc, voxelsize, numpy.zeros(9), numpy.zeros(6))
    gmin = info[have, 2:5].min(axis=0)
    gmax = info[have, 5:8].max(axis=0)
    xmins = info[:, 2]

    # 1. octree box replay, rank by rank (the box grows with the points IN ORDER), then broadcast
    state = numpy.zeros(9)
    if octree:
        if r > 0:
            state = comm.recv(9, r - 1)
        state, _ = ops.replay(pc, float(cs), state)
        if r < G - 1:
            comm.send(state, r + 1)
        if G > 1:
            state = comm.broadcast(state, G - 1)

    # 2. voxel columns -> owners; boundary points move to the owner of their column
    inv = numpy.float32(1.0) / cs
    splits = numpy.full(G + 1, numpy.inf)  # columns [splits[q], splits[q+1]) belong to rank q
    splits[0] = -numpy.inf
    for q in range(1, G):
        splits[q] = numpy.floor(numpy.float32(xmins[q]) * inv) if have[q] else numpy.inf
    for q in range(G - 1, 0, -1):  # an empty part takes the split of the next one (owns nothing); keep the splits monotone
        splits[q] = min(splits[q], splits[q + 1])
    for q in range(1, G):
        splits[q] = max(splits[q], splits[q - 1])
    edges = [column_threshold(v, inv) for v in splits]
    outgoing: List[Optional[object]] = [None] * G
    keep = pc
    if ops.count(pc) > 0 and (b[0] < edges[r] or b[3] >= edges[r + 1]):
        keep = ops.crop_x(pc, edges[r], edges[r + 1])
        for q in range(G):
            if q != r and edges[q] < edges[q + 1] and b[3] >= edges[q] and b[0] < edges[q + 1]:
                outgoing[q] = ops.crop_x(pc, edges[q], edges[q + 1])
    incoming = _exchange_points(comm, ops, outgoing, timestamp, float(ops.cellsize(pc)))
    mine = ops.join([keep] + incoming) if incoming else keep

    # 3. the local reduction, with the whole cloud's octree box, bounding box and point count (the count fixes the
    #    fixed-point scale of the centroid sums: every part then rounds exactly as the one-GPU call does)
    state[8] = float(info[:, 1].sum())
    return ops.downsample_planned(mine, voxelsize, state, numpy.concatenate([gmin, gmax]))


def slab_remove_outliers(pc, k: int, mul: float, comm: TorchComm, ops, halo: Optional[float] = None, timestamp: int = 0):
    """cwipc_remove_outliers(whole cloud, perTile=False) on the partitioned cloud; returns this rank's survivors."""
    G, r = comm.size, comm.rank
    n_local = ops.count(pc)
    _, b = ops.replay(pc, 1.0, numpy.zeros(9))  # only the bounding box is used here
    info = comm.gather_rows([float(ops.cellsize(pc)), float(n_local), b[0] if n_local else numpy.inf, b[3] if n_local else -numpy.inf])
    counts = info[:, 1]
    ext = info[:, 2:4]
    n_total = int(counts.sum())
    if n_total <= k:  # the reference reads past FLANN's results here; defined as keep-all (see outliers.cu)
        return ops.keep_all(pc)
    cs = float(info[:, 0].max())
    if halo is None:
        if cs > 0:
            halo = 3.0 * cs * math.sqrt((k + 1) / math.pi)
        else:  # no spacing hint: a third of the mean slab width
            span = max(ext[counts > 0, 1].max() - ext[counts > 0, 0].min(), 1e-30)
            halo = span / (3.0 * G)
    H = float(halo)

    # 1. halo exchange: rank q needs every point with x in [xmin_q - H, xmax_q + H]
    outgoing: List[Optional[object]] = [None] * G
    if n_local:
        for q in range(G):
            if q == r or counts[q] == 0:
                continue
            lo = float(numpy.nextafter(numpy.float32(ext[q, 0] - H), numpy.float32(-numpy.inf)))
            hi = float(numpy.nextafter(numpy.float32(ext[q, 1] + H), numpy.float32(numpy.inf)))
            if ext[r, 1] >= lo and ext[r, 0] < hi:
                outgoing[q] = ops.crop_x(pc, lo, hi)
    incoming = _exchange_points(comm, ops, outgoing, timestamp, float(ops.cellsize(pc)))
    combined = ops.join([pc] + incoming) if incoming else pc

    # 2. local queries against local + halo points; the distances stay on the device, only the queries whose
    #    neighbourhood leaves the covered interval come back
    others = [q for q in range(G) if q != r and counts[q] > 0]
    lo_lim = ext[r, 0] - H if any(ext[q, 0] < ext[r, 0] - H for q in others) else -numpy.inf
    hi_lim = ext[r, 1] + H if any(ext[q, 1] > ext[r, 1] + H for q in others) else numpy.inf
    dists, open_idx, open_pts, open_kth2 = ops.knn_open(combined, k, n_local, float(lo_lim), float(hi_lim))

    # 3. the open queries: a rank answers, from its own points, those whose current (k+1)-th distance (an upper bound of
    #    the final one) reaches its x-extent; the owner merges the lists
    rec = numpy.zeros(len(open_pts), numpy.dtype([("p", open_pts.dtype), ("b", "<f4")]))
    rec["p"], rec["b"] = open_pts, open_kth2
    all_q = comm.allgather_bytes(rec)
    nq = [len(a) for a in all_q]
    if sum(nq):
        queries = numpy.concatenate(all_q)
        lists = numpy.full((len(queries), k + 1), numpy.inf, numpy.float32)
        if n_local:
            qx = queries["p"]["x"].astype(numpy.float64)
            rk = numpy.sqrt(queries["b"].astype(numpy.float64)) * (1.0 + 1e-6)
            mine_too = (qx + rk >= ext[r, 0]) & (qx - rk <= ext[r, 1])
            sel = numpy.nonzero(mine_too)[0]
            if len(sel):
                lists[sel] = ops.knn_lists(pc, numpy.ascontiguousarray(queries["p"][sel]), k, numpy.ascontiguousarray(queries["b"][sel]))
        all_lists = comm.allgather_bytes(lists)  # one [Q, k+1] block per rank
        if nq[r]:
            off = sum(nq[:r])
            mine = numpy.stack([blk[off:off + nq[r]] for blk in all_lists])
            ops.patch(dists, ops.merge_lists(mine, k))

    # 4. global statistics, local threshold
    s, sq = ops.distance_stats(dists) if n_local else (0.0, 0.0)
    tot = comm.allreduce(numpy.array([s, sq, float(n_local)]), "SUM")
    thr = ops.threshold(float(tot[0]), float(tot[1]), float(tot[2]), mul)
    return ops.filter_by_distance(pc, dists, thr)
