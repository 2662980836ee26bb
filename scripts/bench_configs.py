#!/usr/bin/env python
"""Every BASELINE.json config once, on the GPU(s) of this process group: throughput (CUDA events, L2 flushed between
repetitions) and, where the oracle finishes in seconds, a parity check.  One JSON object per config on stdout.

    python scripts/bench_configs.py                       # configs[0..2] + configs[3] on one GPU (no partition)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_configs.py --slab   # configs[3] as x-slabs
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))


def timed(cw, lib, fn, reps=5, warm=2):
    times, out = [], None
    for i in range(reps + warm):
        lib.cwipc_cuda_flush_l2()
        cw.cuda_synchronize()
        t = lib.cwipc_cuda_timer_create()
        lib.cwipc_cuda_timer_start(t)
        out = fn()
        lib.cwipc_cuda_timer_stop(t)
        cw.cuda_synchronize()
        if i >= warm:
            times.append(lib.cwipc_cuda_timer_elapsed_ms(t))
        lib.cwipc_cuda_timer_destroy(t)
    return float(np.median(times)), out


def single_gpu(args):
    import cwipc_util_b200 as cw
    from cwipc_util_b200 import synthetic
    import oracle
    oracle.load()
    lib = cw.util.cwipc_util_dll_load()
    rows = []

    # configs[0]: default synthetic cloud (160 000 points, cellsize metadata 0 as after a PLY round trip), downsample 0.01
    pts = synthetic.synthetic_cloud(160000)
    pc = cw.cwipc_from_numpy_array(pts, 1)
    ms, out = timed(cw, lib, lambda: cw.cwipc_downsample(pc, 0.01))
    t0 = time.perf_counter()
    want, cs, _, _ = oracle.downsample(pts, 0.01, 0.0)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    got = out.get_numpy_array()
    ok = len(got) == len(want) and np.array_equal(got["tile"], want["tile"]) and all(
        (np.abs(got[a] - want[a]) / np.maximum(np.abs(want[a]), cs)).max() <= 1e-5 for a in "xyz")
    rows.append({"config": 0, "what": "160K synthetic, cwipc_downsample(0.01)", "points": len(pts), "out": out.count(), "gpu_ms": round(ms, 3),
                 "Mpoints_per_s": round(len(pts) / ms / 1e3, 1), "oracle_cpu_ms": round(cpu_ms, 1), "parity": bool(ok)})

    # configs[1]: 1M-point 4-camera cloud, remove_outliers(30, 1.0, perTile=True)
    pts = synthetic.camera_cloud(1000000, seed=0)
    pc = cw.cwipc_from_numpy_array(pts, 1)
    pc._set_cellsize(synthetic.cellsize_of(1000000))
    ms, out = timed(cw, lib, lambda: cw.cwipc_remove_outliers(pc, 30, 1.0, True), reps=3, warm=1)
    row = {"config": 1, "what": "1M 4-camera, cwipc_remove_outliers(30,1.0,perTile)", "points": len(pts), "out": out.count(), "gpu_ms": round(ms, 3),
           "Mpoints_per_s": round(len(pts) / ms / 1e3, 1)}
    if args.parity:
        t0 = time.perf_counter()
        want, _ = oracle.remove_outliers(pts, 30, 1.0, True)
        row["oracle_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        row["parity"] = bool(abs(len(want) - out.count()) <= 4)
        row["count_oracle"] = len(want)
    rows.append(row)

    # configs[2]: tilefilter(1) -> downsample(0.005) -> remove_outliers(30, 1.0) on a 2M-point cloud, device resident
    pts = synthetic.camera_cloud(2000000, seed=1)
    pc = cw.cwipc_from_numpy_array(pts, 1)
    pc._set_cellsize(synthetic.cellsize_of(2000000))

    def chain():
        t = cw.cwipc_tilefilter(pc, 1)
        d = cw.cwipc_downsample(t, 0.005)
        return cw.cwipc_remove_outliers(d, 30, 1.0, False)
    ms, out = timed(cw, lib, chain)
    row = {"config": 2, "what": "2M 4-camera, tilefilter(1) -> downsample(0.005) -> remove_outliers(30,1.0)", "points": len(pts), "out": out.count(),
           "gpu_ms": round(ms, 3), "Mpoints_per_s": round(len(pts) / ms / 1e3, 1)}
    if args.parity:
        t0 = time.perf_counter()
        a = oracle.tilefilter(pts, 1)
        b, cs, _, _ = oracle.downsample(a, 0.005, synthetic.cellsize_of(2000000))
        c, _ = oracle.remove_outliers(b, 30, 1.0, False)
        row["oracle_cpu_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
        row["parity"] = bool(abs(len(c) - out.count()) <= 2)
    rows.append(row)
    for r in rows:
        print(json.dumps(r), flush=True)


def slab(args):
    import torch
    import torch.distributed as dist
    import cwipc_util_b200 as cw
    from cwipc_util_b200 import slab as slabmod, synthetic
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29711")
    dist.init_process_group("nccl", rank=rank, world_size=world)
    lib = cw.util.cwipc_util_dll_load()
    comm, ops = slabmod.TorchComm(f"cuda:{local}"), slabmod.CudaOps(local)
    n = args.points
    pts = synthetic.camera_cloud(n, seed=1)                     # every rank generates the same cloud and keeps its x-slab
    order = np.argsort(pts["x"], kind="stable")
    lo, hi = len(pts) * rank // world, len(pts) * (rank + 1) // world
    part = pts[order[lo:hi]].copy()
    pc = cw.cwipc_from_numpy_array(part, 1)
    pc._set_cellsize(synthetic.cellsize_of(n))
    for vs in (0.002, 0.005, 0.01, 0.02, 0.05):
        times, counts = [], None
        for i in range(args.reps + 1):
            lib.cwipc_cuda_flush_l2()
            cw.cuda_synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            t = lib.cwipc_cuda_timer_create()
            lib.cwipc_cuda_timer_start(t)
            ds = slabmod.slab_downsample(pc, vs, comm, ops)
            lib.cwipc_cuda_timer_stop(t)
            cw.cuda_synchronize()
            t_ds = lib.cwipc_cuda_timer_elapsed_ms(t)
            lib.cwipc_cuda_timer_start(t)
            kept = slabmod.slab_remove_outliers(ds, 30, 1.0, comm, ops)
            lib.cwipc_cuda_timer_stop(t)
            cw.cuda_synchronize()
            t_sor = lib.cwipc_cuda_timer_elapsed_ms(t)
            lib.cwipc_cuda_timer_destroy(t)
            tt = comm.allreduce(np.array([t_ds, t_sor]), "MAX")
            counts = comm.allreduce(np.array([float(ds.count()), float(kept.count())]), "SUM")
            if i >= 1:
                times.append(tt)
        if rank == 0:
            tm = np.median(np.array(times), axis=0)
            print(json.dumps({"config": 3, "what": "8M cloud as x-slabs: slab_downsample -> slab_remove_outliers(30,1.0)", "n_gpus": world, "points": len(pts),
                              "voxelsize": vs, "voxels": int(counts[0]), "kept": int(counts[1]), "downsample_ms": round(float(tm[0]), 3),
                              "remove_outliers_ms": round(float(tm[1]), 3), "Mpoints_per_s": round(len(pts) / float(tm.sum()) / 1e3, 1)}), flush=True)
    # outlier removal of the RAW cloud: the case where a slab's compute (milliseconds) outweighs the protocol
    times, counts = [], None
    for i in range(args.reps + 1):
        lib.cwipc_cuda_flush_l2()
        cw.cuda_synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        t = lib.cwipc_cuda_timer_create()
        lib.cwipc_cuda_timer_start(t)
        kept = slabmod.slab_remove_outliers(pc, 30, 1.0, comm, ops)
        lib.cwipc_cuda_timer_stop(t)
        cw.cuda_synchronize()
        tt = comm.allreduce(np.array([lib.cwipc_cuda_timer_elapsed_ms(t)]), "MAX")
        lib.cwipc_cuda_timer_destroy(t)
        counts = comm.allreduce(np.array([float(kept.count())]), "SUM")
        if i >= 1:
            times.append(float(tt[0]))
    if rank == 0:
        ms = float(np.median(times))
        print(json.dumps({"config": 3, "what": "8M cloud as x-slabs: slab_remove_outliers(30,1.0) of the raw cloud", "n_gpus": world, "points": len(pts),
                          "kept": int(counts[0]), "remove_outliers_ms": round(ms, 3), "Mpoints_per_s": round(len(pts) / ms / 1e3, 1)}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--slab", action="store_true")
    ap.add_argument("--points", type=int, default=8000000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--parity", action="store_true", help="also run the oracle on the 1M/2M configs (tens of seconds of CPU)")
    a = ap.parse_args()
    slab(a) if a.slab else single_gpu(a)
