#!/usr/bin/env python
"""BASELINE.json configs[3] on ONE GPU: an 8M-point synthetic cloud, cwipc_downsample at the voxel sizes of the sweep,
cwipc_remove_outliers(30, 1.0) on the raw cloud and on the 0.005 result.  CUDA-event timed, L2 flushed between runs,
per-kernel breakdown from the library's launch profile.  One JSON object on stdout.

    python scripts/bench_big.py [--points 8000000] [--reps 5]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=8000000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--skip-raw-sor", action="store_true")
    args = ap.parse_args()
    import cwipc_util_b200 as cw
    from cwipc_util_b200 import synthetic
    lib = cw.util.cwipc_util_dll_load()
    t0 = time.perf_counter()
    pts = synthetic.camera_cloud(args.points, seed=1)
    n = len(pts)
    print(f"generated {n} points in {time.perf_counter() - t0:.1f}s", file=sys.stderr)
    pc = cw.cwipc_from_numpy_array(pts, 1)
    pc._set_cellsize(synthetic.cellsize_of(args.points))
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    def timed(fn, reps):
        """median CUDA-event time (ms) of fn(), L2 flushed before every run; returns (ms, last result)"""
        times, out = [], None
        for i in range(reps + 2):
            lib.cwipc_cuda_flush_l2()
            cw.cuda_synchronize()
            t = lib.cwipc_cuda_timer_create()
            lib.cwipc_cuda_timer_start(t)
            out = fn()
            lib.cwipc_cuda_timer_stop(t)
            cw.cuda_synchronize()
            if i >= 2:
                times.append(lib.cwipc_cuda_timer_elapsed_ms(t))
            lib.cwipc_cuda_timer_destroy(t)
        return float(np.median(times)), out

    def profiled(fn):
        lib.cwipc_cuda_profile_reset()
        lib.cwipc_cuda_profile_enable(1)
        lib.cwipc_cuda_flush_l2()
        fn()
        cw.cuda_synchronize()
        lib.cwipc_cuda_profile_enable(0)
        need = lib.cwipc_cuda_profile_report(None, 0)
        buf = ctypes.create_string_buffer(need)
        lib.cwipc_cuda_profile_report(buf, need)
        prof = json.loads(buf.value.decode())
        return {k: {"us": round(v["total_ms"] * 1e3 / max(1, v["launches"]), 1), "launches": v["launches"],
                    "GBps": round(v["bytes"] / max(v["total_ms"], 1e-9) / 1e6, 1), "frac_of_hbm_peak": round(v["bytes"] / max(v["total_ms"], 1e-9) / 1e6 / peak, 4)}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]) if k != "flush_kernel"}

    res = {"points": n, "hbm_peak_GBps": peak, "downsample": [], "remove_outliers": []}
    keep = {}
    for vs in (0.002, 0.005, 0.01, 0.02, 0.05):
        ms, out = timed(lambda: cw.cwipc_downsample(pc, vs), args.reps)
        v = out.count()
        row = {"voxelsize": vs, "voxels": v, "ms": round(ms, 3), "Mpoints_per_s": round(n / ms / 1e3, 1),
               "compulsory_GBps": round(16.0 * (n + v) / ms / 1e6, 1), "compulsory_frac": round(16.0 * (n + v) / ms / 1e6 / peak, 4),
               "kernels": profiled(lambda: cw.cwipc_downsample(pc, vs))}
        res["downsample"].append(row)
        print(json.dumps(row), file=sys.stderr)
        if vs == 0.005:
            keep[vs] = out
    cases = [("downsampled 0.005", keep[0.005])]
    if not args.skip_raw_sor:
        cases.append(("raw", pc))
    for name, cloud in cases:
        m = cloud.count()
        ms, out = timed(lambda: cw.cwipc_remove_outliers(cloud, 30, 1.0, False), max(2, args.reps // 2))
        row = {"input": name, "points": m, "kept": out.count(), "ms": round(ms, 3), "Mpoints_per_s": round(m / ms / 1e3, 1),
               "kernels": profiled(lambda: cw.cwipc_remove_outliers(cloud, 30, 1.0, False))}
        res["remove_outliers"].append(row)
        print(json.dumps(row), file=sys.stderr)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
