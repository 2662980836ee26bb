#!/usr/bin/env python
"""cwipc_downsample of ONE large cloud, voxel-size sweep of BASELINE configs[3], CUDA-event timed with the L2 flushed before
every call: the clean synthetic cloud (as cwipc_synthetic generates it) and the jittered 4-camera
variant (2 mm noise + 0.5 % outliers) that scripts/bench_big.py uses.

    python scripts/bench_ds.py [--points 8000000] [--reps 7]
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=8000000)
    ap.add_argument("--reps", type=int, default=7)
    args = ap.parse_args()
    import cwipc_util_b200 as cw
    from cwipc_util_b200 import synthetic
    lib = cw.util.cwipc_util_dll_load()
    peak = 6544.0
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    clouds = {"synthetic": synthetic.simulate_cameras(synthetic.synthetic_cloud(args.points), 4), "jittered": synthetic.camera_cloud(args.points, seed=1)}
    res = {"hbm_peak_GBps": peak, "rows": []}
    for name, pts in clouds.items():
        n = len(pts)
        pc = cw.cwipc_from_numpy_array(pts, 1)
        pc._set_cellsize(synthetic.cellsize_of(args.points))
        for vs in (0.002, 0.005, 0.01, 0.02, 0.05):
            times, out = [], None
            for i in range(args.reps + 2):
                lib.cwipc_cuda_flush_l2()
                cw.cuda_synchronize()
                t = lib.cwipc_cuda_timer_create()
                lib.cwipc_cuda_timer_start(t)
                out = cw.cwipc_downsample(pc, vs)
                lib.cwipc_cuda_timer_stop(t)
                cw.cuda_synchronize()
                if i >= 2:
                    times.append(lib.cwipc_cuda_timer_elapsed_ms(t))
                lib.cwipc_cuda_timer_destroy(t)
            ms = float(np.median(times))
            v = out.count()
            lib.cwipc_cuda_profile_reset()
            lib.cwipc_cuda_profile_enable(1)
            lib.cwipc_cuda_flush_l2()
            cw.cwipc_downsample(pc, vs)
            cw.cuda_synchronize()
            lib.cwipc_cuda_profile_enable(0)
            need = lib.cwipc_cuda_profile_report(None, 0)
            buf = ctypes.create_string_buffer(need)
            lib.cwipc_cuda_profile_report(buf, need)
            prof = json.loads(buf.value.decode())
            row = {"cloud": name, "points": n, "voxelsize": vs, "voxels": v, "ms": round(ms, 4), "Mpoints_per_s": round(n / ms / 1e3, 1),
                   "compulsory_frac": round(16.0 * (n + v) / ms / 1e6 / peak, 4),
                   "kernels_us": {k: round(x["total_ms"] * 1e3, 1) for k, x in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"]) if k != "flush_kernel"}}
            st = prof.get("voxel_stream_kernel")
            if st:
                row["stream_GBps"] = round(16.0 * n / st["total_ms"] / 1e6, 1)
                row["stream_frac"] = round(16.0 * n / st["total_ms"] / 1e6 / peak, 4)
            res["rows"].append(row)
            print(json.dumps(row), file=sys.stderr)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
