#!/usr/bin/env python
"""Where the streaming downsample kernel spends its time on ONE cloud (single stream): run with CWIPC_CUDA_DEBUG_STREAM=1
(in-kernel globaltimer stamps, csrc/downsample.cu) and CWIPC_CUDA_DEBUG_TAIL=1; prints the event-timed kernel durations beside
them.   python scripts/diag_stream.py [--points 1000000] [--voxel 0.01] [--flush]"""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1000000)
    ap.add_argument("--voxel", type=float, default=0.01)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--clean", action="store_true", help="the clean synthetic cloud instead of a bench frame")
    args = ap.parse_args()
    import cwipc_util_b200 as cw
    from cwipc_util_b200 import synthetic
    lib = cw.util.cwipc_util_dll_load()
    bench.POINTS_PER_FRAME = args.points
    pts = synthetic.simulate_cameras(synthetic.synthetic_cloud(args.points), 4) if args.clean else bench.make_frames(0, 1, 1)[0]
    pc = cw.cwipc_from_numpy_array(pts, 1)
    pc._set_cellsize(synthetic.cellsize_of(args.points))
    for i in range(6):
        if args.flush:
            lib.cwipc_cuda_flush_l2()
        if i == 3:
            lib.cwipc_cuda_profile_reset()
            lib.cwipc_cuda_profile_enable(1)
        cw.cwipc_downsample(pc, args.voxel).free()
        cw.cuda_synchronize()
    lib.cwipc_cuda_profile_enable(0)
    need = lib.cwipc_cuda_profile_report(None, 0)
    buf = ctypes.create_string_buffer(need)
    lib.cwipc_cuda_profile_report(buf, need)
    prof = json.loads(buf.value.decode())
    print(json.dumps({"points": len(pts), "voxel": args.voxel, "flush": args.flush,
                      "kernels_us": {k: round(v["total_ms"] * 1e3 / max(1, v["launches"]), 1) for k, v in prof.items() if k != "flush_kernel"}}))


if __name__ == "__main__":
    main()
