#!/usr/bin/env python
"""Profiling target: cwipc_downsample of one synthetic cloud, a few times (run it under ncu; see scripts/gpu_ncu_ds.sh)."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import cwipc_util_b200 as cw
from cwipc_util_b200 import synthetic

points = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
voxel = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
clean = len(sys.argv) > 4 and sys.argv[4] == "clean"
pts = synthetic.simulate_cameras(synthetic.synthetic_cloud(points), 4) if clean else synthetic.camera_cloud(points, seed=1)
pc = cw.cwipc_from_numpy_array(pts, 1)
pc._set_cellsize(synthetic.cellsize_of(points))
for _ in range(reps):
    out = cw.cwipc_downsample(pc, voxel)
    print(out.count())
    out.free()
    cw.util.cwipc_util_dll_load().cwipc_cuda_flush_l2()
    cw.cuda_synchronize()
