#!/usr/bin/env python
"""Diagnostic: cwipc_remove_outliers on the 8M clean synthetic cloud, plain call and the one-rank slab entry point, with kernel profile."""
import ctypes, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import cwipc_util_b200 as cw
from cwipc_util_b200 import synthetic, util
lib = util.cwipc_util_dll_load()
n_req = 2828 * 2828
pts = synthetic.simulate_cameras(synthetic.synthetic_cloud(n_req), 4)
if len(sys.argv) > 1 and sys.argv[1] == "jittered":
    pts = synthetic.add_outliers(synthetic.add_noise(pts, 0.002, 1), 0.005, 1)
pc = cw.cwipc_from_numpy_array(pts, 3); pc._set_cellsize(synthetic.cellsize_of(n_req))
comm = util.cuda_comm(None, 1, 0)

def run(fn, name):
    for i in range(3):
        cw.cuda_synchronize()
        t = lib.cwipc_cuda_timer_create(); lib.cwipc_cuda_timer_start(t)
        out = fn()
        lib.cwipc_cuda_timer_stop(t); cw.cuda_synchronize()
        ms = lib.cwipc_cuda_timer_elapsed_ms(t); lib.cwipc_cuda_timer_destroy(t)
        print(name, "run", i, "ms", round(ms, 3), "kept", out.count(), flush=True)
    lib.cwipc_cuda_profile_reset(); lib.cwipc_cuda_profile_enable(1)
    fn(); cw.cuda_synchronize(); lib.cwipc_cuda_profile_enable(0)
    need = lib.cwipc_cuda_profile_report(None, 0); buf = ctypes.create_string_buffer(need); lib.cwipc_cuda_profile_report(buf, need)
    prof = json.loads(buf.value.decode())
    print(name, {k: round(v["total_ms"] * 1e3, 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["total_ms"])[:8]}, flush=True)

run(lambda: cw.cwipc_remove_outliers(pc, 30, 1.0, False), "plain")
if len(sys.argv) <= 2:
    run(lambda: comm.remove_outliers(pc, 30, 1.0, False), "slab1")
