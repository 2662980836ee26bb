#!/usr/bin/env python
"""Diagnostic: GPU kNN mean distances vs the oracle on a noisy cloud (run with the CWIPC_CUDA_KNN_* tunables to force a path)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "oracle"))
import cwipc_util_b200 as cw
from cwipc_util_b200 import synthetic
import oracle as orc
orc.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 90000
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 31
pts = synthetic.camera_cloud(n, seed=seed, noise=0.004)
pc = cw.cwipc_from_numpy_array(pts, 1); pc._set_cellsize(synthetic.cellsize_of(n))
got = cw.util.knn_mean_distances(pc, 30)
want = orc.knn_mean_distances(pts, 30)
bad = np.flatnonzero(got != want)
print({k: v for k, v in os.environ.items() if k.startswith("CWIPC_CUDA_KNN")}, "mismatches", len(bad), "of", n)
if len(bad):
    h = synthetic.cellsize_of(n) * np.sqrt(31 / np.pi)
    print("got>want", int((got[bad] > want[bad]).sum()), "got<want", int((got[bad] < want[bad]).sum()), "nan/inf", int((~np.isfinite(got[bad])).sum()))
    print("want/h quantiles", np.quantile(want[bad] / h, [0, .25, .5, .75, 1]))
    print("first", bad[:10], got[bad[:10]], want[bad[:10]])
sys.path.insert(0, os.path.join(REPO, "tests"))
from parity_helpers import per_tile_check
out = cw.cwipc_remove_outliers(pc, 12, 1.5, True).get_numpy_array()
print("per tile kept", len(out))
try:
    per_tile_check(orc, pts, out, 12, 1.5)
    print("per tile OK")
except AssertionError as e:
    print("per tile FAILED", str(e)[:300])
