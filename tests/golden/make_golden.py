"""Generate tests/golden/golden_small.npz: a 6 400-point 4-camera synthetic cloud and the oracle's
outputs for every filter on the hot path.  The reference itself cannot run in this environment (its
arithmetic is PCL's, which is absent), so these vectors pin the ORACLE (regression) and give the GPU
tests a fixture that does not need the oracle library at all.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "oracle"))

import oracle  # noqa: E402
from cwipc_util_b200 import synthetic  # noqa: E402


def raw(a):
    return np.frombuffer(a.tobytes(), np.uint8).reshape(-1, 16)


def main():
    pts = synthetic.camera_cloud(6400, seed=11)
    out = {"points": raw(pts)}
    out["tilefilter_4"] = raw(oracle.tilefilter(pts, 4))
    for name, vs in (("ds_pos", 0.06), ("ds_neg", -0.06)):
        o, _, _, counts = oracle.downsample(pts, vs, 0.0)
        out[name] = raw(o)
        out[name + "_counts"] = counts
    out["knn30"] = oracle.knn_mean_distances(pts, 30)
    out["sor_all"] = raw(oracle.remove_outliers(pts, 30, 1.0, False)[0])
    out["sor_pertile"] = raw(oracle.remove_outliers(pts, 30, 1.0, True)[0])
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
