"""Stand-in for open3d so that the unmodified reference python/cwipc/util.py (which imports open3d at module level,
util.py:24) can be imported on a box without it.  Only the names the module touches at import time exist; the
open3d-based conversions (get_o3d_pointcloud / cwipc_from_o3d_pointcloud) are outside the hot path and raise."""


class _Missing:
    def __init__(self, *a, **k):
        raise RuntimeError("open3d is not available in this environment (tests/stubs/open3d)")


class geometry:
    PointCloud = _Missing


class utility:
    Vector3dVector = _Missing


class visualization:
    Visualizer = _Missing
    VisualizerWithKeyCallback = _Missing
    draw_geometries = _Missing
