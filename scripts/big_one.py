import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cwipc_util_b200 as cw
from cwipc_util_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000000
vs = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
pts = synthetic.camera_cloud(n, seed=1)
pc = cw.cwipc_from_numpy_array(pts, 1)
pc._set_cellsize(synthetic.cellsize_of(n))
for _ in range(3):
    d = cw.cwipc_downsample(pc, vs)
    cw.cuda_synchronize()
print(d.count())
