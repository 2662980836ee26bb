set -x
timeout 900 python -m pytest tests -x -q -m gpu -k "knn or outliers or chain or golden" > gpurun_out/pytest_knn.log 2>&1; echo knn_exit=$?; tail -5 gpurun_out/pytest_knn.log
for cfg in "0.8 1.5" "0.6 1.5" "1.0 1.5" "1.3 1.5" "1.0 1.0" "1.3 1.0" "1.6 1.0"; do
  set -- $cfg
  CWIPC_CUDA_KNN_PITCH=$1 CWIPC_CUDA_KNN_RC=$2 timeout 300 python bench.py --steps 3 --warmup 3 --cpu-frames 1 > gpurun_out/tune_$1_$2.json 2> gpurun_out/tune.err; echo exit=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/tune_$1_$2.json"))
k=d["roofline"]["kernels"]
print("pitch $1 rc $2 value",d["value"],"e2e",d["e2e"]["value"],"tile_ms",k.get("knn_tile_kernel",{}).get("ms"),"far_ms",k.get("knn_far_kernel",{}).get("ms"),"out",d["out_points_per_step"])
PY
done
