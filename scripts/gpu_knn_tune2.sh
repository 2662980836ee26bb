for cfg in "1.0 1.0 128" "1.0 1.0 256" "1.0 1.0 512" "1.0 1.0 1024" "0.8 1.25 256" "1.2 0.85 256" "0.7 1.5 256" "1.4 0.75 256"; do
  set -- $cfg
  CWIPC_CUDA_KNN_PITCH=$1 CWIPC_CUDA_KNN_RC=$2 CWIPC_CUDA_KNN_LEAF=$3 timeout 300 python bench.py --steps 3 --warmup 3 --cpu-frames 1 > gpurun_out/tune2.json 2> gpurun_out/tune2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/tune2.json"))
k=d["roofline"]["kernels"]
print("pitch $1 rc $2 leaf $3 value",d["value"],"tile_us",k["knn_tile_kernel"]["us_per_launch"],"far_us",k["knn_far_kernel"]["us_per_launch"])
PY
done
