# A/B of library builds kept under build/variants/*.so (CWIPC_CUDA_LIBRARY picks one): bench value, per-kernel times, and the
# warp-instruction counts of the kNN kernels from a short ncu metrics pass
TAG=${1:-r02v}
OUT=gpurun_out/ab_${TAG}.txt
: > $OUT
for lib in build/variants/*.so; do
  echo "== $lib" >> $OUT
  CWIPC_CUDA_LIBRARY=$PWD/$lib timeout 300 python scripts/ab_value.py --workers 30 --profile --tag "$lib" 2>> gpurun_out/ab_${TAG}.err | tail -1 | cut -c 1-900 >> $OUT
  CWIPC_CUDA_LIBRARY=$PWD/$lib timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:"knn_tile_kernel|knn_far_kernel|voxel_stream_kernel" -s 6 -c 6 --csv python bench.py --steps 1 --warmup 3 --frames-per-gpu 4 --passes 1 --workers 1 --cpu-frames 1 --skip-config4 2>/dev/null | grep -E "inst_executed|time_duration" | awk -F'","' '{n=$5; sub(/.*unnamed>::/,"",n); print substr(n,1,28), $(NF-2), $NF}' >> $OUT
done
cat $OUT
