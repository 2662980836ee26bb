set -x
TAG=${1:-r02z}
OUT=gpurun_out/ab_${TAG}.txt
: > $OUT
for cfg in "--points 1000000" "--points 1000000 --flush" "--points 1000000 --voxel 0.05" "--points 4000000" "--points 7997584 --clean --flush"; do
  echo "== diag_stream $cfg" >> $OUT
  CWIPC_CUDA_DEBUG_STREAM=1 CWIPC_CUDA_DEBUG_TAIL=1 timeout 200 python scripts/diag_stream.py $cfg 2>&1 | grep -v "^\[" | tail -7 | cut -c 1-600 >> $OUT
done
run() { label=$1; shift; echo "== $label" >> $OUT; ( env "$@" timeout 300 python scripts/ab_value.py --tag "$label" $ABARGS 2>> gpurun_out/ab_${TAG}.err | tail -1 | cut -c 1-1500 ) >> $OUT; }
ABARGS="" run "defaults" X=1
ABARGS="--workers 20" run "carve-out 100, 20 threads" CWIPC_CUDA_CARVEOUT=100
ABARGS="--workers 24" run "carve-out 100, 24 threads" CWIPC_CUDA_CARVEOUT=100
ABARGS="--workers 30" run "carve-out 100, 30 threads" CWIPC_CUDA_CARVEOUT=100
ABARGS="--workers 24" run "24 threads" X=1
ABARGS="" run "carve-out 100" CWIPC_CUDA_CARVEOUT=100
cat $OUT
