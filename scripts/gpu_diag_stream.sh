TAG=${1:-r02z2}
OUT=gpurun_out/diag_stream_${TAG}.txt
: > $OUT
for cfg in "--points 1000000" "--points 1000000 --voxel 0.05" "--points 7997584 --clean --flush" "--points 7997584 --clean --flush --voxel 0.002"; do
  echo "== diag_stream $cfg" >> $OUT
  CWIPC_CUDA_DEBUG_STREAM=1 CWIPC_CUDA_DEBUG_TAIL=1 timeout 200 python scripts/diag_stream.py $cfg 2>&1 | grep -v "^\[" | tail -4 | cut -c 1-600 >> $OUT
done
cat $OUT
