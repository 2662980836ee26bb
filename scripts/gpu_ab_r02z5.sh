# A/B of the voxel list in 64 segments with a counter each (CWIPC_CUDA_DS_COUNTERS): measured, no gain, NOT kept -- the knob no longer exists;
# the output is profiles/r02z_ab_threads_counters.txt
TAG=${1:-r02z5}
OUT=gpurun_out/ab_${TAG}.txt
: > $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -m gpu -k "downsample or golden or chain or config4 or config3 or config5 or full_size or synthetic" > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$? >> $OUT
tail -3 gpurun_out/pytest_${TAG}.log >> $OUT
for k in 64 1 8; do
for cfg in "--points 1000000" "--points 7997584 --clean --flush" "--points 7997584 --clean --flush --voxel 0.002"; do
  echo "== counters $k diag_stream $cfg" >> $OUT
  CWIPC_CUDA_DS_COUNTERS=$k CWIPC_CUDA_DEBUG_STREAM=1 timeout 200 python scripts/diag_stream.py $cfg 2>&1 | grep -v "^\[" | tail -3 | cut -c 1-600 >> $OUT
done
done
run() { label=$1; shift; echo "== $label" >> $OUT; ( env "$@" timeout 300 python scripts/ab_value.py --tag "$label" $ABARGS 2>> gpurun_out/ab_${TAG}.err | tail -1 | cut -c 1-1500 ) >> $OUT; }
ABARGS="--workers 30 --profile" run "64 counters" X=1
ABARGS="--workers 30" run "1 counter" CWIPC_CUDA_DS_COUNTERS=1
ABARGS="--workers 30" run "64 counters again" X=1
cat $OUT
