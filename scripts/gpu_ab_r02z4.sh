TAG=${1:-r02z4}
OUT=gpurun_out/ab_${TAG}.txt
: > $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn or outliers" > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$? >> $OUT
tail -2 gpurun_out/pytest_${TAG}.log >> $OUT
run() { label=$1; shift; echo "== $label" >> $OUT; ( env "$@" timeout 300 python scripts/ab_value.py --tag "$label" $ABARGS 2>> gpurun_out/ab_${TAG}.err | tail -1 | cut -c 1-1500 ) >> $OUT; }
ABARGS="--profile" run "15 threads" X=1
ABARGS="--workers 24" run "24 threads" X=1
ABARGS="--workers 32" run "32 threads" X=1
ABARGS="" run "15 threads again" X=1
ABARGS="--workers 24" run "24 threads again" X=1
ABARGS="--workers 24" run "24 threads, 64 hardware queues" CUDA_DEVICE_MAX_CONNECTIONS=64
cat $OUT
