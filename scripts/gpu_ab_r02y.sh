# round 2, re-entry session: targeted tests of the changed paths + A/B of the new tuning knobs (scripts/ab_value.py)
set -x
TAG=${1:-r02y}
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn or outliers or staging or roundtrip or object or packet" > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$?
tail -5 gpurun_out/pytest_${TAG}.log
OUT=gpurun_out/ab_${TAG}.txt
: > $OUT
run() { # label, env..., -- args
  label=$1; shift
  echo "== $label" >> $OUT
  ( env "$@" timeout 300 python scripts/ab_value.py --tag "$label" $ABARGS 2>> gpurun_out/ab_${TAG}.err | tail -1 | cut -c 1-1500 ) >> $OUT
}
ABARGS="--profile --pageable" run "old: far search from the root, driver-staged pageable copies" CWIPC_CUDA_KNN_FAR_START=0 CWIPC_CUDA_STAGING=0
ABARGS="--profile --pageable" run "new defaults" X=1
ABARGS="" run "all kernels prefer max shared carve-out" CWIPC_CUDA_CARVEOUT=100
ABARGS="" run "stream kernel: one block per SM, driver's carve-out" CWIPC_CUDA_DS_CARVEOUT=-1 CWIPC_CUDA_DS_BLOCKS_PER_SM=1
ABARGS="" run "stream kernel: one block per SM, max carve-out" CWIPC_CUDA_DS_BLOCKS_PER_SM=1
ABARGS="" run "stream kernel: one block per SM, everything prefers 50 %" CWIPC_CUDA_DS_CARVEOUT=50 CWIPC_CUDA_DS_BLOCKS_PER_SM=1 CWIPC_CUDA_CARVEOUT=50
ABARGS="" run "stream kernel: driver's carve-out, two blocks" CWIPC_CUDA_DS_CARVEOUT=-1
ABARGS="" run "250 K-point frames (host-bound or GPU-bound?)" BENCH_POINTS=250000
ABARGS="" run "stage: downsample only" BENCH_STAGE=downsample
ABARGS="" run "stage: outliers only" BENCH_STAGE=outliers
ABARGS="--workers 20" run "20 host threads" X=1
cat $OUT
# fixed cost of the streaming kernel on a 1 M-point frame: time of the last block's octree replay, per-kernel times at 1 M points
CWIPC_CUDA_DEBUG_TAIL=1 timeout 200 python scripts/bench_ds.py --points 1000000 --reps 3 2> gpurun_out/ds1m_${TAG}.err | tail -1 > gpurun_out/ds1m_${TAG}.json
grep "last block" gpurun_out/ds1m_${TAG}.err | sort | uniq -c | sort -rn | head -8
grep '"cloud"' gpurun_out/ds1m_${TAG}.err | cut -c 1-400
