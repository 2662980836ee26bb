# usage: bash scripts/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
# launch list + one full capture of the named kernels on a short single-stream bench run
set -x
TAG=$1; REGEX=$2; SKIP=${3:-40}; COUNT=${4:-12}
CMD="python bench.py --steps 1 --warmup 3 --frames-per-gpu 4 --passes 1 --workers 1 --cpu-frames 1 --skip-config4"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1; echo ncu1_exit=$?
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s $SKIP -c $COUNT -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1; echo ncu2_exit=$?
tail -2 gpurun_out/plain_${TAG}.log
