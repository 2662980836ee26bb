# usage: bash scripts/gpu_test_bench.sh <tag> [pytest -k expr]
set -x
TAG=$1; KEXPR=${2:-}
if [ -n "$KEXPR" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu -k "$KEXPR" > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$?
else
  timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$?
fi
tail -15 gpurun_out/pytest_${TAG}.log
timeout 600 python bench.py --steps 5 --warmup 3 --cpu-frames 2 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo bench_exit=$?; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms_per_step",d["ms_per_step"],"launches",d["gpu_launches"])
for n,v in d["roofline"]["kernels"].items(): print("  ",n,v)
PY
