# last check of a round without ncu: full GPU tests, smoke, bench (both arms)
set -x
TAG=$1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo smoke_exit=$?; tail -1 gpurun_out/smoke_${TAG}.log
timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo bench_exit=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; echo ref_exit=$?
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json")); r=d["roofline"]
print("value",d["value"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"],"clocks",d["clocks"])
print({k:r[k] for k in ("kernel","achieved","frac","traffic","issue","avg_launch_us","share_of_step_kernel_time","op_compulsory_frac")})
print("reference", json.load(open("gpurun_out/bench_ref_${TAG}.json"))["value"], "lines", sum(1 for _ in open("gpurun_out/bench_${TAG}.json")))
PY
