# final evidence of a round: full GPU tests, smoke, bench (both arms), ncu launch list + full capture
set -x
TAG=$1
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$?; tail -3 gpurun_out/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${TAG}.log 2>&1; echo smoke_exit=$?; tail -1 gpurun_out/smoke_${TAG}.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo bench_exit=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; echo ref_exit=$?
bash scripts/gpu_ncu.sh ${TAG} "knn_tile_kernel|knn_far_kernel|voxel_stream_kernel|voxel_words_kernel|radix_cluster_kernel|voxel_emit_kernel|knn_layout_kernel|compact_kernel|stats_threshold_kernel|knn_keygen_kernel" 10 10
python scripts/sass_summary.py > /dev/null
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json")); r=d["roofline"]
print("value",d["value"],"e2e",d["e2e"]["value"],"launches",d["gpu_launches"],"clocks",d["clocks"])
print({k:r[k] for k in ("kernel","achieved","frac","traffic","avg_launch_us","share_of_step_kernel_time","op_compulsory_frac")})
for k,v in r["kernels"].items(): print("  ",k,v)
print("reference", json.load(open("gpurun_out/bench_ref_${TAG}.json"))["value"])
PY
