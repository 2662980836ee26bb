set -x
timeout 600 python -m pytest tests -x -q -m gpu -k "knn or outliers" > gpurun_out/pytest_knn.log 2>&1; echo knn_exit=$?; tail -15 gpurun_out/pytest_knn.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo gpu_exit=$?; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; echo bench_exit=$?; tail -3 gpurun_out/bench_r01b.err; cat gpurun_out/bench_r01b.json
