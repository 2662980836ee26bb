#!/bin/bash
# kNN pitch / cover sweep (results never depend on these; only the split of work between the passes does)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -m gpu -k "outlier or knn or per_tile or config2 or config5 or chain" 2>&1 | tail -5
for far in 2.5 0 4; do
  for p in 1.0 0.8 1.25; do
    echo "== pitch $p rc_far $far"
    CWIPC_CUDA_KNN_PITCH=$p CWIPC_CUDA_KNN_RC_FAR=$far timeout 300 python scripts/bench_sor.py --reps 10 2>&1 | tail -1
  done
done
