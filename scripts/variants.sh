#!/bin/bash
# kNN grid resolution sweep (dense table pyramid: cells per point / absolute cap); results never depend on it
for cfg in "4 25" "16 27" "64 28" "256 29"; do
  set -- $cfg
  echo "== cells_per_point $1 max_log2 $2"
  CWIPC_CUDA_KNN_CELLS_PER_POINT=$1 CWIPC_CUDA_KNN_MAX_CELLS_LOG2=$2 timeout 300 python scripts/bench_sor.py --reps 5 2>&1 | tail -1 | cut -c 1-1500
  CWIPC_CUDA_KNN_CELLS_PER_POINT=$1 CWIPC_CUDA_KNN_MAX_CELLS_LOG2=$2 timeout 300 python scripts/diag_sor8m.py 2>&1 | grep "^plain"
done
