#!/bin/bash
# kNN second-scan sweep: kNN kernel timings (alone) and the bench value (results never depend on it)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_configs.py -x -q -m gpu -k "outlier or knn or per_tile or config2 or config5 or chain" 2>&1 | tail -3
for cfg in "2.0 2 150" "2.0 2 80" "2.0 1 250" "2.5 2 150"; do
  set -- $cfg
  echo "== rc_far $1 second_min $2 max $3"
  CWIPC_CUDA_KNN_RC_FAR=$1 CWIPC_CUDA_KNN_SECOND_MIN=$2 CWIPC_CUDA_KNN_SECOND_MAX=$3 timeout 300 python scripts/bench_sor.py --reps 10 2>&1 | tail -1 | cut -c 1-1300
  CWIPC_CUDA_KNN_RC_FAR=$1 CWIPC_CUDA_KNN_SECOND_MIN=$2 CWIPC_CUDA_KNN_SECOND_MAX=$3 python bench.py --skip-config4 --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', d['value'], 'e2e', d['e2e']['value'], {k:(v['us_per_launch']) for k,v in d['roofline']['kernels'].items() if k.startswith('knn')})"
done
