TAG=${1:-r02v2}
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn or outliers" > gpurun_out/pytest_${TAG}.log 2>&1; echo test_exit=$?
tail -2 gpurun_out/pytest_${TAG}.log
bash scripts/gpu_ab_variants.sh ${TAG} > /dev/null 2>&1
OUT=gpurun_out/ab_${TAG}.txt
for leaf in 64 256; do
  echo "== v2 leaf $leaf" >> $OUT
  CWIPC_CUDA_KNN_LEAF=$leaf timeout 300 python scripts/ab_value.py --workers 30 --profile --tag "leaf $leaf" 2>> gpurun_out/ab_${TAG}.err | tail -1 | cut -c 1-900 >> $OUT
done
cat $OUT
