#!/bin/bash
# ncu comparison of knn_tile_kernel between library builds (variants/*.so)
CMD="python bench.py --steps 1 --warmup 3 --frames-per-gpu 4 --passes 1 --workers 1 --cpu-frames 1 --skip-config4"
for v in B C; do
  CWIPC_CUDA_KNN_RC_FAR=0 CWIPC_CUDA_LIBRARY=$PWD/variants/$v.so ncu --set full --clock-control none --import-source on -k regex:knn_tile -s 4 -c 2 -f -o gpurun_out/prof_knnvar_$v $CMD > gpurun_out/ncu_knnvar_$v.log 2>&1; echo "$v exit=$?"
done
CWIPC_CUDA_KNN_RC_FAR=2.0 CWIPC_CUDA_LIBRARY=$PWD/variants/C.so ncu --set full --clock-control none --import-source on -k regex:"knn_tile|knn_far" -s 8 -c 4 -f -o gpurun_out/prof_knnvar_C20 $CMD > gpurun_out/ncu_knnvar_C20.log 2>&1; echo "C20 exit=$?"
CWIPC_CUDA_KNN_RC_FAR=0 CWIPC_CUDA_LIBRARY=$PWD/variants/B.so ncu --set full --clock-control none --import-source on -k regex:"knn_far" -s 4 -c 2 -f -o gpurun_out/prof_knnvar_Bfar $CMD > gpurun_out/ncu_knnvar_Bfar.log 2>&1; echo "Bfar exit=$?"
