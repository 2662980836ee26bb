OUT=gpurun_out/ab_r02z6.txt
: > $OUT
run() { label=$1; shift; echo "== $label" >> $OUT; ( env "$@" timeout 200 python scripts/ab_value.py --workers 30 --frames 60 --steps 5 --tag "$label" 2>> gpurun_out/ab_r02z6.err | tail -1 | cut -c 1-300 ) >> $OUT; }
run "tiles per warp 1 (default)" X=1
run "tiles per warp 2" CWIPC_CUDA_DS_TILES_PER_WARP=2
run "tiles per warp 1 again" X=1
run "tiles per warp 2 again" CWIPC_CUDA_DS_TILES_PER_WARP=2
cat $OUT
