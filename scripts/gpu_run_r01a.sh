set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo bench_exit=$?
tail -3 gpurun_out/bench_r01.err; cat gpurun_out/bench_r01.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2> gpurun_out/bench_ref_r01.err; echo ref_exit=$?; cat gpurun_out/bench_ref_r01.json
python bench.py --steps 1 --warmup 3 --frames-per-gpu 4 --workers 1 --cpu-frames 1 > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --frames-per-gpu 4 --workers 1 --cpu-frames 1 > gpurun_out/ncu_launches.log 2>&1; echo ncu1_exit=$?
ncu --set full --clock-control none --import-source on -k regex:'knn_cell_kernel|radix_onesweep_kernel|voxel_reduce_kernel|voxel_keygen_kernel|compact_kernel|knn_far_kernel' -s 60 -c 24 -o gpurun_out/prof_r01 python bench.py --steps 1 --warmup 3 --frames-per-gpu 4 --workers 1 --cpu-frames 1 > gpurun_out/ncu_full.log 2>&1; echo ncu2_exit=$?
ls -la gpurun_out
