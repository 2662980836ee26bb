set -x
for st in downsample outliers; do for w in 4 8 16; do
  BENCH_STAGE=$st timeout 300 python bench.py --steps 5 --warmup 3 --workers $w --cpu-frames 1 > gpurun_out/stage_${st}_$w.json 2> gpurun_out/stage.err; echo exit=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/stage_${st}_$w.json"))
print("stage $st workers $w value",d["value"],"ms_per_step",d["ms_per_step"])
PY
done; done
for w in 12 16; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workers $w --cpu-frames 1 > gpurun_out/both_$w.json 2> gpurun_out/stage.err
  python - <<PY
import json
d=json.load(open("gpurun_out/both_$w.json"))
print("both workers $w value",d["value"],"e2e",d["e2e"]["value"],"ms_per_step",d["ms_per_step"])
PY
done
