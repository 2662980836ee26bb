set -x
for w in 1 2 4 8 16; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workers $w --cpu-frames 1 > gpurun_out/sweep_w$w.json 2> gpurun_out/sweep_w$w.err; echo exit=$?
  python - <<PY
import json
d=json.load(open("gpurun_out/sweep_w$w.json"))
print("workers",$w,"value",d["value"],"e2e",d["e2e"]["value"],"ms_per_step",d["ms_per_step"])
PY
done
