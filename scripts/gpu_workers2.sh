for w in 8 12 16 24 32; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workers $w --cpu-frames 1 --frames-per-gpu 96 > gpurun_out/w2.json 2> gpurun_out/w2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/w2.json"))
print("workers $w frames 96 value",d["value"],"e2e",d["e2e"]["value"],"ms_per_step",d["ms_per_step"])
PY
done
