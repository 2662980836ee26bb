#!/usr/bin/env python
"""Top source lines of one kernel by warp-stall samples, from an ncu report captured with
--set full --import-source on (binary built with -lineinfo).

  python profiles/hot_lines.py gpurun_out/prof.ncu-rep knn_tile [N]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, hdr, acc = None, None, {}
first_kernel = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        if first_kernel is None:
            first_kernel = r[1]
        cur_kernel = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue  # SASS rows carry an address; source-line rows have "-"
    if cur_kernel != first_kernel:
        continue  # only the first captured launch of the kernel
    d = dict(zip(hdr[4:], r[4:]))
    try:
        samples = int(d["# Samples"])
        inst = int(d["Instructions Executed"])
    except (KeyError, ValueError):
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    key = (fname, r[0])
    a = acc.setdefault(key, [0, 0, r[1].strip()[:100], {}])
    a[0] += samples
    a[1] += inst
    for k, v in stalls.items():
        a[3][k] = a[3].get(k, 0) + v
tot = sum(a[0] for a in acc.values()) or 1
toti = sum(a[1] for a in acc.values()) or 1
print(f"# {first_kernel[:90]}\n# total samples {tot}, warp instructions {toti}")
print("samples  %smp     inst  %inst  file:line  top-stalls | source")
for (f, ln), (s, i, src, st) in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ",".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{s:7d} {100*s/tot:5.1f} {i:9d} {100*i/toti:5.1f}  {f}:{ln}  [{tops}] | {src}")
