#!/usr/bin/env python
"""Summarise ncu output for profiles/ (run here, no GPU needed).

  python profiles/summarize_ncu.py raw   gpurun_out/prof.ncu-rep    > profiles/rNN_full_summary.csv
  python profiles/summarize_ncu.py list  gpurun_out/launches.csv    > profiles/rNN_launches_summary.csv

`raw`  : one line per profiled launch with the metrics the roofline needs (duration, DRAM bytes read/written,
         DRAM/L2/SM throughput %, warps active, registers, warp instructions).
`list` : the per-launch gpu__time_duration list aggregated per kernel (count, total, mean, share of the total).
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

WANT = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_inst"),
]


def short(name):
    import re
    flat = name.replace("(anonymous namespace)", "anon")
    m = re.search(r"([A-Za-z_]\w*)\s*(<[^(]*>)?\s*\(", flat) or re.search(r"([A-Za-z_]\w*)\s*[<(]", flat)
    if not m:
        return name
    targs = re.sub(r"\(int\)|\(bool\)|cwcu::|anon::|<unnamed>::", "", m.group(2) or "") if m.lastindex and m.lastindex >= 2 else ""
    return m.group(1) + targs


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), label, units[hdr.index(m)]) for m, label in WANT if m in hdr]
    kn = hdr.index("Kernel Name")
    w = csv.writer(sys.stdout)
    w.writerow(["kernel"] + [f"{label}[{unit}]" if unit else label for _, label, unit in cols])
    for r in rows[2:]:
        vals = []
        for i, label, unit in cols:
            v = r[i]
            try:
                v = f"{float(v.replace(',', '')):.6g}"
            except ValueError:
                pass
            vals.append(v)
        w.writerow([short(r[kn])] + vals)


def launches(path):
    text = open(path).read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        t = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        t *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(a[1] for a in agg.values())
    w = csv.writer(sys.stdout)
    w.writerow(["kernel", "launches", "total_us", "mean_us", "share"])
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, c, f"{t:.1f}", f"{t / c:.2f}", f"{t / total:.3f}"])
    w.writerow(["TOTAL", sum(a[0] for a in agg.values()), f"{total:.1f}", "", "1.000"])


if __name__ == "__main__":
    {"raw": raw, "list": launches}[sys.argv[1]](sys.argv[2])
