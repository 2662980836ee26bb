#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of the bench frames (scripts/gpu_ncu.sh): per kernel, the mean over
the captured launches of dram__bytes_read.sum + dram__bytes_write.sum and of smsp__inst_executed.sum.  bench.py reads it for
`roofline.traffic` (a static figure: the bench itself never runs under a profiler).

    python profiles/make_traffic.py gpurun_out/prof_<tag>.ncu-rep <tag>
"""
import csv
import re
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))


def short(name):
    m = re.search(r"([A-Za-z_]\w*)\s*[<(]", name.replace("(anonymous namespace)", "anon"))
    return m.group(1) if m else name


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    cols = {m: hdr.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum")}
    acc = OrderedDict()
    for r in rows[2:]:
        k = short(r[kn])
        rd = float(r[cols["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units[cols["dram__bytes_read.sum"]], 1)
        wr = float(r[cols["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units[cols["dram__bytes_write.sum"]], 1)
        wi = float(r[cols["smsp__inst_executed.sum"]].replace(",", ""))
        a = acc.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += rd + wr
        a[2] += wi
    traffic = OrderedDict((k, int(a[1] / a[0])) for k, a in acc.items())
    traffic["_warp_instructions"] = OrderedDict((k, int(a[2] / a[0])) for k, a in acc.items())
    traffic["_launches_captured"] = OrderedDict((k, a[0]) for k, a in acc.items())
    traffic["_source"] = f"profiles/{tag}_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum per launch (mean), ncu --set full (cold L2), 1M-point frames of the bench workload"
    json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
