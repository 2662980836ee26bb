#!/usr/bin/env python
"""Blackwell evidence from the built library: per kernel, the counts of the SASS mnemonics that prove what the source claims
(TMA bulk copy + mbarrier, cp.async, warp match/vote, L2 reductions without return, cluster barriers, distributed shared
memory).  Runs `cuobjdump -sass` on cwipc_util_b200/lib/libcwipc_util_cuda.so (no GPU needed); writes profiles/sass_summary.txt.

    python scripts/sass_summary.py
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "cwipc_util_b200", "lib", "libcwipc_util_cuda.so")
OUT = os.path.join(REPO, "profiles", "sass_summary.txt")

PATTERNS = [
    ("UBLKCP", r"\bUBLKCP"),                      # cp.async.bulk (TMA 1-D bulk copy)
    ("SYNCS", r"\bSYNCS\."),                      # mbarrier arrive / try_wait
    ("LDGSTS", r"\bLDGSTS"),                      # cp.async (global -> shared without registers)
    ("LDGDEPBAR/DEPBAR", r"\b(LDGDEPBAR|DEPBAR)"),
    ("MATCH", r"\bMATCH\."),                      # __match_any_sync
    ("VOTE", r"\bVOTE\."),                        # __ballot_sync / __any_sync
    ("SHFL", r"\bSHFL\."),
    ("REDG/RED", r"\bRED(G)?\.E"),                # fire-and-forget L2 reductions
    ("ATOMG/ATOM", r"\bATOM(G)?\.E"),             # atomics with a result (CAS, counters)
    ("ATOMS", r"\bATOMS\."),
    ("UCGABAR", r"\bUCGABAR"),                    # cluster barrier
    ("ST/LD .cluster (DSMEM)", r"\b(STS|LDS|ST|LD)\.[A-Z0-9.]*CLUSTER|MAPA"),
    ("LDG.128/STG.128", r"\b(LDG|STG)\.E(\.[A-Z0-9]+)*\.128"),
    ("DADD/DMUL/DFMA", r"\b(DADD|DMUL|DFMA)\b"),
    ("HMMA/UTCHMMA (tensor cores)", r"\b(HMMA|UTC[A-Z]*MMA|IMMA|QMMA)"),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            d = re.sub(r"\(anonymous namespace\)::", "", d)
            d = re.sub(r"^void ", "", d)
            d = re.sub(r"cwcu::", "", d)
            cur = re.sub(r"\(.*", "", d)
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            continue
        per[cur]["instructions"] += 1
        for label, pat in PATTERNS:
            if re.search(pat, line):
                per[cur][label] += 1
    labels = [l for l, _ in PATTERNS]
    with open(OUT, "w") as f:
        f.write("# SASS evidence per kernel of libcwipc_util_cuda.so (cuobjdump -sass; cubins: %s)\n" % ", ".join(archs))
        f.write("# regenerate: python scripts/sass_summary.py   (scripts/gpu_final.sh does)\n")
        f.write("# columns: static instruction counts of the mnemonics; 0 tensor-core instructions by design (no dense contraction on this path)\n\n")
        tot = collections.Counter()
        for k, c in per.items():
            tot.update(c)
        f.write("TOTAL: " + ", ".join(f"{l} {tot[l]}" for l in ["instructions"] + labels if tot[l] or l.startswith("HMMA")) + "\n\n")
        for k, c in sorted(per.items(), key=lambda kv: -kv[1]["instructions"]):
            hits = ", ".join(f"{l} {c[l]}" for l in labels if c[l])
            f.write(f"{k}: {c['instructions']} instr; {hits}\n")
    print(open(OUT).read()[:3000])


if __name__ == "__main__":
    main()
