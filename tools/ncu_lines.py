"""Samples per CUDA source line of one kernel from `ncu -i rep --page source --csv --print-source cuda,sass`.
Usage: python tools/ncu_lines.py view.csv [file-substring] [top-n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur, hdr = None, None
per = defaultdict(lambda: [0, 0, ""])
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 8 and r[0].isdigit() and r[2] in ("", "-"):
        # a CUDA source line row (no SASS address): its counters aggregate the SASS rows below it
        s = int(r[hdr["# Samples"]] or 0)
        ie = int(r[hdr["Instructions Executed"]] or 0)
        tot += s
        k = (cur.split("/")[-1], int(r[0]))
        per[k][0] += s
        per[k][1] += ie
        per[k][2] = r[1][:90]
print("total samples", tot)
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:topn]:
    if want in k[0]:
        print("%-14s %5d  s=%6d (%.3f) inst=%9d  %s" % (k[0], k[1], v[0], v[0] / max(tot, 1), v[1], v[2].strip()))
