"""Stall / opcode breakdown and hottest SASS lines of one kernel from `ncu -i rep --page source --csv --kernel-name ...`."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
I = lambda r, k: int(r[ix[k]] or 0)
tot = sum(I(r, "# Samples") for r in data)
inst = sum(I(r, "Instructions Executed") for r in data)
print("samples", tot, "warp-instr", inst, "sass lines", len(data))
keys = [h for h in hdr if h.startswith("stall_") and "(Not" not in h]
for k, v in sorted(((k, sum(I(r, k) for r in data)) for k in keys), key=lambda x: -x[1])[:9]:
    print("  %-24s %.3f" % (k, v / tot))
c, cs = Counter(), Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += I(r, "Instructions Executed")
    cs[op] += I(r, "# Samples")
for op, v in c.most_common(14):
    print("  %-8s exec %.3f samples %.3f" % (op, v / inst, cs[op] / tot))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
for i in sorted(range(len(data)), key=lambda i: -I(data[i], "# Samples"))[:n]:
    r = data[i]
    print("  L%4d %-58s s=%4d wait=%d ssb=%d lsb=%d mio=%d bar=%d" % (i, r[ix["Source"]][:58], I(r, "# Samples"), I(r, "stall_wait"), I(r, "stall_short_sb"), I(r, "stall_long_sb"), I(r, "stall_mio"), I(r, "stall_barrier")))
for i in sorted(range(len(data)), key=lambda i: -I(data[i], "L1 Wavefronts Shared Excessive"))[:6]:
    r = data[i]
    print("  BANK L%4d %-50s excessive=%d ideal=%d" % (i, r[ix["Source"]][:50], I(r, "L1 Wavefronts Shared Excessive"), I(r, "L1 Wavefronts Shared Ideal")))
