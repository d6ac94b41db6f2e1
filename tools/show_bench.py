"""Condensed view of a bench.py JSON line: headline numbers and the per-kernel roofline table."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value %.4g ms/step %.1f e2e %.4g err %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("error")))
for k in d.get("roofline_kernels", []):
    print("   %s %.1f us %s frac %.3f share %.3f" % (k["kernel"], k["sec"] * 1e6, k["bound"], k["frac"], k["share_of_step"]))
for o in d.get("ot_resample", []) or []:
    print("   OT B=%d N=%d fwd %.0f us bwd %.0f us frac %.3f" % (o["B"], o["N"], o["fwd_us"], o["bwd_us"], o["frac_of_sfu_peak"]))
