"""Print the interesting parts of a bench.py JSON line: python tools/show_bench.py file.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("  value %.4g ms/step %.1f e2e %.4g err %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline_error")))
dp = d.get("e2e_default_path")
if dp:
    print("  default path:", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in dp.items() if k != "gate"})
for r in d.get("roofline_kernels", []):
    print("   %-30s %7.1f us  %-6s frac %.3f share %.3f" % (r["kernel"], r["sec"] * 1e6, r["bound"], r["frac"], r["share_of_step"]))
for o in d.get("ot_resample", []) if isinstance(d.get("ot_resample"), list) else []:
    print("   OT B=%d N=%d fwd %.0f us bwd %.0f us iters %d frac %.3f" % (o["B"], o["N"], o["fwd_us"], o["bwd_us"], o["sinkhorn_iterations"], o["frac_of_sfu_peak"]))
for k, c in (d.get("configs") or {}).items():
    print("   config %s: %s" % (k, {kk: (round(v, 2) if isinstance(v, float) else v) for kk, v in c.items() if kk not in ("workload", "unit", "gate")} if isinstance(c, dict) else c))
print("  roofline:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (d.get("roofline") or {}).items() if k != "peak_source"})
print("  cpu_baseline:", d.get("cpu_baseline"))
