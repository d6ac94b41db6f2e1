"""True per-kernel device time INSIDE one CUDA-graph replay of the headline training step (torch.profiler / CUPTI: kernels are not
serialised or cache-flushed as under ncu), plus the idle time between kernels.  Usage: python tools/graph_breakdown.py [out.json]"""
import json
import os
import sys
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from normalizing_flows_dpfs_b200.graphs import GraphedFilterStep


class A:
    B, N, T, measurement, resampler = 1024, 1024, 50, os.environ.get("MEAS", "gaussian"), "soft"


dev = torch.device("cuda")
a = A()
if a.measurement == "CRNVP":
    from bench_configs import _build
    dpf = _build(["--NF-dyn", "--NF-cond", "--measurement", "CRNVP", "--resampler_type", "soft"], a.B, a.N, a.T, dev, cnf=True)
    dpf.force_resample = True
else:
    dpf = bench.build_b200(a, dev)
    dpf.rng_device = "cuda"
host = bench.synth_batch(a.B, a.T, a.N, 100, pinned=False)
host.pop("noise"), host.pop("offsets")
res = {k: v.to(dev) for k, v in host.items()}
g = GraphedFilterStep(dpf, res)
for _ in range(3):
    g.run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g.run()
    torch.cuda.synchronize()
ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
agg = defaultdict(lambda: [0, 0.0])
for e in ev:
    agg[e.name][0] += 1
    agg[e.name][1] += e.time_range.elapsed_us()
span = ev[-1].time_range.end - ev[0].time_range.start
busy = sum(v[1] for v in agg.values())
rows = [{"kernel": k[:100], "launches": v[0], "total_us": round(v[1], 1), "share_of_span": round(v[1] / span, 4), "avg_us": round(v[1] / v[0], 2)}
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
out = {"span_us": span, "busy_us": round(busy, 1), "idle_share": round(1 - busy / span, 4), "n_kernels": len(ev), "kernels": rows}
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
print("span %.1f us, kernels busy %.1f us, idle %.2f %%, %d kernels" % (span, busy, 100 * (1 - busy / span), len(ev)))
for r in rows[:28]:
    print("%-70s n=%4d total=%9.1f us (%.3f) avg=%7.2f" % (r["kernel"][:70], r["launches"], r["total_us"], r["share_of_span"], r["avg_us"]))
