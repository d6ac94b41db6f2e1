"""Where do the remaining ATen copy kernels of an eager training step come from?  Wraps the library's tensor-normalising helper
(_lib.f32) and reports every call that had to copy (non-contiguous or non-fp32 input) with its call site."""
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synth_batch
from bench_configs import _build
from normalizing_flows_dpfs_b200 import _lib as L
from normalizing_flows_dpfs_b200.losses import supervised_loss

dev = torch.device("cuda")
B, N, T = 64, 256, 3
dpf = _build(["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft"], B, N, T, dev)
dpf.force_resample = True
host = synth_batch(B, T, N, 300, pinned=False)
host.pop("noise"), host.pop("offsets")
d = {k: v.to(dev) for k, v in host.items()}
seen = {}
orig = L.f32


def f32(t):
    if t is not None and (not t.is_contiguous() or t.dtype != torch.float32):
        fr = traceback.extract_stack(limit=4)[:-1]
        key = " <- ".join("%s:%d" % (os.path.basename(f.filename), f.lineno) for f in reversed(fr)) + "  shape %s stride %s" % (tuple(t.shape), t.stride())
        seen[key] = seen.get(key, 0) + 1
    return orig(t)


L.f32 = f32
import normalizing_flows_dpfs_b200.ops as ops_mod
ops_mod.L.f32 = f32
out = dpf.filtering_pos(d["enc"], d["start"], d["vel_in"])
loss, _ = supervised_loss(out[0], out[1], d["state"], 1.0, False)
loss.backward()
torch.cuda.synchronize()
for k, n in sorted(seen.items(), key=lambda kv: -kv[1]):
    print("%3d  %s" % (n, k))
print("T =", T)
