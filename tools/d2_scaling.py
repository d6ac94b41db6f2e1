"""Fixed cost of the D = 2 coupling backward: kernel durations (CUPTI) at N = 128 ... 1024 for B = 1024, C_row = 4 and 36."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nfdpf_oracle as O  # weights initialiser only
from normalizing_flows_dpfs_b200 import ops

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
B = 1024
for C in (4, 36):
    pk = O.init_stack(g, 2, C, std=0.1, bias_std=0.05).to(dev).requires_grad_()
    ctx = torch.randn(B, C, generator=g).to(dev)
    for N in (128, 256, 512, 1024):
        x = (torch.randn(B, N, 2, generator=g) * 2).to(dev).requires_grad_()
        gy, gl = torch.randn(B, N, 2, generator=g).to(dev), torch.randn(B, N, generator=g).to(dev)
        for _ in range(3):
            y, ld = ops.coupling_stack(pk, x, ctx, None, 2, True)
            torch.autograd.backward([y, ld], [gy, gl])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(5):
                y, ld = ops.coupling_stack(pk, x, ctx, None, 2, True)
                torch.autograd.backward([y, ld], [gy, gl])
            torch.cuda.synchronize()
        d = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                k = "bwd" if "coupling_bwd_d2" in e.name else ("fwd" if "coupling_fwd" in e.name else ("reduce" if "d2_reduce" in e.name else None))
                if k:
                    d.setdefault(k, []).append(e.time_range.elapsed_us())
        print("C_row=%2d N=%4d  fwd %.1f us  bwd %.1f us  reduce %.1f us" % (C, N, min(d["fwd"]), min(d["bwd"]), min(d["reduce"])))
