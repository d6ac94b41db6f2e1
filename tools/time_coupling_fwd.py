"""Time the D = 2 forward coupling stacks (C_row = 4 and 36, forward and inverse) at B = N = 1024: 200 back-to-back launches each."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nfdpf_oracle as O  # weights initialiser only
from normalizing_flows_dpfs_b200 import ops

B = N = 1024
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, N, 2, generator=g) * 2).to(dev)
res = {}
for C in (4, 36):
    ctx = torch.randn(B, C, generator=g).to(dev)
    pk = O.init_stack(g, 2, C, std=0.1, bias_std=0.05).to(dev)
    for inv in (False, True):
        with torch.no_grad():
            for _ in range(20):
                ops.coupling_stack(pk, x, ctx, None, 2, inv)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                ops.coupling_stack(pk, x, ctx, None, 2, inv)
            e1.record()
            torch.cuda.synchronize()
        res["C%d_%s_us" % (C, "inv" if inv else "fwd")] = round(e0.elapsed_time(e1) * 1e3 / 200, 2)
print(json.dumps(res))
