"""Time the gaussian measurement forward / backward ops at B = N = 1024 with CUDA events."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from normalizing_flows_dpfs_b200 import ops

B = N = 1024
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, N, 2, generator=g) * 2).to(dev).requires_grad_()
gl = torch.randn(B, N, generator=g).to(dev)
pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))]).to(dev).requires_grad_()
enc = torch.randn(B, 32, generator=g).to(dev)
w = torch.softmax(torch.randn(B, N, generator=g), -1).to(dev)
fwd, bwd = [], []
for it in range(10):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = ops.measure_update(pe, None, enc, x, w.log(), gl, gl, "gaussian", p0=1.0, p1=10.0)
    e[1].record()
    torch.autograd.backward([out[0], out[2]], [gl, gl])
    e[2].record()
    torch.cuda.synchronize()
    fwd.append(e[0].elapsed_time(e[1]))
    bwd.append(e[1].elapsed_time(e[2]))
print(json.dumps({"fwd_ms": min(fwd), "bwd_ms": min(bwd)}))
