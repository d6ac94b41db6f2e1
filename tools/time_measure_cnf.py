"""Time the CRNVP measurement forward / backward ops at B = N = 1024 with CUDA events (NFDPF_MEASURE_BWD_V1=1: the single-role kernel)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nfdpf_oracle as O  # weights initialiser only
from normalizing_flows_dpfs_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, N, 2, generator=g) * 2).to(dev).requires_grad_()
gl = torch.randn(B, N, generator=g).to(dev)
pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))]).to(dev).requires_grad_()
enc = torch.randn(B, 32, generator=g).to(dev)
w = torch.softmax(torch.randn(B, N, generator=g), -1).to(dev)
cnf = O.init_stack(g, 32, 32, std=0.1, bias_std=0.05).to(dev).requires_grad_()
fwd, bwd = [], []
for it in range(8):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    out = ops.measure_update(pe, cnf, enc, x, w.log(), gl, gl, "CRNVP", p0=0.0, p1=2.5)
    e[1].record()
    torch.autograd.backward([out[0], out[2]], [gl, gl])
    e[2].record()
    torch.cuda.synchronize()
    fwd.append(e[0].elapsed_time(e[1]))
    bwd.append(e[1].elapsed_time(e[2]))
print(json.dumps({"B": B, "N": N, "v1": bool(os.environ.get("NFDPF_MEASURE_BWD_V1")), "fwd_ms": min(fwd), "bwd_ms": min(bwd),
                  "grad_checksum": [float(pe.grad.double().sum()), float(cnf.grad.double().sum()), float(x.grad.double().abs().sum())]}))
