"""Run each hot kernel a few times at the headline shape (B = N = 1024) -- the short command ncu profiles."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nfdpf_oracle as O  # weights initialiser only
from normalizing_flows_dpfs_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = N = 1024
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
x = (torch.randn(B, N, 2, generator=g) * 2).to(dev).requires_grad_()
ctx = torch.randn(B, 36, generator=g).to(dev)
pk = O.init_stack(g, 2, 36, std=0.1, bias_std=0.05).to(dev).requires_grad_()
gy, gl = torch.randn(B, N, 2, generator=g).to(dev), torch.randn(B, N, generator=g).to(dev)
for it in range(3):
    if which in ("all", "coupling"):
        y, ld = ops.coupling_stack(pk, x, ctx, None, 2, True)
        torch.autograd.backward([y, ld], [gy, gl])
    if which in ("all", "measure", "measure_cnf"):
        pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))]).to(dev).requires_grad_()
        enc = torch.randn(B, 32, generator=g).to(dev)
        w = torch.softmax(torch.randn(B, N, generator=g), -1).to(dev)
        if which == "measure_cnf":
            cnf = O.init_stack(g, 32, 32, std=0.1, bias_std=0.05).to(dev).requires_grad_()
            out = ops.measure_update(pe, cnf, enc, x, w.log(), gl, gl, "CRNVP", p0=0.0, p1=2.5)
        else:
            out = ops.measure_update(pe, None, enc, x, w.log(), gl, gl, "gaussian", p0=1.0, p1=10.0)
        torch.autograd.backward([out[0], out[2]], [gl, gl])
    if which in ("all", "soft"):
        w = torch.softmax(torch.randn(B, N, generator=g) * 3, -1).to(dev).requires_grad_()
        mk = torch.linspace(0.0, (N - 1.0) / N, N).to(dev)
        off = (torch.rand(B, generator=g) / N).to(dev)
        p, w2, idx = ops.soft_resample(x, w, off, mk, 0.5)
        torch.autograd.backward([p, w2], [gy, gl])
    if which in ("all", "ot"):
        w = torch.softmax(torch.randn(B, N, generator=g) * 2, -1).to(dev)
        p = ops.ot_resample(x.detach() * 10, w.log())
torch.cuda.synchronize()
print("done", which)
