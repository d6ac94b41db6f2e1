"""Time the D = 2 coupling-stack kernels alone at the headline shape (CUDA events, graph replay) and check the backward
against torch autograd of a plain fp64 restatement on a slice.  NFDPF_D2_CFG selects the backward geometry."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nfdpf_oracle as O  # weights initialiser + reference math (checker only)
from normalizing_flows_dpfs_b200 import ops
from bench_extras import _events

B = int(os.environ.get("B", 1024))
N = int(os.environ.get("N", 1024))
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
out = {"cfg": os.environ.get("NFDPF_D2_CFG", "0"), "B": B, "N": N}
for C in (4, 36):
    for inverse in (True, False):
        x = (torch.randn(B, N, 2, generator=g) * 2).to(dev).requires_grad_()
        ctx = torch.randn(B, C, generator=g).to(dev)
        pk = O.init_stack(g, 2, C, std=0.1, bias_std=0.05).to(dev).requires_grad_()
        gy, gl = torch.randn(B, N, 2, generator=g).to(dev), torch.randn(B, N, generator=g).to(dev)
        y, ld = ops.coupling_stack(pk, x, ctx, None, 2, inverse)
        tf = _events(lambda: ops.coupling_stack(pk, x, ctx, None, 2, inverse), n=10)

        def bwd():
            torch.autograd.backward([y, ld], [gy, gl], retain_graph=True)
        tb = _events(bwd, n=10)
        # parity on the first 4 trajectories against the oracle's autograd (fp32 CPU)
        x.grad = pk.grad = None
        torch.autograd.backward([y, ld], [gy, gl], retain_graph=True)
        nb = 4
        xo = x.detach().cpu()[:nb].clone().requires_grad_()
        pko = pk.detach().cpu().clone().requires_grad_()
        W = O.unpack_stack(pko, 2, C)
        ctx_rows = ctx.cpu()[:nb, None, :].expand(nb, N, C).reshape(nb * N, C)
        fn = O.stack_inverse if inverse else O.stack_forward
        yo, ldo = fn(xo.reshape(nb * N, 2), ctx_rows, W)
        torch.autograd.backward([yo, ldo], [gy.cpu()[:nb].reshape(-1, 2), gl.cpu()[:nb].reshape(-1)])
        dx_err = (x.grad.cpu()[:nb] - xo.grad).abs().max().item() / xo.grad.abs().max().item()
        out["C%d_%s" % (C, "inv" if inverse else "fwd")] = {"fwd_us": tf * 1e6, "bwd_us": tb * 1e6, "dx_relmax_err": dx_err}
print(json.dumps(out))
