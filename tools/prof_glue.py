"""Run the HBM-bound glue kernels of a timestep a few times at the headline shape (B = N = 1024) -- the short command ncu profiles."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from normalizing_flows_dpfs_b200 import ops

B = N = 1024
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = (torch.randn(B, N, 2, device=dev, generator=g) * 2).requires_grad_()
vel = torch.randn(B, 2, device=dev, generator=g)
rng = torch.tensor([1234, 0], dtype=torch.int64, device=dev)
gl = torch.randn(B, N, device=dev, generator=g)
gy = torch.randn(B, N, 2, device=dev, generator=g)
w = torch.softmax(torch.randn(B, N, device=dev, generator=g) * 3, -1).requires_grad_()
mk = torch.linspace(0.0, (N - 1.0) / N, N).to(dev)
off = torch.rand(B, device=dev, generator=g) / N
for it in range(3):
    ctx = torch.empty(B, 4, device=dev)
    moved, noise = ops.motion_moments(x, vel, None, ctx, 0, rng, 2.0)
    ops.row_moments(moved.detach(), torch.empty(B, 36, device=dev), 32, head=torch.randn(B, 32, device=dev, generator=g))
    lw = torch.log_softmax(torch.randn(B, N, device=dev, generator=g), -1).requires_grad_()
    lki = torch.randn(B, N, device=dev, generator=g).requires_grad_()
    logw, probs, rs, ess = ops.weight_update(lw, lki, gl, gl, 1e-12)
    torch.autograd.backward([logw, probs], [gl, gl])
    prior, prop = ops.proposal_terms(moved, moved.detach() + 0.1, noise, gl, gl, gl, 2.0)
    torch.autograd.backward([prior, prop], [gl, gl])
    p, w2, idx = ops.soft_resample(x, w, off, mk, 0.5)
    torch.autograd.backward([p, w2], [gy, gl])
torch.cuda.synchronize()
print("done")
