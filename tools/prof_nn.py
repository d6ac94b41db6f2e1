"""Run the NN-likelihood kernels (mode 3) and the block pseudo-likelihood kernels a few times at the headline shape -- the short command
ncu profiles (ncu --set full -k regex:"measure_fwd_kernel<3>|measure_bwd_nn|block_density" ...)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from normalizing_flows_dpfs_b200 import ops

B = N = 1024
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))]).to(dev).requires_grad_()
head = torch.cat([torch.randn(n, generator=g) * s for n, s in ((4096, 0.15), (64, 0.1), (4096, 0.15), (64, 0.1), (64, 0.3), (1, 0.1))]).to(dev).requires_grad_()
enc = torch.randn(B, 32, generator=g).to(dev)
x = (torch.randn(B, N, 2, generator=g) * 3).to(dev).requires_grad_()
lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1).to(dev)
gl = torch.randn(B, N, generator=g).to(dev)
T, bl = 20, 10
w = torch.softmax(torch.randn(T, B, N, generator=g), -1).to(dev)
lik, prior = torch.randn(T, B, N, generator=g).to(dev), torch.randn(T, B, N, generator=g).to(dev)
idx = (torch.sort(torch.randint(0, N, (T, B, N), generator=g), dim=-1).values + N * torch.arange(B)[None, :, None]).to(dev)
lists = [t.transpose(0, 1).requires_grad_() for t in (w, lik, prior)]
for it in range(3):
    out = ops.measure_update(pe, head, enc, x, lw0, gl, gl, "NN")
    torch.autograd.backward([out[0], out[2]], [gl, gl])
    ops.block_density(lists[0], lists[1], lists[2], idx.transpose(0, 1), bl).sum().backward()
torch.cuda.synchronize()
