"""Time one OT resample (forward, and backward) at the BASELINE shapes with CUDA events; report iterations used."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from normalizing_flows_dpfs_b200 import ops

shapes = [(1024, 1024), (256, 4096), (32, 100)] if len(sys.argv) < 3 else [(int(sys.argv[1]), int(sys.argv[2]))]
for B, N in shapes:
    g = torch.Generator().manual_seed(0)
    w = torch.softmax(torch.randn(B, N, generator=g) * 2, -1).cuda()
    x = (torch.randn(B, N, 2, generator=g) * 20).cuda().requires_grad_()
    gy = torch.randn(B, N, 2, generator=g).cuda()
    lw = w.log()
    for _ in range(2):
        p = ops.ot_resample(x, lw)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    p = ops.ot_resample(x, lw)
    e[1].record()
    p.backward(gy)
    e[2].record()
    torch.cuda.synchronize()
    it = int(ops.OtResample.last_iters.item())
    pairs = B * N * N * (2 * (it - 2) + 2 + 2 + 1)   # useful pair evaluations, SURVEY 8(d): loop + init + final + transport
    print("B=%d N=%d iters=%d fwd %.1f us  bwd %.1f us  pair-evals %.3e -> %.2f T pair-evals/s" %
          (B, N, it, e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3, pairs, pairs / (e[0].elapsed_time(e[1]) * 1e-3) / 1e12))
