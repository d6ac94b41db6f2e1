"""Reproduce iteration `it` of tests/test_gpu_crnvp.py::test_crnvp_many_trajectory_randomized_parity under NFDPF_TEST_SEED=base and
locate the parameter-gradient mismatch: which trajectory / particle, and whether it sits on a ReLU kink or an argmax tie."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
from test_gpu_ops import _off_the_argmax_ties, _off_the_relu_kinks, _pe_tuple, cu

base, it = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(9000 + 131 * it + base)
B = int(torch.randint(300, 700, (1,), generator=g))
N = int(torch.randint(40, 140, (1,), generator=g))
pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
cnf = O.init_stack(g, 32, 32, std=0.1, bias_std=0.05)
enc = torch.randn(B, 32, generator=g)
x = _off_the_relu_kinks(torch.randn(B, N, 2, generator=g) * 3, pe)
x = _off_the_argmax_ties(x, lambda xx: O.measurement_cnf(enc, xx, _pe_tuple(pe), O.unpack_stack(cnf, 32, 32), 2.5))
lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
g1, g2, g3 = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g), torch.randn(B, generator=g)
lo = [t.clone().requires_grad_() for t in (pe, cnf, enc, x, lw0, prior, prop)]
lki_o = O.measurement_cnf(lo[2], lo[3], _pe_tuple(lo[0]), O.unpack_stack(lo[1], 32, 32), 2.5)
lw = lo[4] + lki_o + lo[5] - lo[6]
pr = O.normalize_log_probs(lw) + 1e-12
((lki_o * g1).sum() + (pr * g2).sum() * 50 + (lw.sum(-1) * g3).sum() * 0.01).backward()
gt = [cu(t).requires_grad_() for t in (pe, cnf, enc, x, lw0, prior, prop)]
lki, logw, probs, rs, ess = ops.measure_update(*gt, "CRNVP", p0=0.0, p1=2.5)
((lki * cu(g1)).sum() + (probs * cu(g2)).sum() * 50 + (rs * cu(g3)).sum() * 0.01).backward()
print("B, N =", B, N)
dx = (gt[3].grad.cpu() - lo[3].grad).abs().amax(-1)            # (B, N)
b, n = np.unravel_index(int(dx.argmax()), dx.shape)
print("worst d_x at trajectory %d particle %d: %.3e (typical max %.3e)" % (b, n, float(dx.max()), float(lo[3].grad.abs().max())))
print("d_pe max diff %.3e, d_cnf %.3e, d_enc %.3e" % tuple(float((a.grad.cpu() - r.grad).abs().max()) for a, r in zip(gt[:3], lo[:3])))
print("lki max diff %.3e" % float((lki.detach().cpu() - lki_o.detach()).abs().max()))
# the unshifted likelihood of the worst trajectory: top-2 margin
with torch.no_grad():
    raw = O.measurement_cnf(enc[b:b + 1], x[b:b + 1], _pe_tuple(pe), O.unpack_stack(cnf, 32, 32), 2.5)[0]
    top = raw.topk(3)
    print("oracle lki top-3 of that trajectory (shifted):", top.values.tolist(), top.indices.tolist())
    W1, b1, W2, b2, _, _ = _pe_tuple(pe)
    p1 = x[b, n] @ W1.t() + b1
    p2 = torch.relu(p1) @ W2.t() + b2
    print("particle pre-activations: min |layer 1| %.3e, min |layer 2| %.3e" % (float(p1.abs().min()), float(p2.abs().min())))
    allp1 = x[b] @ W1.t() + b1
    allp2 = torch.relu(allp1) @ W2.t() + b2
    print("trajectory-wide: min |layer 1| %.3e, min |layer 2| %.3e" % (float(allp1.abs().min()), float(allp2.abs().min())))
    worst_rows = dx.amax(1).topk(3)
    print("trajectories with the largest d_x error:", worst_rows.indices.tolist(), worst_rows.values.tolist())
    cu_lki = lki.detach().cpu()[b]
    print("cuda argmax %d (lki %.6f), oracle argmax %d" % (int(cu_lki.argmax()), float(cu_lki.max()), int(raw.argmax())))
