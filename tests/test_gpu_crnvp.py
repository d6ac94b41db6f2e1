"""GPU: CRNVP measurement (model/models.py:256-278) parity beyond the fixed cases of test_gpu_ops.py --
a 64-trajectory slice of a full BASELINE-shape launch (B = N = 1024), and a randomized many-trajectory parity loop (every
persistent CTA walks several trajectories; round 1 saw ONE unreproduced d_pe mismatch in such a case: this loop is the hunt)."""
import os

import numpy as np
import pytest
import torch

import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
from test_gpu_ops import _off_the_argmax_ties, _off_the_relu_kinks, _pe_tuple, close, cu, grad_close

pytestmark = pytest.mark.gpu


def _weights(g):
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
    return pe, O.init_stack(g, 32, 32, std=0.1, bias_std=0.05)


def _oracle_case(pe, cnf, enc, x, lw0, prior, prop, g1, g2, g3):
    lo = [t.clone().requires_grad_() for t in (pe, cnf, enc, x, lw0, prior, prop)]
    lki = O.measurement_cnf(lo[2], lo[3], _pe_tuple(lo[0]), O.unpack_stack(lo[1], 32, 32), 2.5)
    lw = lo[4] + lki + lo[5] - lo[6]
    pr = O.normalize_log_probs(lw) + 1e-12
    ((lki * g1).sum() + (pr * g2).sum() * 50 + (lw.sum(-1) * g3).sum() * 0.01).backward()
    return lki, lw, pr, lo


def test_crnvp_slice_of_a_full_size_launch():
    """B = N = 1024 (BASELINE configs[2]); the oracle checks 64 trajectories of it.  Upstream gradients are zero outside the slice,
    so the parameter gradients of the full launch equal those of the slice alone."""
    g = torch.Generator().manual_seed(77)
    B = N = 1024
    rows = torch.arange(0, B, 16)                                 # 64 trajectories spread over the persistent CTAs
    pe, cnf = _weights(g)
    enc = torch.randn(B, 32, generator=g)
    x = _off_the_relu_kinks(torch.randn(B, N, 2, generator=g) * 3, pe)
    x[rows] = _off_the_argmax_ties(x[rows], lambda xx: O.measurement_cnf(enc[rows], xx, _pe_tuple(pe), O.unpack_stack(cnf, 32, 32), 2.5))
    lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
    prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    mask = torch.zeros(B, 1)
    mask[rows] = 1.0
    g1, g2, g3 = torch.randn(B, N, generator=g) * mask, torch.randn(B, N, generator=g) * mask, torch.randn(B, generator=g) * mask[:, 0]
    lki_o, lw_o, pr_o, lo = _oracle_case(pe, cnf, enc[rows], x[rows], lw0[rows], prior[rows], prop[rows], g1[rows], g2[rows], g3[rows])
    gt = [cu(t).requires_grad_() for t in (pe, cnf, enc, x, lw0, prior, prop)]
    lki, logw, probs, rs, ess = ops.measure_update(*gt, "CRNVP", p0=0.0, p1=2.5)
    r = rows.cuda()
    close(lki[r], lki_o, atol=1e-4, what="lki (slice)")
    close(logw[r], lw_o, atol=1e-4, what="logw (slice)")
    close(probs[r], pr_o, atol=1e-8, what="probs (slice)")
    ((lki * cu(g1)).sum() + (probs * cu(g2)).sum() * 50 + (rs * cu(g3)).sum() * 0.01).backward()
    grad_close(gt[3].grad[r], lo[3].grad, "d_x (slice)")
    grad_close(gt[2].grad[r], lo[2].grad, "d_enc (slice)")
    grad_close(gt[0].grad, lo[0].grad, "d_pe")
    grad_close(gt[1].grad, lo[1].grad, "d_cnf")
    others = torch.ones(B, dtype=torch.bool)
    others[rows] = False
    assert float(gt[3].grad[others.cuda()].abs().max()) == 0.0, "rows without upstream gradient must get none"


def test_crnvp_many_trajectory_randomized_parity():
    """Random shapes with more trajectories than persistent CTAs, fresh seeds each: outputs and every gradient against the oracle.
    NFDPF_STRESS_ITERS raises the count (the -m gpu default keeps the suite short)."""
    iters = int(os.environ.get("NFDPF_STRESS_ITERS", "24"))
    base = int(os.environ.get("NFDPF_TEST_SEED", "0"))
    worst = {}
    for it in range(iters):
        g = torch.Generator().manual_seed(9000 + 131 * it + base)
        B = int(torch.randint(300, 700, (1,), generator=g))
        N = int(torch.randint(40, 140, (1,), generator=g))
        pe, cnf = _weights(g)
        enc = torch.randn(B, 32, generator=g)
        x = _off_the_relu_kinks(torch.randn(B, N, 2, generator=g) * 3, pe)
        x = _off_the_argmax_ties(x, lambda xx: O.measurement_cnf(enc, xx, _pe_tuple(pe), O.unpack_stack(cnf, 32, 32), 2.5))
        lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
        prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
        g1, g2, g3 = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g), torch.randn(B, generator=g)
        lki_o, lw_o, pr_o, lo = _oracle_case(pe, cnf, enc, x, lw0, prior, prop, g1, g2, g3)
        gt = [cu(t).requires_grad_() for t in (pe, cnf, enc, x, lw0, prior, prop)]
        lki, logw, probs, rs, ess = ops.measure_update(*gt, "CRNVP", p0=0.0, p1=2.5)
        ((lki * cu(g1)).sum() + (probs * cu(g2)).sum() * 50 + (rs * cu(g3)).sum() * 0.01).backward()
        close(lki, lki_o, atol=1e-4, what="lki")
        for name, a, b in zip(("d_pe", "d_cnf", "d_enc", "d_x"), gt[:4], lo[:4]):
            err = float((a.grad.cpu() - b.grad).abs().max() / b.grad.abs().max())
            worst[name] = max(worst.get(name, 0.0), err)
            grad_close(a.grad, b.grad, "%s (iteration %d, B=%d, N=%d)" % (name, it, B, N))
    print("worst relative-to-max gradient errors over %d launches: %s" % (iters, worst))


@pytest.mark.parametrize("n_flows,B,N", [(1, 5, 129), (1, 300, 64), (3, 4, 200)])
def test_crnvp_other_flow_counts(n_flows, B, N):
    """One flow runs the warp-specialised backward with two stages; three flows fall back to the single-role kernel.  Outputs and all
    gradients against the oracle."""
    g = torch.Generator().manual_seed(4100 + 7 * n_flows + N)
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
    cnf = O.init_stack(g, 32, 32, std=0.1, n_flows=n_flows, bias_std=0.05)
    enc = torch.randn(B, 32, generator=g)
    x = _off_the_relu_kinks(torch.randn(B, N, 2, generator=g) * 3, pe)
    x = _off_the_argmax_ties(x, lambda xx: O.measurement_cnf(enc, xx, _pe_tuple(pe), O.unpack_stack(cnf, 32, 32, n_flows), 2.5))
    lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
    prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    g1, g2, g3 = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g), torch.randn(B, generator=g)
    lo = [t.clone().requires_grad_() for t in (pe, cnf, enc, x, lw0, prior, prop)]
    lki_o = O.measurement_cnf(lo[2], lo[3], _pe_tuple(lo[0]), O.unpack_stack(lo[1], 32, 32, n_flows), 2.5)
    lw_o = lo[4] + lki_o + lo[5] - lo[6]
    pr_o = O.normalize_log_probs(lw_o) + 1e-12
    ((lki_o * g1).sum() + (pr_o * g2).sum() * 50 + (lw_o.sum(-1) * g3).sum() * 0.01).backward()
    gt = [cu(t).requires_grad_() for t in (pe, cnf, enc, x, lw0, prior, prop)]
    lki, logw, probs, rs, ess = ops.measure_update(*gt, "CRNVP", n_flows=n_flows, p0=0.0, p1=2.5)
    close(lki, lki_o, atol=1e-4, what="lki")
    close(probs, pr_o, atol=1e-8, what="probs")
    ((lki * cu(g1)).sum() + (probs * cu(g2)).sum() * 50 + (rs * cu(g3)).sum() * 0.01).backward()
    for name, a, b in zip(("d_pe", "d_cnf", "d_enc", "d_x"), gt[:4], lo[:4]):
        grad_close(a.grad, b.grad, "%s (n_flows=%d)" % (name, n_flows))
