"""GPU tests of the block pseudo-likelihood kernels (csrc/losses.cu, reference losses.py:37-106): against the reference's golden
outputs / gradients, against the oracle at filter-sized lists read in place through the (T,B,N) strides, and end to end through
DPF.forward with trainType SDPF."""
import numpy as np
import pytest
import torch

import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import losses as M
from normalizing_flows_dpfs_b200 import ops
from test_gpu_ops import close, cu, grad_close

pytestmark = pytest.mark.gpu
T_ = lambda a: torch.from_numpy(np.asarray(a))


def test_block_density_golden(golden):
    G = golden("losses")
    for c in range(int(G["n_cases"])):
        g = lambda k: G[f"c{c}_{k}"]
        bl = int(g("block_len"))
        w, lik, prior = (cu(T_(g(k))).requires_grad_() for k in ("w", "lik", "prior"))
        idx, gq = cu(T_(g("idx"))), cu(T_(g("gq")))
        Q = M.compute_block_density_nf(w, None, lik, idx, None, prior, bl)
        if g("Q").size == 0:       # no complete block (the reference divides by zero blocks): zeros, zero gradients
            assert torch.equal(Q, torch.zeros_like(Q))
            (Q * gq).sum().backward()
            assert not w.grad.any() and not lik.grad.any() and not prior.grad.any()
            continue
        close(Q, g("Q"), rtol=1e-5, atol=1e-5, what="block density Q")
        (Q * gq).sum().backward()
        close(w.grad, g("d_w"), rtol=1e-5, atol=1e-5, what="block density d_w")
        close(prior.grad, g("d_prior"), rtol=1e-5, atol=1e-6, what="block density d_prior")
        close(M.pseudolikelihood_loss_nf(w.detach(), None, lik.detach(), idx, None, prior.detach(), bl), g("loss_nf"), rtol=1e-5, atol=1e-5,
              what="pseudolikelihood_loss_nf")
        noise = cu(T_(g("noise"))).requires_grad_()
        close(lik.grad, g("d_lik"), rtol=1e-5, atol=1e-6, what="block density d_lik")
        Q2 = M.compute_block_density(w.detach(), noise, lik.detach(), idx, bl, 2.0, 0.5)
        close(Q2, g("Q_plain"), rtol=1e-5, atol=1e-5, what="block density (plain) Q")
        (Q2 * gq).sum().backward()
        close(noise.grad, g("d_noise"), rtol=1e-5, atol=1e-6, what="block density (plain) d_noise")


@pytest.mark.parametrize("B,T,N,bl,peaked", [(64, 50, 1024, 10, False), (16, 20, 1000, 10, True), (5, 23, 4096, 5, False), (3, 9, 1, 3, False)])
def test_block_density_vs_oracle_on_strided_lists(B, T, N, bl, peaked):
    """Filter-shaped lists: transposed views of (T,B,N) buffers, sorted ancestor rows with repeats (peaked: almost every particle
    descends from a handful of ancestors -- the long-run case of the backward), some identity rows (gate closed)."""
    g = torch.Generator().manual_seed(B * 7 + N)
    buf = lambda *tail, s=1.0: (torch.randn(T, B, N, *tail, generator=g) * s)
    w = torch.softmax(buf(s=6.0 if peaked else 2.0), -1)
    lik, prior = buf(), buf(s=3.0) - 4
    idx = torch.empty(T, B, N, dtype=torch.int64)
    for t in range(T):
        a = torch.sort(torch.multinomial(w[t], N, replacement=True, generator=g), dim=-1).values
        if t % 4 == 1:
            a = torch.arange(N).expand(B, N)
        idx[t] = a + N * torch.arange(B)[:, None]
    gq = torch.randn(B, generator=g)
    wr, lr, pr = (x.transpose(0, 1).clone().requires_grad_() for x in (w, lik, prior))
    Qr = O.block_density(wr, lr, pr, idx.transpose(0, 1), bl)
    (Qr * gq).sum().backward()
    wd, ld, pd = (cu(x).transpose(0, 1).requires_grad_() for x in (w, lik, prior))     # (B,T,N) VIEWS with strides (N, B*N, 1)
    Q = ops.block_density(wd, ld, pd, cu(idx).transpose(0, 1), bl)
    close(Q, Qr, rtol=1e-5, atol=1e-4, what="block density Q")       # |Q| ~ 1e2..1e3 (sums of 10 block terms of magnitude ~4)
    (Q * cu(gq)).sum().backward()
    grad_close(wd.grad, wr.grad, what="block density d_w")
    grad_close(ld.grad, lr.grad, what="block density d_lik")
    grad_close(pd.grad, pr.grad, what="block density d_prior")
    # run-to-run determinism of the backward (fixed-order run sums, no atomics on sorted rows)
    wd2, ld2, pd2 = (cu(x).transpose(0, 1).requires_grad_() for x in (w, lik, prior))
    (ops.block_density(wd2, ld2, pd2, cu(idx).transpose(0, 1), bl) * cu(gq)).sum().backward()
    assert torch.equal(ld.grad, ld2.grad) and torch.equal(wd.grad, wd2.grad)


def test_block_density_rejects_foreign_ancestors():
    B, T, N = 2, 4, 8
    g = torch.Generator().manual_seed(0)
    w = cu(torch.softmax(torch.randn(B, T, N, generator=g), -1)).requires_grad_()
    lik, prior = cu(torch.randn(B, T, N, generator=g)), cu(torch.randn(B, T, N, generator=g))
    idx = (torch.arange(N).expand(B, T, N) + N * torch.arange(B)[:, None, None]).clone()
    idx[0, 3, 2] = N + 1                                           # points into trajectory 1
    Q = ops.block_density(w, lik, prior, cu(idx), 2)
    ref = O.block_density(w.detach().cpu(), lik.cpu(), prior.cpu(), idx, 2)
    close(Q, ref, rtol=1e-5, atol=1e-5, what="block density Q, foreign ancestor (forward follows flat indices)")
    with pytest.raises(RuntimeError, match="another trajectory"):
        Q.sum().backward()
    idx[0, 3, 2] = B * N + 5
    with pytest.raises(ValueError, match="out of range"):
        ops.block_density(w, lik, prior, cu(idx), 2).sum().backward()


def test_sdpf_training_step_uses_the_kernels():
    """DPF.forward with trainType SDPF (supervised + pseudo-likelihood + AE losses, DPFs.py:103-118): the pseudo-likelihood of the
    device lists equals the reference's gather chain evaluated on the same lists, and its gradient reaches the flows."""
    from normalizing_flows_dpfs_b200 import _lib
    from test_gpu_trainer import _filter_dpf
    dpf, dev = _filter_dpf(["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft", "--trainType", "SDPF",
                            "--block-length", "5"], 6, 64, 10)
    out = dpf.filtering_pos(dev["enc"], dev["start"], dev["vel_in"])
    plist, wlist, nlist, llist, _, ilist, jlist, prlist, _ = out
    n0 = _lib.launch_count()
    loss = M.pseudolikelihood_loss_nf(wlist, nlist, llist, ilist, jlist, prlist, 5)
    assert _lib.launch_count() == n0 + 1, "the block pseudo-likelihood must be one libnfdpf launch"
    ref = -O.block_density(wlist.detach().cpu(), llist.detach().cpu(), prlist.detach().cpu(), ilist.cpu(), 5).mean()
    close(loss, ref, rtol=1e-5, atol=1e-4, what="pseudolikelihood_loss_nf on filter lists")
    loss.backward()
    gn = sum(float(p.grad.abs().sum()) for p in dpf.nf_dyn.parameters() if p.grad is not None)
    assert gn > 0 and gn == gn
