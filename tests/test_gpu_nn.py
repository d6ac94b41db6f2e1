"""GPU tests of the NN likelihood (measurement_model_NN, reference model/models.py:221-235): mode 3 of the fused measurement kernel
(forward on tcgen05) against the reference's golden outputs / gradients and against the oracle up to B = N = 1024, with the fused
weight update, and inside the filter loop."""
import numpy as np
import pytest
import torch

import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
from test_gpu_ops import close, cu, grad_close

pytestmark = pytest.mark.gpu
T_ = lambda a: torch.from_numpy(np.asarray(a))


def _pe_tuple(flat):
    sizes, out, o = ((16, 2), (16,), (32, 16), (32,), (32, 32), (32,)), [], 0
    for sh in sizes:
        n = int(np.prod(sh)); out.append(flat[o:o + n].reshape(sh)); o += n
    return tuple(out)


def _off_the_relu_kinks_nn(x, enc, pe, head, margin=2e-5):
    """Same reasoning as test_gpu_ops._off_the_relu_kinks, extended to the head's two ReLU layers: a particle with a pre-activation
    within the fp32 / 3xTF32 evaluation error of zero gets a different mask -- and a finitely different gradient -- from any two correct
    implementations (measured: exactly one such particle among 38 000 with |pre-activation| 6.5e-8 where the typical minimum is 5e-3).
    Those particles are moved; the parity bar stays rtol 1e-4."""
    W = _pe_tuple(pe.double())
    Hd = O.unpack_likelihood_head(head.double())
    x = x.double()
    for _ in range(6):
        p1 = x @ W[0].t() + W[1]
        p2 = torch.relu(p1) @ W[2].t() + W[3]
        e = torch.relu(p2) @ W[4].t() + W[5]
        q1 = torch.cat([enc.double()[:, None, :].expand(-1, x.shape[1], -1), e], -1) @ Hd[0].t() + Hd[1]
        q2 = torch.relu(q1) @ Hd[2].t() + Hd[3]
        near = (p1.abs() < margin).any(-1) | (p2.abs() < margin).any(-1) | (q1.abs() < margin).any(-1) | (q2.abs() < margin).any(-1)
        if not bool(near.any()):
            break
        x = x + near[..., None] * 0.013
    return x.float()


def test_nn_likelihood_golden(golden):
    G = golden("meas_nn")
    for c in range(int(G["n_cases"])):
        g = lambda k: T_(G[f"c{c}_{k}"])
        pe, head, enc, x = (cu(g(k)).requires_grad_() for k in ("pe", "head", "enc", "x"))
        lki = ops.measure(pe, head, enc, x, "NN")
        close(lki, g("lki"), rtol=1e-4, atol=1e-5, what=f"NN lki (golden {c})")
        (lki * cu(g("gl"))).sum().backward()
        grad_close(x.grad, g("d_x"), what="NN d_x"); grad_close(enc.grad, g("d_enc"), what="NN d_enc")
        grad_close(pe.grad, g("d_pe"), what="NN d_pe"); grad_close(head.grad, g("d_head"), what="NN d_head")


@pytest.mark.parametrize("B,N", [(5, 100), (16, 1024), (600, 129), (1024, 1024)])
def test_nn_likelihood_fused_update_vs_oracle(B, N):
    g = torch.Generator().manual_seed(B + 3 * N)
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
    head = torch.cat([torch.randn(n, generator=g) * s for n, s in ((4096, 0.15), (64, 0.1), (4096, 0.15), (64, 0.1), (64, 0.3), (1, 0.1))])
    enc, x = torch.randn(B, 32, generator=g), torch.randn(B, N, 2, generator=g) * 3
    lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
    prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    sl = slice(0, min(B, 8))                                   # the oracle sees the first trajectories (rows are independent)
    with torch.no_grad():
        lk_o = O.measurement_nn(enc[sl], x[sl], _pe_tuple(pe), O.unpack_likelihood_head(head))
        lw_o = lw0[sl] + lk_o + prior[sl] - prop[sl]
        pr_o = O.normalize_log_probs(lw_o) + 1e-12
    lki, logw, probs, row_sum, ess_inv, pred = ops.measure_update(cu(pe), cu(head), cu(enc), cu(x), cu(lw0), cu(prior), cu(prop), "NN", want_pred=True)
    close(lki[sl], lk_o, rtol=1e-4, atol=1e-5, what="NN lki")
    close(logw[sl], lw_o, rtol=1e-4, atol=1e-4, what="NN logw")
    close(probs[sl], pr_o, rtol=1e-4, atol=1e-8, what="NN probs")
    close(pred[sl], (pr_o[..., None] * x[sl]).sum(1), rtol=1e-4, atol=1e-4, what="NN prediction")
    assert bool(torch.isfinite(lki).all()) and bool(torch.isfinite(probs).all())


def test_filter_with_the_nn_likelihood_matches_the_oracle_step():
    """DPF(--measurement NN) through filtering_pos: the fused branch is taken (one measurement launch per step) and training works."""
    from normalizing_flows_dpfs_b200 import _lib
    from normalizing_flows_dpfs_b200.losses import supervised_loss
    from normalizing_flows_dpfs_b200.model.models import _FusedMeasurement
    from test_gpu_trainer import _filter_dpf
    dpf, dev = _filter_dpf(["--NF-dyn", "--NF-cond", "--measurement", "NN", "--resampler_type", "soft"], 6, 128, 4)
    assert isinstance(dpf.measurement_model, _FusedMeasurement)
    dpf.injected = dict(init_particles=dev["init_particles"], noise=dev["noise"], offsets=dev["offsets"])
    dpf.force_resample = True
    out = dpf.filtering_pos(dev["enc"], dev["start"], dev["vel_in"])
    loss, _ = supervised_loss(out[0], out[1], dev["state"], 1.0, False)
    loss.backward()
    assert bool(torch.isfinite(loss))
    for mod in (dpf.likelihood_est, dpf.particle_encoder, dpf.nf_dyn, dpf.cond_model):
        gn = sum(float(p.grad.abs().sum()) for p in mod.parameters() if p.grad is not None)
        assert gn > 0 and gn == gn
    # the step's likelihoods equal the reference formula on the step's own particles
    with torch.no_grad():
        x_last, enc_last = out[0][:, -1].cpu(), dev["enc"][:, -1].cpu()
        pe = tuple(p.detach().cpu() for p in dpf.particle_encoder.parameters())
        head = tuple(p.detach().cpu() for p in dpf.likelihood_est.parameters())
        close(out[3][:, -1], O.measurement_nn(enc_last, x_last, pe, head), rtol=1e-4, atol=1e-5, what="NN lki inside the filter")


@pytest.mark.parametrize("B,N", [(9, 300), (16, 1024), (300, 129), (150, 256), (300, 128)])
def test_nn_likelihood_backward_vs_oracle(B, N):
    """The fused mode-3 backward (tcgen05 data path, register-tiled head weight gradients, per-trajectory observation half) through the
    fused weight update and prediction, against torch autograd of the oracle; many trajectories per persistent CTA; run twice (bitwise)."""
    g = torch.Generator().manual_seed(7 * B + N)
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
    head = torch.cat([torch.randn(n, generator=g) * s for n, s in ((4096, 0.15), (64, 0.1), (4096, 0.15), (64, 0.1), (64, 0.3), (1, 0.1))])
    enc, x = torch.randn(B, 32, generator=g), torch.randn(B, N, 2, generator=g) * 3
    x = _off_the_relu_kinks_nn(x, enc, pe, head)
    lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
    prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    gp, gpr, gl = torch.randn(B, N, generator=g), torch.randn(B, 2, generator=g), torch.randn(B, N, generator=g)
    ol = [t.clone().requires_grad_() for t in (pe, head, enc, x, lw0, prior, prop)]      # (the oracle's encoder casts to fp32, like the reference)
    lk = O.measurement_nn(ol[2], ol[3], _pe_tuple(ol[0]), O.unpack_likelihood_head(ol[1]))
    lw = ol[4] + lk + ol[5] - ol[6]
    pr = O.normalize_log_probs(lw) + 1e-12
    ((pr * gp).sum() + ((pr[..., None] * ol[3]).sum(1) * gpr).sum() + (lk * gl).sum()).backward()
    grads = []
    for _ in range(2):
        cl = [cu(t).requires_grad_() for t in (pe, head, enc, x, lw0, prior, prop)]
        lki, logw, probs, row_sum, ess_inv, pred = ops.measure_update(*cl, "NN", want_pred=True)
        ((probs * cu(gp)).sum() + (pred * cu(gpr)).sum() + (lki * cu(gl)).sum()).backward()
        grads.append([t.grad.clone() for t in cl])
    for a, b_ in zip(grads[0], grads[1]):
        assert torch.equal(a, b_), "the NN backward must be run-to-run deterministic"
    for a, o, name in zip(grads[0], ol, ("pe", "head", "enc", "x", "lw0", "prior", "prop")):
        grad_close(a, o.grad, what="NN backward: d_" + name)
