"""GPU tests of the host-free step pieces (csrc/step.cu): the device-side ESS gate and its pass-through branches, the in-kernel
Philox draws (distribution tests: they are not the reference's CPU stream), the fused prediction of the supervised loss, and
CUDA-graph execution of the filter with the REAL gate."""
import math

import numpy as np
import pytest
import torch

import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
from normalizing_flows_dpfs_b200.graphs import GraphedFilterStep
from normalizing_flows_dpfs_b200.losses import supervised_loss
from test_gpu_ops import close, cu, grad_close
from test_gpu_trainer import _filter_dpf

pytestmark = pytest.mark.gpu


def _gate(v):
    return torch.tensor([v], dtype=torch.int32, device="cuda")


def test_ess_gate_matches_the_host_rule():
    g = torch.Generator().manual_seed(1)
    for B, N, scale in ((7, 100, 1.0), (1024, 1024, 1.0), (33, 64, 0.45), (5, 10, 0.55)):
        stats = torch.rand(B, 2, generator=g).cuda()
        stats[:, 1] = stats[:, 1] * N * scale                      # 1 / sum p^2 in (0, N)
        gate, off = ops.ess_gate(stats[:, 1], B, N)
        assert off is None
        assert bool(gate.item()) == bool(stats[:, 1].double().mean().item() < 0.5 * N)
        assert int(ops.ess_gate(stats[:, 1], B, N, force=True)[0].item()) == 1
        assert int(ops.ess_gate(stats[:, 1], B, N, force=False)[0].item()) == 0


@pytest.mark.parametrize("B,N", [(5, 257), (64, 1024)])
def test_soft_resample_closed_gate_is_the_else_branch(B, N):
    """gate == 0: particles / weights pass through, identity ancestors, log-weights = log(probs) (DPFs.py:168-170), with backward."""
    g = torch.Generator().manual_seed(B + N)
    w = torch.softmax(torch.randn(B, N, generator=g) * 2, -1)
    p = torch.randn(B, N, 2, generator=g) * 10
    off, mk = torch.rand(B, generator=g) / N, torch.linspace(0.0, (N - 1.0) / N, N)
    pg, wg = cu(p).requires_grad_(), cu(w).requires_grad_()
    p2, w2, idx, lw2 = ops.soft_resample(pg, wg, cu(off), cu(mk), 0.5, want_log=True, gate=_gate(0))
    assert torch.equal(p2, pg.detach()) and torch.equal(w2, wg.detach())
    assert torch.equal(idx, torch.arange(B * N, device="cuda").reshape(B, N))
    close(lw2, w.log(), what="log weights")
    g1, g2, g3 = torch.randn(B, N, 2, generator=g), torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    ((p2 * cu(g1)).sum() + (w2 * cu(g2)).sum() + (lw2 * cu(g3)).sum()).backward()
    close(pg.grad, g1, what="d_particles")
    close(wg.grad, g2 + g3 / w, rtol=1e-5, what="d_probs")
    # an open gate equals the ungated call bit for bit
    a = ops.soft_resample(cu(p), cu(w), cu(off), cu(mk), 0.5, want_log=True, gate=_gate(1))
    b = ops.soft_resample(cu(p), cu(w), cu(off), cu(mk), 0.5, want_log=True)
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_ot_resample_gate():
    g = torch.Generator().manual_seed(3)
    B, N = 4, 300
    w = torch.softmax(torch.randn(B, N, generator=g) * 2, -1)
    x = torch.randn(B, N, 2, generator=g) * 20
    xg, wg = cu(x).requires_grad_(), cu(w).requires_grad_()
    out = ops.ot_resample(xg, wg.log(), gate=_gate(0))
    w2, lw2 = ops.gate_weights(wg, _gate(0))
    assert torch.equal(out, xg.detach()) and torch.equal(w2, wg.detach())
    g1, g2 = torch.randn(B, N, 2, generator=g), torch.randn(B, N, generator=g)
    ((out * cu(g1)).sum() + (lw2 * cu(g2)).sum()).backward()
    close(xg.grad, g1, what="dx through a closed gate")
    close(wg.grad, g2 / w, rtol=1e-5, what="d_probs through a closed gate")
    ref = ops.ot_resample(cu(x), cu(w).log())
    assert torch.equal(ops.ot_resample(cu(x), cu(w).log(), gate=_gate(1)), ref)
    w3, lw3 = ops.gate_weights(cu(w), _gate(1))
    assert float((w3 - 1.0 / N).abs().max()) == 0 and float((lw3 + math.log(N)).abs().max()) < 1e-6


def test_weighted_mean_vs_torch():
    g = torch.Generator().manual_seed(5)
    for B, N in ((3, 100), (1024, 1024), (7, 4096)):
        x, w = torch.randn(B, N, 2, generator=g) * 30, torch.softmax(torch.randn(B, N, generator=g) * 3, -1)
        gp = torch.randn(B, 2, generator=g)
        xo, wo = x.double().requires_grad_(), w.double().requires_grad_()
        po = (xo * wo[..., None]).sum(1)
        (po * gp.double()).sum().backward()
        xg, wg = cu(x).requires_grad_(), cu(w).requires_grad_()
        pg = ops.weighted_mean(xg, wg)
        close(pg, po, rtol=1e-5, atol=1e-4, what="prediction")
        (pg * cu(gp)).sum().backward()
        close(xg.grad, xo.grad, rtol=1e-6, atol=1e-9, what="d_particles")
        close(wg.grad, wo.grad, rtol=1e-5, atol=1e-4, what="d_probs")


def test_device_rng_draws_have_the_reference_distributions():
    """model/models.py:199-200 (N(0, pos_noise^2) motion noise), resamplers.py:43 (U(0, 1/N) offsets), utils.py:46-62 (initial cloud)."""
    dev = torch.device("cuda")
    B, N, sigma = 512, 1024, 20.0
    rng = torch.tensor([1234, 0], dtype=torch.int64, device=dev)
    x = torch.zeros(B, N, 2, device=dev)
    vel = torch.zeros(B, 2, device=dev)
    ctx = torch.zeros(B, 4, device=dev)
    moved, noise = ops.motion_moments(x, vel, None, ctx, 0, rng, sigma)
    assert torch.equal(moved, noise)
    n = noise.double()
    se = sigma / math.sqrt(n.numel())
    assert abs(float(n.mean())) < 5 * se and abs(float(n.std()) - sigma) < 5 * se
    assert abs(float((n[..., 0] * n[..., 1]).mean())) < 5 * sigma * sigma / math.sqrt(B * N)       # x / y uncorrelated
    z = (n / sigma).flatten()
    assert abs(float((z ** 3).mean())) < 0.02 and abs(float((z ** 4).mean()) - 3.0) < 0.05        # skewness, kurtosis
    assert abs(float((z.abs() < 1).double().mean()) - 0.6827) < 0.003
    close(ctx[:, 2:], noise.std(1), rtol=1e-4, what="fused moments")
    # same state -> same draws; advanced step counter -> fresh draws
    _, again = ops.motion_moments(x, vel, None, None, 0, rng, sigma)
    assert torch.equal(again, noise)
    gate, off = ops.ess_gate(None, B, N, force=True, rng_state=rng, advance=True, want_offsets=True)
    assert int(rng[1].item()) == 1 and float(off.min()) > 0 and float(off.max()) < 1.0 / N
    assert abs(float(off.mean()) * N - 0.5) < 5 / math.sqrt(12 * B)
    _, fresh = ops.motion_moments(x, vel, None, None, 0, rng, sigma)
    assert not torch.equal(fresh, noise) and abs(float((fresh * noise).mean())) < 5 * sigma * sigma / math.sqrt(B * N)
    start = torch.randn(B, 4, device=dev) * 10
    box = ops.init_particles(start[:, :2], 128.0, N, rng)
    assert float(box.min()) >= -64 and float(box.max()) < 64 and abs(float(box.mean())) < 5 * 128 / math.sqrt(12 * B * N * 2)
    assert abs(float(box.std()) - 128 / math.sqrt(12)) < 0.1
    near = ops.init_particles(start[:, :2], 128.0, N, rng, init_with_true_state=True)
    assert abs(float((near - start[:, None, :2]).std()) - 1.0) < 0.01


def test_seeded_particle_initialization_matches_reference(golden):
    """utils.py:46-62 under torch.manual_seed: same cloud, same RNG consumption (the unused velocity draw included)."""
    from normalizing_flows_dpfs_b200.utils import particle_initialization
    G = golden("init")
    for c in range(int(G["n_cases"])):
        g = lambda k: G[f"c{c}_{k}"]
        torch.manual_seed(int(g("seed")))
        p, lw = particle_initialization(cu(g("start")), float(g("width")), int(g("N")), 2, init_with_true_state=bool(g("true_state")))
        nxt = torch.rand(3)
        assert np.array_equal(p.cpu().numpy(), g("particles")), f"case {c}: initial cloud differs from the reference"
        assert np.array_equal(lw.cpu().numpy(), g("logw"))
        assert np.array_equal(nxt.numpy(), g("next_draw")), f"case {c}: the CPU generator was advanced differently"


@pytest.mark.parametrize("resampler", ["soft", "ot"])
def test_graph_with_the_real_gate_equals_eager(resampler):
    """The reference's ESS rule (not forced) inside a captured CUDA graph: same gate decisions as the eager run, gradients bit for bit."""
    flags = ["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", resampler]
    dpf, batch = _filter_dpf(flags, 8, 256, 6)
    dpf.injected = dict(init_particles=batch["init_particles"], noise=batch["noise"], offsets=batch["offsets"])
    out = dpf.filtering_pos(batch["enc"], batch["start"], batch["vel_in"])
    fired = list(dpf.fired)
    assert any(fired) and not all(fired), "the case should exercise both branches of the gate: %r" % (fired,)
    loss, _ = supervised_loss(out[0], out[1], batch["state"], 1.0, False)
    dpf.zero_grad(set_to_none=True)
    loss.backward()
    eager = [p.grad.clone() for p in dpf.nf_dyn.parameters()]
    loss = loss.detach().clone()
    del out
    step = GraphedFilterStep(dpf, batch)
    for _ in range(2):
        g_loss = step.run(batch)
        torch.cuda.synchronize()
        assert dpf.fired == fired
        assert torch.equal(g_loss, loss)
        for a, p in zip(eager, dpf.nf_dyn.parameters()):
            assert torch.equal(a, p.grad), "graph replay must reproduce the eager gradients bit for bit"


def test_fused_prediction_equals_the_list_formula():
    """supervised_loss on the lists of filtering_pos (fused per-step predictions ride along) == the reference formula on plain copies."""
    flags = ["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft"]
    dpf, batch = _filter_dpf(flags, 6, 200, 4)
    dpf.injected = dict(init_particles=batch["init_particles"], noise=batch["noise"], offsets=batch["offsets"])
    grads = []
    for plain in (False, True):
        out = dpf.filtering_pos(batch["enc"], batch["start"], batch["vel_in"])
        pl, wl = (out[0] + 0.0, out[1] + 0.0) if plain else (out[0], out[1])     # "+ 0" drops the attached predictions
        loss, pred = supervised_loss(pl, wl, batch["state"], 1.0, False)
        dpf.zero_grad(set_to_none=True)
        loss.backward()
        grads.append((loss.detach(), pred.detach(), [p.grad.clone() for p in dpf.cond_model.parameters()]))
    close(grads[0][0], grads[1][0], rtol=1e-6, what="loss")
    close(grads[0][1], grads[1][1], rtol=1e-5, atol=1e-4, what="predictions")
    for a, b in zip(grads[0][2], grads[1][2]):
        grad_close(a, b, "cond_model gradient")


def test_ot_far_outlier_row_takes_the_safe_path(golden):
    """A cloud with a far outlier (scaled extent 8.5, met in a C5-shaped run): at the final eps the column exponents span hundreds of
    log2 units, the shared-exponential sums of most rows underflow and must be recomputed by the per-row online LSE -- the first
    version of the round-2 pass kernel turned such a row into NaNs (a denormal sum passed `> 0`, lg2.approx.ftz mapped it to -inf).
    Oracle (fp64): never converges below the threshold, runs to max_iter."""
    G = golden("ot_outlier")
    x, lw = cu(G["x"])[None].contiguous(), cu(G["logw"])[None].contiguous()
    out = ops.ot_resample(x, lw)
    assert int(ops.OtResample.last_iters.item()) == int(G["iters"]) == 101
    assert bool(torch.isfinite(out).all())
    close(out[0], G["p_oracle"], rtol=1e-4, atol=1e-4 * float(np.abs(G["p_oracle"]).max()), what="OT particles, outlier row")
    # the same row inside a batch must not disturb its neighbours' stop rule bookkeeping: finite everywhere
    g = torch.Generator().manual_seed(0)
    xb = torch.cat([x, cu(torch.randn(3, 1024, 2, generator=g) * 20)])
    lwb = torch.cat([lw, cu(torch.log_softmax(torch.randn(3, 1024, generator=g), -1))])
    assert bool(torch.isfinite(ops.ot_resample(xb, lwb)).all())


def test_fanout_sums_the_consumers_gradients_in_one_launch():
    """ops.fanout: n aliases forward (no copy), ONE nfdpf_sum4 launch backward -- equal to autograd's own accumulation."""
    from normalizing_flows_dpfs_b200 import _lib
    g = torch.Generator().manual_seed(4)
    for shape, n in (((64, 100, 2), 4), ((3, 7), 2), ((5, 33, 2), 3), ((2, 2), 6)):
        x = cu(torch.randn(*shape, generator=g)).requires_grad_()
        ws = [cu(torch.randn(*shape, generator=g)) for _ in range(n)]
        parts = ops.fanout(x, n)
        assert all(p.data_ptr() == x.data_ptr() for p in parts)
        n0 = _lib.launch_count()
        sum((p * w).sum() for p, w in zip(parts, ws)).backward()
        assert _lib.launch_count() - n0 == (1 if n <= 4 else 2)        # four operands per summing pass
        close(x.grad, sum(ws), rtol=1e-6, atol=1e-6, what="fanout gradient")
    y = cu(torch.randn(4, 4))
    assert ops.fanout(y, 3)[1] is y          # nothing to do without a gradient


def test_weight_update_backward_writes_the_negated_copy():
    g = torch.Generator().manual_seed(6)
    B, N = 9, 257
    terms = [cu(torch.randn(B, N, generator=g)).requires_grad_() for _ in range(4)]
    logw, probs, row_sum, _ = ops.weight_update(terms[0], terms[1], terms[2], terms[3], 1e-12)
    ((probs * cu(torch.randn(B, N, generator=g))).sum() + (logw * cu(torch.randn(B, N, generator=g))).sum() + row_sum.sum()).backward()
    assert torch.equal(terms[3].grad, -terms[1].grad) and torch.equal(terms[1].grad, terms[2].grad)


@pytest.mark.parametrize("mode,B,N", [("gaussian", 9, 300), ("cos", 5, 129), ("CRNVP", 3, 200), ("gaussian", 64, 1024)])
def test_measurement_epilogue_prediction_equals_the_separate_kernel(mode, B, N):
    """want_pred: the supervised-loss prediction sum_n probs particles (losses.py:22) formed in the measurement kernel's epilogue, its
    gradient folded into the measurement / weight-update backward -- against weighted_mean on the same tensors through autograd."""
    g = torch.Generator().manual_seed(B + N)
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
    cnf = O.init_stack(g, 32, 32, std=0.1, bias_std=0.05) if mode == "CRNVP" else None
    enc, x = torch.randn(B, 32, generator=g), torch.randn(B, N, 2, generator=g) * 2
    lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
    prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    gpred, gprobs = torch.randn(B, 2, generator=g), torch.randn(B, N, generator=g)
    p0, p1 = (0.0, 2.5) if mode == "CRNVP" else (1.0, 10.0)
    res = []
    for fused in (True, False):
        leaves = [cu(t).requires_grad_() if t is not None else None for t in (pe, cnf, enc, x, lw0, prior, prop)]
        out = ops.measure_update(*leaves, mode, p0=p0, p1=p1, want_pred=fused)
        pred = out[5] if fused else ops.weighted_mean(leaves[3], out[2])
        ((pred * cu(gpred)).sum() + (out[2] * cu(gprobs)).sum()).backward()
        res.append((pred.detach(), [l.grad for l in leaves if l is not None]))
    close(res[0][0], res[1][0], rtol=1e-5, atol=1e-5, what="fused prediction")
    for a, b, name in zip(res[0][1], res[1][1], [n for n, t in zip(("pe", "cnf", "enc", "x", "lw0", "prior", "prop"), (pe, cnf, enc, x, lw0, prior, prop)) if t is not None]):
        grad_close(a, b, what="fused prediction: d_" + name)


@pytest.mark.parametrize("B,N", [(7, 100), (64, 1024), (5, 2048), (3, 4096), (4, 2)])
def test_row_moments_vs_torch(B, N):
    """[mean | unbiased std] over the particle axis (model/models.py:309-310, 338-339): the single-pass d = 2 kernel (N <= 2048) and
    the generic one, written into a wider context row at an offset."""
    g = torch.Generator().manual_seed(N)
    x = torch.randn(B, N, 2, generator=g) * 30 + 5
    out = torch.full((B, 9), -7.0, device="cuda")
    ops.row_moments(cu(x), out, 3)
    close(out[:, 3:5], x.double().mean(1), rtol=1e-5, atol=1e-5, what="row mean")
    close(out[:, 5:7], x.double().std(1), rtol=1e-5, atol=1e-5, what="row std")
    assert bool((out[:, :3] == -7.0).all()) and bool((out[:, 7:] == -7.0).all())


@pytest.mark.parametrize("B,N", [(7, 100), (64, 1024), (3, 4096)])
def test_row_moments_writes_the_context_head(B, N):
    """The proposal's context row [observation encoding | mean | std] (model/models.py:360-361) from ONE launch."""
    g = torch.Generator().manual_seed(N + 1)
    x = torch.randn(B, N, 2, generator=g) * 30 + 5
    head = torch.randn(B, 32, generator=g)
    out = torch.full((B, 36), -7.0, device="cuda")
    ops.row_moments(cu(x), out, 32, head=cu(head))
    assert torch.equal(out[:, :32].cpu(), head), "encoding columns must be copied bit for bit"
    close(out[:, 32:34], x.double().mean(1), rtol=1e-5, atol=1e-5, what="row mean")
    close(out[:, 34:36], x.double().std(1), rtol=1e-5, atol=1e-5, what="row std")


def test_grad_slab_equals_autograd_accumulation():
    """Parameter gradients of a module's packed vector: kernel calls write rows of the pack's GradSlab and the pack's backward sums
    them once (nf/flows.py) -- same gradients as autograd's pairwise accumulation over plain leaf copies, also when one call's output
    never reaches the loss (its row is never written) and when backward runs twice over a retained graph."""
    from normalizing_flows_dpfs_b200.nf.flows import RealNVP_cond
    from normalizing_flows_dpfs_b200.nf.models import NormalizingFlowModel_cond
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(11)
    flows = [RealNVP_cond(dim=2, obser_dim=4) for _ in range(2)]
    for f in flows:
        f.zero_initialization(var=0.3)
    model = NormalizingFlowModel_cond(None, flows, device="cuda").cuda()
    B, N = 6, 300
    xs = [cu(torch.randn(B, N, 2, generator=g)) for _ in range(3)]
    ctx = cu(torch.randn(B, 4, generator=g))
    gy = cu(torch.randn(B, N, 2, generator=g))

    def run(packed_of):
        outs = [ops.coupling_stack(packed_of(), x, ctx, None, 2, inverse=bool(i & 1)) for i, x in enumerate(xs)]
        return (outs[0][0] * gy).sum() + (outs[2][1]).sum() * 0.1        # the second call's outputs are unused
    packed = model.packed()
    assert getattr(packed, "_nfdpf_slab", None) is not None
    loss = run(model.packed)
    loss.backward(retain_graph=True)
    first = [p.grad.clone() for p in model.parameters()]
    for p in model.parameters():
        p.grad = None
    loss.backward()
    again = [p.grad.clone() for p in model.parameters()]
    leaf = packed.detach().clone().requires_grad_()
    run(lambda: leaf).backward()
    flat = torch.cat([t.reshape(-1) for t in first])
    assert torch.equal(flat, torch.cat([t.reshape(-1) for t in again])), "a second backward over the retained graph must reproduce the first"
    close(flat, leaf.grad.cpu().double(), rtol=1e-6, atol=1e-7 * float(leaf.grad.abs().max()), what="slab sum vs autograd accumulation")
