"""GPU parity tests: CUDA kernels (through the C-ABI) vs the CPU oracle and the reference-generated goldens.
Tolerances (BASELINE.json north_star): indices bit-exact; everything else fp32 rtol 1e-4 / atol 1e-5."""
import os

import numpy as np
import pytest
import torch

import ledger
import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5
T = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a))
cu = lambda a: T(a).cuda()


def close(a, b, rtol=RTOL, atol=ATOL, what=""):
    a = T(a).detach().cpu().double().numpy()
    b = T(b).detach().cpu().double().numpy()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    ledger.record(what, a, b, rtol, atol)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), "%s: max abs diff %.3e (worst excess %.3e)" % (what, np.abs(a - b).max(), err.max())


GRAD_FLOOR = float(os.environ.get("NFDPF_GRAD_FLOOR", "0.2"))      # atol = 2e-5 of the gradient tensor's largest magnitude (round 1: 1e-4); measured worst: 6.5e-6


def grad_close(a, b, what="", floor=None):
    """Gradients are sums of 1e3..1e6 fp32 products.  Every entry must hold |err| <= rtol * max(|ref|, floor * max|ref|) with
    rtol = 1e-4: entries above `floor` of the tensor's largest magnitude are checked at the north-star relative tolerance, smaller
    ones absolutely against that floor.  The floor is set from the measured errors (profiles/r2_parity_errors.json lists, per
    comparison, the CUDA-vs-oracle error next to the error of torch's own fp32 evaluation against fp64)."""
    floor = GRAD_FLOOR if floor is None else floor
    b_ = T(b).detach().cpu().double().numpy()
    close(a, b, rtol=RTOL, atol=max(1e-12, RTOL * floor * float(np.abs(b_).max())), what=what)


# ------------------------------------------------------------------------------------------ soft resampling
def _soft_case(p, w, off, alpha):
    N = w.shape[1]
    markers = torch.linspace(0.0, (N - 1.0) / N, N)
    pg, wg = cu(p).requires_grad_(), cu(w).requires_grad_()
    return pg, wg, ops.soft_resample(pg, wg, cu(off), cu(markers), alpha)


def test_soft_resample_golden(golden):
    G = golden("soft_resample")
    for c in range(int(G["n_cases"])):
        g = lambda k: G[f"c{c}_{k}"]
        alpha = float(g("alpha"))
        pg, wg, (p_res, w_res, idx) = _soft_case(g("particles"), g("probs"), g("offsets"), alpha)
        assert np.array_equal(idx.cpu().numpy(), g("idx")), f"case {c}: resampled indices must be bit-exact"
        assert np.array_equal(p_res.detach().cpu().numpy(), g("p_res"))
        close(w_res, g("w_res"), what=f"case {c} weights")
        tot = (p_res * cu(g("gp"))).sum()
        if alpha < 1.0:
            tot = tot + (w_res * cu(g("gw"))).sum()
        tot.backward()
        grad_close(pg.grad, g("d_particles"), f"case {c} d_particles")
        grad_close(wg.grad, g("d_probs"), f"case {c} d_probs")


@pytest.mark.parametrize("B,N,alpha,temp", [(1024, 1024, 0.5, 4.0), (256, 4096, 0.5, 8.0), (512, 100, 0.5, 2.0),
                                            (37, 1000, 0.7, 3.0), (64, 1024, 1.0, 3.0), (3, 1, 0.5, 1.0)])
def test_soft_resample_vs_oracle_full_size(B, N, alpha, temp):
    g = torch.Generator().manual_seed(B + N)
    w = torch.softmax(torch.randn(B, N, generator=g) * temp, -1)
    p = torch.randn(B, N, 2, generator=g) * 20
    off = torch.rand(B, generator=g) / N
    pg, wg, (p_res, w_res, idx) = _soft_case(p, w, off, alpha)
    po, wo, io = O.soft_resample(p, w, alpha, off)
    assert np.array_equal(idx.cpu().numpy(), io.numpy()), "indices must be bit-exact (%d mismatches)" % int(
        (idx.cpu() != io).sum())
    assert torch.equal(p_res.detach().cpu(), po)
    close(w_res, wo, what="weights")
    # size-independent properties: weights sum to one, indices monotone and inside the row
    assert torch.allclose(w_res.sum(-1), torch.ones(B, device="cuda"), atol=1e-5)
    loc = idx - N * torch.arange(B, device="cuda")[:, None]
    assert bool((loc[:, 1:] >= loc[:, :-1]).all()) and int(loc.min()) >= 0 and int(loc.max()) < N


# --------------------------------------------------------------------------------------- weight update / norm
@pytest.mark.parametrize("B,N", [(3, 40), (64, 1024), (5, 4096), (7, 1), (1024, 1024)])
def test_weight_update_vs_oracle(B, N):
    g = torch.Generator().manual_seed(B * 7 + N)
    lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
    lki, prior, prop = (torch.randn(B, N, generator=g) * s for s in (3.0, 1.0, 1.0))
    leaves = [t.clone().requires_grad_() for t in (lw0, lki, prior, prop)]
    lw_o = leaves[0] + leaves[1] + leaves[2] - leaves[3]
    pr_o = O.normalize_log_probs(lw_o) + 1e-12
    gl = [cu(t).requires_grad_() for t in (lw0, lki, prior, prop)]
    logw, probs, row_sum, ess_inv = ops.weight_update(gl[0], gl[1], gl[2], gl[3], 1e-12)
    close(logw, lw_o, what="logw")
    close(probs, pr_o, rtol=1e-4, atol=1e-9, what="probs")
    close(row_sum, lw_o.sum(-1), rtol=1e-4, atol=1e-3, what="row_sum")
    close(ess_inv, 1.0 / (pr_o ** 2).sum(-1), what="ess")
    gp, gs = torch.randn(B, N, generator=g), torch.randn(B, generator=g)
    ((pr_o * gp).sum() + (lw_o.sum(-1) * gs).sum()).backward()
    ((probs * cu(gp)).sum() + (row_sum * cu(gs)).sum()).backward()
    for a, b, n in zip(gl, leaves, ("lw0", "lki", "prior", "propose")):
        grad_close(a.grad, b.grad, "d_" + n)


def test_normalize_golden(golden):
    G = golden("glue")
    _, probs, _, _ = ops.weight_update(cu(G["lw"]))
    close(probs, G["normalize"], rtol=1e-4, atol=1e-9)


# ------------------------------------------------------------------------------------------- coupling stacks
def test_coupling_golden(golden):
    """per-particle context exactly as the reference passes it (P,C): kernels <HALF,CP> with C_row = 0."""
    G = golden("flows")
    for c in range(int(G["n_cases"])):
        g = lambda k: G[f"c{c}_{k}"]
        D, C = int(g("D")), int(g("C"))
        P = g("x").shape[0]
        for direction in ("forward", "inverse"):
            pk = cu(g("params")).requires_grad_()
            x = cu(g("x")).reshape(1, P, D).requires_grad_()
            ctx = cu(g("ctx")).reshape(1, P, C).requires_grad_() if C else None
            y, ld = ops.coupling_stack(pk, x, None, ctx, 2, direction == "inverse")
            close(y.reshape(P, D), g(f"{direction}_y"), what=f"case {c} {direction} y")
            close(ld.reshape(P), g(f"{direction}_ld"), what=f"case {c} {direction} log_det")
            ((y.reshape(P, D) * cu(g(f"{direction}_gy"))).sum() + (ld.reshape(P) * cu(g(f"{direction}_gl"))).sum()).backward()
            grad_close(x.grad.reshape(P, D), g(f"{direction}_dx"), f"case {c} {direction} dx")
            grad_close(pk.grad, g(f"{direction}_dW"), f"case {c} {direction} dW")
            if C:
                grad_close(ctx.grad.reshape(P, C), g(f"{direction}_dctx"), f"case {c} {direction} dctx")


@pytest.mark.parametrize("D,C_row,C_part,B,N,inverse", [(2, 4, 0, 5, 300, True), (2, 4, 0, 5, 300, False), (2, 36, 0, 4, 1024, True),
                                                        (32, 0, 32, 3, 200, False), (2, 0, 0, 2, 64, False), (4, 2, 3, 3, 129, True),
                                                        (32, 5, 0, 2, 100, False),
                                                        # D = 2 backward, resident-set geometry: several passes over the CTA's entries with a
                                                        # ragged 128-particle iteration; several 1024-particle entries per trajectory (ragged
                                                        # last one); both at once with more than one trajectory per CTA
                                                        (2, 4, 0, 1500, 130, True), (2, 36, 0, 3, 2500, False), (2, 4, 0, 300, 1100, False)])
def test_coupling_row_context_vs_oracle(D, C_row, C_part, B, N, inverse):
    """row-constant context hoisted into the layer-1 bias == the reference's materialised (P,C) concat."""
    g = torch.Generator().manual_seed(D * 1000 + C_row * 10 + C_part + int(inverse))
    C = C_row + C_part
    pk = O.init_stack(g, D, C, std=0.3, bias_std=0.1)
    x = torch.randn(B, N, D, generator=g) * 1.5
    rc = torch.randn(B, C_row, generator=g) if C_row else None
    pc = torch.randn(B, N, C_part, generator=g) if C_part else None
    gy, gl = torch.randn(B, N, D, generator=g), torch.randn(B, N, generator=g)
    # oracle: materialise the context like model/models.py:309-315 does
    lo = [t.clone().requires_grad_() if t is not None else None for t in (pk, x, rc, pc)]
    parts = ([lo[2][:, None, :].expand(B, N, C_row)] if C_row else []) + ([lo[3]] if C_part else [])
    ctx = torch.cat(parts, -1).reshape(B * N, C) if parts else None
    fn = O.stack_inverse if inverse else O.stack_forward
    yo, ldo = fn(lo[1].reshape(B * N, D), ctx, O.unpack_stack(lo[0], D, C))
    ((yo.reshape(B, N, D) * gy).sum() + (ldo.reshape(B, N) * gl).sum()).backward()
    gt = [cu(t).requires_grad_() if t is not None else None for t in (pk, x, rc, pc)]
    y, ld = ops.coupling_stack(gt[0], gt[1], gt[2], gt[3], 2, inverse)
    close(y, yo.reshape(B, N, D), what="y")
    close(ld, ldo.reshape(B, N), what="log_det")
    ((y * cu(gy)).sum() + (ld * cu(gl)).sum()).backward()
    for a, b, n in zip(gt, lo, ("dW", "dx", "d_row_ctx", "d_part_ctx")):
        if a is not None:
            grad_close(a.grad, b.grad, n)


def test_coupling_roundtrip_full_size():
    """size-independent property at BASELINE shape (B=N=1024): inverse(forward(x)) == x, log-dets cancel."""
    g = torch.Generator().manual_seed(5)
    B = N = 1024
    pk = cu(O.init_stack(g, 2, 36, std=0.2, bias_std=0.1))
    x = cu(torch.randn(B, N, 2, generator=g) * 2)
    rc = cu(torch.randn(B, 36, generator=g))
    z, ld_f = ops.coupling_stack(pk, x, rc, None, 2, False)
    xr, ld_i = ops.coupling_stack(pk, z, rc, None, 2, True)
    close(xr, x, rtol=1e-4, atol=1e-4, what="roundtrip")
    close(ld_i, -ld_f, rtol=1e-4, atol=1e-5, what="log-det antisymmetry")


def test_unsupported_shape_raises():
    pk = torch.zeros(4 * 2 * (8 * (3 + 0) + 8 + 64 + 8 + 3 * 8 + 3), device="cuda")
    with pytest.raises(RuntimeError):
        ops.coupling_stack(pk, torch.zeros(1, 4, 6, device="cuda"), None, None, 2, False)


# --------------------------------------------------------------------------------------------- measurement
def _pe_tuple(flat):
    sizes = [(16, 2), (16,), (32, 16), (32,), (32, 32), (32,)]
    o, parts = 0, []
    for s in sizes:
        n = int(np.prod(s))
        parts.append(flat[o:o + n].reshape(*s))
        o += n
    return tuple(parts)


def test_measurement_golden(golden):
    G = golden("glue")
    pe, cnf, enc, x = cu(G["pe"]), cu(G["cnf"]), cu(G["enc"]), cu(G["x"])
    close(ops.measure(pe, None, enc, x, "gaussian", p0=1.0, p1=10.0), G["meas_gauss"], atol=1e-4, what="gaussian")
    close(ops.measure(pe, None, enc, x, "cos"), G["meas_cos"], what="cos")
    close(ops.measure(pe, cnf, enc, x, "CRNVP", p0=0.0, p1=2.5), G["meas_cnf"], atol=1e-4, what="CRNVP")


def _off_the_relu_kinks(x, pe, margin=2e-5):
    """The encoder's ReLU masks are discontinuous in the pre-activations: a particle whose pre-activation sits within the
    fp32 / 3xTF32 evaluation error (~5e-7) of zero gets a different mask -- and a finitely different gradient -- from any two
    correct implementations.  With ~1e5 particles x 48 units a handful always do, so the parity inputs are nudged off the kinks."""
    W1, b1, W2, b2, _, _ = _pe_tuple(pe)
    for _ in range(4):
        p1 = x @ W1.t() + b1
        p2 = torch.relu(p1) @ W2.t() + b2
        near = (p1.abs() < margin).any(-1) | (p2.abs() < margin).any(-1)
        if not bool(near.any()):
            break
        x = x + near[..., None] * 0.01
    return x


def _off_the_argmax_ties(x, lki_fn, margin=1e-3):
    """`likelihood - likelihood.max(-1)` (models.py:252, 276) routes -sum(g) to the argmax particle: two particles within the
    evaluation error of the row maximum make the two implementations pick different ones.  Runner-ups closer than `margin` are moved."""
    if x.shape[1] < 2:
        return x
    for it in range(40):
        with torch.no_grad():
            top2 = lki_fn(x).topk(2, dim=-1)
        tie = (top2.values[:, 0] - top2.values[:, 1]) < margin
        if not bool(tie.any()):
            return x
        x = x.clone()
        rows = torch.nonzero(tie).flatten()
        # growing displacement: where the likelihood is nearly flat in x (freshly initialised CRNVP stacks) a fixed 0.05 step does not
        # separate the pair -- three such rounds left a 3.8e-6 tie in NFDPF_TEST_SEED=1, iteration 12 of the randomized CRNVP loop,
        # the "5e-2 d_pe mismatch" of round 1 (CUDA and oracle picked different argmax particles)
        x[rows, top2.indices[rows, 1]] += 0.05 * (1 + it)
    raise AssertionError("could not separate the two most likely particles of every trajectory by %g" % margin)


@pytest.mark.parametrize("mode,B,N,fused", [("gaussian", 3, 200, True), ("cos", 3, 200, True), ("CRNVP", 3, 200, True),
                                            ("gaussian", 2, 50, False), ("CRNVP", 2, 129, False), ("gaussian", 16, 1024, True),
                                            ("CRNVP", 8, 1024, True),
                                            # more trajectories than persistent CTAs (2 x 148): every CTA walks several trajectories, its
                                            # tensor-memory gradient accumulators and the per-trajectory d_enc increments carry across them
                                            ("gaussian", 700, 200, True), ("cos", 650, 130, False), ("CRNVP", 600, 129, True)])
def test_measure_update_vs_oracle(mode, B, N, fused):
    # deterministic (hash() of a str is salted per process); NFDPF_TEST_SEED shifts it for seed sweeps
    g = torch.Generator().manual_seed((len(mode) * 7919 + B * 31 + N) % 1000 + int(os.environ.get("NFDPF_TEST_SEED", "0")))
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
    cnf = O.init_stack(g, 32, 32, std=0.1, bias_std=0.05) if mode == "CRNVP" else None
    enc = torch.randn(B, 32, generator=g)
    x = _off_the_relu_kinks(torch.randn(B, N, 2, generator=g) * 3, pe)
    lw0 = torch.log_softmax(torch.randn(B, N, generator=g), -1)
    prior, prop = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    p0, p1 = {"gaussian": (1.0, 10.0), "cos": (0.0, 1.0), "CRNVP": (0.0, 2.5)}[mode]
    if mode == "gaussian":
        x = _off_the_argmax_ties(x, lambda xx: O.measurement_gaussian(enc, xx, _pe_tuple(pe)))
    elif mode == "CRNVP":
        x = _off_the_argmax_ties(x, lambda xx: O.measurement_cnf(enc, xx, _pe_tuple(pe), O.unpack_stack(cnf, 32, 32), 2.5))
    names = ("pe", "cnf", "enc", "x", "lw0", "prior", "prop")
    lo = {k: (v.clone().requires_grad_() if v is not None else None) for k, v in zip(names, (pe, cnf, enc, x, lw0, prior, prop))}
    if mode == "gaussian":
        lki_o = O.measurement_gaussian(lo["enc"], lo["x"], _pe_tuple(lo["pe"]))
    elif mode == "cos":
        lki_o = O.measurement_cos(lo["enc"], lo["x"], _pe_tuple(lo["pe"]))
    else:
        lki_o = O.measurement_cnf(lo["enc"], lo["x"], _pe_tuple(lo["pe"]), O.unpack_stack(lo["cnf"], 32, 32), 2.5)
    gt = {k: (cu(v).requires_grad_() if v is not None else None) for k, v in zip(names, (pe, cnf, enc, x, lw0, prior, prop))}
    g1, g2, g3 = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g), torch.randn(B, generator=g)
    if fused:
        lw_o = lo["lw0"] + lki_o + lo["prior"] - lo["prop"]
        pr_o = O.normalize_log_probs(lw_o) + 1e-12
        ((lki_o * g1).sum() + (pr_o * g2).sum() * 50 + (lw_o.sum(-1) * g3).sum() * 0.01).backward()
        lki, logw, probs, rs, ess = ops.measure_update(gt["pe"], gt["cnf"], gt["enc"], gt["x"], gt["lw0"], gt["prior"], gt["prop"], mode,
                                                       p0=p0, p1=p1)
        close(logw, lw_o, atol=1e-4, what="logw")
        close(probs, pr_o, atol=1e-8, what="probs")
        close(ess, 1.0 / (pr_o ** 2).sum(-1), what="ess")
        ((lki * cu(g1)).sum() + (probs * cu(g2)).sum() * 50 + (rs * cu(g3)).sum() * 0.01).backward()
    else:
        (lki_o * g1).sum().backward()
        lki = ops.measure(gt["pe"], gt["cnf"], gt["enc"], gt["x"], mode, p0=p0, p1=p1)
        (lki * cu(g1)).sum().backward()
    close(lki, lki_o, atol=1e-4, what="lki")
    for k in names:
        if gt[k] is not None and lo[k].grad is not None:
            grad_close(gt[k].grad, lo[k].grad, "d_" + k)


# ------------------------------------------------------------------------------------------- OT resampling
def test_ot_resample_golden(golden):
    from normalizing_flows_dpfs_b200.resamplers.resamplers import resampler_ot
    G = golden("ot_resample")
    for c in range(int(G["n_cases"])):
        g = lambda k: G[f"c{c}_{k}"]
        x, w = cu(g("x")).requires_grad_(), cu(g("w")).requires_grad_()
        p_res, w_res, idx = resampler_ot(x, w, eps=float(g("eps")))
        assert int(ops.OtResample.last_iters.item()) == int(g("iters")), (c, int(ops.OtResample.last_iters.item()), int(g("iters")))
        # particles here are O(50): rtol 1e-4 / atol 1e-5 on the cloud's own scale (the reference runs this in fp64)
        close(p_res, g("p_res"), rtol=1e-4, atol=1e-5 * float(np.abs(g("x")).max()) * 10, what=f"case {c} particles")
        assert np.array_equal(w_res.cpu().numpy(), g("w_res")) and np.array_equal(idx.cpu().numpy(), g("idx"))
        (p_res * cu(g("gp"))).sum().backward()
        grad_close(x.grad, g("dx"), f"case {c} dx")
        assert w.grad is None or float(w.grad.abs().max()) == 0.0


@pytest.mark.parametrize("B,N", [(4, 100), (8, 1024), (2, 2500), (3, 37)])
def test_ot_resample_vs_oracle(B, N):
    g = torch.Generator().manual_seed(B * 31 + N)
    w = torch.softmax(torch.randn(B, N, generator=g) * 2.0, -1)
    x = torch.randn(B, N, 2, generator=g) * torch.tensor([15.0, 6.0]) + torch.tensor([10.0, -30.0])
    xo = x.clone().requires_grad_()
    po, wo, io, iters = O.ot_resample(xo, w, return_iters=True)
    xg = cu(x).requires_grad_()
    pg = ops.ot_resample(xg, cu(w).log())
    assert int(ops.OtResample.last_iters.item()) == iters
    close(pg, po, rtol=1e-4, atol=2e-3, what="particles")
    gp = torch.randn(B, N, 2, generator=g)
    (po * gp).sum().backward()
    (pg * cu(gp)).sum().backward()
    grad_close(xg.grad, xo.grad, "dx")


def test_ot_resample_properties_full_size():
    """B = N = 1024 (BASELINE config 3 shape): plan column sums = N w_j, so the weighted mean is preserved:
    (1/N) sum_i x'_i = sum_j w_j x_j (row sums are only ~1, so there is no bounding-box guarantee)."""
    g = torch.Generator().manual_seed(11)
    B = N = 1024
    w = cu(torch.softmax(torch.randn(B, N, generator=g) * 2.0, -1))
    x = cu(torch.randn(B, N, 2, generator=g) * 20.0)
    p = ops.ot_resample(x, w.log())
    close(p.mean(1), (w[..., None] * x).sum(1), rtol=1e-3, atol=2e-2, what="weighted mean preserved")
    assert bool(torch.isfinite(p).all())
    it = int(ops.OtResample.last_iters.item())
    assert 10 < it <= 100, it


# ------------------------------------------------------------------------------------------------- step glue
def test_motion_moments_and_proposal_terms_vs_oracle():
    g = torch.Generator().manual_seed(9)
    B, N = 5, 300
    x, noise = torch.randn(B, N, 2, generator=g) * 30, torch.randn(B, N, 2, generator=g) * 20
    vel = torch.randn(B, 2, generator=g) * 3
    ctx = torch.zeros(B, 9, device="cuda")
    out, _ = ops.motion_moments(cu(x), cu(vel), cu(noise), ctx, 3)
    ref = O.motion_update(x, vel, noise)
    close(out, ref, atol=1e-4, what="motion")
    mean, std = O.row_stats(ref)
    close(ctx[:, 3:5], mean[:, 0], atol=1e-4, what="mean")
    close(ctx[:, 5:7], std[:, 0], atol=1e-4, what="std")
    assert float(ctx[:, :3].abs().max()) == 0 and float(ctx[:, 7:].abs().max()) == 0
    back = torch.randn(B, N, 2, generator=g) * 30
    jb, jd, jp = (torch.randn(B, N, generator=g) for _ in range(3))
    lo = [t.clone().requires_grad_() for t in (back, ref, jb, jd, jp)]
    prior_o = O.normal_density(lo[0] - (lo[1] - noise), 20.0) - lo[2]
    prop_o = O.normal_density(noise, 20.0) + lo[3] + lo[4]
    gt = [cu(t).requires_grad_() for t in (back, ref, jb, jd, jp)]
    prior, prop = ops.proposal_terms(gt[0], gt[1], cu(noise), gt[2], gt[3], gt[4], 20.0)
    close(prior, prior_o, atol=1e-4, what="prior")
    close(prop, prop_o, atol=1e-4, what="propose")
    g1, g2 = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    ((prior_o * g1).sum() + (prop_o * g2).sum()).backward()
    ((prior * cu(g1)).sum() + (prop * cu(g2)).sum()).backward()
    for a, b, n in zip(gt, lo, ("back", "phys", "jac_back", "jac_dyn", "jac_prop")):
        grad_close(a.grad, b.grad, "d_" + n)


def test_soft_resample_log_output_gradient():
    g = torch.Generator().manual_seed(4)
    B, N = 6, 257
    w = torch.softmax(torch.randn(B, N, generator=g) * 2, -1)
    p = torch.randn(B, N, 2, generator=g)
    off = torch.rand(B, generator=g) / N
    mk = torch.linspace(0.0, (N - 1.0) / N, N)
    wo = w.clone().requires_grad_()
    _, w_res, _ = O.soft_resample(p, wo, 0.5, off)
    g1, g2 = torch.randn(B, N, generator=g), torch.randn(B, N, generator=g)
    ((w_res * g1).sum() + (w_res.log() * g2).sum()).backward()
    wg = cu(w).requires_grad_()
    _, w2, _, lw2 = ops.soft_resample(cu(p), wg, cu(off), cu(mk), 0.5, want_log=True)
    close(lw2, w_res.log(), what="log weights")
    ((w2 * cu(g1)).sum() + (lw2 * cu(g2)).sum()).backward()
    grad_close(wg.grad, wo.grad, "d_probs through log output")
