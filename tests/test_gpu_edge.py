"""GPU edge cases and the full BASELINE configs[0] shape (bootstrap DPF, Gaussian, soft, N=100, B=32, T=50) against the oracle."""
import numpy as np
import pytest
import torch

import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
from normalizing_flows_dpfs_b200.arguments import parse_args
from normalizing_flows_dpfs_b200.DPFs import DPF
from normalizing_flows_dpfs_b200.losses import supervised_loss
from normalizing_flows_dpfs_b200.nf.flows import RealNVP, RealNVP_cond, pack_parameters
from normalizing_flows_dpfs_b200.nf.models import NormalizingFlowModel_cond
from normalizing_flows_dpfs_b200.utils import compute_normal_density, normalize_log_probs
from test_gpu_ops import close, cu, grad_close

pytestmark = pytest.mark.gpu


def _pe_tuple(mod):
    return tuple(p.detach().cpu().clone() for p in mod.parameters())


@pytest.mark.parametrize("flags,B,N,T", [(["--measurement", "gaussian", "--resampler_type", "soft"], 32, 100, 50),
                                         (["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft"], 3, 1000, 4),
                                         (["--NF-dyn", "--measurement", "cos", "--resampler_type", "soft"], 2, 130, 3),
                                         (["--NF-cond", "--measurement", "CRNVP", "--resampler_type", "ot"], 2, 96, 3)])
def test_filter_vs_oracle_with_reference_gate(flags, B, N, T):
    """Whole filter with the reference's own ESS gate (not forced), ragged N, every NF / NF-cond combination."""
    g = torch.Generator().manual_seed(len(flags) * 100 + N)
    args = parse_args(["--num-particles", str(N), "--batchsize", str(B), "--sequence-length", str(T)] + flags)
    dpf = DPF(args)
    with torch.no_grad():
        for mod, ws in ((dpf.nf_dyn, 0.1), (dpf.cond_model, 0.05), (dpf.particle_encoder, 0.4)) + (
                ((dpf.cnf_measurement, 0.1),) if args.measurement == "CRNVP" else ()):
            for p in mod.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (ws if p.dim() > 1 else 0.05))
    dpf.encoder = torch.nn.Identity()
    cpu = lambda t: t.detach().cpu().clone()
    W = {"dyn": O.unpack_stack(cpu(pack_parameters([dpf.nf_dyn])), 2, 4), "cond": O.unpack_stack(cpu(pack_parameters([dpf.cond_model])), 2, 36),
         "pe": _pe_tuple(dpf.particle_encoder)}
    if args.measurement == "CRNVP":
        W["cnf"] = O.unpack_stack(cpu(pack_parameters([dpf.cnf_measurement])), 32, 32)
    host = dict(enc=torch.randn(B, T, 32, generator=g) * 2, start=torch.randn(B, 4, generator=g) * 10, vel_in=torch.randn(B, T, 2, generator=g) * 3,
                init_particles=torch.rand(B, N, 2, generator=g) * 128 - 64, noise=torch.randn(B, T, N, 2, generator=g) * 20,
                offsets=torch.rand(B, T, generator=g) / N, state=torch.randn(B, T, 4, generator=g) * 20)
    cfg = dict(NF=args.NF_dyn, NF_cond=args.NF_cond, measurement=args.measurement, resampler=args.resampler_type, alpha=0.5, pos_noise=20.0,
               eps=0.1, scaling=0.75, threshold=1e-3, max_iter=100)
    ref = O.filtering(cfg, W, host["init_particles"], host["start"][:, 2:], host["vel_in"], host["enc"], host["noise"], host["offsets"])
    dpf = dpf.cuda()
    dev = {k: v.cuda() for k, v in host.items()}
    dpf.injected = dict(init_particles=dev["init_particles"], noise=dev["noise"], offsets=dev["offsets"])
    out = dpf.filtering_pos(dev["enc"], dev["start"], dev["vel_in"])
    assert dpf.fired == list(ref["fired"]), "ESS gate decisions differ from the oracle"
    assert torch.equal(out[5].cpu(), ref["index"]), "ancestor indices must be bit-exact"
    ot = args.resampler_type == "ot"
    close(out[0], ref["particles"], rtol=1e-4, atol=2e-2 if ot else 2e-3, what="particles")
    close(out[1], ref["probs"], rtol=5e-3 if ot else 5e-4, atol=1e-7, what="probs")
    loss, _ = supervised_loss(out[0], out[1], dev["state"], 1.0, False)
    loss_o, _ = O.supervised_rmse(ref["particles"], ref["probs"], host["state"][:, :, :2])
    close(loss, loss_o, rtol=1e-3, what="RMSE (north star: within 1%)")


def test_module_api_shapes_and_roundtrip():
    """Reference-style module calls: (P,D) inputs, per-sample context, state_dict-compatible submodules."""
    torch.manual_seed(0)
    flow = RealNVP_cond(dim=2, obser_dim=4).cuda()
    flow.zero_initialization(var=0.3)
    x, c = torch.randn(777, 2, device="cuda"), torch.randn(777, 4, device="cuda")
    z, ld = flow.forward(x, c)
    xr, ldi = flow.inverse(z, c)
    assert z.shape == (777, 2) and ld.shape == (777,)
    close(xr, x, atol=1e-5, what="inverse(forward(x))")
    close(ldi, -ld, atol=1e-5, what="log-det antisymmetry")
    plain = RealNVP(dim=4).cuda()
    plain.zero_initialization(var=0.3)
    y = torch.randn(33, 4, device="cuda")
    close(plain.inverse(plain.forward(y)[0])[0], y, atol=1e-5, what="RealNVP round trip")
    prior = torch.distributions.MultivariateNormal(torch.zeros(2, device="cuda"), torch.eye(2, device="cuda"))
    model = NormalizingFlowModel_cond(prior, [RealNVP_cond(dim=2, obser_dim=4) for _ in range(2)], device="cuda")
    for f in model.flows:
        f.zero_initialization(var=0.3)
    zz, logp, ld2 = model.forward(x, c)
    close(logp, prior.log_prob(zz), what="prior log-prob")
    close(model.inverse(zz, c)[0], x, atol=1e-5, what="container round trip")
    assert model.sample(5, c[:5]).shape == (5, 2)
    with pytest.raises((RuntimeError, ValueError)):
        RealNVP_cond(dim=2, hidden_dim=16, obser_dim=4)


def test_utils_match_reference_formulas():
    lw = torch.randn(7, 300, device="cuda") * 4
    close(normalize_log_probs(lw), torch.softmax(lw, 1), rtol=1e-4, atol=1e-9)
    noise = torch.randn(4, 50, 2, device="cuda") * 20
    dens = compute_normal_density(pos_noise=20.0, vel_noise=20.0)(noise)
    close(dens, O.normal_density(noise.cpu(), 20.0), atol=1e-5)


@pytest.mark.parametrize("B,N", [(1, 1), (1, 5), (2, 129), (3, 4096), (1, 5000)])
def test_soft_resample_edge_shapes(B, N):
    g = torch.Generator().manual_seed(N)
    w = torch.softmax(torch.randn(B, N, generator=g) * 5, -1)
    p = torch.randn(B, N, 2, generator=g)
    off = torch.rand(B, generator=g) / N
    mk = torch.linspace(0.0, (N - 1.0) / N, N)
    wg, pg = cu(w).requires_grad_(), cu(p).requires_grad_()
    p2, w2, idx = ops.soft_resample(pg, wg, cu(off), cu(mk), 0.5)
    wo, po = w.clone().requires_grad_(), p.clone().requires_grad_()
    p_o, w_o, i_o = O.soft_resample(po, wo, 0.5, off)
    assert torch.equal(idx.cpu(), i_o)
    close(w2, w_o, what="weights")
    g1, g2 = torch.randn(B, N, 2, generator=g), torch.randn(B, N, generator=g)
    ((p_o * g1).sum() + (w_o * g2).sum()).backward()
    ((p2 * cu(g1)).sum() + (w2 * cu(g2)).sum()).backward()
    grad_close(pg.grad, po.grad, "d_particles")
    grad_close(wg.grad, wo.grad, "d_probs")


def test_cpu_tensors_and_bad_arguments_raise():
    with pytest.raises(RuntimeError):
        ops.weight_update(torch.zeros(2, 3))                      # CPU tensor: there is no CPU fallback
    with pytest.raises((ValueError, AssertionError)):
        ops.soft_resample(torch.zeros(1, 4, 2, device="cuda"), torch.full((1, 4), 0.25, device="cuda"), torch.zeros(1, device="cuda"),
                          torch.zeros(4, device="cuda"), 1.5)     # alpha outside (0, 1]
    with pytest.raises(ValueError):
        ops.coupling_stack(torch.zeros(10, device="cuda"), torch.zeros(1, 4, 2, device="cuda"), None, None, 2, False)  # wrong parameter count
