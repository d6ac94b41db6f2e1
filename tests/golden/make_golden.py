"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference).

Run in the build container only (the reference mount does not exist on the GPU box):
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
The reference has no tests or golden vectors of its own; these fixtures are what pins both the CPU
oracle (oracle/nfdpf_oracle.py) and the CUDA path to the reference's behaviour.
Every fixture stores the inputs (including every random draw the reference made, captured by
wrapping torch.normal / Tensor.uniform_ / torch.rand) and the reference's outputs.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NFDPF_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):  # absent here; plot.py only
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)

import utils as ref_utils  # noqa: E402
from nf import flows as ref_flows  # noqa: E402
from model import models as ref_models  # noqa: E402
from resamplers import resamplers as ref_rs  # noqa: E402

torch.set_num_threads(8)


def npy(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def flat_params(module):
    return torch.cat([p.detach().reshape(-1) for p in module.parameters()])


def randomise(module, gen, w_std, b_std):
    with torch.no_grad():
        for n, p in module.named_parameters():
            p.copy_(torch.randn(p.shape, generator=gen) * (w_std if p.dim() > 1 else b_std))


def flat_grads(module):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in module.parameters()])


# ------------------------------------------------------------------ soft resampler
def gen_soft():
    out = {}
    cases = [(2, 8, 0.5, 1.0), (4, 100, 0.5, 2.0), (8, 1024, 0.5, 4.0), (2, 4096, 0.3, 3.0), (3, 100, 1.0, 2.0),
             (5, 7, 0.5, 1.0), (16, 1024, 0.5, 6.0)]
    for ci, (B, N, alpha, temp) in enumerate(cases):
        g = torch.Generator().manual_seed(100 + ci)
        probs = torch.softmax(torch.randn(B, N, generator=g) * temp, -1)
        particles = torch.randn(B, N, 2, generator=g) * 20
        seed = 7 + ci
        torch.manual_seed(seed)
        offsets = torch.FloatTensor(B).uniform_(0.0, 1.0 / N)  # the draw soft_resampler makes (resamplers.py:43)
        torch.manual_seed(seed)
        pp = particles.clone().requires_grad_()
        ww = probs.clone().requires_grad_()
        p_res, w_res, idx = ref_rs.soft_resampler(pp, ww, alpha, N, index=True, device="cpu")
        gp = torch.randn(p_res.shape, generator=g)
        gw = torch.randn(w_res.shape, generator=g)
        tot = (p_res * gp).sum()
        if w_res.requires_grad:  # alpha == 1 returns constant 1/N weights
            tot = tot + (w_res * gw).sum()
        tot.backward()
        out.update({f"c{ci}_{k}": v for k, v in npy(dict(
            alpha=alpha, probs=probs, particles=particles, offsets=offsets, p_res=p_res, w_res=w_res, idx=idx,
            gp=gp, gw=gw, d_particles=pp.grad,
            d_probs=ww.grad if ww.grad is not None else torch.zeros_like(ww))).items()})
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "soft_resample.npz"), **out)


# ------------------------------------------------------------------ coupling flows
def gen_flows():
    out = {}
    cases = [("cond", 2, 4, 64), ("cond", 2, 36, 64), ("cond", 32, 32, 48), ("plain", 2, 0, 64), ("plain", 4, 0, 33),
             ("cond", 4, 3, 17)]
    for ci, (kind, D, C, P) in enumerate(cases):
        g = torch.Generator().manual_seed(200 + ci)
        if kind == "cond":
            flows = [ref_flows.RealNVP_cond(dim=D, obser_dim=C) for _ in range(2)]
        else:
            flows = [ref_flows.RealNVP(dim=D) for _ in range(2)]
        mod = torch.nn.ModuleList(flows)
        randomise(mod, g, 0.3, 0.1)
        x = torch.randn(P, D, generator=g)
        ctx = torch.randn(P, C, generator=g) if kind == "cond" else None
        rec = dict(kind=kind, D=D, C=C, params=flat_params(mod), x=x)
        if ctx is not None:
            rec["ctx"] = ctx
        for direction in ("forward", "inverse"):
            mod.zero_grad()
            xx = x.clone().requires_grad_()
            cc = ctx.clone().requires_grad_() if ctx is not None else None
            y, ld = xx, torch.zeros(P)
            seq = flows if direction == "forward" else flows[::-1]
            for f in seq:
                fn = getattr(f, direction)
                y, l1 = fn(y, cc) if cc is not None else fn(y)
                ld = ld + l1
            gy = torch.randn(y.shape, generator=g)
            gl = torch.randn(ld.shape, generator=g)
            ((y * gy).sum() + (ld * gl).sum()).backward()
            rec.update({f"{direction}_y": y, f"{direction}_ld": ld, f"{direction}_gy": gy, f"{direction}_gl": gl,
                        f"{direction}_dx": xx.grad, f"{direction}_dW": flat_grads(mod)})
            if cc is not None:
                rec[f"{direction}_dctx"] = cc.grad
        out.update({f"c{ci}_{k}": v for k, v in npy(rec).items()})
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "flows.npz"), **out)


# ------------------------------------------------------------------ per-step glue
def gen_glue():
    g = torch.Generator().manual_seed(300)
    B, N, h = 3, 40, 32
    out = {}
    dyn = ref_models.build_conditional_nf(2, 4, 2, init_var=0.01)
    cond = ref_models.build_conditional_nf(2, 36, 2, init_var=0.01)
    cnf = ref_models.build_conditional_nf(2, h, h, init_var=0.01, prior_std=2.5)
    pe = ref_models.build_particle_encoder(h, 2)
    randomise(dyn, g, 0.2, 0.1), randomise(cond, g, 0.1, 0.1), randomise(cnf, g, 0.1, 0.05), randomise(pe, g, 0.3, 0.1)
    x = torch.randn(B, N, 2, generator=g) * 20 + 5
    noise = torch.randn(B, N, 2, generator=g) * 20
    enc = torch.randn(B, h, generator=g)
    lw = torch.randn(B, N, generator=g) * 3
    dens = ref_utils.compute_normal_density(pos_noise=20.0, vel_noise=20.0)
    gauss = torch.distributions.MultivariateNormal(torch.ones(h), 100 * torch.eye(h))
    m_gauss = ref_models.measurement_model_Gaussian(pe, gauss)
    m_cnf = ref_models.measurement_model_cnf(pe, cnf)
    m_cos = ref_models.measurement_model_cosine_distance(pe)
    rec = dict(x=x, noise=noise, enc=enc, lw=lw, dyn=flat_params(dyn), cond=flat_params(cond), cnf=flat_params(cnf),
               pe=flat_params(pe))
    rec["normalize"] = ref_utils.normalize_log_probs(lw)
    rec["density"] = dens(noise)
    xd, jac = ref_models.nf_dynamic_model(dyn, x, (B, N), NF=True)
    rec["dyn_x"], rec["dyn_jac"] = xd, jac
    xp, jp = ref_models.normalising_flow_propose(cond, x, enc)
    rec["prop_x"], rec["prop_jac"] = xp, jp
    rec["meas_gauss"] = m_gauss(enc, x)
    rec["meas_cnf"] = m_cnf(enc, x)
    rec["meas_cos"] = m_cos(enc, x)
    for tag, NF, NFc, meas in (("pl_full_gauss", True, True, m_gauss), ("pl_full_cnf", True, True, m_cnf),
                               ("pl_cond_only", False, True, m_gauss), ("pl_boot", False, False, m_gauss),
                               ("pl_dyn_only_cos", True, False, m_cos)):
        xd_, jac_ = ref_models.nf_dynamic_model(dyn, x, (B, N), NF=NF)
        prop, lki, prior, plog = ref_models.proposal_likelihood(cond, dyn, meas, xd_, x, enc, noise, jac_, NF, NFc, dens)
        rec.update({f"{tag}_prop": prop, f"{tag}_lki": lki, f"{tag}_prior": prior, f"{tag}_plog": plog})
    out.update(npy(rec))
    np.savez_compressed(os.path.join(OUT, "glue.npz"), **out)


# ------------------------------------------------------------------ OT resampler
def gen_ot():
    out = {}
    cases = [(3, 16, 1.0, 0.1), (4, 100, 2.0, 0.1), (2, 256, 3.0, 0.1), (2, 64, 2.0, 0.5), (1, 33, 1.0, 0.05)]
    for ci, (B, N, temp, eps) in enumerate(cases):
        g = torch.Generator().manual_seed(400 + ci)
        w = torch.softmax(torch.randn(B, N, generator=g) * temp, -1)
        x = torch.randn(B, N, 2, generator=g) * torch.tensor([20.0, 7.0]) + torch.tensor([3.0, -40.0])
        iters = {}
        orig = ref_rs.sinkhorn_loop

        def spy(*a, **k):
            r = orig(*a, **k)
            iters["n"] = r[-1]
            return r

        ref_rs.sinkhorn_loop = spy
        xx = x.clone().requires_grad_()
        ww = w.clone().requires_grad_()
        p_res, w_res, idx = ref_rs.resampler_ot(xx, ww, eps=eps, scaling=0.75, threshold=1e-3, max_iter=100, device="cpu")
        ref_rs.sinkhorn_loop = orig
        gp = torch.randn(p_res.shape, generator=g)
        (p_res * gp).sum().backward()
        with torch.no_grad():
            T = ref_rs.transport_function(x, w.log(), eps, 0.75, 1e-3, 100, N, "cpu")
        rec = dict(eps=eps, w=w, x=x, p_res=p_res, w_res=w_res, idx=idx, iters=iters["n"], gp=gp, dx=xx.grad,
                   dw_is_none=ww.grad is None, T_colsum=T.sum(1), T_rowsum=T.sum(2))
        if N <= 100:
            rec["T"] = T
        out.update({f"c{ci}_{k}": v for k, v in npy(rec).items()})
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "ot_resample.npz"), **out)


# ------------------------------------------------------------------ whole filter (DPF.filtering_pos)
class DrawLog:
    """Capture every random draw the reference makes, in order."""

    def __init__(self):
        self.normal, self.uniform, self.rand = [], [], []

    def __enter__(self):
        self._n, self._u, self._r = torch.normal, torch.Tensor.uniform_, torch.rand
        log = self

        def normal(*a, **k):
            r = log._n(*a, **k)
            log.normal.append(r.clone())
            return r

        def uniform_(t, *a, **k):
            r = log._u(t, *a, **k)
            log.uniform.append(r.clone())
            return r

        def rand(*a, **k):
            r = log._r(*a, **k)
            log.rand.append(r.clone())
            return r

        torch.normal, torch.Tensor.uniform_, torch.rand = normal, uniform_, rand
        return self

    def __exit__(self, *a):
        torch.normal, torch.Tensor.uniform_, torch.rand = self._n, self._u, self._r


def gen_filter():
    sys.argv = ["main.py"]
    import arguments as ref_args
    import DPFs as ref_dpfs
    import losses as ref_losses
    out = {}
    cases = [
        ("boot_gauss_soft", ["--measurement", "gaussian", "--resampler_type", "soft"]),
        ("cnf_gauss_soft", ["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft"]),
        ("crnvp_soft", ["--measurement", "CRNVP", "--resampler_type", "soft"]),
        ("full_crnvp_ot", ["--NF-dyn", "--NF-cond", "--measurement", "CRNVP", "--resampler_type", "ot"]),
        ("cnf_cos_soft", ["--NF-dyn", "--NF-cond", "--measurement", "cos", "--resampler_type", "soft"]),
    ]
    B, N, T = 4, 32, 6
    for ci, (tag, flags) in enumerate(cases):
        sys.argv = ["main.py", "--num-particles", str(N), "--batchsize", str(B), "--sequence-length", str(T)] + flags
        args = ref_args.parse_args()
        torch.manual_seed(500 + ci)
        dpf = ref_dpfs.DPF(args)
        g = torch.Generator().manual_seed(600 + ci)
        randomise(dpf.nf_dyn, g, 0.1, 0.05), randomise(dpf.cond_model, g, 0.05, 0.05)
        randomise(dpf.particle_encoder, g, 0.4, 0.2)
        if hasattr(dpf, "cnf_measurement"):
            randomise(dpf.cnf_measurement, g, 0.1, 0.05)
        dpf.encoder = torch.nn.Identity()  # obs[:, t] IS the precomputed encoding (CNN out of scope)
        enc = torch.randn(B, T, 32, generator=g) * (3.0 if "gauss" in tag else 1.0)
        state = torch.cat([torch.randn(B, T, 2, generator=g) * 20, torch.randn(B, T, 2, generator=g) * 3], -1)
        start = torch.cat([torch.randn(B, 2, generator=g) * 20, torch.randn(B, 2, generator=g) * 3], -1)
        vel_in = state[:, :, 2:] + torch.randn(B, T, 2, generator=g) * 4
        fired = []
        rs_fwd = dpf.resampler.forward

        def spy(p, w, _f=rs_fwd, _l=fired):
            _l.append(len(log.normal))  # number of motion-noise draws so far == step index
            return _f(p, w)

        dpf.resampler.forward = spy
        with DrawLog() as log:
            res = dpf.filtering_pos(enc, start, vel_in)
        (pl, wl, nl, ll, iwl, il, jl, prl, obs_lik) = res
        loss, pred = ref_losses.supervised_loss(pl, wl, state, 1.0, False)
        dpf.zero_grad()
        loss.backward()
        fired_mask = np.zeros(T, bool)
        fired_mask[fired] = True
        offs = torch.zeros(B, T)
        for k, t in enumerate(fired):
            if args.resampler_type == "soft":
                offs[:, t] = log.uniform[k]
        rec = dict(flags=" ".join(flags), B=B, N=N, T=T, enc=enc, state=state, start=start, vel_in=vel_in,
                   init_particles=128.0 * log.rand[0] - 64.0, noise=torch.stack(log.normal, 1), offsets=offs,
                   fired=fired_mask, particles=pl, probs=wl, lki=ll, index=il, obs_likelihood=obs_lik, loss=loss, pred=pred,
                   dyn=flat_params(dpf.nf_dyn), cond=flat_params(dpf.cond_model), pe=flat_params(dpf.particle_encoder),
                   d_dyn=flat_grads(dpf.nf_dyn), d_cond=flat_grads(dpf.cond_model), d_pe=flat_grads(dpf.particle_encoder))
        assert torch.equal(rec["init_particles"], pl.new_tensor(0) + rec["init_particles"])
        if jl is not None:
            rec["jac"], rec["prior"] = jl, prl
        if hasattr(dpf, "cnf_measurement"):
            rec["cnf"], rec["d_cnf"] = flat_params(dpf.cnf_measurement), flat_grads(dpf.cnf_measurement)
        out.update({f"{tag}_{k}": v for k, v in npy(rec).items()})
        print(tag, "loss", float(loss), "fired", fired_mask.astype(int))
    out["tags"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(OUT, "filter.npz"), **out)


def gen_init():
    """utils.py:46-62 under a seed: the cloud, the log-weights, and the generator state afterwards (next draw)."""
    out = {}
    cases = [(4, 100, 128.0, False, 3), (3, 64, 128.0, True, 5), (2, 1024, 64.0, False, 11)]
    for ci, (B, N, width, true_state, seed) in enumerate(cases):
        g = torch.Generator().manual_seed(50 + ci)
        start = torch.randn(B, 2, generator=g) * 20
        torch.manual_seed(seed)
        p, lw = ref_utils.particle_initialization(start, width, N, 2, init_with_true_state=true_state)
        nxt = torch.rand(3)
        out.update({f"c{ci}_{k}": v for k, v in npy(dict(start=start, width=width, N=N, true_state=true_state, seed=seed, particles=p,
                                                         logw=lw, next_draw=nxt)).items()})
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "init.npz"), **out)


def gen_losses():
    """losses.py: block pseudo-likelihood (NF and plain variants) and the masked supervised loss, with gradients."""
    import losses as ref_losses
    out = {}
    cases = [(3, 20, 16, 10, "sorted"), (2, 25, 64, 5, "sorted"), (4, 12, 33, 4, "mixed"), (2, 7, 8, 10, "sorted"), (2, 10, 12, 5, "shuffled")]
    for ci, (B, T, N, bl, kind) in enumerate(cases):
        g = torch.Generator().manual_seed(900 + ci)
        w = torch.softmax(torch.randn(B, T, N, generator=g) * 2, -1)
        lik = torch.randn(B, T, N, generator=g)
        prior = torch.randn(B, T, N, generator=g) * 3 - 4
        jac = torch.randn(B, T, N, generator=g)
        noise = torch.randn(B, T, N, 4, generator=g)
        idx = torch.empty(B, T, N, dtype=torch.int64)
        for b in range(B):
            for t in range(T):
                if kind == "shuffled":
                    a = torch.randint(0, N, (N,), generator=g)
                elif kind == "mixed" and t % 3 == 0:
                    a = torch.arange(N)                       # gate closed: identity ancestors
                else:                                         # what soft resampling returns: sorted, with repeats
                    a = torch.sort(torch.multinomial(w[b, t], N, replacement=True, generator=g)).values
                idx[b, t] = a + N * b
        ww, ll, pp, nn = (x.clone().requires_grad_() for x in (w, lik, prior, noise))
        Q = ref_losses.compute_block_density_nf(ww, None, ll, idx, jac, pp, bl) if T >= bl else None
        gq = torch.randn(B, generator=g)
        if Q is not None:
            (Q * gq).sum().backward()
            loss = ref_losses.pseudolikelihood_loss_nf(w, None, lik, idx, jac, prior, bl)
            Q2 = ref_losses.compute_block_density(w, nn, ll.detach(), idx, bl, 2.0, 0.5)
            (Q2 * gq).sum().backward()
        else:   # no complete block: the reference divides 0 by 0 blocks
            Q, loss, Q2 = torch.zeros(0), torch.zeros(()), torch.zeros(0)
        z = torch.zeros
        out.update({f"c{ci}_{k}": v for k, v in npy(dict(
            block_len=bl, w=w, lik=lik, prior=prior, noise=noise, idx=idx, gq=gq, Q=Q, loss_nf=loss, Q_plain=Q2,
            d_w=ww.grad if ww.grad is not None else z(0), d_lik=ll.grad if ll.grad is not None else z(0),
            d_prior=pp.grad if pp.grad is not None else z(0), d_noise=nn.grad if nn.grad is not None else z(0))).items()})
    out["n_cases"] = len(cases)
    # supervised loss, train branch with a label mask (losses.py:18-27)
    g = torch.Generator().manual_seed(990)
    B, T, N = 4, 6, 32
    x = (torch.randn(B, T, N, 2, generator=g) * 10).requires_grad_()
    w = torch.softmax(torch.randn(B, T, N, generator=g), -1).requires_grad_()
    state = torch.randn(B, T, 4, generator=g) * 10
    mask = (torch.rand(B, T, generator=g) < 0.6).float()
    loss, pred = ref_losses.supervised_loss(x, w, state, mask, True, labeledRatio=0.6)
    loss.backward()
    loss_eval, _ = ref_losses.supervised_loss(x.detach(), w.detach(), state, 1.0, False)
    out.update({f"sup_{k}": v for k, v in npy(dict(x=x, w=w, state=state, mask=mask, loss=loss, pred=pred, d_x=x.grad, d_w=w.grad,
                                                    loss_eval=loss_eval)).items()})
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def gen_nn():
    """measurement_model_NN (model/models.py:221-235) on seeded inputs: log-likelihoods and gradients (a file of its own, so that
    glue.npz keeps regenerating bit for bit)."""
    out = {}
    cases = [(3, 40, 0.3, 0.12), (4, 129, 0.5, 0.2), (2, 300, 0.2, 0.1)]
    for ci, (B, N, pe_std, head_std) in enumerate(cases):
        g = torch.Generator().manual_seed(700 + ci)
        pe = ref_models.build_particle_encoder(32, 2)
        head = ref_models.build_likelihood(32, 2)
        randomise(pe, g, pe_std, 0.1), randomise(head, g, head_std, 0.1)
        m = ref_models.measurement_model_NN(pe, head)
        x = (torch.randn(B, N, 2, generator=g) * 3 + 1).requires_grad_()
        enc = torch.randn(B, 32, generator=g).requires_grad_()
        lki = m(enc, x)
        gl = torch.randn(B, N, generator=g)
        (lki * gl).sum().backward()
        out.update({f"c{ci}_{k}": v for k, v in npy(dict(x=x, enc=enc, pe=flat_params(pe), head=flat_params(head), lki=lki, gl=gl,
                                                         d_x=x.grad, d_enc=enc.grad, d_pe=flat_grads(pe), d_head=flat_grads(head))).items()})
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "meas_nn.npz"), **out)


def gen_state_dict():
    """state_dict keys + shapes of the reference DPF for the accelerated configurations (checkpoint compatibility)."""
    import json
    sys.argv = ["main.py"]
    import arguments as ref_args
    import DPFs as ref_dpfs
    out = {}
    for meas in ("gaussian", "cos", "CRNVP", "NN"):
        sys.argv = ["main.py", "--measurement", meas]
        dpf = ref_dpfs.DPF(ref_args.parse_args())
        out[meas] = {k: list(v.shape) for k, v in dpf.state_dict().items()}
    with open(os.path.join(OUT, "state_dict.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["soft", "flows", "glue", "ot", "filter", "state_dict", "init", "losses", "nn"]
    sys.argv = sys.argv[:1]
    for w in which:
        {"soft": gen_soft, "flows": gen_flows, "glue": gen_glue, "ot": gen_ot, "filter": gen_filter, "state_dict": gen_state_dict, "init": gen_init, "losses": gen_losses, "nn": gen_nn}[w]()
        print("wrote", w)
