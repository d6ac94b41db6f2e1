"""CPU oracle for the NF-DPF per-timestep particle update.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU tensors + numpy) of the algorithm the
reference `xiongjiechen/Normalizing-Flows-DPFs` runs on its hot path.  It exists so the CUDA
path can be checked where `/root/reference` is not mounted (the GPU box).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
it; the product package `normalizing_flows_dpfs_b200` never does.

Pinning: the reference has no tests or golden vectors of its own, so this
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, imported unmodified in the build
container by `tests/golden/make_golden.py`; the resulting fixtures live in `tests/golden/*.npz`
and `tests/test_oracle_golden.py` replays them (soft-resample indices bit-exact, everything
else at rtol 1e-5 or tighter).

All citations are `file:line` in the reference repository.

Conventions
  B trajectories, N particles, P = B*N, d = 2 state dims, h = 32 encoding dims, H = 8 FCNN width.
  A coupling *stack* is a list of flows; a flow is a dict {"t1","s1","t2","s2"} of FCNN
  parameter 6-tuples (W1,b1,W2,b2,W3,b3) in torch.nn.Linear layout (out,in).
"""
from __future__ import annotations

import math

import numpy as np
import torch

FCNN_ORDER = ("t1", "s1", "t2", "s2")  # registration order, nf/flows.py:185-188


# ----------------------------------------------------------------------------------------
# parameter packing (shared convention with the C-ABI: include/nfdpf.h "packed stack")
# ----------------------------------------------------------------------------------------
def stack_param_count(D: int, C: int, n_flows: int = 2, H: int = 8) -> int:
    half = D // 2
    per_fcnn = H * (half + C) + H + H * H + H + half * H + half
    return n_flows * 4 * per_fcnn


def unpack_stack(flat: torch.Tensor, D: int, C: int, n_flows: int = 2, H: int = 8):
    """Split a flat parameter vector (state_dict order: flows.k.{t1,s1,t2,s2}.network.{0,2,4}.{weight,bias})."""
    half = D // 2
    fin = half + C
    flows, o = [], 0

    def take(*shape):
        nonlocal o
        n = int(np.prod(shape))
        v = flat[o:o + n].reshape(*shape)
        o += n
        return v

    for _ in range(n_flows):
        flow = {}
        for name in FCNN_ORDER:
            flow[name] = (take(H, fin), take(H), take(H, H), take(H), take(half, H), take(half))
        flows.append(flow)
    assert o == flat.numel()
    return flows


def init_stack(gen: torch.Generator, D: int, C: int, std: float = 0.01, n_flows: int = 2, H: int = 8,
               bias_std: float = 0.0) -> torch.Tensor:
    """N(0,std) weights, zero (or small random) biases -- zero_initialization, nf/flows.py:191-211."""
    flat = torch.zeros(stack_param_count(D, C, n_flows, H))
    half, fin = D // 2, D // 2 + C
    o = 0
    for _ in range(n_flows * 4):
        for n_w, n_b in ((H * fin, H), (H * H, H), (half * H, half)):
            flat[o:o + n_w] = torch.randn(n_w, generator=gen) * std
            o += n_w
            if bias_std > 0:
                flat[o:o + n_b] = torch.randn(n_b, generator=gen) * bias_std
            o += n_b
    return flat


# ----------------------------------------------------------------------------------------
# flows -- nf/flows.py:101-114 (FCNN), 215-239 (RealNVP_cond), 155-179 (RealNVP)
# ----------------------------------------------------------------------------------------
def fcnn(x: torch.Tensor, p) -> torch.Tensor:
    W1, b1, W2, b2, W3, b3 = p
    h = torch.tanh(x.float() @ W1.t() + b1)  # `.float()` cast: nf/flows.py:114
    h = torch.tanh(h @ W2.t() + b2)
    return h @ W3.t() + b3


def _join(half, ctx):
    return half if ctx is None else torch.cat([half, ctx], dim=-1)


def coupling_forward(x, ctx, flow, D):
    """nf/flows.py:215-226 (ctx=None gives the unconditional RealNVP.forward, 155-166)."""
    lower, upper = x[:, :D // 2], x[:, D // 2:]
    t1, s1 = fcnn(_join(lower, ctx), flow["t1"]), fcnn(_join(lower, ctx), flow["s1"])
    upper = t1 + upper * torch.exp(s1)
    t2, s2 = fcnn(_join(upper, ctx), flow["t2"]), fcnn(_join(upper, ctx), flow["s2"])
    lower = t2 + lower * torch.exp(s2)
    return torch.cat([lower, upper], dim=1), s1.sum(dim=1) + s2.sum(dim=1)


def coupling_inverse(z, ctx, flow, D):
    """nf/flows.py:228-239 (ctx=None gives RealNVP.inverse, 168-179)."""
    lower, upper = z[:, :D // 2], z[:, D // 2:]
    t2, s2 = fcnn(_join(upper, ctx), flow["t2"]), fcnn(_join(upper, ctx), flow["s2"])
    lower = (lower - t2) * torch.exp(-s2)
    t1, s1 = fcnn(_join(lower, ctx), flow["t1"]), fcnn(_join(lower, ctx), flow["s1"])
    upper = (upper - t1) * torch.exp(-s1)
    return torch.cat([lower, upper], dim=1), (-s1).sum(dim=1) + (-s2).sum(dim=1)


def stack_forward(x, ctx, flows):
    """NormalizingFlowModel_cond.forward without the prior term, nf/models.py:45-52."""
    D = x.shape[1]
    log_det = torch.zeros(x.shape[0])
    for flow in flows:
        x, ld = coupling_forward(x, ctx, flow, D)
        log_det = log_det + ld
    return x, log_det


def stack_inverse(z, ctx, flows):
    """NormalizingFlowModel_cond.inverse, nf/models.py:54-61 (flows applied last to first)."""
    D = z.shape[1]
    log_det = torch.zeros(z.shape[0])
    for flow in flows[::-1]:
        z, ld = coupling_inverse(z, ctx, flow, D)
        log_det = log_det + ld
    return z, log_det


def diag_normal_logprob(z, mean: float, std: float):
    """MultivariateNormal(mean*1, std^2*I).log_prob -- the priors at model/models.py:167-168, DPFs.py:84-85."""
    k = z.shape[-1]
    return -0.5 * (((z - mean) / std) ** 2).sum(-1) - k * (math.log(std) + 0.5 * math.log(2 * math.pi))


# ----------------------------------------------------------------------------------------
# filter glue -- model/models.py
# ----------------------------------------------------------------------------------------
def row_stats(x):
    """Detached per-trajectory mean and unbiased std, model/models.py:309-310, 338-339."""
    return x.mean(dim=1, keepdim=True).detach().clone(), x.std(dim=1, keepdim=True).detach().clone()


def motion_update(particles, vel, noise):
    """model/models.py:191-204 with the N(0,pos_noise^2) draw injected by the caller."""
    return particles + vel[:, None, :] + noise


def nf_dynamic(flows, x, forward=False, mean=None, std=None):
    """nf_dynamic_model with NF=True, model/models.py:305-329.  Returns (particles, jac=-log_det)."""
    B, N, d = x.shape
    if mean is None:
        mean, std = row_stats(x)
    ctx = torch.cat([mean.detach().repeat(1, N, 1).reshape(-1, d), std.detach().repeat(1, N, 1).reshape(-1, d)], -1)
    flat = x.reshape(-1, d)
    out, ld = stack_forward(flat, ctx, flows) if forward else stack_inverse(flat, ctx, flows)
    return out.reshape(B, N, d), (-ld).reshape(B, N)


def nf_propose(flows, x, enc):
    """normalising_flow_propose, model/models.py:334-356: context = [enc_b, mean_b, std_b]."""
    B, N, d = x.shape
    mean, std = row_stats(x)
    ctx = torch.cat([enc[:, None, :].repeat(1, N, 1).reshape(B * N, -1),
                     mean.repeat(1, N, 1).reshape(-1, d), std.repeat(1, N, 1).reshape(-1, d)], -1)
    out, ld = stack_inverse(x.reshape(-1, d), ctx, flows)
    return out.reshape(B, N, d), (-ld).reshape(B, N)


def normal_density(noise, std_pos: float):
    """compute_normal_density.forward for d=2 (velocity terms vanish), utils.py:17-37."""
    d = noise.shape[-1]
    log_c = -0.5 * math.log(2 * math.pi)
    return d * log_c - 2 * math.log(std_pos) - (noise[..., :2] ** 2 / (2 * std_pos ** 2)).sum(-1)


def particle_encoder(x, pe):
    """build_particle_encoder: Linear(2,16)-ReLU-Linear(16,32)-ReLU-Linear(32,h), model/models.py:130-139."""
    W1, b1, W2, b2, W3, b3 = pe
    a = torch.relu(x.float() @ W1.t() + b1)
    a = torch.relu(a @ W2.t() + b2)
    return a @ W3.t() + b3


def measurement_gaussian(enc, x, pe):
    """measurement_model_Gaussian with MVN(loc=1, cov=100 I), model/models.py:237-254, DPFs.py:84-86."""
    e = particle_encoder(x, pe)
    ll = diag_normal_logprob(enc[:, None, :] - e, 1.0, 10.0)
    return ll - ll.max(dim=-1, keepdim=True)[0]


def measurement_cnf(enc, x, pe, cnf_flows, prior_std: float = 2.5):
    """measurement_model_cnf, model/models.py:256-278; prior N(0, 2.5^2 I) from DPFs.py:75-76."""
    B, N, _ = x.shape
    h = enc.shape[-1]
    e = particle_encoder(x, pe).reshape(-1, h)
    obs = enc[:, None, :].repeat(1, N, 1).reshape(-1, h)
    z, ld = stack_forward(obs, e, cnf_flows)
    ll = (diag_normal_logprob(z.float(), 0.0, prior_std) + ld).reshape(B, N)
    return ll - ll.max(dim=-1, keepdim=True)[0]


def measurement_cos(enc, x, pe):
    """measurement_model_cosine_distance, model/models.py:206-219 + et_distance utils.py:8-15."""
    e = particle_encoder(x, pe)
    a = torch.nn.functional.normalize(enc[:, None, :].expand_as(e), p=2, dim=-1, eps=1e-12)
    b = torch.nn.functional.normalize(e, p=2, dim=-1, eps=1e-12)
    return (1.0 / (1e-7 + (1.0 - (a * b).sum(-1)))).log()


def unpack_likelihood_head(flat):
    """build_likelihood, model/models.py:119-128: Linear(64,64)-ReLU-Linear(64,64)-ReLU-Linear(64,1)-Sigmoid in state_dict order."""
    sizes = ((64, 64), (64,), (64, 64), (64,), (1, 64), (1,))
    out, o = [], 0
    for sh in sizes:
        n = int(np.prod(sh))
        out.append(flat[o:o + n].reshape(sh))
        o += n
    assert o == flat.numel()
    return tuple(out)


def measurement_nn(enc, x, pe, head):
    """measurement_model_NN, model/models.py:221-235: sigmoid MLP on [observation encoding | particle encoding], then log."""
    W1, b1, W2, b2, W3, b3 = head
    e = particle_encoder(x, pe)
    inp = torch.cat([enc[:, None, :].expand(-1, e.shape[1], -1), e], dim=-1)
    h = torch.relu(inp @ W1.t() + b1)
    h = torch.relu(h @ W2.t() + b2)
    return torch.sigmoid(h @ W3.t() + b3)[..., 0].log()


def normalize_log_probs(lw):
    """utils.py:39-44."""
    e = (lw - lw.max(dim=1, keepdim=True)[0]).exp()
    return e / e.sum(dim=1, keepdim=True)


def proposal_likelihood(cfg, W, x_dyn, x_phys, enc, noise, jac_dyn):
    """model/models.py:358-379.  cfg keys: NF, NF_cond, measurement, pos_noise."""
    dens = lambda n: normal_density(n, cfg["pos_noise"])
    enc_prop = enc.detach().clone()
    if cfg["NF_cond"]:
        prop, jac_prop = nf_propose(W["cond"], x_dyn, enc_prop)
        if cfg["NF"]:
            mean, std = x_phys.mean(dim=1, keepdim=True), x_phys.std(dim=1, keepdim=True)
            back, jac_back = nf_dynamic(W["dyn"], prop, forward=True, mean=mean, std=std)
            prior = dens(back - (x_phys - noise)) - jac_back
        else:
            prior = dens(prop - (x_phys - noise))
        propose_log = dens(noise) + jac_dyn + jac_prop
    else:
        prop = x_dyn
        prior = dens(noise) + jac_dyn
        propose_log = dens(noise) + jac_dyn
    if cfg["measurement"] == "gaussian":
        lki = measurement_gaussian(enc, prop, W["pe"])
    elif cfg["measurement"] == "CRNVP":
        lki = measurement_cnf(enc, prop, W["pe"], W["cnf"])
    elif cfg["measurement"] == "cos":
        lki = measurement_cos(enc, prop, W["pe"])
    else:
        raise ValueError(cfg["measurement"])
    return prop, lki, prior, propose_log


# ----------------------------------------------------------------------------------------
# soft resampling -- resamplers/resamplers.py:20-60
# ----------------------------------------------------------------------------------------
def _ceil_log2(x: int) -> int:
    return 0 if x <= 1 else (x - 1).bit_length()


def _multi_row_sum(a: np.ndarray) -> np.ndarray:
    """ATen's multi_row_sum (aten/src/ATen/native/cpu/SumKernel.cpp, torch 2.x): a four-level cascade
    with level_step = 2^max(4, ceil_log2(size)/4).  `a` is (size, ...) float32; the sum runs over axis 0
    elementwise for every trailing index."""
    size = a.shape[0]
    level_power = max(4, _ceil_log2(size) // 4)
    level_step = 1 << level_power
    level_mask = level_step - 1
    acc = np.zeros((4,) + a.shape[1:], np.float32)
    i = 0
    while i + level_step <= size:
        for _ in range(level_step):
            acc[0] = acc[0] + a[i]
            i += 1
        for j in range(1, 4):
            acc[j] = acc[j] + acc[j - 1]
            acc[j - 1] = 0
            if (i & (level_mask << (j * level_power))) != 0:
                break
    while i < size:
        acc[0] = acc[0] + a[i]
        i += 1
    for j in range(1, 4):
        acc[0] = acc[0] + acc[j]
    return acc[0]


def cascade_row_sum(q: np.ndarray) -> np.ndarray:
    """Bit-exact emulation of `torch.sum(q, dim=-1)` for a contiguous float32 (B,N) CPU tensor
    (resamplers.py:33).  ATen's vectorised inner sum = 8 SIMD lanes x 4 ILP accumulators, each a
    sequential cascade over elements congruent mod 32, then ILP fold, scalar tail, lane fold.
    Verified equal to torch.sum on the build host for N in {1..8, 37, 100, 128, 1000, 1024, 4096, 5000}."""
    q = np.ascontiguousarray(q, np.float32)
    Bn, n = q.shape
    lanes, ilp = (8, 4) if n >= 8 else (1, 4)
    vec = n // lanes
    rows = vec // ilp
    main = q[:, :rows * ilp * lanes].reshape(Bn, rows, ilp, lanes).transpose(1, 0, 2, 3)
    ps = _multi_row_sum(main)  # (B, ilp, lanes)
    for i in range(rows * ilp, vec):
        ps[:, 0] = q[:, i * lanes:(i + 1) * lanes] + ps[:, 0]
    for k in range(1, ilp):
        ps[:, 0] = ps[:, 0] + ps[:, k]
    out = np.zeros(Bn, np.float32)
    for k in range(vec * lanes, n):
        out = out + q[:, k]
    for k in range(lanes):
        out = out + ps[:, 0, k]
    return out


def soft_resample(particles, probs, alpha: float, offsets, markers=None):
    """soft_resampler with the U(0,1/N) offsets injected.  Returns (particles', probs', flat idx int64).

    Index semantics (resamplers.py:42-52): idx[b,i] = #{j : markers[b,i] > cum[b,j]} with
    cum = cumsum(q) accumulated in fp64 and rounded per prefix to fp32 (ATen CPU cumsum), cum[:, -1] = 1.
    """
    assert 0.0 < alpha <= 1.0
    B, N = probs.shape
    uniform = torch.ones(B, N) / N
    if alpha < 1.0:
        q = torch.stack((probs * alpha, uniform * (1.0 - alpha)), dim=-1).sum(dim=-1)
        s = q.sum(dim=-1, keepdim=True)
        s_exact = torch.from_numpy(cascade_row_sum(q.detach().numpy()))[:, None]
        s = s + (s_exact - s).detach()  # canonical (host independent) value, same autograd graph
        q = q / s
        w_is = probs / q
    else:
        q = probs
        w_is = uniform
    if markers is None:
        markers = torch.linspace(0.0, (N - 1.0) / N, N)
    mk = (offsets[:, None].float() + markers[None, :]).numpy()
    cum = np.cumsum(q.detach().numpy().astype(np.float64), axis=1).astype(np.float32)
    cum[:, -1] = 1.0
    samples = np.empty((B, N), np.int64)
    for b in range(B):  # strict '>' count; rows are sorted except possibly the forced last entry
        samples[b] = np.searchsorted(cum[b, :-1], mk[b], side="left") + (mk[b] > cum[b, -1])
    idx = torch.from_numpy(samples) + N * torch.arange(B)[:, None]
    p_res = particles.reshape(B * N, -1)[idx, :]
    w_res = w_is.reshape(B * N)[idx]
    w_res = w_res / w_res.sum(dim=-1, keepdim=True)
    return p_res, w_res, idx


# ----------------------------------------------------------------------------------------
# entropy-regularised OT resampling -- resamplers/resamplers.py:62-277
# ----------------------------------------------------------------------------------------
def _softmin(eps, cost, h):
    """resamplers.py:94-110: -eps * LSE_j(h_j - C_ij/eps); eps is (B,) or 0-dim."""
    e = eps.reshape(-1, 1, 1)
    return -eps.reshape(-1, 1) * torch.logsumexp(h[:, None, :] - cost / e, dim=2)


def ot_transport(x, logw, eps: float = 0.1, scaling: float = 0.75, threshold: float = 1e-3, max_iter: int = 100):
    """transport_function, resamplers.py:211-227, restated for x == y (one symmetric cost matrix, the two
    live potential chains a_y / b_x; the a_x / b_y chains of the reference never reach the output).
    Follows the reference's dtype flow: everything after `diameter` is float64.  Returns (T fp64, iterations)."""
    B, N, d = x.shape
    eps_t = torch.tensor(eps, dtype=torch.float)
    log_n = torch.log(torch.tensor(float(N)))
    log_beta = -log_n * torch.ones_like(logw)
    centered = x - x.mean(dim=1, keepdim=True)
    diam = x.std(dim=1, unbiased=False).max(dim=-1)[0]
    diam = torch.where(diam == 0.0, 1.0, diam.double())
    scale = diam.reshape(-1, 1, 1) * torch.sqrt(torch.tensor(d))
    sx = centered / scale
    cost = torch.cdist(sx, sx, p=2.0) ** 2 / 2.0
    # max_min, resamplers.py:87-91 (mirrored literally, including the max-then-min asymmetry)
    max_max = sx.max(dim=1)[0].max(dim=1)[0]
    min_min = torch.minimum(sx.max(dim=1)[0].min(dim=1)[0], sx.min(dim=1)[0].min(dim=1)[0])
    eps_run = (max_max - min_min) ** 2
    s2 = scaling ** 2
    a_y = _softmin(eps_run, cost, logw)
    b_x = _softmin(eps_run, cost, log_beta)
    cont = torch.ones(B, dtype=torch.bool)
    it = 0
    while it < max_iter - 1 and bool(cont.all()):
        er = eps_run.reshape(-1, 1)
        at_y = _softmin(eps_run, cost, logw + b_x / er)
        bt_x = _softmin(eps_run, cost, log_beta + a_y / er)
        a_new, b_new = (a_y + at_y) / 2, (b_x + bt_x) / 2
        local = ((a_new - a_y).abs().max(dim=1)[0] > threshold) | ((b_new - b_x).abs().max(dim=1)[0] > threshold)
        eps_new = torch.maximum(eps_run * s2, eps_t)
        cont = (eps_new < eps_run) | local
        a_y, b_x, eps_run, it = a_new, b_new, eps_new, it + 1
    f = _softmin(eps_t, cost, logw + b_x / eps_t)       # alpha = final_a_y
    g = _softmin(eps_t, cost, log_beta + a_y / eps_t)   # beta  = final_b_x
    temp = (f[:, :, None] + g[:, None, :] - cost) / eps_t
    temp = temp - torch.logsumexp(temp, dim=1, keepdim=True) + log_n
    T = torch.exp(temp + logw[:, None, :])
    return T, it + 2


def ot_resample(particles, weights, eps=0.1, scaling=0.75, threshold=1e-3, max_iter=100, return_iters=False):
    """resampler_ot / OT_resampling, resamplers.py:62-70, 267-277.  The transport matrix is a constant for
    autograd (transport.backward returns None, resamplers.py:240-245): d particles' / d particles = T only."""
    B, N, _ = particles.shape
    with torch.no_grad():
        T, iters = ot_transport(particles.detach(), weights.detach().log(), eps, scaling, threshold, max_iter)
    p_res = torch.matmul(T.float(), particles.float())
    w_res = torch.ones_like(weights) / float(N)
    idx = torch.arange(N)[None, :] + N * torch.arange(B)[:, None]
    if return_iters:
        return p_res, w_res, idx, iters
    return p_res, w_res, idx


# ----------------------------------------------------------------------------------------
# the filter loop -- DPFs.py:144-216
# ----------------------------------------------------------------------------------------
def filtering(cfg, W, init_particles, start_vel, vel_input, enc_seq, noise_seq, offsets_seq, force_resample=None):
    """DPF.filtering_pos with every random draw injected and the CNN encoder replaced by precomputed
    encodings enc_seq (B,T,h).  cfg keys: NF, NF_cond, measurement, resampler ('soft'|'ot'), alpha,
    pos_noise, eps, scaling, threshold, max_iter.  `force_resample`: None = the reference's ESS gate
    (DPFs.py:163-165), True/False = override (benchmarks state which).
    Returns dict of the per-step lists (B,T,...) plus obs_likelihood and the gate decisions."""
    B, N, _ = init_particles.shape
    T = enc_seq.shape[1]
    particles = init_particles
    probs = normalize_log_probs(torch.full((B, N), -math.log(N)))
    vel = start_vel
    out = {k: [] for k in ("particles", "probs", "noise", "lki", "index", "jac", "prior")}
    fired, obs_lik = [], 0.0
    for t in range(T):
        idx = torch.arange(N)[None, :] + N * torch.arange(B)[:, None]
        ess = torch.mean(1.0 / torch.sum(probs ** 2, dim=-1))
        fire = bool(ess < 0.5 * N) if force_resample is None else bool(force_resample)
        fired.append(fire)
        if fire:
            if cfg["resampler"] == "soft":
                particles, pr, idx = soft_resample(particles, probs, cfg["alpha"], offsets_seq[:, t])
            else:
                particles, pr, idx = ot_resample(particles, probs, cfg["eps"], cfg["scaling"], cfg["threshold"],
                                                 cfg["max_iter"])
            logw = pr.log()
        else:
            logw = probs.log()
        noise = noise_seq[:, t]
        x_phys = motion_update(particles, vel, noise)
        vel = vel_input[:, t]
        if cfg["NF"]:
            x_dyn, jac = nf_dynamic(W["dyn"], x_phys)
        else:
            x_dyn, jac = x_phys, torch.zeros(B, N)
        prop, lki, prior, propose_log = proposal_likelihood(cfg, W, x_dyn, x_phys, enc_seq[:, t], noise, jac)
        logw = logw + lki + prior - propose_log
        particles = prop
        obs_lik = obs_lik + logw.mean()
        probs = normalize_log_probs(logw) + 1e-12
        for k, v in (("particles", particles), ("probs", probs), ("noise", noise), ("lki", lki), ("index", idx),
                     ("jac", jac), ("prior", prior)):
            out[k].append(v)
    res = {k: torch.stack(v, dim=1) for k, v in out.items()}
    res["obs_likelihood"] = obs_lik
    res["fired"] = fired
    return res


def supervised_rmse(particle_list, probs_list, true_xy):
    """supervised_loss with mask=1 / eval branch, losses.py:18-31."""
    pred = torch.sum(particle_list * probs_list[..., None], dim=2)
    return torch.sqrt(torch.mean((pred - true_xy) ** 2)), pred


# ----------------------------------------------------------------------------------------
# Semi-supervised objective: block pseudo-likelihood (losses.py:33-70, 73-110)
# ----------------------------------------------------------------------------------------
def block_density(weights, lik, prior, index, block_len: int):
    """compute_block_density_nf, losses.py:37-70, written per trajectory-particle instead of as (B*N,) gather chains:
    for the end k of every block each particle follows its flat ancestor pointer back through the block and adds the
    prior + likelihood terms it meets (losses.py:51-64); the running sum is never reset between blocks (losses.py:47, 64);
    Q[b] = mean over blocks of sum_n w[b,k,n] * running[b,n] (losses.py:65-68).  weights / lik / prior (B,T,N), index (B,T,N)
    flat int64 ancestors (j + N*b).  Differentiable (torch) in weights, lik and prior."""
    B, T, N = weights.shape
    flat = lambda lst, j: lst[:, j, :].reshape(B * N)
    running = torch.zeros(B, N, dtype=weights.dtype)
    Q = torch.zeros(B, dtype=weights.dtype)
    n_blocks = 0
    for k in range(block_len - 1, T, block_len):
        pos = torch.arange(B * N)                           # where each walker currently stands (flat)
        for j in range(k, k - block_len, -1):
            running = (running + flat(prior, j)[pos].reshape(B, N)) + flat(lik, j)[pos].reshape(B, N)
            pos = flat(index, j)[pos]
        Q = Q + (weights[:, k, :] * running).sum(-1)
        n_blocks += 1
    return Q / n_blocks


def block_prior_from_noise(noise, std_pos: float, std_vel: float):
    """log-prior term of compute_block_density, losses.py:93-96: two diagonal Gaussians over noise[..., :2] and noise[..., 2:]
    (the second sum is empty when the noise has two columns)."""
    log_c = -0.5 * math.log(2 * math.pi)
    pos, vel = noise[..., :2], noise[..., 2:]
    return (2 * log_c - 2 * math.log(std_pos) - (pos ** 2 / (2 * std_pos ** 2)).sum(-1)) + \
           (2 * log_c - 2 * math.log(std_vel) - (vel ** 2 / (2 * std_vel ** 2)).sum(-1))


def supervised_loss_train(particle_list, probs_list, true_state, mask, labeled_ratio: float = 1.0):
    """supervised_loss, train branch with a label mask, losses.py:18-27."""
    pred = torch.sum(particle_list * probs_list[..., None], dim=2)
    return torch.sqrt(torch.mean(mask[:, :, None] * (pred - true_state[:, :, :2]) ** 2) / labeled_ratio), pred
