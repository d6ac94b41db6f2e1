"""Model builders and per-step glue with the reference's names (reference model/models.py).

Hot path: motion_update, nf_dynamic_model, normalising_flow_propose, proposal_likelihood and the
Gaussian / cosine / conditional-RealNVP measurement models run on the fused sm_100a kernels.  The CNN image
encoder / decoder and the NN / cGlow likelihoods are outside the hot path and stay stock PyTorch modules."""
import torch
from torch import nn
from torch.distributions import MultivariateNormal

from .. import ops
from ..nf.flows import FCNN, RealNVP, RealNVP_cond, pack_parameters, _PackCache  # noqa: F401
from ..nf.models import NormalizingFlowModel, NormalizingFlowModel_cond
from ..utils import et_distance  # noqa: F401

device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


# ------------------------------------------------------------------------------- out-of-scope CNN (cuDNN) parts
def _conv_stack(chans):
    layers = []
    for cin, cout in zip(chans[:-1], chans[1:]):
        layers += [nn.Conv2d(cin, cout, kernel_size=4, stride=2, padding=1, bias=False), nn.ReLU(True), nn.BatchNorm2d(cout)]
    return layers


def build_encoder(hidden_size):
    """3x128x128 image -> hidden_size encoding: five stride-2 conv blocks + Linear (reference models.py:10-34)."""
    return nn.Sequential(*_conv_stack([3, 16, 32, 64, 128, 256]), nn.Flatten(), nn.Linear(256 * 4 * 4, hidden_size))


def build_encoder_cglow(hidden_size):
    return nn.Sequential(*_conv_stack([3, 16, 32, 64, 128, 256]), nn.Flatten(), nn.Linear(256 * 4 * 4, 192))


def _deconv_stack(chans):
    layers = []
    for cin, cout in zip(chans[:-1], chans[1:]):
        layers += [nn.ConvTranspose2d(cin, cout, kernel_size=4, padding=1, stride=2, bias=False), nn.ReLU(True), nn.BatchNorm2d(cout)]
    return layers


def build_decoder(hidden_size):
    """Mirror of the encoder (reference models.py:62-89): Linear -> 256x4x4 -> transposed-conv blocks -> 3x128x128 in [0,1]."""
    return nn.Sequential(nn.Linear(hidden_size, 256 * 4 * 4), nn.Unflatten(-1, (256, 4, 4)), *_deconv_stack([256, 128, 64, 32, 16]),
                         nn.ConvTranspose2d(16, 3, kernel_size=4, padding=1, stride=2, bias=False), nn.BatchNorm2d(3), nn.Sigmoid())


def build_decoder_cglow(hidden_size):
    return build_decoder(192)


def build_likelihood(hidden_size, state_dim):
    return nn.Sequential(nn.Linear(2 * hidden_size, 64), nn.ReLU(True), nn.Linear(64, 64), nn.ReLU(True), nn.Linear(64, 1), nn.Sigmoid())


def build_particle_encoder(hidden_size, state_dim):
    return nn.Sequential(nn.Linear(state_dim, 16), nn.ReLU(), nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, hidden_size))


def build_particle_encoder_cglow(hidden_size, state_dim):
    return nn.Sequential(nn.Linear(state_dim, 16), nn.ReLU(), nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 192))


def build_transition_model(state_dim):
    return nn.Sequential(nn.Linear(state_dim, 64), nn.ReLU(), nn.Linear(64, 64), nn.ReLU(), nn.Linear(64, state_dim))


# ------------------------------------------------------------------------------------------- flow builders
def build_conditional_nf(n_sequence, hidden_size, state_dim, init_var=0.01, prior_mean=0.0, prior_std=1.0):
    """n_sequence conditional couplings with N(prior_mean, prior_std^2 I) base density (reference models.py:161-172)."""
    flows = [RealNVP_cond(dim=state_dim, obser_dim=hidden_size) for _ in range(n_sequence)]
    for f in flows:
        f.zero_initialization(var=init_var)
    prior = MultivariateNormal(torch.zeros(state_dim).to(device) + prior_mean, torch.eye(state_dim).to(device) * prior_std ** 2)
    model = NormalizingFlowModel_cond(prior, flows, device=device)
    model.prior_mean, model.prior_std = float(prior_mean), float(prior_std)
    return model


def build_dyn_nf(n_sequence, hidden_size, state_dim, init_var=0.01):
    flows = [RealNVP(dim=state_dim) for _ in range(n_sequence)]
    for f in flows:
        f.zero_initialization(var=init_var)
    prior = MultivariateNormal(torch.zeros(state_dim).to(device), torch.eye(state_dim).to(device))
    return NormalizingFlowModel(prior, flows, device=device)


def build_conditional_glow(args):
    raise NotImplementedError("conditional Glow (--measurement CGLOW) is outside the accelerated hot path (SURVEY section 2, row 10)")


# ---------------------------------------------------------------------------------------------- per-step glue
def motion_update(particles, vel, pos_noise=20.0, noise=None):
    """x + vel_b + eps, eps ~ N(0, pos_noise^2) (reference models.py:191-204).  The draw comes from the CPU generator
    like the reference's unless `noise` (B,N,2) is injected."""
    B, N, _ = particles.shape
    if noise is None:
        noise = torch.normal(mean=0.0, std=pos_noise, size=(B, N, 2)).to(particles.device, non_blocking=True)
    return particles + vel[:, None, :] + noise, noise


def _moments_context(x, mean=None, std=None, lead=None):
    """(B, [lead] + 2d) row context [lead | mean | std]; moments detached like the reference (models.py:309-313)."""
    B, N, d = x.shape
    w = 0 if lead is None else lead.shape[-1]
    ctx = torch.empty(B, w + 2 * d, dtype=torch.float32, device=x.device)
    if lead is not None:
        ctx[:, :w] = lead.detach()
    if mean is None:
        ops.row_moments(x, ctx, w)
    else:
        ctx[:, w:w + d] = mean.detach().reshape(B, d)
        ctx[:, w + d:] = std.detach().reshape(B, d)
    return ctx


def nf_dynamic_model(dynamical_nf, dynamic_particles, jac_shape, NF=False, forward=False, mean=None, std=None):
    """Dynamics flow on (B,N,d) particles with context [mean_N, std_N] (reference models.py:305-332).
    Returns (particles, jac = -log_det)."""
    if not NF:
        return dynamic_particles, torch.zeros(jac_shape, device=dynamic_particles.device)
    ctx = _moments_context(dynamic_particles, mean, std)
    return dynamical_nf.run_stack(dynamic_particles, row_ctx=ctx, inverse=not forward, neg_logdet=True)


def normalising_flow_propose(cond_model, particles_pred, obs, flow=RealNVP_cond, n_sequence=2, hidden_dimension=8, obser_dim=None):
    """Proposal flow with context [obs encoding, mean_N, std_N] (reference models.py:334-356)."""
    ctx = _moments_context(particles_pred, lead=obs)
    return cond_model.run_stack(particles_pred, row_ctx=ctx, inverse=True, neg_logdet=True)


def proposal_likelihood(cond_model, dynamical_nf, measurement_model, particles_dynamic, particles_physical, encodings, noise,
                        jac_dynamic, NF, NF_cond, prototype_density):
    """Proposal, measurement log-likelihood and the prior / proposal log-densities (reference models.py:358-379)."""
    enc_detached = encodings.detach()
    if NF_cond:
        propose_particle, jac_prop = normalising_flow_propose(cond_model, particles_dynamic, enc_detached)
        if NF:
            back, jac_back = nf_dynamic_model(dynamical_nf, propose_particle, jac_dynamic.shape, NF=True, forward=True,
                                              mean=particles_physical.mean(dim=1, keepdim=True),
                                              std=particles_physical.std(dim=1, keepdim=True))
            prior_log = prototype_density(back - (particles_physical - noise)) - jac_back
        else:
            prior_log = prototype_density(propose_particle - (particles_physical - noise))
        propose_log = prototype_density(noise) + jac_dynamic + jac_prop
    else:
        propose_particle = particles_dynamic
        prior_log = prototype_density(noise) + jac_dynamic
        propose_log = prototype_density(noise) + jac_dynamic
    lki_log = measurement_model(encodings, propose_particle)
    return propose_particle, lki_log, prior_log, propose_log


# ------------------------------------------------------------------------------------------ measurement models
class _FusedMeasurement(nn.Module):
    mode = None

    def __init__(self, particle_encoder):
        super().__init__()
        self.particle_encoder = particle_encoder
        self._pe_cache = _PackCache()

    def _check_encoder(self):
        lin = [m for m in self.particle_encoder if isinstance(m, nn.Linear)]
        shape = [tuple(m.weight.shape) for m in lin]
        if shape != [(16, 2), (32, 16), (32, 32)]:
            raise ValueError("fused measurement kernels need the 2-16-32-32 particle encoder (hiddensize 32); got %s" % (shape,))

    def pe_packed(self):
        self._check_encoder()
        return self._pe_cache.get([self.particle_encoder])

    def cnf_packed(self):
        return None

    def params(self):
        return 0.0, 1.0, 2

    def forward(self, encodings, update_particles):
        p0, p1, n_flows = self.params()
        return ops.measure(self.pe_packed(), self.cnf_packed(), encodings, update_particles, self.mode, n_flows, p0, p1)

    def forward_update(self, encodings, update_particles, logw_prev, prior_log, propose_log, add_eps=1e-12, out=None, want_pred=False):
        """measurement + DPFs.py:187-192 in one kernel: (lki, logw, probs, row_sum_logw, ess_inv[, prediction of losses.py:22])."""
        p0, p1, n_flows = self.params()
        return ops.measure_update(self.pe_packed(), self.cnf_packed(), encodings, update_particles, logw_prev, prior_log, propose_log,
                                  self.mode, n_flows, p0, p1, add_eps, out, want_pred)


class measurement_model_cosine_distance(_FusedMeasurement):
    mode = "cos"


class measurement_model_Gaussian(_FusedMeasurement):
    mode = "gaussian"

    def __init__(self, particle_encoder, gaussian_distribution):
        super().__init__(particle_encoder)
        self.gaussian_distribution = gaussian_distribution
        loc, cov = gaussian_distribution.loc, gaussian_distribution.covariance_matrix
        var = torch.diagonal(cov)
        if not (torch.allclose(loc, loc[0].expand_as(loc)) and torch.allclose(cov, torch.diag(var)) and torch.allclose(var, var[0].expand_as(var))):
            raise ValueError("the fused Gaussian likelihood supports MultivariateNormal(c*1, s^2*I) (DPFs.py:84-86)")
        self._p = (float(loc[0]), float(var[0]) ** 0.5)

    def params(self):
        return self._p[0], self._p[1], 2


class measurement_model_cnf(_FusedMeasurement):
    mode = "CRNVP"

    def __init__(self, particle_encoder, CNF):
        super().__init__(particle_encoder)
        self.CNF = CNF

    def cnf_packed(self):
        return self.CNF.packed()

    def params(self):
        mean = getattr(self.CNF, "prior_mean", None)
        if mean is None:
            mean, std = float(self.CNF.prior.loc[0]), float(self.CNF.prior.covariance_matrix[0, 0]) ** 0.5
        else:
            std = self.CNF.prior_std
        return mean, std, len(self.CNF.flows)


class measurement_model_NN(_FusedMeasurement):
    """Learned likelihood head on [obs encoding | particle encoding] (reference models.py:221-235): Sigmoid MLP 64-64-64-1, then log.
    Mode 3 of the fused measurement kernels: the head's 64-wide layers are tcgen05 rounds behind the encoder's, forward and backward;
    its weight gradients are register tiles on the CUDA cores (csrc/measure.cu, measure_bwd_nn_kernel)."""
    mode = "NN"

    def __init__(self, particle_encoder, likelihood_estimator):
        super().__init__(particle_encoder)
        self.likelihood_estimator = likelihood_estimator
        self._head_cache = _PackCache()

    def cnf_packed(self):     # the kernels' second parameter slot carries the packed head
        lin = [m for m in self.likelihood_estimator if isinstance(m, nn.Linear)]
        shape = [tuple(m.weight.shape) for m in lin]
        if shape != [(64, 64), (64, 64), (1, 64)]:
            raise ValueError("the fused NN likelihood needs build_likelihood's 64-64-64-1 head (hiddensize 32); got %s" % (shape,))
        return self._head_cache.get([self.likelihood_estimator])


class measurement_model_cglow(nn.Module):
    def __init__(self, particle_encoder, CGLOW):
        super().__init__()
        raise NotImplementedError("conditional Glow likelihood is outside the accelerated hot path (SURVEY section 2, row 10)")
