"""Filter utilities with the reference's names and argument meaning (reference utils.py).

`normalize_log_probs` runs the fused max-shift softmax kernel (nfdpf_weight_update_fwd); the rest is one-off
host glue (particle initialisation, checkpoint dicts) kept in plain PyTorch."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import ops

device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def et_distance(encoding_input, e_t):
    """1 - cosine similarity along the last axis (reference utils.py:8-15)."""
    a = F.normalize(encoding_input, p=2, dim=-1, eps=1e-12)
    b = F.normalize(e_t, p=2, dim=-1, eps=1e-12)
    return 1.0 - (a * b).sum(dim=-1)


class compute_normal_density(nn.Module):
    """Diagonal-Gaussian log-density of the motion noise (reference utils.py:17-37): the first two noise
    dimensions use `pos_noise`, any remaining ones `vel_noise`."""

    def __init__(self, pos_noise=1.0, vel_noise=1.0):
        super().__init__()
        self.pos_noise = pos_noise
        self.vel_noise = vel_noise

    def forward(self, noise, std_pos=None, std_vel=None):
        sp = self.pos_noise if std_pos is None else std_pos
        sv = self.vel_noise if std_vel is None else std_vel
        d = noise.shape[-1]
        const = -0.5 * d * math.log(2.0 * math.pi) - 2.0 * math.log(sp) - (d - 2) * math.log(sv)
        quad = (noise[..., :2] ** 2).sum(-1) / (2.0 * sp ** 2)
        if d > 2:
            quad = quad + (noise[..., 2:] ** 2).sum(-1) / (2.0 * sv ** 2)
        return const - quad


def normalize_log_probs(probs):
    """softmax over the particle axis (reference utils.py:39-44), one fused kernel with backward."""
    return ops.weight_update(probs)[1]


def particle_initialization(start_state, width, num_particles, state_dim=2, init_with_true_state=False):
    """Initial particle cloud and log-weights (reference utils.py:46-62).  Draws come from the CPU generator in
    the reference's order -- including the velocity draw it never uses -- so seeded runs line up."""
    batch = start_state.shape[0]
    dev = start_state.device
    if init_with_true_state:
        particles = start_state[:, None, :] + torch.randn(batch, num_particles, state_dim).to(dev)
    else:
        half = width / 2.0
        particles = (2.0 * half) * torch.rand(batch, num_particles, 2).to(dev) - half
        torch.randn(batch, num_particles, 2)  # unused velocity draw of the reference (utils.py:58)
    log_w = torch.log(torch.ones(batch, num_particles, device=dev) / num_particles)
    return particles, log_w


def _set_trainable(model, flag):
    for tensor in model.parameters():
        tensor.requires_grad = flag


def freeze_model(model):
    _set_trainable(model, False)


def unfreeze_model(model):
    _set_trainable(model, True)


_CKPT_PARTS = (("model", lambda m: m), ("model_optim", lambda m: m.optim), ("model_optim_scheduler", lambda m: m.optim_scheduler))


def checkpoint_state(model, epoch):
    """{'model', 'model_optim', 'model_optim_scheduler', 'epoch'}: the checkpoint layout of the reference (utils.py:72-79)."""
    state = {key: pick(model).state_dict() for key, pick in _CKPT_PARTS}
    state["epoch"] = epoch
    return state


def load_model(model, ckpt_e2e):
    for key, pick in _CKPT_PARTS:
        pick(model).load_state_dict(ckpt_e2e[key])
