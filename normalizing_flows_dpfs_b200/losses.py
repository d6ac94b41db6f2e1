"""Losses consuming the per-step lists of the filter (reference losses.py), run once per batch after the filter loop.
On the device both hot ones are kernels: the supervised loss picks up the per-step predictions a fused kernel formed inside the
loop, the block pseudo-likelihood is one launch that walks the ancestry of every block (csrc/losses.cu).  Lists on the CPU
(oracle-style checks of the mirror against the reference, no GPU involved) take the reference's own gather chain."""
import math

import torch
from torch import nn


def autoencoder_loss(image, train, encoder, decoder):
    """MSE of decoder(encoder(img)) over all B*T frames (reference losses.py:5-16)."""
    b, t, c, h, w = image.shape
    frames = image.reshape(b * t, c, h, w)
    return nn.functional.mse_loss(decoder(encoder(frames)), frames)


def supervised_loss(particle_list, particle_weight_list, true_state, mask, train, labeledRatio=1.0, group=None):
    """RMSE between the weighted particle mean and the true position (reference losses.py:18-31).  Lists that come straight from
    DPF.filtering_pos carry the per-step predictions a fused kernel already formed (sum_n w_n x_n, with backward): they are used
    instead of multiplying and reducing the (B,T,N,2) lists again."""
    fusedp = getattr(particle_list, "_nfdpf_pred", None)
    if fusedp is not None and fusedp[1] is particle_weight_list:
        prediction = fusedp[0]
    else:
        prediction = (particle_list * particle_weight_list[..., None]).sum(dim=2)
    err2 = (prediction - true_state[:, :, :2]) ** 2
    if group is not None:
        # batch-sharded run (extension): the mean under the root is taken over the GLOBAL batch -- the squared-error sum is
        # all-reduced differentiably, so every rank returns the same RMSE and backpropagates its local part of the global gradient
        # (reduce the parameter gradients with SUM, distributed.GradBucket.allreduce(average=False))
        from .distributed import global_sum
        w = err2 if not train else mask[:, :, None] * err2
        count = global_sum(torch.tensor(float(err2.numel()), device=err2.device), group)
        mean = global_sum(w.sum(), group) / count
        return torch.sqrt(mean if not train else mean / labeledRatio), prediction
    if not train:
        return torch.sqrt(err2.mean()), prediction
    if labeledRatio > 0:
        return torch.sqrt((mask[:, :, None] * err2).mean() / labeledRatio), prediction
    return 0


def _trace_blocks(weights, seq_len, block_len, terms_at):
    """Block pseudo-likelihood: at the end k of each block, walk the ancestry back through index_list and add the
    per-step log terms of every ancestor, then weight by the step-k particle weights (reference losses.py:37-106).
    Like the reference, the running sum is NOT reset between blocks."""
    q_total, running, n_blocks = 0.0, 0.0, 0
    for k in range(seq_len):
        if (k + 1) % block_len:
            continue
        anc = None
        for j in range(k, k - block_len, -1):
            term, idx_j = terms_at(j, anc)
            running = running + term
            anc = idx_j if anc is None else idx_j.reshape(-1)[anc]
        q_total = q_total + (weights[:, k, :] * running).sum(dim=-1)
        n_blocks += 1
    return q_total / n_blocks


def compute_block_density_nf(particle_weight_list, noise_list, likelihood_list, index_list, jac_list, prior_list, block_len=10):
    B, T, N = particle_weight_list.shape
    if particle_weight_list.is_cuda:
        from . import ops
        return ops.block_density(particle_weight_list, likelihood_list, prior_list, index_list, block_len)

    def terms_at(j, anc):
        lik, prior, idx = likelihood_list[:, j, :], prior_list[:, j, :], index_list[:, j, :]
        if anc is not None:
            lik, prior = lik.reshape(-1)[anc], prior.reshape(-1)[anc]
        return prior + lik, idx

    return _trace_blocks(particle_weight_list, T, block_len, terms_at)


def compute_block_density(particle_weight_list, noise_list, likelihood_list, index_list, block_len=10, std_pos=1.0, std_vel=1.0):
    B, T, N = particle_weight_list.shape
    log_c = -0.5 * math.log(2 * math.pi)
    if particle_weight_list.is_cuda:
        from . import ops
        pos, vel = noise_list[..., :2], noise_list[..., 2:]
        log_prior = (2 * log_c - 2 * math.log(std_pos) - (pos ** 2 / (2 * std_pos ** 2)).sum(-1)) + \
                    (2 * log_c - 2 * math.log(std_vel) - (vel ** 2 / (2 * std_vel ** 2)).sum(-1))
        return ops.block_density(particle_weight_list, likelihood_list, log_prior, index_list, block_len)

    def terms_at(j, anc):
        lik, idx = likelihood_list[:, j, :], index_list[:, j, :]
        noise = noise_list[:, j]
        if anc is not None:
            lik, noise = lik.reshape(-1)[anc], noise.reshape(B * N, -1)[anc, :]
        pos, vel = noise[..., :2], noise[..., 2:]
        log_prior = (2 * log_c - 2 * math.log(std_pos) - (pos ** 2 / (2 * std_pos ** 2)).sum(-1)) + \
                    (2 * log_c - 2 * math.log(std_vel) - (vel ** 2 / (2 * std_vel ** 2)).sum(-1))
        return log_prior + lik, idx

    return _trace_blocks(particle_weight_list, T, block_len, terms_at)


def pseudolikelihood_loss_nf(particle_weight_list, noise_list, likelihood_list, index_list, jac_list, prior_list, block_len=10):
    return -1.0 * torch.mean(compute_block_density_nf(particle_weight_list, noise_list, likelihood_list, index_list, jac_list,
                                                       prior_list, block_len))


def pseudolikelihood_loss(particle_weight_list, noise_list, likelihood_list, index_list, block_len=10, std_pos=1.0, std_vel=1.0):
    return -1.0 * torch.mean(compute_block_density(particle_weight_list, noise_list, likelihood_list, index_list, block_len, std_pos,
                                                    std_vel))
