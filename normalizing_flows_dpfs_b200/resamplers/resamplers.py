"""Resamplers with the reference's call signatures (reference resamplers/resamplers.py).

soft_resampler  -> nfdpf_soft_resample_{fwd,bwd}: block scan + search, no (B,N,N) temporaries, indices bit-exact
                   with the reference CPU path for the same uniforms.
resampler_ot    -> nfdpf_sinkhorn_ot_{fwd,bwd}: log-domain Sinkhorn recomputing cost tiles from shared memory."""
import torch
import torch.nn as nn

from .. import ops

device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


class resampler(nn.Module):
    def __init__(self, param):
        super().__init__()
        if param.resampler_type == "ot":
            self.kargs = {"eps": param.epsilon, "scaling": param.scaling, "threshold": param.threshold, "max_iter": param.max_iter,
                          "device": device}
            self.resampling = resampler_ot
        elif param.resampler_type == "soft":
            self.kargs = {"num_resampled": param.num_particles, "index": True, "alpha": param.alpha, "device": device}
            self.resampling = soft_resampler
        else:
            raise ValueError("resampler_type must be 'ot' or 'soft', got %r" % (param.resampler_type,))

    def forward(self, particles, particle_probs):
        return self.resampling(particles, particle_probs, **self.kargs)


_marker_cache = {}


def _markers(n, dev):
    key = (n, str(dev))
    if key not in _marker_cache:  # torch's own linspace: not bitwise arange/n unless n is a power of two
        _marker_cache[key] = torch.linspace(0.0, (n - 1.0) / n, n).to(dev)
    return _marker_cache[key]


def soft_resampler(particles, particle_probs, alpha, num_resampled, index=True, device="cuda", random_offset=None, want_log=False,
                   gate=None, out=None):
    """Soft (mixture-with-uniform) systematic resampling.  `random_offset` (B,) may be injected; by default it is
    drawn exactly like the reference does (CPU generator, U(0, 1/N), resamplers.py:43).  `gate` (device int32 flag, the
    filter loop's ESS decision) and `out` (pre-allocated outputs) are extensions used by DPF.filtering_pos."""
    assert 0.0 < alpha <= 1.0
    batch, n = particle_probs.shape
    if num_resampled != n:
        raise ValueError("soft_resampler resamples N -> N particles (num_resampled=%d, N=%d)" % (num_resampled, n))
    if random_offset is None:
        random_offset = torch.FloatTensor(batch).uniform_(0.0, 1.0 / num_resampled)
    off = random_offset.to(particles.device, non_blocking=True)
    res = ops.soft_resample(particles, particle_probs, off, _markers(n, particles.device), alpha, want_log, gate, out)
    if want_log:  # (particles, probs, idx, log probs): the filter loop's fused path
        return res
    return res if index else res[:2]


def resampler_ot(particles, weights, eps=0.1, scaling=0.75, threshold=1e-3, max_iter=100, device="cuda",
                 flag=torch.tensor(True, requires_grad=False)):
    batch, n, _ = particles.shape
    p, w, _ = OT_resampling(particles, logw=weights.log(), eps=eps, scaling=scaling, threshold=threshold, max_iter=max_iter, n=n,
                            device=device, flag=flag)
    idx = torch.arange(batch * n, device=particles.device, dtype=torch.int64).reshape(batch, n)
    return p, w, idx


def OT_resampling(x, logw, eps, scaling, threshold, max_iter, n, device="cuda", flag=torch.tensor(True, requires_grad=False)):
    """Entropy-regularised OT resampling: particles' = T x with T the Sinkhorn transport plan; uniform weights.
    Like the reference (resamplers.py:234-245) the plan is a constant for autograd: d particles' / d particles = T."""
    if not bool(flag):
        return x.float(), logw.exp().float(), logw.float()
    p = ops.ot_resample(x, logw, eps, scaling, threshold, max_iter)
    w = torch.full_like(logw, 1.0 / n, dtype=torch.float32)
    return p, w, w.log()
