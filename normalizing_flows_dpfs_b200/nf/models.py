"""Flow containers (reference nf/models.py:5-66).  When every flow is one of the fused coupling types the whole
chain runs as a single kernel launch (n_flows couplings + summed log-det); otherwise flows run one by one."""
import torch
from torch import nn

from .. import ops
from .flows import RealNVP, RealNVP_cond, _PackCache


class _FlowChain(nn.Module):
    _conditional = False

    def __init__(self, prior, flows, device="cuda"):
        super().__init__()
        self.prior = prior
        self.device = device
        self.flows = nn.ModuleList(flows).to(self.device)
        self._cache = _PackCache()

    def _fusable(self):
        kind = RealNVP_cond if self._conditional else RealNVP
        return 1 <= len(self.flows) <= 4 and all(type(f) is kind for f in self.flows) and len({f.dim for f in self.flows}) == 1

    def packed(self):
        return self._cache.get(list(self.flows))

    def run_stack(self, x, row_ctx=None, part_ctx=None, inverse=False, neg_logdet=False, out=None):
        """(B,N,D) entry point used by the filter: row-constant context stays (B,C) and is hoisted in-kernel."""
        return ops.coupling_stack(self.packed(), x, row_ctx, part_ctx, len(self.flows), inverse, neg_logdet, out)

    def _chain(self, x, obser, inverse):
        P, D = x.shape
        if self._fusable():
            ctx = obser.reshape(1, P, -1) if obser is not None else None
            y, ld = ops.coupling_stack(self.packed(), x.reshape(1, P, D), None, ctx, len(self.flows), inverse)
            return y.reshape(P, D), ld.reshape(P)
        log_det = torch.zeros(P, device=x.device)
        for flow in (self.flows[::-1] if inverse else self.flows):
            fn = flow.inverse if inverse else flow.forward
            x, ld = fn(x, obser) if self._conditional else fn(x)
            log_det = log_det + ld
        return x, log_det


class NormalizingFlowModel(_FlowChain):
    def forward(self, x):
        z, log_det = self._chain(x, None, False)
        return z, None, log_det  # the reference returns no prior term here (nf/models.py:19-20)

    def inverse(self, z):
        return self._chain(z, None, True)

    def sample(self, n_samples):
        z = self.prior.sample((n_samples,)).to(self.device)
        return self.inverse(z)[0]


class NormalizingFlowModel_cond(_FlowChain):
    _conditional = True

    def forward(self, x, obser):
        z, log_det = self._chain(x, obser, False)
        return z, self.prior.log_prob(z.float()), log_det

    def inverse(self, z, obser):
        return self._chain(z, obser, True)

    def sample(self, n_samples, obser):
        z = self.prior.sample((n_samples,)).to(self.device)
        return self.inverse(z, obser)[0]
