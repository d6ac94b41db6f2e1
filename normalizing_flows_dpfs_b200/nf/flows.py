"""RealNVP / conditional RealNVP with the reference's module layout (reference nf/flows.py:101-239):
submodules t1, s1, t2, s2, each an FCNN whose `.network` is Sequential(Linear, Tanh, Linear, Tanh, Linear), so
state_dict keys match `...t1.network.{0,2,4}.{weight,bias}`.  The parameters are ordinary nn.Parameters; the
forward / inverse are ONE fused sm_100a kernel each (nfdpf_coupling_fwd / _bwd) instead of ~45 ATen calls."""
import torch
from torch import nn

from .. import ops

HIDDEN = 8  # the only width the kernels are built for (and the only one the reference uses)


class FCNN(nn.Module):
    def __init__(self, in_dim, out_dim, hidden_dim):
        super().__init__()
        self.network = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.Tanh(), nn.Linear(hidden_dim, hidden_dim), nn.Tanh(),
                                     nn.Linear(hidden_dim, out_dim))

    def forward(self, x):  # plain PyTorch; the fused kernels read the parameters directly
        return self.network(x.float())


def pack_parameters(modules):
    """Flat fp32 vector of the modules' parameters in registration (= state_dict) order; differentiable."""
    return torch.cat([p.reshape(-1) for m in modules for p in m.parameters()])


class GradSlab:
    """Gradient rows of one packed parameter vector.  Every kernel call that takes the vector owns one row (`take` in its forward);
    its backward writes the row in place and hands autograd NOTHING for the vector; the pack's own backward sums the rows with one
    launch.  Without it autograd adds the per-call gradients pairwise: one tiny ATen launch per call (four per timestep, 1.2 % of
    the device time of a training step)."""
    CHUNK = 256
    BLOCK_CHUNK = 64     # calls per block tensor of the deferred D = 2 reduction (1.9 MB per call at 148 CTAs, C_row = 36)

    def __init__(self, numel, device):
        self.numel, self.device, self.count = numel, device, 0
        self.chunks, self.written = [], set()
        # deferred reduction (ops.CouplingStack, D = 2): per kernel shape key = (n_flows, D, C_row, B) the calls' partial-row blocks
        self.bcount, self.bchunks, self.bwritten = {}, {}, {}

    def take(self):
        self.count += 1
        return self.count - 1

    def row(self, i):
        """Row i (contiguous, uninitialised until the caller's kernel has written every entry); marks it as written."""
        c = i // self.CHUNK
        while len(self.chunks) <= c:
            self.chunks.append(torch.empty(self.CHUNK, self.numel, dtype=torch.float32, device=self.device))
        self.written.add(i)
        return self.chunks[c][i % self.CHUNK]

    def take_block(self, key):
        i = self.bcount.get(key, 0)
        self.bcount[key] = i + 1
        return i

    def block(self, key, i, block_floats):
        """Block i of shape `key` (uninitialised until the backward kernel has written it); marks it as written."""
        chunks = self.bchunks.setdefault(key, [])
        c = i // self.BLOCK_CHUNK
        while len(chunks) <= c:
            chunks.append(torch.empty(self.BLOCK_CHUNK, block_floats, dtype=torch.float32, device=self.device))
        self.bwritten.setdefault(key, set()).add(i)
        return chunks[c][i % self.BLOCK_CHUNK]

    def _blocks_total(self):
        """One reduce launch per run of consecutive written blocks (normally one per 64 calls)."""
        out = None
        for key, written in self.bwritten.items():
            for c, chunk in enumerate(self.bchunks[key]):
                rows = sorted(i - c * self.BLOCK_CHUNK for i in written if i // self.BLOCK_CHUNK == c)
                k = 0
                while k < len(rows):
                    j = k
                    while j + 1 < len(rows) and rows[j + 1] == rows[j] + 1:
                        j += 1
                    part = ops.coupling_bwd_reduce(key, chunk[rows[k]:rows[j] + 1], self.numel)
                    out = part if out is None else out + part
                    k = j + 1
        self.bwritten = {}
        return out

    def total(self):
        """Sum of the written rows and blocks (None if there are none).  They are consumed: a second backward writes them again."""
        blocks = self._blocks_total() if self.bwritten else None
        if not self.written:
            return blocks
        out = blocks
        for c, chunk in enumerate(self.chunks):
            rows = sorted(i - c * self.CHUNK for i in self.written if i // self.CHUNK == c)
            if not rows:
                continue
            if rows == list(range(rows[0], rows[-1] + 1)):
                part = chunk[rows[0]:rows[-1] + 1].sum(0)
            else:
                part = chunk.index_select(0, torch.tensor(rows, device=self.device)).sum(0)
            out = part if out is None else out + part
        self.written = set()
        return out


class _Pack(torch.autograd.Function):
    """torch.cat of the flattened parameters whose backward adds the rows of the vector's GradSlab to the incoming gradient."""

    @staticmethod
    def forward(ctx, slab, *params):
        ctx.slab, ctx.shapes = slab, [p.shape for p in params]
        ctx.set_materialize_grads(False)
        return torch.cat([p.reshape(-1) for p in params])

    @staticmethod
    def backward(ctx, g):
        total = ctx.slab.total()
        if g is not None:
            total = g if total is None else total + g
        if total is None:
            return (None,) * (1 + len(ctx.shapes))
        outs, o = [], 0
        for k, shp in enumerate(ctx.shapes):
            n = 1
            for d in shp:
                n *= d
            outs.append(total[o:o + n].view(shp) if ctx.needs_input_grad[1 + k] else None)
            o += n
        return (None, *outs)


class _PackCache:
    """Re-pack only when a parameter changed (optimizer steps bump tensor versions)."""

    def __init__(self):
        self.key, self.value = None, None

    def get(self, modules):
        params = [p for m in modules for p in m.parameters()]
        # the packed tensor carries an autograd node bound to the stream it was built on: never reuse it across
        # streams (CUDA-graph capture runs on a side stream)
        stream = torch.cuda.current_stream().cuda_stream if params and params[0].is_cuda else 0
        key = (torch.is_grad_enabled(), stream) + tuple((p.data_ptr(), p._version, p.requires_grad) for p in params)
        if key != self.key:
            # drop the old packed tensor FIRST: its cat node keeps the parameters' AccumulateGrad nodes alive, and those
            # are bound to the stream they were first used on (a legacy-stream accumulator breaks graph capture)
            self.key = self.value = None
            if params and params[0].is_cuda and torch.is_grad_enabled() and any(p.requires_grad for p in params):
                slab = GradSlab(sum(p.numel() for p in params), params[0].device)
                value = _Pack.apply(slab, *params)
                value._nfdpf_slab = slab
            else:
                value = torch.cat([p.reshape(-1) for p in params])
            self.key, self.value = key, value
        return self.value

    def clear(self):
        self.key = self.value = None


class _Coupling(nn.Module):
    def _init_nets(self, dim, hidden_dim, base_network, obser_dim):
        if hidden_dim != HIDDEN or base_network is not FCNN:
            raise ValueError("the fused coupling kernels implement FCNN nets of hidden width 8 (reference default)")
        self.dim, self.obser_dim = dim, obser_dim
        fin = dim // 2 + (obser_dim or 0)
        for name in ("t1", "s1", "t2", "s2"):
            setattr(self, name, base_network(fin, dim // 2, hidden_dim))
        self._cache = _PackCache()

    def zero_initialization(self, var=0.1):
        """N(0, var) weights and zero biases on every Linear (reference nf/flows.py:127-150, 191-211)."""
        for net in (self.t1, self.s1, self.t2, self.s2):
            for layer in net.network:
                if isinstance(layer, nn.Linear):
                    nn.init.normal_(layer.weight, std=var)
                    layer.bias.data.fill_(0)

    def packed(self):
        return self._cache.get([self])

    def _run(self, x, obser, inverse):
        P, D = x.shape
        ctx = obser.reshape(1, P, -1) if obser is not None else None
        y, ld = ops.coupling_stack(self.packed(), x.reshape(1, P, D), None, ctx, 1, inverse)
        return y.reshape(P, D), ld.reshape(P)


class RealNVP(_Coupling):
    """Non-volume-preserving coupling flow [Dinh et al. 2017] (reference nf/flows.py:117-179)."""

    def __init__(self, dim, hidden_dim=8, base_network=FCNN):
        super().__init__()
        self._init_nets(dim, hidden_dim, base_network, None)

    def forward(self, x):
        return self._run(x, None, False)

    def inverse(self, z):
        return self._run(z, None, True)


class RealNVP_cond(_Coupling):
    """Coupling flow whose s/t nets also see a per-sample context `obser` (reference nf/flows.py:181-239)."""

    def __init__(self, dim, hidden_dim=8, base_network=FCNN, obser_dim=None):
        super().__init__()
        self._init_nets(dim, hidden_dim, base_network, obser_dim)

    def forward(self, x, obser):
        return self._run(x, obser, False)

    def inverse(self, z, obser):
        return self._run(z, obser, True)
