"""Batch-sharded data parallelism for the filter: one process per GPU, independent trajectories are
split contiguously by batch, parameters are replicated, and the only LARGE collective is one all-reduce of a flat fp32
gradient bucket per optimiser step (the reference has no distributed code at all).

Three quantities of the reference couple the trajectories of a batch; what happens to each under sharding:
  * the ESS gate is a mean over the WHOLE batch (DPFs.py:163-165).  `DPF.dist_group = group` makes the gate use the global
    mean: one 8-byte all-reduce per timestep (`global_ess_mean`), so an N-GPU run takes the same resampling decisions as
    a 1-GPU run on the union batch.  Without it the gate is per shard.
  * the supervised loss is the root of a batch mean (losses.py:25): `supervised_loss(..., group=group)` all-reduces the
    squared-error SUM (differentiably), so every rank holds the global RMSE and its local gradients are the local part of the
    global gradient; the bucket then SUMS (`GradBucket.allreduce(average=False)`).  Averaging per-shard RMSE gradients instead
    is a different (per-shard) objective.
  * the Sinkhorn loop stops as soon as ANY trajectory of the batch has converged (resamplers.py:126-129).  That rule stays
    SHARD-LOCAL: it is evaluated on the device once per iteration, a global version would need ~100 more collectives per
    resample.  OT results of an N-GPU run therefore match the oracle run on each shard, not a 1-GPU run of the union
    (SURVEY 8e measured how much the iteration count matters: 10 % median / 38 % max on the outputs)."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous [start, stop) of `total` trajectories owned by `rank` (remainder spread over the first ranks)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class _AllReduceSum(torch.autograd.Function):
    """sum over ranks; every rank ends up with the same value, so the upstream gradient passes through unchanged."""

    @staticmethod
    def forward(ctx, x, group):
        out = x.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


def global_sum(x, group=None):
    """Differentiable sum of a (small) tensor over the ranks of `group`; identity without an initialised process group."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    return _AllReduceSum.apply(x, group)


def global_ess_mean(ess_inv, group=None):
    """Whole-batch mean of 1 / sum_n p^2 over every shard (DPFs.py:163-164) as a 1-element tensor: [sum, count] is all-reduced."""
    acc = torch.stack([ess_inv.detach().double().sum(), torch.tensor(float(ess_inv.numel()), dtype=torch.float64, device=ess_inv.device)])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return (acc[0] / acc[1]).to(torch.float32).reshape(1)


class GradBucket:
    """Flat gradient bucket: after backward(), `allreduce()` reduces every parameter's .grad over the process group with a
    single collective.  The bucket holds only the parameters that actually received a gradient at the first call (the filter's
    hot path trains ~10.6 k of the module's 1.68 M parameters when the image encoder is not in the graph): 42 KB instead of
    6.7 MB on the wire.  A parameter that starts receiving gradients later makes the bucket rebuild itself."""

    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        self.group = process_group
        self.active, self.flat, self.views = None, None, None

    def _build(self, active):
        self.active = active
        ps = [self.params[i] for i in active]
        self.flat = torch.zeros(sum(p.numel() for p in ps), dtype=torch.float32, device=ps[0].device)
        self.views, o = [], 0
        for p in ps:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def allreduce(self, average=True):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        active = [i for i, p in enumerate(self.params) if p.grad is not None]
        if world > 1:      # every rank must agree on the bucket layout: union of the ranks' active sets
            mask = torch.zeros(len(self.params), dtype=torch.int32, device=self.params[0].device)
            if self.active is None or active != self.active:
                mask[active] = 1
                dist.all_reduce(mask, op=dist.ReduceOp.MAX, group=self.group)
                active = torch.nonzero(mask).flatten().tolist()
        if not active:
            return None
        if self.active is None or not set(active) <= set(self.active):
            self._build(sorted(set(active) | set(self.active or [])))
        ps = [self.params[i] for i in self.active]
        have = [(v, p.grad) for v, p in zip(self.views, ps) if p.grad is not None]
        if len(have) < len(ps):
            self.flat.zero_()
        torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(world)
        for v, p in zip(self.views, ps):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat
