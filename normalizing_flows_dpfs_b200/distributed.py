"""Batch-sharded data parallelism for the filter: one process per GPU, independent trajectories are
split contiguously by batch, parameters are replicated, and the ONLY collective is one all-reduce of a flat fp32
gradient bucket per optimiser step (the reference has no distributed code at all).  Nothing in the forward /
backward of the filter communicates; the ESS gate and the Sinkhorn stop rule are evaluated per shard."""
import torch
import torch.distributed as dist


def shard_bounds(total, rank, world):
    """Contiguous [start, stop) of `total` trajectories owned by `rank` (remainder spread over the first ranks)."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GradBucket:
    """Flat gradient bucket: after backward(), `allreduce()` averages every parameter's .grad over the process
    group with a single collective (1.68 M parameters = 6.7 MB: latency bound on NVLink, so one bucket)."""

    def __init__(self, module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        self.group = process_group
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def allreduce(self, average=True):
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        have = [(v, p.grad) for v, p in zip(self.views, self.params)]
        self.flat.zero_()
        torch._foreach_copy_([v for v, g in have if g is not None], [g for v, g in have if g is not None])
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
            if average:
                self.flat.div_(world)
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat
