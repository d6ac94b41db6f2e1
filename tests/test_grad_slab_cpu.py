"""Host logic of the parameter-gradient slab (nf/flows.py: GradSlab, _Pack) on CPU tensors: rows are handed out in forward order,
only written rows are summed, a second `total()` starts from scratch, and autograd runs the pack's backward after every consumer
even though the consumers return no gradient for the packed vector."""
import torch

from normalizing_flows_dpfs_b200.nf.flows import GradSlab, _Pack


def test_rows_written_or_not_and_consumed():
    slab = GradSlab(5, torch.device("cpu"))
    ids = [slab.take() for _ in range(300)]                 # more than one chunk
    assert ids == list(range(300))
    want = torch.zeros(5)
    for i in (0, 1, 2, 7, 256, 299):                        # a contiguous run, a hole, the second chunk
        r = slab.row(i)
        r.copy_(torch.full((5,), float(i + 1)))
        want += i + 1
    assert torch.equal(slab.total(), want)
    assert slab.total() is None                             # consumed
    slab.row(3).copy_(torch.ones(5))
    assert torch.equal(slab.total(), torch.ones(5))         # a second backward rewrites what it needs


class _Consumer(torch.autograd.Function):
    """Stands in for a kernel call: writes its parameter gradient into its slab row, returns None for the packed vector."""

    @staticmethod
    def forward(ctx, packed, x, slab):
        ctx.slab, ctx.row = slab, slab.take()
        ctx.save_for_backward(packed, x)
        return x * packed.sum()

    @staticmethod
    def backward(ctx, g):
        packed, x = ctx.saved_tensors
        ctx.slab.row(ctx.row).copy_(torch.full_like(packed, float((g * x).sum())))
        return None, g * packed.sum(), None


def test_pack_backward_runs_after_all_consumers_and_splits_the_sum():
    a = torch.tensor([1.0, 2.0], requires_grad=True)
    b = torch.tensor([[3.0], [4.0], [5.0]], requires_grad=True)
    slab = GradSlab(5, torch.device("cpu"))
    packed = _Pack.apply(slab, a, b)
    x = torch.tensor([0.5, -1.0, 2.0], requires_grad=True)
    y = _Consumer.apply(packed, x, slab)
    z = _Consumer.apply(packed, y, slab)
    unused = _Consumer.apply(packed, x, slab)               # never reaches the loss: its row must not be summed
    z.sum().backward()
    ar, br = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    pr = torch.cat([ar.reshape(-1), br.reshape(-1)])
    xr = x.detach().clone().requires_grad_()
    ((xr * pr.sum()) * pr.sum()).sum().backward()
    assert torch.allclose(a.grad, ar.grad) and torch.allclose(b.grad, br.grad) and torch.allclose(x.grad, xr.grad)
    assert b.grad.shape == b.shape and unused is not None
