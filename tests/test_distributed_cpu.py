"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: batch sharding and the flat gradient bucket."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from normalizing_flows_dpfs_b200.distributed import GradBucket, global_ess_mean, shard_bounds
from normalizing_flows_dpfs_b200.losses import supervised_loss


def test_shard_bounds_cover_batch():
    for total in (1, 7, 16, 16384, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    frozen = torch.nn.Linear(2, 2)          # a parameter that never gets a gradient
    model = torch.nn.ModuleList([net, frozen])
    x = torch.full((3, 4), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    bucket = GradBucket(model)
    bucket.allreduce()
    gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(world)]
    dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
    expect = sum(gathered) / world
    got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    ok = torch.allclose(got, expect, atol=1e-6) and all(p.grad is None for p in frozen.parameters())      # never in the bucket
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


# ---- world-2 vs world-1 parity of the batch-coupled quantities (ESS gate input, RMSE loss, parameter gradients) -----------------
def _toy_filter(net, x, w_logits):
    """stand-in for the filter outputs: particles (B,T,N,2) from a tiny net, weights (B,T,N) -- plain torch, CPU."""
    B, T, N, _ = x.shape
    particles = net(x.reshape(-1, 2)).reshape(B, T, N, 2) + x
    weights = torch.softmax(w_logits, -1)
    return particles, weights


def _full_batch_reference(seed, B, T, N):
    g = torch.Generator().manual_seed(seed)
    x, wl = torch.randn(B, T, N, 2, generator=g), torch.randn(B, T, N, generator=g)
    state = torch.randn(B, T, 4, generator=g)
    ess = torch.rand(B, generator=g) * N
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(2, 8), torch.nn.Tanh(), torch.nn.Linear(8, 2))
    return net, x, wl, state, ess


def _parity_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, T, N = 6, 3, 16
    net, x, wl, state, ess = _full_batch_reference(5, B, T, N)
    lo, hi = shard_bounds(B, rank, world)
    p, w = _toy_filter(net, x[lo:hi], wl[lo:hi])
    loss, _ = supervised_loss(p, w, state[lo:hi], 1.0, False, group=dist.group.WORLD)
    loss.backward()
    GradBucket(net).allreduce(average=False)
    out[rank] = (float(loss), torch.cat([q.grad.reshape(-1) for q in net.parameters()]).tolist(), float(global_ess_mean(ess[lo:hi])))
    dist.destroy_process_group()


def test_world2_matches_world1_on_the_union_batch():
    B, T, N = 6, 3, 16
    net, x, wl, state, ess = _full_batch_reference(5, B, T, N)
    p, w = _toy_filter(net, x, wl)
    loss, _ = supervised_loss(p, w, state, 1.0, False)
    loss.backward()
    ref_grad = torch.cat([q.grad.reshape(-1) for q in net.parameters()])
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_parity_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    for r in range(world):
        l, g, m = res[r]
        assert abs(l - float(loss)) < 1e-6, "global RMSE differs from the single-process run"
        assert torch.allclose(torch.tensor(g), ref_grad, rtol=1e-5, atol=1e-7), "summed shard gradients != gradient of the global loss"
        assert abs(m - float(ess.mean())) < 1e-4, "gate input is not the whole-batch ESS mean"
        assert (m < 0.5 * N) == (float(ess.mean()) < 0.5 * N)
