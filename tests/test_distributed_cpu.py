"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: batch sharding and the flat gradient bucket."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from normalizing_flows_dpfs_b200.distributed import GradBucket, shard_bounds


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
    ok = torch.allclose(got, expect, atol=1e-6) and all(float(p.grad.abs().max()) == 0 for p in frozen.parameters())
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo_world2():
    world, port = 2, _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
