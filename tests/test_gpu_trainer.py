"""GPU: the out-of-hot-path glue still works end to end with the accelerated filter inside it -- the flow of the
reference's main.py (dataset -> DataLoader -> DPF(args).to(device) -> train_val -> testing) on a tiny synthetic
dataset, CUDA-graph execution of a whole training step, and the on-device RNG path."""
import os

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from normalizing_flows_dpfs_b200.arguments import parse_args
from normalizing_flows_dpfs_b200.dataset import ToyDiskDataset
from normalizing_flows_dpfs_b200.DPFs import DPF
from normalizing_flows_dpfs_b200.graphs import GraphedFilterStep
from normalizing_flows_dpfs_b200.losses import supervised_loss

pytestmark = pytest.mark.gpu


def _toy_npz(path, n_seq, T, rng):
    def split(n):
        return {"start_image": rng.random((n, 128, 128, 3), dtype=np.float32), "start_state": rng.normal(0, 20, (n, 4)),
                "image": rng.random((n, T, 128, 128, 3), dtype=np.float32), "state": rng.normal(0, 20, (n, T, 4)),
                "q": rng.normal(0, 1, (n, T, 4)), "visible": np.ones((n, T), np.int64)}
    for name in ("train", "val", "test"):
        np.savez(os.path.join(path, "toy_pn=2.0_d=25_const_%s.npz" % name), **{name + "_data": split(n_seq)})


@pytest.mark.parametrize("flags", [["--measurement", "gaussian", "--resampler_type", "soft"],
                                   ["--NF-dyn", "--NF-cond", "--measurement", "CRNVP", "--resampler_type", "ot", "--trainType", "SDPF"],
                                   []])   # [] = the CLI defaults: cos measurement, OT resampling
def test_main_flow_trains_and_tests(tmp_path, monkeypatch, flags):
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(0)
    os.makedirs("data")
    _toy_npz("data", 8, 4, rng)
    args = parse_args(["--batchsize", "4", "--num-particles", "32", "--sequence-length", "4", "--num-epochs", "2", "--block-length", "2"] + flags)
    torch.manual_seed(args.seed)
    train = DataLoader(ToyDiskDataset("data", "toy_pn=2.0_d=25_const", "train_data"), batch_size=4, shuffle=True, drop_last=True)
    val = DataLoader(ToyDiskDataset("data", "toy_pn=2.0_d=25_const", "val_data"), batch_size=4, shuffle=False, drop_last=True)
    dpf = DPF(args).to("cuda")
    run_id = "run"
    for d in ("logs", "logs/run", "logs/run/models", "logs/run/data"):
        os.makedirs(d, exist_ok=True)
    before = [p.detach().clone() for p in dpf.particle_encoder.parameters()]
    dpf.train_val(train, val, run_id)
    assert any(not torch.equal(a, b) for a, b in zip(before, dpf.particle_encoder.parameters())), "optimizer did not update the hot-path weights"
    assert os.path.exists("logs/run/models/e2e_model_bestval_e2e.pth")
    ckpt = torch.load("logs/run/models/e2e_model_bestval_e2e.pth", weights_only=False)
    fresh = DPF(args).to("cuda")
    fresh.load_state_dict(ckpt["model"])                     # checkpoint round trip through the reference's key names
    test = DataLoader(ToyDiskDataset("data", "toy_pn=2.0_d=25_const", "test_data"), batch_size=4, shuffle=False, drop_last=True)
    dpf.testing(test, run_id=run_id)
    assert os.path.exists("logs/run/data/test_result.npz")
    out = np.load("logs/run/data/test_result.npz")
    assert out["particle_list"].shape == (4, 4, 32, 2) and np.isfinite(out["particle_list"]).all()


def _filter_dpf(flags, B, N, T, seed=0):
    g = torch.Generator().manual_seed(seed)
    dpf = DPF(parse_args(["--num-particles", str(N), "--batchsize", str(B), "--sequence-length", str(T)] + flags))
    with torch.no_grad():
        for mod, ws in ((dpf.nf_dyn, 0.1), (dpf.cond_model, 0.05), (dpf.particle_encoder, 0.4)):
            for p in mod.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (ws if p.dim() > 1 else 0.05))
    dpf.encoder = torch.nn.Identity()
    batch = dict(enc=torch.randn(B, T, 32, generator=g), start=torch.randn(B, 4, generator=g) * 10, vel_in=torch.randn(B, T, 2, generator=g) * 3,
                 init_particles=torch.rand(B, N, 2, generator=g) * 128 - 64, noise=torch.randn(B, T, N, 2, generator=g) * 20,
                 offsets=torch.rand(B, T, generator=g) / N, state=torch.randn(B, T, 4, generator=g) * 20)
    return dpf.cuda(), {k: v.cuda() for k, v in batch.items()}


def test_cuda_graph_step_equals_eager():
    flags = ["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft"]
    dpf, batch = _filter_dpf(flags, 8, 256, 5)
    dpf.force_resample = True
    dpf.injected = dict(init_particles=batch["init_particles"], noise=batch["noise"], offsets=batch["offsets"])
    out = dpf.filtering_pos(batch["enc"], batch["start"], batch["vel_in"])
    loss, _ = supervised_loss(out[0], out[1], batch["state"], 1.0, False)
    dpf.zero_grad(set_to_none=True)
    loss.backward()
    eager = [p.grad.clone() for p in dpf.nf_dyn.parameters()]
    loss = loss.detach().clone()
    del out                                              # nothing of the eager autograd graph may outlive this point
    step = GraphedFilterStep(dpf, batch)
    for _ in range(2):                                   # replays are idempotent for fixed inputs
        g_loss = step.run(batch)
        torch.cuda.synchronize()
        assert torch.allclose(g_loss, loss, rtol=1e-6, atol=0)
        for a, p in zip(eager, dpf.nf_dyn.parameters()):
            assert torch.equal(a, p.grad), "graph replay must reproduce the eager gradients bit for bit (deterministic reductions)"


def test_device_rng_path_runs_and_is_seed_reproducible():
    flags = ["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft"]
    dpf, batch = _filter_dpf(flags, 4, 128, 4)
    dpf.rng_device = "cuda"
    dpf.injected = dict(init_particles=batch["init_particles"])
    runs = []
    for _ in range(2):
        torch.manual_seed(11)
        out = dpf.filtering_pos(batch["enc"], batch["start"], batch["vel_in"])
        runs.append(out[0].detach().clone())
    assert torch.equal(runs[0], runs[1]) and torch.isfinite(runs[0]).all()
    noise = out[2]
    assert abs(float(noise.std()) - 20.0) < 1.0            # N(0, pos_noise^2) drawn on the device


def test_encoder_hoist_is_exact_in_eval_mode_and_trains():
    """one encoder call on all B*T frames == T calls on B frames when BatchNorm uses its running statistics."""
    torch.manual_seed(0)
    B, N, T = 3, 64, 4
    dpf = DPF(parse_args(["--NF-dyn", "--NF-cond", "--measurement", "gaussian", "--resampler_type", "soft", "--num-particles", str(N),
                          "--batchsize", str(B), "--sequence-length", str(T)])).cuda()
    g = torch.Generator().manual_seed(1)
    obs = torch.rand(B, T, 3, 128, 128, generator=g).cuda()
    start, vel_in = (torch.randn(B, 4, generator=g) * 10).cuda(), (torch.randn(B, T, 2, generator=g) * 3).cuda()
    dpf.injected = dict(init_particles=(torch.rand(B, N, 2, generator=g) * 128 - 64).cuda(), noise=(torch.randn(B, T, N, 2, generator=g) * 20).cuda(),
                        offsets=(torch.rand(B, T, generator=g) / N).cuda())
    dpf.eval()
    with torch.no_grad():
        ref = dpf.filtering_pos(obs, start, vel_in)
        dpf.hoist_encoder = True
        out = dpf.filtering_pos(obs, start, vel_in)
    assert torch.allclose(out[0], ref[0], rtol=1e-4, atol=1e-3) and torch.allclose(out[1], ref[1], rtol=1e-3, atol=1e-7)
    dpf.train()
    res = dpf.filtering_pos(obs, start, vel_in)
    supervised_loss(res[0], res[1], torch.randn(B, T, 4, generator=g).cuda(), 1.0, False)[0].backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in dpf.encoder.parameters())
