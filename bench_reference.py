"""`bench.py --impl reference`: the reference's own CPU algorithm for the benchmarked path, timed on the box's
host cores.  The reference is pure Python/PyTorch and is not present on the GPU box (/root/reference cannot travel),
so this arm runs the CPU oracle port (oracle/nfdpf_oracle.py: a restatement of the reference's filter step, pinned to
reference-generated goldens) with all host threads.  Each step is a BOUNDED SAMPLE of the workload: the same N, the
same per-step algorithm, fewer trajectories and timesteps (stated in cpu_baseline.sample)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def oracle_weights(gen, measurement):
    import nfdpf_oracle as O
    W = {"dyn": O.init_stack(gen, 2, 4, std=0.1, bias_std=0.05), "cond": O.init_stack(gen, 2, 36, std=0.05, bias_std=0.05)}
    pe = [torch.randn(s, generator=gen) * (0.4 if len(s) > 1 else 0.2) for s in ((16, 2), (16,), (32, 16), (32,), (32, 32), (32,))]
    W["pe"] = pe
    if measurement == "CRNVP":
        W["cnf"] = O.init_stack(gen, 32, 32, std=0.1, bias_std=0.05)
    return W


def cpu_filter_step(a, B, T, seed=0):
    """One oracle forward + backward at (B, N=a.N, T); returns seconds."""
    import nfdpf_oracle as O
    from bench import synth_batch
    gen = torch.Generator().manual_seed(7)
    flat = oracle_weights(gen, a.measurement)
    leaves = {k: (v.clone().requires_grad_() if torch.is_tensor(v) else [t.clone().requires_grad_() for t in v]) for k, v in flat.items()}
    W = {"dyn": O.unpack_stack(leaves["dyn"], 2, 4), "cond": O.unpack_stack(leaves["cond"], 2, 36), "pe": tuple(leaves["pe"])}
    if "cnf" in leaves:
        W["cnf"] = O.unpack_stack(leaves["cnf"], 32, 32)
    d = synth_batch(B, T, a.N, 100 + seed, pinned=False)
    cfg = dict(NF=True, NF_cond=True, measurement=a.measurement, resampler=a.resampler, alpha=0.5, pos_noise=20.0, eps=0.1, scaling=0.75,
               threshold=1e-3, max_iter=100)
    t0 = time.perf_counter()
    res = O.filtering(cfg, W, d["init_particles"], d["start"][:, 2:], d["vel_in"], d["enc"], d["noise"], d["offsets"], force_resample=True)
    loss, _ = O.supervised_rmse(res["particles"], res["probs"], d["state"][:, :, :2])
    loss.backward()
    return time.perf_counter() - t0


def sample_shape(a):
    # ~4.5e5 particle-steps/s on 16 host threads (soft resampling): 256 x 1024 x 10 = 2.6 M particle-steps ~ 6 s per timed step;
    # the fp64 (B,N,N) Sinkhorn of the reference is ~1 s per trajectory at N = 1024, so OT samples are much smaller
    return (min(a.B, 256), min(a.T, 10)) if a.resampler == "soft" else (min(a.B, 2), min(a.T, 2))


def run_reference(a, rank, world):
    if rank != 0:
        return
    from bench import workload_string
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, T = sample_shape(a)
    for _ in range(min(a.warmup, 1)):
        cpu_filter_step(a, B, T)
    t = [cpu_filter_step(a, B, T, s) for s in range(a.steps)]
    sec = sum(t)
    value = B * a.N * T * a.steps / sec
    sample = "oracle port of the reference filter step, B=%d of %d trajectories x T=%d of %d steps at N=%d per timed step" % (B, a.B, T, a.T, a.N)
    line = {"impl": "reference", "metric": "particle-steps/sec, NF-DPF filter fwd+bwd, N=%d" % a.N, "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sec / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(a)},
            "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
