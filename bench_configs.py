"""The other BASELINE.json configurations, measured inside the default `python bench.py` run (bounded: a few steps each) and
reported under `configs` in the same JSON line:

  c3  configs[2]: DPF-CM (--measurement CRNVP), soft resampling, B = N = 1024, T = 50 -- training step fwd + bwd
  c4  configs[3]: OT (Sinkhorn) resampling at N = 4096, B = 256 (T = 5 of the 50 steps: every step is the same work)
  c5  configs[4]: NF-dyn + NF-cond + CRNVP + OT, batch 16384 sharded over the ranks (B = 16384 / world per GPU), N = 1024, T = 5,
      with the gradient all-reduce; the Sinkhorn iteration counts per rank are the interesting scaling limiter (stop-rule skew)

All use the product's host-free path (in-kernel draws, device-side gate) replayed as one CUDA graph per step, timed with CUDA
events, max over ranks.  Resampling is FORCED every step, like the headline: with freshly initialised CRNVP / cos likelihoods the
reference's ESS rule fires 0-2 times in 50 steps and the step would skip the resampler these configurations are about."""
import torch


def _build(flags, B, N, T, dev, cnf=False, nn_head=False):
    from normalizing_flows_dpfs_b200.arguments import parse_args
    from normalizing_flows_dpfs_b200.DPFs import DPF
    torch.manual_seed(1234)
    dpf = DPF(parse_args(flags + ["--num-particles", str(N), "--batchsize", str(B), "--sequence-length", str(T)]))
    dpf.encoder = torch.nn.Identity()
    gen = torch.Generator().manual_seed(7)
    mods = [(dpf.nf_dyn, 0.1, 0.05), (dpf.cond_model, 0.05, 0.05), (dpf.particle_encoder, 0.4, 0.2)]
    if cnf:
        mods.append((dpf.cnf_measurement, 0.1, 0.05))
    if nn_head:
        mods.append((dpf.likelihood_est, 0.15, 0.1))
    with torch.no_grad():
        for mod, ws, bs in mods:
            for p in mod.parameters():
                p.copy_(torch.randn(p.shape, generator=gen) * (ws if p.dim() > 1 else bs))
    dpf = dpf.to(dev)
    dpf.rng_device = "cuda"
    return dpf


def _time_config(name, flags, B, N, T, dev, world, timed, steps, bucket_cls=None, cnf=False, force=None, nn_head=False):
    from bench import synth_batch
    from normalizing_flows_dpfs_b200 import ops
    from normalizing_flows_dpfs_b200.graphs import GraphedFilterStep
    rank = torch.distributed.get_rank() if world > 1 else 0
    dpf = _build(flags, B, N, T, dev, cnf, nn_head)
    dpf.force_resample = force
    host = synth_batch(B, T, N, 300 + rank, pinned=False)
    host.pop("noise"), host.pop("offsets")
    resident = {k: v.to(dev) for k, v in host.items()}
    bucket = bucket_cls(dpf) if (bucket_cls is not None and world > 1) else None
    g = GraphedFilterStep(dpf, resident, warmup=1)

    def run():
        g.run()
        if bucket is not None:
            bucket.allreduce()
    run()
    ms = timed(run, steps)
    fired = dpf.fired
    out = {"workload": name, "B_per_gpu": B, "N": N, "T": T, "steps": steps, "ms_per_step": ms / steps,
           "value": B * N * T * world * steps / (ms / 1e3), "unit": "particle-steps/s", "resampled_steps": int(sum(fired)), "of_steps": len(fired),
           "gate": "forced" if force else "reference ESS rule on the device"}
    if "ot" in flags:
        it = ops.OtResample.last_iters
        iters = torch.tensor([int(it.item()) if it is not None else 0], device=dev)
        if world > 1:
            allit = [torch.zeros_like(iters) for _ in range(world)]
            torch.distributed.all_gather(allit, iters)
            out["sinkhorn_iterations_last_resample_per_rank"] = [int(t.item()) for t in allit]
        else:
            out["sinkhorn_iterations_last_resample"] = int(iters.item())
    del g, dpf, resident
    torch.cuda.empty_cache()
    return out


def extra_configs(a, dev, world, timed, peaks=None):
    """Runs on EVERY rank (collectives inside); the returned dict is printed by rank 0."""
    from normalizing_flows_dpfs_b200.distributed import GradBucket
    cfgs = {}
    nf = ["--NF-dyn", "--NF-cond"]
    try:
        cfgs["c3"] = _time_config("DPF-CM: --NF-dyn --NF-cond --measurement CRNVP, soft resampling", nf + ["--measurement", "CRNVP", "--resampler_type", "soft"],
                                  1024, 1024, 50, dev, world, timed, 2, GradBucket, cnf=True, force=True)
    except Exception as e:   # never lose the headline line
        cfgs["c3"] = {"error": repr(e)}
    try:
        cfgs["c4"] = _time_config("NF-DPF with OT (Sinkhorn) resampling", nf + ["--measurement", "gaussian", "--resampler_type", "ot"],
                                  256, 4096, 5, dev, world, timed, 2, GradBucket, force=True)
    except Exception as e:
        cfgs["c4"] = {"error": repr(e)}
    try:
        cfgs["c5"] = _time_config("batch-sharded NF-DPF: --NF-dyn --NF-cond --measurement CRNVP, OT resampling, global batch 16384",
                                  nf + ["--measurement", "CRNVP", "--resampler_type", "ot"], 16384 // world, 1024, 5, dev, world, timed, 2,
                                  GradBucket, cnf=True, force=True)
        cfgs["c5"]["global_batch"] = 16384
    except Exception as e:
        cfgs["c5"] = {"error": repr(e)}
    try:   # not a BASELINE configuration: the NN likelihood, same shape as the headline, T = 10 of the 50 steps
        cfgs["nn"] = _time_config("--NF-dyn --NF-cond --measurement NN (fused mode 3), soft resampling", nf + ["--measurement", "NN", "--resampler_type", "soft"],
                                  1024, 1024, 10, dev, world, timed, 2, GradBucket, force=True, nn_head=True)
    except Exception as e:
        cfgs["nn"] = {"error": repr(e)}
    return cfgs
