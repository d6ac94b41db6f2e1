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


def crnvp_roofline(dev):
    """c3's own kernels (BASELINE configs[2]): the CRNVP measurement forward / backward timed alone at B = N = 1024 (CUDA events).
    Algorithmic FLOP per particle: encoder 3136 + stack 9216 forward, twice that backward.  EXECUTED tensor FLOP per
    particle (3xTF32: every product is issued three times): forward = encoder layers 2-3 + layer 1 of the eight nets on tcgen05
    (1536 + 3072 FMA); backward = those again (recompute) + W3^T d3, W2^T d2 + delta1 W1cat of the four stages on tcgen05
    (4608 + 1536 + 3072 FMA) + the weight-gradient contractions on mma.sync (8 x 608 + 1648 FMA).  Both kernels are bound by the
    latency of their issue -> commit -> wait rounds and of the legacy mma.sync path, not by either peak: the fractions say how far."""
    from bench_extras import _events, measured_peaks
    from normalizing_flows_dpfs_b200 import ops
    from normalizing_flows_dpfs_b200.nf.flows import pack_parameters
    B = N = 1024
    g = torch.Generator(device=dev).manual_seed(3)
    x = (torch.randn(B, N, 2, device=dev, generator=g) * 2).requires_grad_()
    gl = torch.randn(B, N, device=dev, generator=g)
    enc = torch.randn(B, 32, device=dev, generator=g)
    lw0 = torch.log_softmax(torch.randn(B, N, device=dev, generator=g), -1)
    dpf = _build(["--NF-dyn", "--NF-cond", "--measurement", "CRNVP", "--resampler_type", "soft"], 8, N, 2, dev, cnf=True)
    with torch.no_grad():
        pe = pack_parameters([dpf.particle_encoder]).detach().clone().requires_grad_()
        cnf = pack_parameters([dpf.cnf_measurement]).detach().clone().requires_grad_()
    out = ops.measure_update(pe, cnf, enc, x, lw0, gl, gl, "CRNVP", p0=0.0, p1=2.5)
    tf = _events(lambda: ops.measure_update(pe, cnf, enc, x, lw0, gl, gl, "CRNVP", p0=0.0, p1=2.5), n=5)
    tb = _events(lambda: torch.autograd.backward([out[0], out[2]], [gl, gl], retain_graph=True), n=5)
    peaks = measured_peaks(dev)
    P = B * N
    alg_f, alg_b = 3136 + 9216, 2 * (3136 + 9216)
    ten_f = 2 * 3 * (1536 + 3072)
    ten_b = 2 * 3 * (4608 + 1536 + 3072 + 8 * 608 + 1648)
    rows = []
    for name, sec, alg, ten in (("measure_update_CRNVP_fwd", tf, alg_f, ten_f), ("measure_update_CRNVP_bwd", tb, alg_b, ten_b)):
        rows.append({"kernel": name, "sec": round(sec, 7), "bound": "tensor", "achieved": round(ten * P / sec / 1e12, 3),
                     "peak": peaks["tf32_tensor_tflops"], "unit": "TFLOP/s (executed 3xTF32 tensor FLOP; peak = measured bf16 / 2)",
                     "frac": round(ten * P / sec / 1e12 / peaks["tf32_tensor_tflops"], 4),
                     "frac_fp32_algorithmic": round(alg * P / sec / 1e12 / peaks["fp32_tflops"], 4), "launches_per_step": 50})
    return rows


def extra_configs(a, dev, world, timed, peaks=None):
    """Runs on EVERY rank (collectives inside); the returned dict is printed by rank 0."""
    from normalizing_flows_dpfs_b200.distributed import GradBucket
    cfgs = {}
    nf = ["--NF-dyn", "--NF-cond"]
    try:
        cfgs["c3"] = _time_config("DPF-CM: --NF-dyn --NF-cond --measurement CRNVP, soft resampling", nf + ["--measurement", "CRNVP", "--resampler_type", "soft"],
                                  1024, 1024, 50, dev, world, timed, 2, GradBucket, cnf=True, force=True)
        if (torch.distributed.get_rank() if world > 1 else 0) == 0:
            cfgs["c3"]["roofline_kernels"] = crnvp_roofline(dev)
    except Exception as e:   # never lose the headline line
        cfgs["c3"] = {"error": repr(e)}
    try:
        cfgs["c4"] = _time_config("NF-DPF with OT (Sinkhorn) resampling", nf + ["--measurement", "gaussian", "--resampler_type", "ot"],
                                  256, 4096, 5, dev, world, timed, 2, GradBucket, force=True)
        cfgs["c4"]["roofline"] = "ot_resample[1] of this line (B = 256, N = 4096): SFU-bound, frac_sfu_executed"
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
