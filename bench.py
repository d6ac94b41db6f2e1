#!/usr/bin/env python
"""Benchmark of the NF-DPF per-timestep particle update (BASELINE.json metric: particle-steps/sec, filter fwd+bwd).

Workload (config.workload): CNF-DPF (--NF-dyn --NF-cond), Gaussian measurement, soft resampling forced every step,
N=1024 particles, B=1024 trajectories, T=50 steps, precomputed observation encodings (the CNN encoder is outside
the hot path).  One step = filtering_pos forward + supervised-loss backward over one synthetic batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--B .. --N .. --T ..]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--T", type=int, default=50)
    ap.add_argument("--measurement", default="gaussian")
    ap.add_argument("--resampler", default="soft")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the roofline / OT / cpu_baseline legs (profiling runs)")
    ap.add_argument("--inject-noise", action="store_true",
                    help="feed the motion noise / resampling offsets from the host batch (as the parity tests do) instead of drawing them on the device")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying one CUDA graph per step")
    return ap.parse_args()


FLAGS = lambda a: ["--NF-dyn", "--NF-cond", "--measurement", a.measurement, "--resampler_type", a.resampler, "--num-particles",
                   str(a.N), "--batchsize", str(a.B), "--sequence-length", str(a.T)]


def synth_batch(B, T, N, seed, pinned):
    """Synthetic disk-tracking-shaped inputs: encodings, states, start state, action velocities, and every random
    draw of the filter (initial cloud, motion noise, resampling offsets) so both arms consume identical numbers."""
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s: torch.empty(*s, pin_memory=pinned)
    d = dict(enc=mk(B, T, 32).normal_(generator=g) * 3.0,
             state=torch.cat([mk(B, T, 2).normal_(generator=g) * 20, mk(B, T, 2).normal_(generator=g) * 3], -1),
             start=torch.cat([mk(B, 2).normal_(generator=g) * 20, mk(B, 2).normal_(generator=g) * 3], -1),
             init_particles=mk(B, N, 2).uniform_(-64, 64, generator=g),
             noise=mk(B, T, N, 2).normal_(generator=g) * 20.0,
             offsets=mk(B, T).uniform_(0, 1.0 / N, generator=g))
    d["vel_in"] = d["state"][:, :, 2:] + torch.randn(B, T, 2, generator=g) * 4
    if pinned:
        d = {k: v.contiguous().pin_memory() for k, v in d.items()}
    return d


def workload_string(a):
    return ("CNF-DPF (--NF-dyn --NF-cond) %s measurement, %s resampling forced every step, N=%d, B=%d per GPU, T=%d, precomputed encodings "
            "(CNN encoder excluded), %s" % (a.measurement, a.resampler, a.N, a.B, a.T,
                                            "motion noise / offsets injected from the batch" if a.inject_noise else
                                            "motion noise / offsets drawn on the device"))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            pass

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v == "Active":
                    reasons.add(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


def build_b200(a, dev):
    from normalizing_flows_dpfs_b200.arguments import parse_args
    from normalizing_flows_dpfs_b200.DPFs import DPF
    torch.manual_seed(1234)
    dpf = DPF(parse_args(FLAGS(a)))
    dpf.encoder = torch.nn.Identity()
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():  # non-trivial weights (the reference's 0.01 init makes the flows almost the identity)
        for mod, ws, bs in ((dpf.nf_dyn, 0.1, 0.05), (dpf.cond_model, 0.05, 0.05), (dpf.particle_encoder, 0.4, 0.2)):
            for p in mod.parameters():
                p.copy_(torch.randn(p.shape, generator=gen) * (ws if p.dim() > 1 else bs))
    dpf = dpf.to(dev)
    dpf.force_resample = True
    return dpf


def step_b200(dpf, batch, dev, host_inputs, bucket=None):
    """One filter forward + backward.  host_inputs=True: inputs start in pinned host memory and the loss is read
    back (the e2e number); False: inputs are resident device tensors (the kernel-side `value`)."""
    from normalizing_flows_dpfs_b200.losses import supervised_loss
    d = {k: v.to(dev, non_blocking=True) for k, v in batch.items()} if host_inputs else batch
    dpf.injected = {k: d[k] for k in ("init_particles", "noise", "offsets") if k in d}
    dpf.zero_grad(set_to_none=True)
    out = dpf.filtering_pos(d["enc"], d["start"], d["vel_in"])
    loss, _ = supervised_loss(out[0], out[1], d["state"], 1.0, False)
    loss.backward()
    if bucket is not None:      # multi-GPU training step: the one collective (flat gradient all-reduce over NCCL)
        bucket.allreduce()
    return loss.item() if host_inputs else loss


def main():
    a = parse()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if a.impl == "reference":
        from bench_reference import run_reference
        return run_reference(a, rank, world)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from normalizing_flows_dpfs_b200 import _lib
    dpf = build_b200(a, dev)
    host = synth_batch(a.B, a.T, a.N, 100 + rank, pinned=True)     # weak scaling: every rank owns B trajectories
    if not a.inject_noise:       # production setting: random draws on the device (Philox), nothing but observations cross PCIe
        dpf.rng_device = "cuda"
        host.pop("noise"), host.pop("offsets")
    resident = {k: v.to(dev) for k, v in host.items()}
    bucket = None
    if world > 1:
        from normalizing_flows_dpfs_b200.distributed import GradBucket
        bucket = GradBucket(dpf)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms)

    for _ in range(a.warmup):
        step_b200(dpf, resident, dev, False, bucket)
    l0 = _lib.launch_count()
    step_b200(dpf, resident, dev, False, bucket)
    launches_per_step = _lib.launch_count() - l0          # kernels of ours in one step (a graph replay re-issues all of them)
    graphed = None
    if not a.no_graph:
        from normalizing_flows_dpfs_b200.graphs import GraphedFilterStep
        graphed = GraphedFilterStep(dpf, resident)

    def run_resident():
        if graphed is None:
            return step_b200(dpf, resident, dev, False, bucket)
        graphed.run()
        if bucket is not None:
            bucket.allreduce()

    def run_host():
        if graphed is None:
            return step_b200(dpf, host, dev, True, bucket)
        loss = graphed.run(host)                       # H2D of every input of the step from pinned memory, then replay
        if bucket is not None:
            bucket.allreduce()
        return loss.item()                             # D2H read of the step's result

    for _ in range(a.warmup):
        run_resident()
    clocks = ClockSampler(local)
    ms = timed(run_resident, a.steps)
    launches = launches_per_step * a.steps
    clk = clocks.stop()
    for _ in range(min(a.warmup, 2)):
        run_host()
    ms_e2e = timed(run_host, a.steps)
    # the product's fast path with the REFERENCE's gate: DPF(--fast) semantics (device-side ESS gate + in-kernel draws), one graph
    # replay per step, host inputs copied in and the loss read back -- what a user of DPF.filtering_pos gets without forcing anything
    default_path = None
    try:
        if not a.no_graph and not a.inject_noise:
            dpf.force_resample = None
            g2 = GraphedFilterStep(dpf, resident)

            def run_default():
                loss = g2.run(host)
                if bucket is not None:
                    bucket.allreduce()
                return loss.item()
            for _ in range(min(a.warmup, 2)):
                run_default()
            ms_def = timed(run_default, a.steps)
            fired = dpf.fired
            default_path = {"value": a.B * a.N * a.T * world * a.steps / (ms_def / 1e3), "unit": "particle-steps/s", "ms_per_step": ms_def / a.steps,
                            "gate": "reference ESS rule evaluated on the device (DPFs.py:163-165)", "resampled_steps": int(sum(fired)), "of_steps": len(fired)}
            dpf.force_resample = True
    except Exception as e:
        default_path = {"error": repr(e)}
        dpf.force_resample = True
    configs = None
    if not a.no_extras and not a.no_graph and (a.B, a.N, a.T, a.measurement, a.resampler) == (1024, 1024, 50, "gaussian", "soft"):
        try:     # the other BASELINE configurations (C3, C4, C5 shard), a few steps each -- every rank takes part
            graphed = g2 = None
            torch.cuda.empty_cache()
            from bench_configs import extra_configs
            configs = extra_configs(a, dev, world, timed)
        except Exception as e:
            configs = {"error": repr(e)}
    units = a.B * a.N * a.T * world
    if rank != 0:
        return
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    line = {
        "metric": "particle-steps/sec, NF-DPF filter fwd+bwd, N=%d" % a.N, "value": units * a.steps / (ms / 1e3), "unit": "particle-steps/s",
        "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(a),
                   "l2": "per-step working set (particles, noise, lists: >300 MB) exceeds the 126 MB L2", "parallelism": "batch-sharded x%d%s" % (world, ", NCCL flat-gradient all-reduce per step" if world > 1 else ""),
                   "execution": "eager launches" if a.no_graph else "one CUDA graph replay per step (forward over T + loss + backward)"},
        "e2e": {"value": units * a.steps / (ms_e2e / 1e3), "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "e2e_default_path": default_path, "configs": configs,
        "gpu_launches": int(launches), "clocks": clk,
    }
    try:
        if not a.no_extras:
            from bench_extras import roofline_and_cpu
            line.update(roofline_and_cpu(a, dpf, resident, dev, ms / a.steps))
    except Exception as e:  # never lose the headline line
        line["roofline_error"] = repr(e)
    print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
