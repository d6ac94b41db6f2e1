"""Roofline and CPU-baseline legs of bench.py (kept apart so a failure here can never lose the headline line).

Algorithmic work per particle per launch (restated in DESIGN.md):
  coupling stack pass, D=2, row context hoisted: 1280 FLOP forward, 2x that backward; 20 B of HBM traffic
  particle encoder + Gaussian likelihood + weight update: 3136 + 96 FLOP forward; 2x backward
  soft resampling: 32 B forward (4 w + 8 gather in, 8 + 4 + 8 out), 36 B backward; weight update: 20 B
The kernels of the headline workload are FP32-pipe / SFU bound, not HBM or tensor bound (hidden width 8), so the
denominator for them is the FP32 FFMA peak MEASURED on this GPU by libnfdpf's probe kernel; the HBM-bound kernels use
MEASURED_PEAKS.json's copy bandwidth (fallback 6650 GB/s)."""
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def _events(fn, n=10, warm=3):
    """Mean device time of fn() in seconds.  The launches are captured into a CUDA graph and replayed so that the
    ~30 us of Python/ctypes launch overhead does not pollute kernels that run for 10-100 us; falls back to plain
    back-to-back launches if capture is not possible."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    graph = None
    for attempt in range(2):      # the first capture of a process can be invalidated by one-time lazy initialisation: retry once
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fn()
            graph.replay()
            torch.cuda.synchronize()
            break
        except Exception:
            graph = None
            torch.cuda.synchronize()
            for _ in range(2):    # a failed capture leaves the next launches slow: re-warm before anything is timed
                fn()
            torch.cuda.synchronize()
    best = None
    for run in ([graph.replay] if graph is not None else []) + [fn]:   # min of graph replay and plain launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / n * 1e-3
        best = t if best is None else min(best, t)
    return best


def _traffic(kernel):
    """dram__bytes_read + dram__bytes_write per launch from the committed ncu capture (profiles/r2_traffic.json), or None."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel)
        except OSError:
            continue
    return None


def measured_peaks(dev):
    from normalizing_flows_dpfs_b200 import _lib as L
    out = torch.zeros(4, device=dev)
    peaks = {}
    for kind, name, mult in ((0, "fp32_tflops", 2.0), (1, "sfu_tops", 1.0)):
        ops = [0]

        def run():
            ops[0] = L.load().nfdpf_peak_probe(kind, 4096, L.ptr(out), L.stream())
        sec = _events(run, n=5, warm=2)
        peaks[name] = ops[0] * mult / sec / 1e12
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        mp = json.load(open(path))
        peaks["hbm_gbs"], peaks["hbm_source"] = mp["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
        bf16 = mp.get("bf16_tflops", 1590.0)
    else:
        peaks["hbm_gbs"], peaks["hbm_source"], bf16 = 6650.0, "fallback (B200_PROFILING.md)", 1590.0
    # dense TF32 runs at half the bf16 rate on the 5th-generation tensor cores (1.1 vs 2.25 PFLOP/s nominal): the measured
    # cuBLAS bf16 figure / 2 is the denominator for the 3xTF32 tcgen05 / mma.sync products of the measurement kernels
    peaks["tf32_tensor_tflops"] = bf16 / 2.0
    return peaks


def kernel_table(a, dpf, dev):
    """Time each libnfdpf kernel of the workload alone at the workload's shapes (CUDA events on the current stream)."""
    from normalizing_flows_dpfs_b200 import ops
    from normalizing_flows_dpfs_b200.nf.flows import pack_parameters
    B, N = a.B, a.N
    P = B * N
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(B, N, 2, device=dev, generator=g) * 20
    w = torch.softmax(torch.randn(B, N, device=dev, generator=g) * 3, -1)
    ctx36, enc = torch.randn(B, 36, device=dev, generator=g), torch.randn(B, 32, device=dev, generator=g)
    off = torch.rand(B, device=dev, generator=g) / N
    mk = torch.linspace(0.0, (N - 1.0) / N, N).to(dev)
    gy, gl = torch.randn(B, N, 2, device=dev, generator=g), torch.randn(B, N, device=dev, generator=g)
    with torch.no_grad():
        pk_c, pk_d = pack_parameters([dpf.cond_model]), pack_parameters([dpf.nf_dyn])
        pe = pack_parameters([dpf.particle_encoder])
    rows = []

    def add(name, fwd, flops_f, bytes_f, flops_b, bytes_b, grads, per_step_f, per_step_b, sfu=0, tensor_f=0, tensor_b=0):
        out = fwd()
        outs = [o for o in (out if isinstance(out, tuple) else (out,)) if o is not None and o.requires_grad]
        tf = _events(lambda: fwd(), n=5)
        rows.append(dict(kernel=name + "_fwd", sec=tf, flops=flops_f * P, bytes=bytes_f * P, sfu_ops=sfu * P, tensor_flops=tensor_f * P,
                         launches_per_step=per_step_f))
        if outs:
            gouts = [grads[tuple(o.shape)] for o in outs]

            def bwd():
                torch.autograd.backward(outs, gouts, retain_graph=True)
            tb = _events(bwd, n=5)
            rows.append(dict(kernel=name + "_bwd", sec=tb, flops=flops_b * P, bytes=bytes_b * P, sfu_ops=sfu * P, tensor_flops=tensor_b * P,
                             launches_per_step=per_step_b))

    grads = {(B, N, 2): gy, (B, N): gl}
    pkc, pkd = pk_c.clone().requires_grad_(), pk_d.clone().requires_grad_()
    xr = x.clone().requires_grad_()
    # MUFU ops actually needed per stack pass: 128 tanh at 1.5 each (two ex2 + one shared rcp per pair of activations) + 4 exp
    # (ex2); the backward re-evaluates every FCNN once and needs e^s and e^-s per stage
    add("coupling_D2_C36", lambda: ops.coupling_stack(pkc, xr, ctx36, None, 2, True), 1280, 20, 2560, 40, grads, 1, 1, sfu=196)
    add("coupling_D2_C4", lambda: ops.coupling_stack(pkd, xr, ctx36[:, :4].contiguous(), None, 2, True), 1280, 20, 2560, 40, grads, 2, 2, sfu=196)
    per = pe.clone().requires_grad_()
    lw0 = w.log()
    mode = a.measurement
    p0, p1 = {"gaussian": (1.0, 10.0), "cos": (0.0, 1.0), "CRNVP": (0.0, 2.5)}[mode]
    cnf = pack_parameters([dpf.cnf_measurement]).detach().clone().requires_grad_() if mode == "CRNVP" else None
    mflop = 3136 + 96 + (9216 if mode == "CRNVP" else 0)

    def meas():
        o = ops.measure_update(per, cnf, enc, xr, lw0, gl, gl, mode, p0=p0, p1=p1)
        return o[0], o[2]
    # EXECUTED tensor-core FLOP per particle (3xTF32: every product is issued three times): forward = encoder layers 2-3
    # (1536 FMA) [+ CRNVP layer 1 of 8 nets, 16 x 48 per stage pair: 4 x 768 FMA]; backward = the two recomputed forward rounds + W3^T d3 +
    # W2^T d2 (tcgen05) + the dW2 / dW3 contractions (mma.sync), 4608 FMA.  They bound these kernels by tensor-ROUND latency, not by
    # tensor throughput: the fraction below is what ncu's sm__pipe_tensor_cycles_active shows (19 % / 36 %, profiles/).
    tf_f = 2 * 3 * (1536 + (3072 if mode == "CRNVP" else 0))
    tf_b = 2 * 3 * 4608
    add("measure_update_" + mode, meas, mflop, 28, 2 * mflop, 24, grads, 1, 1, tensor_f=tf_f, tensor_b=tf_b)
    wr = w.clone().requires_grad_()
    if a.resampler == "soft":
        add("soft_resample", lambda: ops.soft_resample(xr, wr, off, mk, 0.5)[:2], 0, 32, 0, 36, grads, 1, 1)
    else:
        add("ot_resample", lambda: ops.ot_resample(xr, lw0), 0, 0, 0, 0, grads, 1, 1)
    return rows


def ot_metric(dev, peaks):
    """Second half of the BASELINE metric: OT-resample time per call (forward and backward) at the BASELINE shapes,
    with the SFU roofline: useful pair evaluations (2 per loop iteration + 2 init + 2 final + 1 transport,
    one ex2 each) / duration against the ex2 rate measured by the probe kernel."""
    from normalizing_flows_dpfs_b200 import ops
    out = []
    for B, N in ((1024, 1024), (256, 4096)):
        g = torch.Generator(device=dev).manual_seed(5)
        w = torch.softmax(torch.randn(B, N, device=dev, generator=g) * 2, -1)
        x = (torch.randn(B, N, 2, device=dev, generator=g) * 20).requires_grad_()
        gy = torch.randn(B, N, 2, device=dev, generator=g)
        lw = w.log()
        for _ in range(2):      # two warm-up rounds: the first launch of each kernel pays one-time driver costs
            p = ops.ot_resample(x, lw)
            p.backward(gy)
            x.grad = None
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        p = ops.ot_resample(x, lw)
        ev[1].record()
        p.backward(gy)
        ev[2].record()
        torch.cuda.synchronize()
        iters = int(ops.OtResample.last_iters.item())
        pairs = B * N * N * (2 * (iters - 2) + 2 + 2 + 1)
        fwd_s = ev[0].elapsed_time(ev[1]) * 1e-3
        # round 2: one exponential serves BOTH chains of a pass, so the kernels execute B N^2 (iters + 2) MUFU ops, about half of
        # the algorithmic pair-evaluation count of SURVEY 8d (one exp per chain and pair): the algorithmic fraction can exceed 1
        executed = B * N * N * ((iters - 2) + 1 + 1 + 1 + 1)
        out.append({"B": B, "N": N, "sinkhorn_iterations": iters, "fwd_us": fwd_s * 1e6, "bwd_us": ev[1].elapsed_time(ev[2]) * 1e3,
                    "pair_evals": pairs, "achieved_tops": pairs / fwd_s / 1e12, "sfu_peak_tops": peaks["sfu_tops"],
                    "frac_of_sfu_peak": pairs / fwd_s / 1e12 / peaks["sfu_tops"], "mufu_ops_executed": executed,
                    "frac_sfu_executed": executed / fwd_s / 1e12 / peaks["sfu_tops"]})
    return out


def block_density_metric(dev):
    """the block pseudo-likelihood of the semi-supervised objective (losses.py:37-70) on filter-shaped lists --
    (T,B,N) buffers read through transposed views, sorted ancestor rows -- forward + backward: the kernel pair of csrc/losses.cu
    next to the reference's gather chain (same tensors, same GPU, plain PyTorch)."""
    from normalizing_flows_dpfs_b200 import losses, ops
    B, N, T, bl = 1024, 1024, 50, 10
    g = torch.Generator(device=dev).manual_seed(9)
    w = torch.softmax(torch.randn(T, B, N, device=dev, generator=g) * 2, -1)
    lik, prior = torch.randn(T, B, N, device=dev, generator=g), torch.randn(T, B, N, device=dev, generator=g) * 3 - 4
    idx = torch.sort(torch.randint(0, N, (T, B, N), device=dev, generator=g), dim=-1).values + N * torch.arange(B, device=dev)[None, :, None]
    lists = [x.transpose(0, 1).requires_grad_() for x in (w, lik, prior)]
    iv = idx.transpose(0, 1)

    def kern():
        losses.pseudolikelihood_loss_nf(lists[0], None, lists[1], iv, None, lists[2], bl).backward()

    def chain():
        Q = losses._trace_blocks(lists[0], T, bl, lambda j, anc: ((lists[2][:, j] + lists[1][:, j]) if anc is None else
                                                                  (lists[2][:, j].reshape(-1)[anc] + lists[1][:, j].reshape(-1)[anc]),
                                                                  iv[:, j]))
        (-Q.mean()).backward()
    out = {"B": B, "N": N, "T": T, "block_len": bl}
    for name, fn in (("kernel_us", kern), ("torch_gather_chain_us", chain)):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 3 * 1e3
    out["bytes_per_particle_step"] = 16 + 12
    out["kernel_gbs"] = B * N * T * 28 / (out["kernel_us"] * 1e-6) / 1e9
    return out


def roofline_and_cpu(a, dpf, resident, dev, ms_per_step):
    peaks = measured_peaks(dev)
    rows = kernel_table(a, dpf, dev)
    step_sec = ms_per_step * 1e-3 / a.T          # one filter timestep, forward + backward
    for r in rows:
        r["share_of_step"] = r["sec"] * r["launches_per_step"] / step_sec
        if r.get("tensor_flops", 0) > 0:     # tcgen05 / mma.sync kernels: executed 3xTF32 FLOP against the TF32 tensor peak
            r["frac_fp32_algorithmic"] = r["flops"] / r["sec"] / 1e12 / peaks["fp32_tflops"]   # (round 1 reported this one: most of these
            r.update(bound="tensor", achieved=r["tensor_flops"] / r["sec"] / 1e12, peak=peaks["tf32_tensor_tflops"],   # FMAs do not run on the FP32 pipe)
                     unit="TFLOP/s (executed 3xTF32 tensor FLOP; peak = measured bf16 / 2)")
        elif r["flops"] > 0 and r["flops"] / max(r["bytes"], 1) > 10:   # arithmetic intensity >> machine balance: compute bound
            f_fp32 = r["flops"] / r["sec"] / 1e12 / peaks["fp32_tflops"]
            f_sfu = r.get("sfu_ops", 0) / r["sec"] / 1e12 / peaks["sfu_tops"]
            r["frac_fp32"], r["frac_sfu"] = f_fp32, f_sfu
            if f_sfu > f_fp32:   # the MUFU pipe (16 lanes/clk/SM) binds before the FP32 pipe (128 lanes/clk/SM)
                r.update(bound="sfu", achieved=r["sfu_ops"] / r["sec"] / 1e12, peak=peaks["sfu_tops"], unit="Tops/s (MUFU)")
            else:
                r.update(bound="fp32", achieved=r["flops"] / r["sec"] / 1e12, peak=peaks["fp32_tflops"], unit="TFLOP/s")
        else:
            r.update(bound="hbm", achieved=r["bytes"] / r["sec"] / 1e9, peak=peaks["hbm_gbs"], unit="GB/s")
        r["frac"] = r["achieved"] / r["peak"]
    top = max(rows, key=lambda r: r["share_of_step"])
    out = {"roofline": {"kernel": top["kernel"], "bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
                        "frac": top["frac"], "traffic": _traffic(top["kernel"]), "peak_source": peaks["hbm_source"] if top["bound"] == "hbm"
                        else "FFMA / MUFU probe kernel (nfdpf_peak_probe) measured in this run", "share_of_step": top["share_of_step"]},
           "roofline_kernels": [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()} for r in rows],
           "peaks": peaks}
    try:
        out["ot_resample"] = ot_metric(dev, peaks)
    except Exception as e:
        out["ot_resample"] = {"error": repr(e)}
    try:
        out["block_density"] = block_density_metric(dev)
    except Exception as e:  # the headline must not die on an auxiliary measurement
        out["block_density"] = {"error": repr(e)}
    if not a.no_cpu_baseline:
        from bench_reference import cpu_filter_step, sample_shape
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        B, T = sample_shape(a)
        cpu_filter_step(a, min(B, 2), 1)   # warm-up
        t0 = time.perf_counter()
        n = 0
        while n < 2 and time.perf_counter() - t0 < 25:
            cpu_filter_step(a, B, T, n)
            n += 1
        sec = (time.perf_counter() - t0) / n
        out["cpu_baseline"] = {"value": B * a.N * T / sec, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                               "sample": "oracle port (torch CPU, %d threads) of the same filter step fwd+bwd on B=%d of %d trajectories x T=%d of %d "
                                         "steps at N=%d, mean of %d runs" % (cores, B, a.B, T, a.T, a.N, n)}
    return out
