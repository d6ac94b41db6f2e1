"""CUDA-graph execution of a whole filter training step (forward over T timesteps + loss + backward).

The per-timestep update is a chain of ~15 short libnfdpf launches (forward) plus as many in backward; at
B = N = 1024 each runs 20-400 us, so Python/ctypes/autograd dispatch (~40 us per launch) is comparable to the GPU
time of the forward pass.  Capturing the step once and replaying it removes the host from the loop: one
cudaGraphLaunch per training step.  Requirement (checked): random draws must be injected or generated on the device
(`dpf.rng_device = "cuda"`); the reference's ESS gate is then evaluated on the device (DPF.filtering_pos), so the REAL gate
is captured -- `force_resample` is not needed.  Drop every reference to the outputs / loss of earlier eager steps before
capturing (they keep autograd nodes bound to the default stream)."""
import torch
from torch.nn.utils.stateless import _reparametrize_module

from .losses import supervised_loss


class GraphedFilterStep:
    """graph = GraphedFilterStep(dpf, batch)   # batch: dict of device tensors enc, start, vel_in, state,
                                               #        init_particles, noise, offsets (static shapes)
       loss = graph.run(new_batch)              # copies new_batch into the static buffers, replays, returns the
                                               # (static) loss tensor; parameter .grad tensors hold the gradients."""

    KEYS = ("enc", "start", "vel_in", "state", "init_particles", "noise", "offsets")

    def __init__(self, dpf, batch, warmup=2):
        self.KEYS = tuple(k for k in self.KEYS if k in batch)   # noise / offsets may be left to the device RNG
        if ("noise" not in batch or "offsets" not in batch) and dpf.rng_device != "cuda":
            raise ValueError("graph capture without injected noise / offsets needs dpf.rng_device = 'cuda'")
        soft = dpf.param.resampler_type == "soft"
        if dpf.force_resample is None and soft and dpf.rng_device != "cuda" and "offsets" not in batch:
            raise ValueError("graph capture needs a host-free ESS gate: inject the resampling offsets or set dpf.rng_device = 'cuda'")
        self.dpf = dpf
        self.static = {k: batch[k].clone() for k in self.KEYS}
        for m in dpf.modules():   # packed-parameter caches pin autograd state of the stream they were built on
            for name in ("_cache", "_pe_cache"):
                if hasattr(m, name):
                    getattr(m, name).clear()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        dpf.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(self.graph):
                self.loss = self._step()
        except RuntimeError as err:
            raise RuntimeError(
                "CUDA-graph capture of the filter step failed.  The usual cause: tensors from an earlier eager step (loss, "
                "filter outputs) are still alive and pin the parameters' AccumulateGrad nodes to the default stream -- delete "
                "them before constructing GraphedFilterStep.  Original error: %s" % (err,)) from err

    def _step(self):
        d, dpf = self.static, self.dpf
        dpf.injected = {k: d[k] for k in ("init_particles", "noise", "offsets") if k in d}
        # Run on fresh leaf views of the parameters (same storage): their AccumulateGrad nodes are created on the current
        # (side / capture) stream.  The real parameters' accumulators may be pinned to the legacy default stream by any
        # still-alive tensor of an earlier eager step, which would make the capture illegal.
        names = [n for n, p in dpf.named_parameters() if p.requires_grad]
        real = dict(dpf.named_parameters())
        proxies = {n: real[n].detach().requires_grad_() for n in names}
        with _reparametrize_module(dpf, proxies):
            out = dpf.filtering_pos(d["enc"], d["start"], d["vel_in"])
            loss, _ = supervised_loss(out[0], out[1], d["state"], 1.0, False)
            grads = torch.autograd.grad(loss, [proxies[n] for n in names], allow_unused=True)
        for n, g in zip(names, grads):
            real[n].grad = g          # static graph outputs: every replay refreshes them in place
        return loss

    def run(self, batch=None):
        if batch is not None:
            for k in self.KEYS:
                self.static[k].copy_(batch[k], non_blocking=True)
        self.graph.replay()
        return self.loss
