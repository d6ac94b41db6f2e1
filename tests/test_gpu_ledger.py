"""GPU: the measured side of the tolerance ledger.  For the main kernels the same quantity is evaluated three ways -- CUDA (fp32),
the oracle in torch fp32 on the CPU (what the parity tests compare against) and the oracle in fp64 -- and the errors of the first
two against fp64 are recorded (gpurun_out/parity_errors.json -> profiles/r2_parity_errors.json).  A CUDA error of the order of
torch's own fp32 error is as close as two correct fp32 implementations can get; the asserts only require that (factor 4)."""
import contextlib

import numpy as np
import pytest
import torch

import ledger
import nfdpf_oracle as O
from normalizing_flows_dpfs_b200 import ops
from test_gpu_ops import _off_the_argmax_ties, _off_the_relu_kinks, _pe_tuple, cu

pytestmark = pytest.mark.gpu


@contextlib.contextmanager
def fp64_oracle():
    """The oracle follows the reference's `.float()` casts (nf/flows.py:114); for the fp64 yardstick they become no-ops."""
    orig = torch.Tensor.float
    torch.Tensor.float = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.float = orig


def _three_way(name, fn_oracle, fn_cuda, inputs, grads_wanted):
    """fn_oracle(*tensors) -> tuple of outputs (torch, CPU); fn_cuda likewise on CUDA tensors.  inputs: dict name -> fp32 CPU tensor;
    the scalar sum_i <out_i, g_i> with fixed random g is differentiated w.r.t. grads_wanted."""
    gen = torch.Generator().manual_seed(1)
    res = {}
    for kind in ("fp64", "fp32", "cuda"):
        if kind == "cuda":
            t = {k: v.detach().clone().cuda().requires_grad_(k in grads_wanted) for k, v in inputs.items()}
            outs = fn_cuda(**t)
        else:
            dt = torch.float64 if kind == "fp64" else torch.float32
            t = {k: v.detach().clone().to(dt).requires_grad_(k in grads_wanted) for k, v in inputs.items()}
            with (fp64_oracle() if kind == "fp64" else contextlib.nullcontext()):
                outs = fn_oracle(**t)
        gen.manual_seed(1)
        gs = [torch.randn(o.shape, generator=gen) for o in outs]
        sum((o * g.to(o.device, o.dtype)).sum() for o, g in zip(outs, gs)).backward()
        res[kind] = ([o.detach().cpu().double().numpy() for o in outs], {k: t[k].grad.detach().cpu().double().numpy() for k in grads_wanted})
    ref_o, ref_g = res["fp64"]
    for kind in ("fp32", "cuda"):
        tag = "torchfp32_vs_fp64" if kind == "fp32" else "cuda_vs_fp64"
        for i, (a, b) in enumerate(zip(res[kind][0], ref_o)):
            ledger.record("%s out%d" % (name, i), a, b, 1e-4, 1e-5, kind=tag)
        for k in grads_wanted:
            ledger.record("%s d_%s" % (name, k), res[kind][1][k], ref_g[k], 1e-4, 0.0, kind=tag)
    for k in grads_wanted:       # CUDA is as good an fp32 evaluation as torch's: same order of error against fp64
        scale = np.abs(ref_g[k]).max()
        e_cuda = np.abs(res["cuda"][1][k] - ref_g[k]).max() / scale
        e_fp32 = np.abs(res["fp32"][1][k] - ref_g[k]).max() / scale
        assert e_cuda <= 4 * e_fp32 + 2e-6, "%s d_%s: CUDA error %.2e vs torch fp32 error %.2e (relative to the tensor's max)" % (name, k, e_cuda, e_fp32)


@pytest.mark.parametrize("C,inverse", [(4, True), (36, True), (36, False)])
def test_ledger_coupling_stack(C, inverse):
    g = torch.Generator().manual_seed(C)
    B, N = 8, 1024
    inputs = dict(pk=O.init_stack(g, 2, C, std=0.3, bias_std=0.1), x=torch.randn(B, N, 2, generator=g) * 1.5, rc=torch.randn(B, C, generator=g))

    def oracle(pk, x, rc):
        ctx = rc[:, None, :].expand(B, N, C).reshape(B * N, C)
        y, ld = (O.stack_inverse if inverse else O.stack_forward)(x.reshape(B * N, 2), ctx, O.unpack_stack(pk, 2, C))
        return y.reshape(B, N, 2), ld.reshape(B, N)

    _three_way("coupling D=2 C=%d %s" % (C, "inverse" if inverse else "forward"), oracle,
               lambda pk, x, rc: ops.coupling_stack(pk, x, rc, None, 2, inverse), inputs, ("pk", "x"))


@pytest.mark.parametrize("mode,B,N", [("gaussian", 16, 1024), ("CRNVP", 4, 512)])
def test_ledger_measurement(mode, B, N):
    g = torch.Generator().manual_seed(B)
    pe = torch.cat([torch.randn(n, generator=g) * s for n, s in ((32, 0.3), (16, 0.1), (512, 0.3), (32, 0.1), (1024, 0.2), (32, 0.1))])
    cnf = O.init_stack(g, 32, 32, std=0.1, bias_std=0.05)
    enc = torch.randn(B, 32, generator=g)
    x = _off_the_relu_kinks(torch.randn(B, N, 2, generator=g) * 3, pe)
    fn = (lambda e, xx, p, c: O.measurement_gaussian(e, xx, _pe_tuple(p))) if mode == "gaussian" else (
        lambda e, xx, p, c: O.measurement_cnf(e, xx, _pe_tuple(p), O.unpack_stack(c, 32, 32), 2.5))
    x = _off_the_argmax_ties(x, lambda xx: fn(enc, xx, pe, cnf))
    p0, p1 = (1.0, 10.0) if mode == "gaussian" else (0.0, 2.5)
    inputs = dict(pe=pe, cnf=cnf, enc=enc, x=x)
    _three_way("measurement " + mode, lambda pe, cnf, enc, x: (fn(enc, x, pe, cnf),),
               lambda pe, cnf, enc, x: (ops.measure(pe, cnf if mode == "CRNVP" else None, enc, x, mode, p0=p0, p1=p1),), inputs,
               ("pe", "x") + (("cnf",) if mode == "CRNVP" else ()))
