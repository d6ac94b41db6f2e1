"""torch.autograd.Function wrappers over the C-ABI.  Each forward/backward is one (or two) kernel launches on
the current CUDA stream; tensors are allocated by torch, the library only sees raw pointers."""
import os

import torch

from . import _lib as L


_DEFER_D2 = os.environ.get("NFDPF_DEFER_D2", "1") != "0"     # A/B switch of the deferred D = 2 parameter-gradient reduction


def _slab_row(packed):
    """(GradSlab, row) when `packed` comes from a module's pack cache (nf/flows.py): the backward then writes its parameter gradient
    into that row and returns None for the vector -- the rows are summed once, in the pack's own backward."""
    if packed is None:
        return None
    slab = getattr(packed, "_nfdpf_slab", None)
    if slab is None or not packed.requires_grad:
        return None
    return slab, slab.take()


def _out(out, key, like=None, shape=None, dtype=torch.float32, device=None):
    """Output tensor of a kernel: the caller's pre-allocated slice `out[key]` (the per-step (T,B,...) list buffers of the filter
    loop: no torch.stack afterwards) or a fresh allocation.  `out` is a plain dict, invisible to autograd; the kernels write
    through raw pointers, so returning such a view from a Function is legal."""
    if out is not None and key in out:
        t = out[key]
        assert t.is_contiguous() and t.dtype == dtype, "pre-allocated output %r must be contiguous %s" % (key, dtype)
        return t
    if like is not None:
        return torch.empty_like(like)
    return torch.empty(shape, dtype=dtype, device=device)


def _stack_dims(packed, D, C, n_flows):
    half = D // 2
    per = 8 * (half + C) + 8 + 64 + 8 + half * 8 + half
    if packed.numel() != 4 * n_flows * per:
        raise ValueError("packed stack has %d parameters, expected %d for D=%d C=%d n_flows=%d" %
                         (packed.numel(), 4 * n_flows * per, D, C, n_flows))


class SoftResample(torch.autograd.Function):
    """soft_resampler, resamplers/resamplers.py:20-60 (+ the log of the new weights, DPFs.py:167, when want_log).
    gate: device int32 tensor or None -- the ESS decision taken on the device (ess_gate); closed gate = pass-through."""

    @staticmethod
    def forward(ctx, particles, probs, offsets, markers, alpha, want_log, gate=None, out=None):
        B, N, d = particles.shape
        p, w = L.f32(particles), L.f32(probs)
        off, mk = L.f32(offsets), L.f32(markers)
        p_out = torch.empty_like(p)
        w_out = torch.empty_like(w)
        lw_out = torch.empty_like(w) if want_log else None
        idx = _out(out, "index", shape=(B, N), dtype=torch.int64, device=p.device)
        saved = torch.empty(B, 2, dtype=torch.float32, device=p.device)
        L.call("nfdpf_soft_resample_fwd", L.ptr(p), L.ptr(w), L.ptr(off), L.ptr(mk), float(alpha), B, N, d, L.ptr(p_out),
               L.ptr(w_out), L.ptr(idx), L.ptr(saved), L.ptr(lw_out), L.ptr(gate), L.stream())
        ctx.save_for_backward(w, idx, saved, gate)
        ctx.set_materialize_grads(False)     # unused outputs arrive as None, not as zero-filled tensors
        ctx.alpha, ctx.shape = float(alpha), (B, N, d)
        ctx.mark_non_differentiable(idx)
        return p_out, w_out, idx, lw_out

    @staticmethod
    def backward(ctx, g_p, g_w, _g_idx, g_lw):
        w, idx, saved, gate = ctx.saved_tensors
        B, N, d = ctx.shape
        g_p = L.f32(g_p) if g_p is not None else None
        g_w = L.f32(g_w) if g_w is not None else None
        g_lw = L.f32(g_lw) if g_lw is not None else None
        d_p = torch.empty(B, N, d, dtype=torch.float32, device=w.device)
        d_w = torch.empty(B, N, dtype=torch.float32, device=w.device)
        L.call("nfdpf_soft_resample_bwd", L.ptr(g_p), L.ptr(g_w), L.ptr(w), L.ptr(idx), L.ptr(saved), ctx.alpha, B, N, d,
               L.ptr(d_p), L.ptr(d_w), L.ptr(g_lw), L.ptr(gate), L.stream())
        return d_p, d_w, None, None, None, None, None, None


class WeightUpdate(torch.autograd.Function):
    """logw = logw_prev + lki + prior - propose; probs = softmax(logw) + eps; row stats (DPFs.py:187-192)."""

    @staticmethod
    def forward(ctx, logw_prev, lki, prior, propose, add_eps, out=None):
        B, N = logw_prev.shape
        a = L.f32(logw_prev)
        terms = [L.f32(t) if t is not None else None for t in (lki, prior, propose)]
        logw = torch.empty_like(a)
        probs = _out(out, "probs", like=a)
        stats = torch.empty(B, 2, dtype=torch.float32, device=a.device)
        L.call("nfdpf_weight_update_fwd", L.ptr(a), L.ptr(terms[0]), L.ptr(terms[1]), L.ptr(terms[2]), float(add_eps), B, N,
               L.ptr(logw), L.ptr(probs), L.ptr(stats), L.stream())
        ctx.save_for_backward(probs)
        ctx.set_materialize_grads(False)
        ctx.add_eps = float(add_eps)
        ctx.has = [t is not None for t in (lki, prior, propose)]
        row_sum, ess_inv = stats[:, 0], stats[:, 1]
        ctx.mark_non_differentiable(ess_inv)
        return logw, probs, row_sum, ess_inv

    @staticmethod
    def backward(ctx, g_logw, g_probs, g_rowsum, _g_ess):
        (probs,) = ctx.saved_tensors
        B, N = probs.shape
        g_logw = L.f32(g_logw) if g_logw is not None else None
        g_probs = L.f32(g_probs) if g_probs is not None else None
        g_rowsum = L.f32(g_rowsum) if g_rowsum is not None else None
        d = torch.empty_like(probs)
        dn = torch.empty_like(probs) if ctx.has[2] else None
        L.call("nfdpf_weight_update_bwd", L.ptr(g_probs), L.ptr(g_logw), L.ptr(g_rowsum), L.ptr(probs), ctx.add_eps, B, N,
               L.ptr(d), L.ptr(dn), None, None, L.stream())
        return d, (d if ctx.has[0] else None), (d if ctx.has[1] else None), dn, None, None


class CouplingStack(torch.autograd.Function):
    """Fused (conditional) RealNVP stack, forward or inverse (nf/flows.py:215-239, nf/models.py:45-61).

    x (B,N,D); row_ctx (B,C_row) or None; part_ctx (B,N,C_part) or None; packed = flat parameters."""

    @staticmethod
    def forward(ctx, packed, x, row_ctx, part_ctx, n_flows, inverse, neg_logdet=False, out=None):
        B, N, D = x.shape
        C_row = 0 if row_ctx is None else row_ctx.shape[-1]
        C_part = 0 if part_ctx is None else part_ctx.shape[-1]
        _stack_dims(packed, D, C_row + C_part, n_flows)
        pk, xx = L.f32(packed), L.f32(x)
        rc = L.f32(row_ctx) if row_ctx is not None else None
        pc = L.f32(part_ctx) if part_ctx is not None else None
        y = _out(out, "y", like=xx)
        ld = _out(out, "log_det", shape=(B, N), device=xx.device)
        flags = int(bool(inverse)) | (2 if neg_logdet else 0)
        L.call("nfdpf_coupling_fwd", L.ptr(pk), n_flows, D, C_row, C_part, L.ptr(xx), L.ptr(rc), L.ptr(pc), flags, B, N,
               L.ptr(y), L.ptr(ld), L.stream())
        ctx.save_for_backward(pk, y, rc, pc)
        ctx.set_materialize_grads(False)
        ctx.meta = (n_flows, D, C_row, C_part, flags, B, N)
        ctx.slab = ctx.block = None
        slab = getattr(packed, "_nfdpf_slab", None) if packed.requires_grad else None
        if slab is not None:
            # D = 2 stacks with row context: the call's partial gradient rows wait in a block of the slab for ONE reduce launch per
            # training step (nfdpf_coupling_bwd_deferred / _reduce) instead of one small reduce launch per call
            bf = L.load().nfdpf_coupling_bwd_block_floats(n_flows, D, C_row, C_part, B) if _DEFER_D2 else 0
            if bf > 0:
                key = (n_flows, D, C_row, B)
                ctx.block = (slab, key, slab.take_block(key), bf)
            else:
                ctx.slab = (slab, slab.take())
        return y, ld

    @staticmethod
    def backward(ctx, g_y, g_ld):
        pk, y, rc, pc = ctx.saved_tensors
        n_flows, D, C_row, C_part, inverse, B, N = ctx.meta
        g_y = L.f32(g_y) if g_y is not None else None
        g_ld = L.f32(g_ld) if g_ld is not None else None
        d_x = torch.empty_like(y)
        need_rc = rc is not None and ctx.needs_input_grad[2]
        need_pc = pc is not None and ctx.needs_input_grad[3]
        d_rc = torch.empty_like(rc) if need_rc else None
        d_pc = torch.empty_like(pc) if need_pc else None
        ws_bytes = L.load().nfdpf_coupling_bwd_workspace(n_flows, D, C_row, C_part, B, N)
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=y.device)
        if ctx.block:
            slab, key, i, bf = ctx.block
            blk = slab.block(key, i, bf)
            L.call("nfdpf_coupling_bwd_deferred", L.ptr(pk), n_flows, D, C_row, C_part, L.ptr(y), L.ptr(rc), inverse, B, N,
                   L.ptr(g_y), L.ptr(g_ld), L.ptr(d_x), L.ptr(d_rc), L.ptr(blk), L.ptr(ws), L.stream())
            return None, d_x, d_rc, d_pc, None, None, None, None
        d_pk = ctx.slab[0].row(ctx.slab[1]) if ctx.slab else torch.empty_like(pk)        # every entry is written by the reduce kernel
        L.call("nfdpf_coupling_bwd", L.ptr(pk), n_flows, D, C_row, C_part, L.ptr(y), L.ptr(rc), L.ptr(pc), inverse, B, N,
               L.ptr(g_y), L.ptr(g_ld), L.ptr(d_x), L.ptr(d_rc), L.ptr(d_pc), L.ptr(d_pk), L.ptr(ws), L.stream())
        return (None if ctx.slab else d_pk), d_x, d_rc, d_pc, None, None, None, None


def coupling_bwd_reduce(key, blocks, numel):
    """Parameter gradient of `blocks.shape[0]` deferred D = 2 backward calls of shape key = (n_flows, D, C_row, B): one launch."""
    n_flows, D, C_row, B = key
    assert blocks.is_contiguous()
    d_pk = torch.empty(numel, dtype=torch.float32, device=blocks.device)
    L.call("nfdpf_coupling_bwd_reduce", n_flows, D, C_row, 0, B, L.ptr(blocks), blocks.shape[0], L.ptr(d_pk), L.stream())
    return d_pk


def soft_resample(particles, probs, offsets, markers, alpha, want_log=False, gate=None, out=None):
    """(particles', probs', flat idx) or, with want_log, (particles', probs', flat idx, log probs')."""
    res = SoftResample.apply(particles, probs, offsets, markers, alpha, want_log, gate, out)
    return res if want_log else res[:3]


def weight_update(logw_prev, lki=None, prior=None, propose=None, add_eps=0.0, out=None):
    return WeightUpdate.apply(logw_prev, lki, prior, propose, add_eps, out)


def coupling_stack(packed, x, row_ctx=None, part_ctx=None, n_flows=2, inverse=False, neg_logdet=False, out=None):
    """(y, log_det) -- or (y, jac = -log_det) with neg_logdet (model/models.py:325, 350).  out: optional dict with
    pre-allocated "y" / "log_det" tensors."""
    return CouplingStack.apply(packed, x, row_ctx, part_ctx, n_flows, inverse, neg_logdet, out)


MEASURE_MODES = {"gaussian": 0, "cos": 1, "CRNVP": 2, "NN": 3}



class MeasureUpdate(torch.autograd.Function):
    """Measurement log-likelihood (model/models.py:206-278) fused with DPFs.py:187-192 when logw_prev is given.

    pe: packed particle encoder (1648,), cnf: packed D=32/C=32 stack or None, enc (B,32), particles (B,N,2).
    Returns (lki, logw, probs, row_sum, ess_inv, pred); all but lki are None when not fused, pred (B,2) = sum_n probs particles
    (the prediction of the supervised loss, losses.py:22, formed in the kernel's epilogue) is None unless want_pred."""

    @staticmethod
    def forward(ctx, pe, cnf, enc, particles, logw_prev, prior, propose, mode, n_flows, p0, p1, add_eps, out=None, want_pred=False):
        B, N, d = particles.shape
        if d != 2:
            raise ValueError("measurement kernels take 2-d particle states (DPFs.py:31), got d=%d" % d)
        hidden = enc.shape[-1]
        pe_, enc_, x_ = L.f32(pe), L.f32(enc), L.f32(particles)
        cnf_ = L.f32(cnf) if cnf is not None else None
        if pe_.numel() != 1648:
            raise ValueError("particle encoder must be Linear(2,16)-Linear(16,32)-Linear(32,32): 1648 parameters, got %d" % pe_.numel())
        if mode == 3 and (cnf_ is None or cnf_.numel() != 8385):
            raise ValueError("the NN likelihood head must be Linear(64,64)-Linear(64,64)-Linear(64,1): 8385 parameters")
        fused = logw_prev is not None
        lw0 = L.f32(logw_prev) if fused else None
        pr = L.f32(prior) if prior is not None else None
        pp = L.f32(propose) if propose is not None else None
        dev = x_.device
        lki = _out(out, "lki", shape=(B, N), device=dev)
        argmax = torch.empty(B, dtype=torch.int32, device=dev)
        logw = torch.empty(B, N, dtype=torch.float32, device=dev) if fused else None
        probs = _out(out, "probs", shape=(B, N), device=dev) if fused else None
        stats = torch.empty(B, 2, dtype=torch.float32, device=dev) if fused else None
        pred = _out(out, "pred", shape=(B, 2), device=dev) if (fused and want_pred) else None
        need_grad = any(ctx.needs_input_grad[:4])
        z = torch.empty(B, N, hidden, dtype=torch.float32, device=dev) if (mode == 2 and need_grad) else None   # flow output, for the backward
        L.call("nfdpf_measure_fwd", mode, L.ptr(pe_), L.ptr(cnf_), n_flows, float(p0), float(p1), L.ptr(enc_), L.ptr(x_), B, N, hidden,
               L.ptr(lw0), L.ptr(pr), L.ptr(pp), float(add_eps), L.ptr(lki), L.ptr(argmax), L.ptr(logw), L.ptr(probs), L.ptr(stats),
               L.ptr(z), L.ptr(pred), L.stream())
        ctx.save_for_backward(pe_, cnf_, enc_, x_, argmax, probs, z)
        ctx.set_materialize_grads(False)
        ctx.meta = (mode, n_flows, float(p0), float(p1), float(add_eps), B, N, hidden, fused, prior is not None, propose is not None)
        ctx.slabs = (_slab_row(pe), _slab_row(cnf))
        if not fused:
            return lki, None, None, None, None, None
        row_sum, ess_inv = stats[:, 0], stats[:, 1]
        ctx.mark_non_differentiable(ess_inv)
        return lki, logw, probs, row_sum, ess_inv, pred

    @staticmethod
    def backward(ctx, g_lki, g_logw, g_probs, g_rowsum, _g_ess, g_pred):
        pe_, cnf_, enc_, x_, argmax, probs, z = ctx.saved_tensors
        mode, n_flows, p0, p1, add_eps, B, N, hidden, fused, has_prior, has_prop = ctx.meta
        dev = x_.device
        d_logw = d_neg = None
        g_total = L.f32(g_lki) if g_lki is not None else None
        g_pred = L.f32(g_pred) if g_pred is not None else None
        if fused and any(g is not None for g in (g_logw, g_probs, g_rowsum, g_pred)):
            d_logw = torch.empty(B, N, dtype=torch.float32, device=dev)
            d_neg = torch.empty(B, N, dtype=torch.float32, device=dev) if has_prop else None
            L.call("nfdpf_weight_update_bwd", L.ptr(L.f32(g_probs) if g_probs is not None else None),
                   L.ptr(L.f32(g_logw) if g_logw is not None else None),
                   L.ptr(L.f32(g_rowsum) if g_rowsum is not None else None), L.ptr(probs), add_eps, B, N, L.ptr(d_logw), L.ptr(d_neg),
                   L.ptr(x_) if g_pred is not None else None, L.ptr(g_pred), L.stream())
            g_total = d_logw if g_total is None else g_total + d_logw
        if g_total is None:
            g_total = torch.zeros(B, N, dtype=torch.float32, device=dev)
        d_x = torch.empty_like(x_)
        d_enc = torch.empty_like(enc_) if ctx.needs_input_grad[2] else None
        s_pe, s_cnf = ctx.slabs
        d_pe = s_pe[0].row(s_pe[1]) if s_pe else torch.empty_like(pe_)       # every entry is written by the reduce kernels
        d_cnf = (s_cnf[0].row(s_cnf[1]) if s_cnf else torch.empty_like(cnf_)) if cnf_ is not None else None
        ws = torch.empty(L.load().nfdpf_measure_bwd_workspace(mode, n_flows, B, N) // 4, dtype=torch.float32, device=dev)
        L.call("nfdpf_measure_bwd", mode, L.ptr(pe_), L.ptr(cnf_), n_flows, p0, p1, L.ptr(enc_), L.ptr(x_), B, N, hidden,
               L.ptr(g_total.contiguous()), L.ptr(argmax), L.ptr(d_x), L.ptr(d_enc), L.ptr(d_pe), L.ptr(d_cnf), L.ptr(ws), L.ptr(z),
               L.ptr(g_pred), L.ptr(probs) if g_pred is not None else None, L.stream())
        return (None if s_pe else d_pe, None if s_cnf else d_cnf, d_enc, d_x, d_logw if fused else None, d_logw if has_prior else None,
                d_neg if has_prop else None, None, None, None, None, None, None, None)


def measure(pe, cnf, enc, particles, mode, n_flows=2, p0=0.0, p1=1.0):
    """lki (B,N) only -- the measurement_model_*.forward of the reference."""
    return MeasureUpdate.apply(pe, cnf, enc, particles, None, None, None, MEASURE_MODES[mode], n_flows, p0, p1, 0.0, None, False)[0]


def measure_update(pe, cnf, enc, particles, logw_prev, prior, propose, mode, n_flows=2, p0=0.0, p1=1.0, add_eps=1e-12, out=None,
                   want_pred=False):
    """(lki, logw, probs, row_sum_logw, ess_inv) -- measurement + DPFs.py:187-192 in one kernel; with want_pred a sixth value,
    the supervised-loss prediction sum_n probs particles (losses.py:22)."""
    res = MeasureUpdate.apply(pe, cnf, enc, particles, logw_prev, prior, propose, MEASURE_MODES[mode], n_flows, p0, p1, add_eps, out, want_pred)
    return res if want_pred else res[:5]


def row_moments(x, out=None, out_off=0, head=None):
    """[mean | unbiased std] over the particle axis, (B,N,d) -> (B,2d), no autograd (the reference detaches it).  head (B, out_off):
    written into the leading columns of `out` by the same launch (the proposal context's observation encoding)."""
    B, N, d = x.shape
    xx = L.f32(x)
    if out is None:
        out = torch.empty(B, out_off + 2 * d, dtype=torch.float32, device=xx.device)
    if head is not None:
        hh = L.f32(head)
        assert hh.shape == (B, out_off), "head must be (B, out_off)"
        L.call("nfdpf_row_moments_head", L.ptr(xx), B, N, d, L.ptr(hh), out_off, L.ptr(out), out.shape[1], L.stream())
    else:
        L.call("nfdpf_row_moments", L.ptr(xx), B, N, d, L.ptr(out), out.shape[1], out_off, L.stream())
    return out


class OtResample(torch.autograd.Function):
    """particles' = T particles, T = Sinkhorn plan of (particles, logw) (resamplers.py:62-277); d out/d particles = T.
    gate: device int32 tensor or None (closed gate = identity plan, every Sinkhorn launch exits at once)."""

    last_iters = None  # device int32 tensor of the most recent call (the reference's total_iter + 2), for tests / reports

    @staticmethod
    def forward(ctx, particles, logw, eps, scaling, threshold, max_iter, gate=None):
        B, N, d = particles.shape
        x, lw = L.f32(particles), L.f32(logw)
        out = torch.empty_like(x)
        saved = torch.empty(B, N, 4, dtype=torch.float32, device=x.device)
        iters = torch.zeros(1, dtype=torch.int32, device=x.device)
        ws = torch.empty(L.load().nfdpf_ot_workspace(B, N) // 4 + 1, dtype=torch.float32, device=x.device)
        L.call("nfdpf_ot_resample_fwd", L.ptr(x), L.ptr(lw), float(eps), float(scaling), float(threshold), int(max_iter), B, N, d,
               L.ptr(out), L.ptr(saved), L.ptr(iters), L.ptr(ws), L.ptr(gate), L.stream())
        OtResample.last_iters = iters
        ctx.save_for_backward(saved, gate)
        ctx.meta = (float(eps), B, N, d)
        return out

    @staticmethod
    def backward(ctx, g_out):
        saved, gate = ctx.saved_tensors
        eps, B, N, d = ctx.meta
        g = L.f32(g_out)
        dx = torch.empty_like(g)
        L.call("nfdpf_ot_resample_bwd", L.ptr(g), L.ptr(saved), eps, B, N, d, L.ptr(dx), L.ptr(gate), L.stream())
        return dx, None, None, None, None, None, None


def ot_resample(particles, logw, eps=0.1, scaling=0.75, threshold=1e-3, max_iter=100, gate=None):
    return OtResample.apply(particles, logw, eps, scaling, threshold, max_iter, gate)


class GateWeights(torch.autograd.Function):
    """Weights after an OT resample under the device gate: (w, log w) = (1/N, -log N) if it fired, (probs, log probs) otherwise
    (DPFs.py:166-170)."""

    @staticmethod
    def forward(ctx, probs, gate):
        B, N = probs.shape
        p = L.f32(probs)
        w, lw = torch.empty_like(p), torch.empty_like(p)
        L.call("nfdpf_gate_weights_fwd", L.ptr(p), L.ptr(gate), B, N, L.ptr(w), L.ptr(lw), L.stream())
        ctx.save_for_backward(p, gate)
        ctx.set_materialize_grads(False)
        return w, lw

    @staticmethod
    def backward(ctx, g_w, g_lw):
        p, gate = ctx.saved_tensors
        B, N = p.shape
        d = torch.empty_like(p)
        L.call("nfdpf_gate_weights_bwd", L.ptr(L.f32(g_w) if g_w is not None else None), L.ptr(L.f32(g_lw) if g_lw is not None else None),
               L.ptr(p), L.ptr(gate), B, N, L.ptr(d), L.stream())
        return d, None


def gate_weights(probs, gate):
    return GateWeights.apply(probs, gate)


def ess_gate(ess_inv, B, N, force=None, rng_state=None, advance=False, want_offsets=False, gate_out=None):
    """Device-side ESS gate (DPFs.py:163-165): int32 flag tensor (and, if asked, U(0, 1/N) offsets (B,) drawn from rng_state).
    ess_inv: the (B,) 1 / sum p^2 column of the previous weight update (a strided view is fine); force: None (the rule) / bool."""
    dev = ess_inv.device if ess_inv is not None else rng_state.device
    gate = gate_out if gate_out is not None else torch.empty(1, dtype=torch.int32, device=dev)
    offsets = torch.empty(B, dtype=torch.float32, device=dev) if want_offsets else None
    L.call("nfdpf_ess_gate", None if ess_inv is None else ess_inv.data_ptr(), 0 if ess_inv is None else ess_inv.stride(0), B, N,
           -1 if force is None else int(bool(force)), L.ptr(rng_state), int(bool(advance)), L.ptr(gate), L.ptr(offsets), L.stream())
    return gate, offsets


class WeightedMean(torch.autograd.Function):
    """pred (B,2) = sum_n probs[b,n] particles[b,n,:] -- the prediction of the supervised loss for one timestep (losses.py:22)."""

    @staticmethod
    def forward(ctx, particles, probs, out=None):
        B, N, d = particles.shape
        x, w = L.f32(particles), L.f32(probs)
        pred = _out(out, "pred", shape=(B, d), device=x.device)
        L.call("nfdpf_weighted_mean_fwd", L.ptr(x), L.ptr(w), B, N, d, L.ptr(pred), L.stream())
        ctx.save_for_backward(x, w)
        return pred

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        B, N, d = x.shape
        d_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        d_w = torch.empty_like(w) if ctx.needs_input_grad[1] else None
        if d_x is None and d_w is None:
            return None, None, None
        L.call("nfdpf_weighted_mean_bwd", L.ptr(L.f32(g)), L.ptr(x), L.ptr(w), B, N, d, L.ptr(d_x), L.ptr(d_w), L.stream())
        return d_x, d_w, None


def weighted_mean(particles, probs, out=None):
    return WeightedMean.apply(particles, probs, out)


class Fanout(torch.autograd.Function):
    """n aliases of one tensor, one per consumer.  Forward: nothing is computed or copied.  Backward: the consumers' gradients
    are summed in ONE launch (nfdpf_sum4) -- autograd's own accumulation would chain n - 1 two-operand adds, each a launch and a
    full read-modify-write of the (B,N,2) gradient."""

    @staticmethod
    def forward(ctx, x, n):
        ctx.set_materialize_grads(False)
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *gs):
        gs = [L.f32(g) for g in gs if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        out = torch.empty_like(gs[0])
        while len(gs) > 1:       # up to four operands per pass
            take, gs = gs[:4], gs[4:]
            take += [None] * (4 - len(take))
            L.call("nfdpf_sum4", L.ptr(take[0]), L.ptr(take[1]), L.ptr(take[2]), L.ptr(take[3]), out.numel(), L.ptr(out), L.stream())
            gs = [out] + gs
        return out, None


def fanout(x, n):
    if not (torch.is_tensor(x) and x.requires_grad and torch.is_grad_enabled()):
        return (x,) * n
    return Fanout.apply(x, n)


def _list3(t, dtype):
    """(tensor, sb, st) of a (B,T,N) list whose rows the kernels can read in place (last dimension contiguous); the transposed
    views of the filter's (T,B,N) buffers qualify, anything else is copied once."""
    if not t.is_cuda:
        raise RuntimeError("libnfdpf has no CPU path: tensor is on %s" % t.device)
    t = t.detach()
    if t.dtype != dtype or t.stride(2) != 1:
        t = t.to(dtype).contiguous()
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError("libnfdpf: tensor lives on cuda:%d but the current device is cuda:%d" % (t.device.index, torch.cuda.current_device()))
    return t, t.stride(0), t.stride(1)


class BlockDensity(torch.autograd.Function):
    """Q (B,) of compute_block_density_nf (reference losses.py:37-70): one launch walks the ancestry of every block; the backward
    pushes the block-end weights down the (sorted) ancestor runs in a fixed order."""

    @staticmethod
    def forward(ctx, weights, lik, prior, index, block_len):
        B, T, N = weights.shape
        w, wsb, wst = _list3(weights, torch.float32)
        l, lsb, lst = _list3(lik, torch.float32)
        p, psb, pst = _list3(prior, torch.float32)
        i, isb, ist = _list3(index, torch.int64)
        nb = T // block_len
        Q = torch.empty(B, device=w.device, dtype=torch.float32)
        run = torch.empty(max(nb, 1), B, N, device=w.device, dtype=torch.float32)
        bad = torch.zeros(1, device=w.device, dtype=torch.int32)
        L.call("nfdpf_block_density_fwd", w.data_ptr(), wsb, wst, l.data_ptr(), lsb, lst, p.data_ptr(), psb, pst, i.data_ptr(), isb, ist,
               B, T, N, int(block_len), L.ptr(Q), L.ptr(run), L.ptr(bad), L.stream())
        ctx.save_for_backward(w, i, run, bad)
        ctx.block_len = int(block_len)
        return Q

    @staticmethod
    def backward(ctx, gQ):
        w, i, run, bad = ctx.saved_tensors
        B, T, N = w.shape
        if not torch.cuda.is_current_stream_capturing():
            flag = int(bad.item())
            if flag & 2:
                raise ValueError("block density: ancestor index out of range [0, B*N)")
            if flag & 1:
                raise RuntimeError("block density backward: an ancestor index points into another trajectory (unsupported)")
        # gradients as (T,B,N) buffers: ListView's backward hands contiguous (B,N) slices to the step kernels
        d_w, d_l, d_p = (torch.empty(T, B, N, device=w.device, dtype=torch.float32) for _ in range(3))
        L.call("nfdpf_block_density_bwd", L.ptr(L.f32(gQ)), w.data_ptr(), w.stride(0), w.stride(1), i.data_ptr(), i.stride(0), i.stride(1),
               L.ptr(run), B, T, N, ctx.block_len, L.ptr(d_w), L.ptr(d_l), L.ptr(d_p), N, B * N, L.stream())
        return d_w.transpose(0, 1), d_l.transpose(0, 1), d_p.transpose(0, 1), None, None


def block_density(weights, lik, prior, index, block_len):
    return BlockDensity.apply(weights, lik, prior, index, block_len)


class MotionMoments(torch.autograd.Function):
    """x' = (x + vel_b) + noise (model/models.py:191-204); writes the detached [mean | std] context of x' into ctx.
    noise given: injected draws (parity tests).  noise None: drawn in-kernel from rng_state (Philox) as sigma N(0,1) and
    returned as the second output."""

    @staticmethod
    def forward(ctx, particles, vel, noise, ctx_buf, ctx_off, rng_state=None, sigma=1.0, out=None):
        B, N, d = particles.shape
        x, v = L.f32(particles), L.f32(vel)
        res = _out(out, "moved", like=x)
        cs = 0 if ctx_buf is None else ctx_buf.shape[1]
        if noise is not None:
            e = L.f32(noise)
            L.call("nfdpf_motion_moments", L.ptr(x), L.ptr(v), L.ptr(e), B, N, d, L.ptr(res), L.ptr(ctx_buf), cs, ctx_off, L.stream())
        else:
            e = _out(out, "noise", like=x)
            L.call("nfdpf_motion_moments_rng", L.ptr(x), L.ptr(v), L.ptr(rng_state), float(sigma), B, N, d, L.ptr(res), L.ptr(e),
                   L.ptr(ctx_buf), cs, ctx_off, L.stream())
        ctx.mark_non_differentiable(e)
        ctx.set_materialize_grads(False)     # (autograd zero-filled a (B,N,2) gradient for the noise output on every backward call)
        return res, e

    @staticmethod
    def backward(ctx, g, _g_noise):
        return g, None, None, None, None, None, None, None


class ProposalTerms(torch.autograd.Function):
    """prior / proposal log-densities of proposal_likelihood (model/models.py:369-376)."""

    @staticmethod
    def forward(ctx, back, phys, noise, jac_back, jac_dyn, jac_prop, sigma, out=None):
        bk, ph, nz = L.f32(back), L.f32(phys), L.f32(noise)
        B, N, _ = bk.shape
        jb, jd, jp = (L.f32(t) if t is not None else None for t in (jac_back, jac_dyn, jac_prop))
        prior = _out(out, "prior", shape=(B, N), device=bk.device)
        propose = _out(out, "propose", shape=(B, N), device=bk.device)
        L.call("nfdpf_proposal_terms_fwd", L.ptr(bk), L.ptr(ph), L.ptr(nz), L.ptr(jb), L.ptr(jd), L.ptr(jp), float(sigma), B * N,
               L.ptr(prior), L.ptr(propose), L.stream())
        ctx.save_for_backward(bk, ph, nz)
        ctx.set_materialize_grads(False)
        ctx.sigma, ctx.has = float(sigma), (jac_back is not None, jac_dyn is not None, jac_prop is not None)
        return prior, propose

    @staticmethod
    def backward(ctx, g_prior, g_prop):
        bk, ph, nz = ctx.saved_tensors
        B, N, _ = bk.shape
        has_b, has_d, has_p = ctx.has
        d_back = d_phys = neg = None
        if g_prior is not None:
            gp = L.f32(g_prior)
            d_back, d_phys = torch.empty_like(bk), torch.empty_like(bk)
            neg = torch.empty_like(gp) if has_b else None
            L.call("nfdpf_proposal_terms_bwd", L.ptr(gp), L.ptr(bk), L.ptr(ph), L.ptr(nz), ctx.sigma, B * N, L.ptr(d_back), L.ptr(d_phys),
                   L.ptr(neg), L.stream())
        return d_back, d_phys, None, neg, (g_prop if has_d else None), (g_prop if has_p else None), None, None


def motion_moments(particles, vel, noise, ctx_buf=None, ctx_off=0, rng_state=None, sigma=1.0, out=None):
    """(particles', noise).  noise=None draws sigma N(0,1) in-kernel from rng_state (device int64[2] = seed, step counter)."""
    return MotionMoments.apply(particles, vel, noise, ctx_buf, ctx_off, rng_state, sigma, out)


def proposal_terms(back, phys, noise, jac_back, jac_dyn, jac_prop, sigma, out=None):
    return ProposalTerms.apply(back, phys, noise, jac_back, jac_dyn, jac_prop, sigma, out)


def init_particles(start_state, width, N, rng_state, init_with_true_state=False):
    """particle_initialization (utils.py:46-62) on the device: (B,N,2) uniform box or start + N(0,1)."""
    B = start_state.shape[0]
    st = L.f32(start_state)
    out = torch.empty(B, N, 2, dtype=torch.float32, device=st.device)
    L.call("nfdpf_init_particles_rng", L.ptr(st), st.shape[1], L.ptr(rng_state), float(width), int(bool(init_with_true_state)), B, N, 2,
           L.ptr(out), L.stream())
    return out


class ListView(torch.autograd.Function):
    """The (B,T,...) list of the filter outputs (DPFs.py:207-214) as a transposed VIEW of the (T,B,...) buffer the step kernels
    already wrote into -- no torch.stack / torch.cat copy.  Autograd-wise it is stack(steps, dim=1): the backward hands slice t of
    the incoming gradient to step t."""

    @staticmethod
    def forward(ctx, buf_holder, *steps):
        ctx.n = len(steps)
        return buf_holder[0].transpose(0, 1)

    @staticmethod
    def backward(ctx, g):
        # the step kernels want contiguous gradients: transpose the (B,T,...) gradient ONCE instead of one strided-slice copy per step
        if ctx.n > 1 and not g[:, 0].is_contiguous():
            gt = g.transpose(0, 1).contiguous()
            return (None,) + tuple(gt[t] for t in range(ctx.n))
        return (None,) + tuple(g[:, t] for t in range(ctx.n))


def list_view(buf, steps):
    if not any(torch.is_tensor(s) and s.requires_grad for s in steps):
        return buf.transpose(0, 1)
    return ListView.apply((buf,), *steps)
