"""torch.autograd.Function wrappers over the C-ABI.  Each forward/backward is one (or two) kernel launches on
the current CUDA stream; tensors are allocated by torch, the library only sees raw pointers."""
import torch

from . import _lib as L


def _stack_dims(packed, D, C, n_flows):
    half = D // 2
    per = 8 * (half + C) + 8 + 64 + 8 + half * 8 + half
    if packed.numel() != 4 * n_flows * per:
        raise ValueError("packed stack has %d parameters, expected %d for D=%d C=%d n_flows=%d" %
                         (packed.numel(), 4 * n_flows * per, D, C, n_flows))


class SoftResample(torch.autograd.Function):
    """soft_resampler, resamplers/resamplers.py:20-60."""

    @staticmethod
    def forward(ctx, particles, probs, offsets, markers, alpha):
        B, N, d = particles.shape
        p, w = L.f32(particles), L.f32(probs)
        off, mk = L.f32(offsets), L.f32(markers)
        p_out = torch.empty_like(p)
        w_out = torch.empty_like(w)
        idx = torch.empty(B, N, dtype=torch.int64, device=p.device)
        saved = torch.empty(B, 2, dtype=torch.float32, device=p.device)
        L.call("nfdpf_soft_resample_fwd", L.ptr(p), L.ptr(w), L.ptr(off), L.ptr(mk), float(alpha), B, N, d, L.ptr(p_out),
               L.ptr(w_out), L.ptr(idx), L.ptr(saved), L.stream())
        ctx.save_for_backward(w, idx, saved)
        ctx.alpha, ctx.shape = float(alpha), (B, N, d)
        ctx.mark_non_differentiable(idx)
        return p_out, w_out, idx

    @staticmethod
    def backward(ctx, g_p, g_w, _g_idx):
        w, idx, saved = ctx.saved_tensors
        B, N, d = ctx.shape
        g_p = L.f32(g_p) if g_p is not None else None
        g_w = L.f32(g_w) if g_w is not None else None
        d_p = torch.empty(B, N, d, dtype=torch.float32, device=w.device)
        d_w = torch.empty(B, N, dtype=torch.float32, device=w.device)
        L.call("nfdpf_soft_resample_bwd", L.ptr(g_p), L.ptr(g_w), L.ptr(w), L.ptr(idx), L.ptr(saved), ctx.alpha, B, N, d,
               L.ptr(d_p), L.ptr(d_w), L.stream())
        return d_p, d_w, None, None, None


class WeightUpdate(torch.autograd.Function):
    """logw = logw_prev + lki + prior - propose; probs = softmax(logw) + eps; row stats (DPFs.py:187-192)."""

    @staticmethod
    def forward(ctx, logw_prev, lki, prior, propose, add_eps):
        B, N = logw_prev.shape
        a = L.f32(logw_prev)
        terms = [L.f32(t) if t is not None else None for t in (lki, prior, propose)]
        logw = torch.empty_like(a)
        probs = torch.empty_like(a)
        stats = torch.empty(B, 2, dtype=torch.float32, device=a.device)
        L.call("nfdpf_weight_update_fwd", L.ptr(a), L.ptr(terms[0]), L.ptr(terms[1]), L.ptr(terms[2]), float(add_eps), B, N,
               L.ptr(logw), L.ptr(probs), L.ptr(stats), L.stream())
        ctx.save_for_backward(probs)
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
        L.call("nfdpf_weight_update_bwd", L.ptr(g_probs), L.ptr(g_logw), L.ptr(g_rowsum), L.ptr(probs), ctx.add_eps, B, N,
               L.ptr(d), L.stream())
        return d, (d if ctx.has[0] else None), (d if ctx.has[1] else None), (-d if ctx.has[2] else None), None


class CouplingStack(torch.autograd.Function):
    """Fused (conditional) RealNVP stack, forward or inverse (nf/flows.py:215-239, nf/models.py:45-61).

    x (B,N,D); row_ctx (B,C_row) or None; part_ctx (B,N,C_part) or None; packed = flat parameters."""

    @staticmethod
    def forward(ctx, packed, x, row_ctx, part_ctx, n_flows, inverse):
        B, N, D = x.shape
        C_row = 0 if row_ctx is None else row_ctx.shape[-1]
        C_part = 0 if part_ctx is None else part_ctx.shape[-1]
        _stack_dims(packed, D, C_row + C_part, n_flows)
        pk, xx = L.f32(packed), L.f32(x)
        rc = L.f32(row_ctx) if row_ctx is not None else None
        pc = L.f32(part_ctx) if part_ctx is not None else None
        y = torch.empty_like(xx)
        ld = torch.empty(B, N, dtype=torch.float32, device=xx.device)
        L.call("nfdpf_coupling_fwd", L.ptr(pk), n_flows, D, C_row, C_part, L.ptr(xx), L.ptr(rc), L.ptr(pc), int(inverse), B, N,
               L.ptr(y), L.ptr(ld), L.stream())
        ctx.save_for_backward(pk, y, rc, pc)
        ctx.meta = (n_flows, D, C_row, C_part, int(inverse), B, N)
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
        d_pk = torch.zeros_like(pk)
        ws_bytes = L.load().nfdpf_coupling_bwd_workspace(n_flows, D, C_row, C_part, B, N)
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=y.device)
        L.call("nfdpf_coupling_bwd", L.ptr(pk), n_flows, D, C_row, C_part, L.ptr(y), L.ptr(rc), L.ptr(pc), inverse, B, N,
               L.ptr(g_y), L.ptr(g_ld), L.ptr(d_x), L.ptr(d_rc), L.ptr(d_pc), L.ptr(d_pk), L.ptr(ws), L.stream())
        return d_pk, d_x, d_rc, d_pc, None, None


def soft_resample(particles, probs, offsets, markers, alpha):
    return SoftResample.apply(particles, probs, offsets, markers, alpha)


def weight_update(logw_prev, lki=None, prior=None, propose=None, add_eps=0.0):
    return WeightUpdate.apply(logw_prev, lki, prior, propose, add_eps)


def coupling_stack(packed, x, row_ctx=None, part_ctx=None, n_flows=2, inverse=False):
    return CouplingStack.apply(packed, x, row_ctx, part_ctx, n_flows, inverse)
