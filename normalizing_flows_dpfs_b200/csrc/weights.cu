// (K2, weight half) log-weight update + max-shift softmax normalisation + ESS / log-likelihood row statistics.
// Replaces DPFs.py:187-192 + utils.py:39-44 (about ten ATen launches) with one pass: 20 B / particle.
// One warp per row when N <= 1024 (pure shuffle reductions), else one CTA per row.
#include "common.cuh"

namespace nfdpf {

template <bool BLOCK_PER_ROW>
__global__ void weight_update_fwd_kernel(const float* __restrict__ lw0, const float* __restrict__ lki,
                                         const float* __restrict__ prior, const float* __restrict__ propose, float add_eps,
                                         int B, int N, float* __restrict__ logw_out, float* __restrict__ probs_out,
                                         float* __restrict__ row_stats) {
    __shared__ float s_red[33];
    const int lane = threadIdx.x & 31;
    const int row = BLOCK_PER_ROW ? blockIdx.x : blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (!BLOCK_PER_ROW && row >= B) return;
    const int t0 = BLOCK_PER_ROW ? threadIdx.x : lane, stride = BLOCK_PER_ROW ? blockDim.x : 32;
    const size_t base = (size_t)row * N;
    auto red = [&](float v, auto op, float id) {
        if (BLOCK_PER_ROW) return block_allreduce(v, s_red, op, id);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL, v, o));
        return v;
    };
    float mx = -INFINITY, sl = 0.f;
    for (int n = t0; n < N; n += stride) {
        float v = lw0[base + n];
        if (lki) v += lki[base + n];          // (logw + lki + prior) - propose, DPFs.py:187 evaluation order
        if (prior) v += prior[base + n];
        if (propose) v -= propose[base + n];
        if (logw_out) logw_out[base + n] = v;
        probs_out[base + n] = v;              // stash; re-read by the same thread below
        mx = fmaxf(mx, v);
        sl += v;
    }
    mx = red(mx, OpMax(), -INFINITY);
    sl = red(sl, OpSum(), 0.f);
    float se = 0.f;
    for (int n = t0; n < N; n += stride) {
        const float e = expf(probs_out[base + n] - mx);
        probs_out[base + n] = e;
        se += e;
    }
    se = red(se, OpSum(), 0.f);
    float s2 = 0.f;
    for (int n = t0; n < N; n += stride) {
        const float p = __fdiv_rn(probs_out[base + n], se) + add_eps;   // utils.py:43, DPFs.py:192
        probs_out[base + n] = p;
        s2 += p * p;
    }
    s2 = red(s2, OpSum(), 0.f);
    if (t0 == 0 && row_stats) { row_stats[2 * row] = sl; row_stats[2 * row + 1] = 1.0f / s2; }
}

// N <= 1024, N % 4 == 0: one warp per row, the row in registers (a lane owns four consecutive entries of each 128-entry chunk): 128-bit
// loads issued up front, every tensor read and written ONCE (the kernel above stashes the row in probs_out and re-reads it twice:
// 25 us cold for 12.6 MB at B = N = 1024).
__global__ void __launch_bounds__(256)
weight_update_fwd_reg_kernel(const float* __restrict__ lw0, const float* __restrict__ lki, const float* __restrict__ prior,
                             const float* __restrict__ propose, float add_eps, int B, int N, float* __restrict__ logw_out,
                             float* __restrict__ probs_out, float* __restrict__ row_stats) {
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= B) return;
    const size_t base = (size_t)row * N;
    float v[8][4];
    float mx = -INFINITY, sl = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = 128 * i + 4 * lane;
        if (n < N) {
            float4 a = *reinterpret_cast<const float4*>(lw0 + base + n);
            if (lki) { const float4 t = *reinterpret_cast<const float4*>(lki + base + n); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
            if (prior) { const float4 t = *reinterpret_cast<const float4*>(prior + base + n); a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w; }
            if (propose) { const float4 t = *reinterpret_cast<const float4*>(propose + base + n); a.x -= t.x; a.y -= t.y; a.z -= t.z; a.w -= t.w; }
            if (logw_out) *reinterpret_cast<float4*>(logw_out + base + n) = a;     // (logw + lki + prior) - propose, DPFs.py:187
            v[i][0] = a.x; v[i][1] = a.y; v[i][2] = a.z; v[i][3] = a.w;
            mx = fmaxf(fmaxf(mx, fmaxf(a.x, a.y)), fmaxf(a.z, a.w));
            sl += (a.x + a.y) + (a.z + a.w);
        } else {
            v[i][0] = v[i][1] = v[i][2] = v[i][3] = -INFINITY;
        }
    }
    mx = warp_max(mx);
    sl = warp_sum(sl);
    float se = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float e = 128 * i + 4 * lane < N ? expf(v[i][u] - mx) : 0.f;
            v[i][u] = e;
            se += e;
        }
    se = warp_sum(se);
    float s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = 128 * i + 4 * lane;
        if (n < N) {
            float4 p;
            p.x = __fdiv_rn(v[i][0], se) + add_eps; p.y = __fdiv_rn(v[i][1], se) + add_eps;      // utils.py:43, DPFs.py:192
            p.z = __fdiv_rn(v[i][2], se) + add_eps; p.w = __fdiv_rn(v[i][3], se) + add_eps;
            *reinterpret_cast<float4*>(probs_out + base + n) = p;
            s2 += (p.x * p.x + p.y * p.y) + (p.z * p.z + p.w * p.w);
        }
    }
    s2 = warp_sum(s2);
    if (lane == 0 && row_stats) { row_stats[2 * row] = sl; row_stats[2 * row + 1] = 1.0f / s2; }
}

// d logw = p (g - sum g p) + g_logw + g_rowsum, with p = softmax (the forward output minus add_eps).
template <bool BLOCK_PER_ROW>
__global__ void weight_update_bwd_kernel(const float* __restrict__ g_probs, const float* __restrict__ g_logw,
                                         const float* __restrict__ g_rowsum, const float* __restrict__ probs, float add_eps,
                                         int B, int N, float* __restrict__ d_logw, float* __restrict__ d_neg,
                                         const float* __restrict__ particles, const float* __restrict__ g_pred) {
    __shared__ float s_red[33];
    const int lane = threadIdx.x & 31;
    const int row = BLOCK_PER_ROW ? blockIdx.x : blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (!BLOCK_PER_ROW && row >= B) return;
    const int t0 = BLOCK_PER_ROW ? threadIdx.x : lane, stride = BLOCK_PER_ROW ? blockDim.x : 32;
    const size_t base = (size_t)row * N;
    // gradient reaching probs[n]: the caller's g_probs plus, with the fused prediction (losses.py:22), <g_pred[row], particles[n]>
    float gx = 0.f, gy = 0.f;
    if (g_pred) { gx = g_pred[2 * row]; gy = g_pred[2 * row + 1]; }
    auto gp = [&](int n) {
        float g = g_probs ? g_probs[base + n] : 0.f;
        if (g_pred) { const float2 x = *reinterpret_cast<const float2*>(particles + (base + n) * 2); g = fmaf(gx, x.x, fmaf(gy, x.y, g)); }
        return g;
    };
    const bool has_gp = g_probs || g_pred;
    float dot = 0.f;
    if (has_gp)
        for (int n = t0; n < N; n += stride) dot += gp(n) * (probs[base + n] - add_eps);
    if (BLOCK_PER_ROW) dot = block_allreduce(dot, s_red, OpSum(), 0.f);
    else dot = warp_sum(dot);
    const float gr = g_rowsum ? g_rowsum[row] : 0.f;
    for (int n = t0; n < N; n += stride) {
        float v = gr;
        if (has_gp) v += (probs[base + n] - add_eps) * (gp(n) - dot);
        if (g_logw) v += g_logw[base + n];
        d_logw[base + n] = v;
        if (d_neg) d_neg[base + n] = -v;      // the proposal term enters with a minus sign (DPFs.py:187)
    }
}

// N <= 1024, N % 4 == 0: one warp per row with the row in registers (a lane owns four consecutive entries of each 128-entry chunk):
// every load is a 128-bit one issued before the first use, probs / g_probs / particles are read ONCE (the kernel above walks the row
// twice with scalar loads, one dependent round trip per iteration).
__global__ void __launch_bounds__(256)
weight_update_bwd_reg_kernel(const float* __restrict__ g_probs, const float* __restrict__ g_logw, const float* __restrict__ g_rowsum,
                             const float* __restrict__ probs, float add_eps, int B, int N, float* __restrict__ d_logw,
                             float* __restrict__ d_neg, const float* __restrict__ particles, const float* __restrict__ g_pred) {
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= B) return;
    const size_t base = (size_t)row * N;
    float gx = 0.f, gy = 0.f;
    if (g_pred) { gx = g_pred[2 * row]; gy = g_pred[2 * row + 1]; }
    float pm[8][4], g[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = 128 * i + 4 * lane;
        float4 p4 = make_float4(add_eps, add_eps, add_eps, add_eps), g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (n < N) {
            p4 = *reinterpret_cast<const float4*>(probs + base + n);
            if (g_probs) g4 = *reinterpret_cast<const float4*>(g_probs + base + n);
            if (g_pred) {   // the fused prediction (losses.py:22): d pred / d probs[n] = particles[n]
                const float4 x0 = *reinterpret_cast<const float4*>(particles + (base + n) * 2);
                const float4 x1 = *reinterpret_cast<const float4*>(particles + (base + n) * 2 + 4);
                g4.x = fmaf(gx, x0.x, fmaf(gy, x0.y, g4.x)); g4.y = fmaf(gx, x0.z, fmaf(gy, x0.w, g4.y));
                g4.z = fmaf(gx, x1.x, fmaf(gy, x1.y, g4.z)); g4.w = fmaf(gx, x1.z, fmaf(gy, x1.w, g4.w));
            }
        }
        pm[i][0] = p4.x - add_eps; pm[i][1] = p4.y - add_eps; pm[i][2] = p4.z - add_eps; pm[i][3] = p4.w - add_eps;
        g[i][0] = g4.x; g[i][1] = g4.y; g[i][2] = g4.z; g[i][3] = g4.w;
    }
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int u = 0; u < 4; ++u) dot = fmaf(g[i][u], pm[i][u], dot);
    dot = warp_sum(dot);
    const float gr = g_rowsum ? g_rowsum[row] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = 128 * i + 4 * lane;
        if (n < N) {
            float4 v = make_float4(gr, gr, gr, gr);
            if (g_logw) {
                const float4 l = *reinterpret_cast<const float4*>(g_logw + base + n);
                v.x += l.x; v.y += l.y; v.z += l.z; v.w += l.w;
            }
            v.x += pm[i][0] * (g[i][0] - dot); v.y += pm[i][1] * (g[i][1] - dot);
            v.z += pm[i][2] * (g[i][2] - dot); v.w += pm[i][3] * (g[i][3] - dot);
            *reinterpret_cast<float4*>(d_logw + base + n) = v;
            if (d_neg) *reinterpret_cast<float4*>(d_neg + base + n) = make_float4(-v.x, -v.y, -v.z, -v.w);   // DPFs.py:187: minus sign
        }
    }
}

}  // namespace nfdpf

using namespace nfdpf;

extern "C" int nfdpf_weight_update_fwd(const float* logw_prev, const float* lki, const float* prior, const float* propose,
                                       float add_eps, int B, int N, float* logw_out, float* probs_out, float* row_stats,
                                       void* stream) {
    NFDPF_REQUIRE(logw_prev && probs_out, "weight_update_fwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0, "weight_update_fwd: B and N must be positive (got %d, %d)", B, N);
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 1024 && N % 4 == 0) {
        weight_update_fwd_reg_kernel<<<(B + 7) / 8, 256, 0, st>>>(logw_prev, lki, prior, propose, add_eps, B, N, logw_out, probs_out, row_stats);
    } else if (N <= 1024) {
        const int wpb = 8;
        weight_update_fwd_kernel<false><<<(B + wpb - 1) / wpb, wpb * 32, 0, st>>>(logw_prev, lki, prior, propose, add_eps, B, N,
                                                                                  logw_out, probs_out, row_stats);
    } else {
        weight_update_fwd_kernel<true><<<B, 512, 0, st>>>(logw_prev, lki, prior, propose, add_eps, B, N, logw_out, probs_out,
                                                         row_stats);
    }
    return check_launch("weight_update_fwd");
}

extern "C" int nfdpf_weight_update_bwd(const float* g_probs, const float* g_logw, const float* g_rowsum, const float* probs,
                                       float add_eps, int B, int N, float* d_logw, float* d_neg, const float* particles, const float* g_pred,
                                       void* stream) {
    NFDPF_REQUIRE(probs && d_logw, "weight_update_bwd: null pointer");
    NFDPF_REQUIRE(!g_pred || particles, "weight_update_bwd: the fused prediction gradient needs the particles");
    NFDPF_REQUIRE(B > 0 && N > 0, "weight_update_bwd: B and N must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 1024 && N % 4 == 0) {
        weight_update_bwd_reg_kernel<<<(B + 7) / 8, 256, 0, st>>>(g_probs, g_logw, g_rowsum, probs, add_eps, B, N, d_logw, d_neg, particles, g_pred);
    } else if (N <= 1024) {
        const int wpb = 8;
        weight_update_bwd_kernel<false><<<(B + wpb - 1) / wpb, wpb * 32, 0, st>>>(g_probs, g_logw, g_rowsum, probs, add_eps, B, N,
                                                                                  d_logw, d_neg, particles, g_pred);
    } else {
        weight_update_bwd_kernel<true><<<B, 512, 0, st>>>(g_probs, g_logw, g_rowsum, probs, add_eps, B, N, d_logw, d_neg, particles, g_pred);
    }
    return check_launch("weight_update_bwd");
}

// ---- per-trajectory moments: the (detached) flow context of model/models.py:309-310, 338-339 -------------------
// out[b, off + k] = mean_n x[b,n,k];  out[b, off + d + k] = std_n x[b,n,k] (unbiased, N-1), k < d.  One warp per row pair.
namespace nfdpf {
__global__ void row_moments_kernel(const float* __restrict__ x, int B, int N, int d, float* __restrict__ out, int out_stride,
                                   int out_off, const float* __restrict__ head) {
    __shared__ float s_red[33];
    const int b = blockIdx.x;
    if (head) for (int k = threadIdx.x; k < out_off; k += blockDim.x) out[(size_t)b * out_stride + k] = head[(size_t)b * out_off + k];
    const float* xr = x + (size_t)b * N * d;
    for (int k = 0; k < d; ++k) {
        float s = 0.f;
        for (int n = threadIdx.x; n < N; n += blockDim.x) s += xr[(size_t)n * d + k];
        const float mean = block_allreduce(s, s_red, OpSum(), 0.f) / (float)N;
        float v = 0.f;
        for (int n = threadIdx.x; n < N; n += blockDim.x) { const float t = xr[(size_t)n * d + k] - mean; v = fmaf(t, t, v); }
        v = block_allreduce(v, s_red, OpSum(), 0.f);
        if (threadIdx.x == 0) {
            out[(size_t)b * out_stride + out_off + k] = mean;
            out[(size_t)b * out_stride + out_off + d + k] = sqrtf(v / (float)(N - 1));   // N == 1 -> NaN like torch.std
        }
    }
}
// d == 2, N <= 8 * 256: the row is read ONCE (float2 per particle, up to eight per thread, kept in registers); both means come out of
// one pair of block reductions, both variances out of a second one -- two round trips instead of the generic kernel's four passes.
__global__ void __launch_bounds__(256) row_moments2_kernel(const float* __restrict__ x, int N, float* __restrict__ out, int out_stride,
                                                           int out_off, const float* __restrict__ head) {
    __shared__ float s_red[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    // the leading out_off columns of the context row that are not moments (the observation encoding of the proposal's context,
    // model/models.py:360-361) ride along instead of an ATen slice copy per timestep
    if (head && tid < out_off) out[(size_t)b * out_stride + tid] = head[(size_t)b * out_off + tid];
    const float2* xr = reinterpret_cast<const float2*>(x) + (size_t)b * N;
    float2 v[8];
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = tid + 256 * i;
        v[i] = n < N ? xr[n] : make_float2(0.f, 0.f);
        sx += v[i].x; sy += v[i].y;
    }
    __shared__ float s_a[16], s_b[16];
    block_sum2_256(sx, sy, s_a);
    const float mx = sx / (float)N, my = sy / (float)N;
    float vx = 0.f, vy = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (tid + 256 * i < N) { const float tx = v[i].x - mx, ty = v[i].y - my; vx = fmaf(tx, tx, vx); vy = fmaf(ty, ty, vy); }
    block_sum2_256(vx, vy, s_b);
    if (tid == 0) {
        float* o = out + (size_t)b * out_stride + out_off;
        o[0] = mx; o[1] = my;
        o[2] = sqrtf(vx / (float)(N - 1)); o[3] = sqrtf(vy / (float)(N - 1));     // N == 1 -> NaN like torch.std
    }
}
}  // namespace nfdpf

extern "C" int nfdpf_row_moments(const float* x, int B, int N, int d, float* out, int out_stride, int out_off, void* stream) {
    NFDPF_REQUIRE(x && out, "row_moments: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && d > 0 && out_stride >= out_off + 2 * d && out_off >= 0, "row_moments: bad sizes");
    if (d == 2 && N <= 8 * 256) nfdpf::row_moments2_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, N, out, out_stride, out_off, nullptr);
    else nfdpf::row_moments_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, B, N, d, out, out_stride, out_off, nullptr);
    return nfdpf::check_launch("row_moments");
}

extern "C" int nfdpf_row_moments_head(const float* x, int B, int N, int d, const float* head, int head_dim, float* out, int out_stride,
                                      void* stream) {
    NFDPF_REQUIRE(x && out && head, "row_moments_head: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && d > 0 && head_dim > 0 && head_dim <= 256 && out_stride >= head_dim + 2 * d, "row_moments_head: bad sizes");
    if (d == 2 && N <= 8 * 256) nfdpf::row_moments2_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, N, out, out_stride, head_dim, head);
    else nfdpf::row_moments_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x, B, N, d, out, out_stride, head_dim, head);
    return nfdpf::check_launch("row_moments_head");
}
