// Per-step pieces that keep the host out of the filter loop (DPFs.py:160-214):
//   ess_gate        : the whole-batch resampling decision  mean_b(1 / sum_n p^2) < N/2  (DPFs.py:163-165) as a DEVICE flag that the
//                     resampling kernels read -- no device-to-host synchronisation, so the real gate is CUDA-graph capturable;
//                     also draws the soft resampler's U(0, 1/N) offsets (resamplers.py:43) and advances the step counter of the RNG;
//   motion_moments_rng : motion_update (model/models.py:191-204) with the N(0, sigma^2) noise drawn in-kernel (Philox4x32-10,
//                     counter = particle index, step counter, launch tag) and written out once, fused with the context moments;
//   init_particles_rng : particle_initialization (utils.py:46-62) on the device;
//   gate_weights    : weights after an OT resample under the device gate: 1/N when it fired, the old ones otherwise (DPFs.py:166-170);
//   weighted_mean   : the supervised loss' prediction  sum_n w_n x_n  (losses.py:22) per step, with backward, so the (B,T,N,2) lists
//                     never enter the loss graph.
#include "common.cuh"

namespace nfdpf {

// ---- Philox4x32-10 (Salmon et al. 2011), the counter-based generator cuRAND / torch use --------------------------------
struct Philox {
    static constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __device__ static uint4 gen(uint4 c, uint2 k) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x, hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += W0; k.y += W1;
        }
        return c;
    }
};
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 9) + 0.5f) * (1.0f / 8388608.0f); }   // strictly inside (0, 1): 23 bits + half a step, exact in fp32
// two independent N(0, 1) draws from two 32-bit words (Box-Muller)
__device__ __forceinline__ float2 normal2(unsigned a, unsigned b) {
    const float r = sqrtf(-2.0f * __logf(u01(a)));
    float s, c;
    __sincosf(6.283185307179586f * u01(b), &s, &c);
    return make_float2(r * c, r * s);
}
// rng_state: int64[2] = {seed, step counter}; `tag` separates the consumers inside one step
__device__ __forceinline__ uint4 philox_at(const long long* __restrict__ rng_state, unsigned long long idx, unsigned tag) {
    const unsigned long long seed = (unsigned long long)rng_state[0], step = (unsigned long long)rng_state[1];
    return Philox::gen(make_uint4((unsigned)idx, (unsigned)(idx >> 32), (unsigned)step, tag ^ (unsigned)(step >> 32) * 0x9E3779B9u),
                       make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
}

// One block.  gate = (mean_b ess_inv[b] < N / 2), fp64 fixed-order sum (the reference's fp32 mean differs from it by an ulp at
// most -- decisions can only differ when the mean sits within 1e-7 relative of the threshold).  force: -1 = the rule, 0 / 1 = constant.
__global__ void __launch_bounds__(1024)
ess_gate_kernel(const float* __restrict__ ess_inv, int ess_stride, int B, int N, int force, long long* __restrict__ rng_state, int advance,
                int* __restrict__ gate_out, float* __restrict__ offsets_out) {
    __shared__ double s_red[33];
    double a = 0.0;
    if (force < 0)
        for (int b = threadIdx.x; b < B; b += blockDim.x) a += (double)ess_inv[(size_t)b * ess_stride];
    if (offsets_out && rng_state)
        for (int b = threadIdx.x; b < B; b += blockDim.x) offsets_out[b] = u01(philox_at(rng_state, (unsigned long long)b, 0x0ff5e7u).x) / (float)N;
    a = block_allreduce(a, s_red, OpSum(), 0.0);
    if (threadIdx.x == 0) {
        *gate_out = force < 0 ? ((a / (double)B) < 0.5 * (double)N ? 1 : 0) : force;
        if (rng_state && advance) rng_state[1] += 1;     // every consumer of this step has been enqueued BEFORE the next step's gate
    }
}

__global__ void __launch_bounds__(256)
motion_moments_rng_kernel(const float* __restrict__ x, const float* __restrict__ vel, const long long* __restrict__ rng_state,
                          float sigma, int N, float* __restrict__ out, float* __restrict__ noise_out, float* __restrict__ ctx,
                          int ctx_stride, int ctx_off) {
    __shared__ float s_red[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float2 v = reinterpret_cast<const float2*>(vel)[b];
    const float2* xr = reinterpret_cast<const float2*>(x) + (size_t)b * N;
    float2* nr = reinterpret_cast<float2*>(noise_out) + (size_t)b * N;
    float2* orow = reinterpret_cast<float2*>(out) + (size_t)b * N;
    float sx = 0.f, sy = 0.f;
    for (int n = tid; n < N; n += 256) {
        const uint4 r = philox_at(rng_state, (unsigned long long)b * N + n, 0x6d0710u);
        float2 e = normal2(r.x, r.y);
        e.x *= sigma; e.y *= sigma;
        const float2 p = xr[n];
        const float2 o = make_float2((p.x + v.x) + e.x, (p.y + v.y) + e.y);   // (particles + vel) + noise, models.py:196-202
        nr[n] = e;
        orow[n] = o;
        sx += o.x; sy += o.y;
    }
    if (!ctx) return;
    const float mx = block_allreduce(sx, s_red, OpSum(), 0.f) / (float)N;
    const float my = block_allreduce(sy, s_red, OpSum(), 0.f) / (float)N;
    float vx = 0.f, vy = 0.f;
    for (int n = tid; n < N; n += 256) {
        const float2 o = orow[n];
        vx = fmaf(o.x - mx, o.x - mx, vx); vy = fmaf(o.y - my, o.y - my, vy);
    }
    vx = block_allreduce(vx, s_red, OpSum(), 0.f);
    vy = block_allreduce(vy, s_red, OpSum(), 0.f);
    if (tid == 0) {
        float* c = ctx + (size_t)b * ctx_stride + ctx_off;
        c[0] = mx; c[1] = my; c[2] = sqrtf(vx / (float)(N - 1)); c[3] = sqrtf(vy / (float)(N - 1));
    }
}

// N <= 256 NI: the row stays in registers -- every load is issued before the first Philox round (the generic kernel walks the row
// with one DRAM round trip per iteration), the moved particles are not read back for the variances, two barriers instead of twelve.
template <int NI>
__global__ void __launch_bounds__(256)
motion_moments_rng_reg_kernel(const float* __restrict__ x, const float* __restrict__ vel, const long long* __restrict__ rng_state,
                              float sigma, int N, float* __restrict__ out, float* __restrict__ noise_out, float* __restrict__ ctx,
                              int ctx_stride, int ctx_off) {
    __shared__ float s_a[16], s_b[16];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float2 v = reinterpret_cast<const float2*>(vel)[b];
    const float2* xr = reinterpret_cast<const float2*>(x) + (size_t)b * N;
    float2* nr = reinterpret_cast<float2*>(noise_out) + (size_t)b * N;
    float2* orow = reinterpret_cast<float2*>(out) + (size_t)b * N;
    float2 p[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { const int n = tid + 256 * i; p[i] = n < N ? xr[n] : make_float2(0.f, 0.f); }
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const int n = tid + 256 * i;
        if (n < N) {
            const uint4 r = philox_at(rng_state, (unsigned long long)b * N + n, 0x6d0710u);
            float2 e = normal2(r.x, r.y);
            e.x *= sigma; e.y *= sigma;
            const float2 o = make_float2((p[i].x + v.x) + e.x, (p[i].y + v.y) + e.y);   // (particles + vel) + noise, models.py:196-202
            nr[n] = e;
            orow[n] = o;
            p[i] = o;
            sx += o.x; sy += o.y;
        }
    }
    if (!ctx) return;
    block_sum2_256(sx, sy, s_a);
    const float mx = sx / (float)N, my = sy / (float)N;
    float vx = 0.f, vy = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i)
        if (tid + 256 * i < N) { vx = fmaf(p[i].x - mx, p[i].x - mx, vx); vy = fmaf(p[i].y - my, p[i].y - my, vy); }
    block_sum2_256(vx, vy, s_b);
    if (tid == 0) {
        float* c = ctx + (size_t)b * ctx_stride + ctx_off;
        c[0] = mx; c[1] = my; c[2] = sqrtf(vx / (float)(N - 1)); c[3] = sqrtf(vy / (float)(N - 1));
    }
}

// utils.py:46-62: uniform box [-width/2, width/2)^2, or start_state + N(0, 1) when init_with_true_state
__global__ void init_particles_rng_kernel(const float* __restrict__ start, int start_stride, const long long* __restrict__ rng_state,
                                          float width, int true_state, size_t P, int N, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const uint4 r = philox_at(rng_state, i, 0x1417u);
    float2 o;
    if (true_state) {
        const float2 e = normal2(r.x, r.y);
        const float* s = start + (i / N) * start_stride;
        o = make_float2(s[0] + e.x, s[1] + e.y);
    } else {
        o = make_float2(width * u01(r.x) - 0.5f * width, width * u01(r.y) - 0.5f * width);
    }
    reinterpret_cast<float2*>(out)[i] = o;
}

__global__ void gate_weights_fwd_kernel(const float* __restrict__ probs, const int* __restrict__ gate, size_t P, int N,
                                        float* __restrict__ w_out, float* __restrict__ logw_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float w = *gate ? 1.0f / (float)N : probs[i];
    w_out[i] = w;
    logw_out[i] = logf(w);
}
__global__ void gate_weights_bwd_kernel(const float* __restrict__ g_w, const float* __restrict__ g_logw, const float* __restrict__ probs,
                                        const int* __restrict__ gate, size_t P, float* __restrict__ d_probs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    float d = 0.f;
    if (!*gate) d = (g_w ? g_w[i] : 0.f) + (g_logw ? g_logw[i] / probs[i] : 0.f);
    d_probs[i] = d;
}

// pred[b] = sum_n w[b,n] x[b,n,:]  (losses.py:22 for one timestep); one warp per row when N <= 2048, else one CTA
__global__ void __launch_bounds__(256)
weighted_mean_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int B, int N, float* __restrict__ pred) {
    __shared__ float s_red[33];
    const int b = blockIdx.x;
    const float2* xr = reinterpret_cast<const float2*>(x) + (size_t)b * N;
    const float* wr = w + (size_t)b * N;
    float ax = 0.f, ay = 0.f;
    for (int n = threadIdx.x; n < N; n += 256) { const float2 p = xr[n]; const float q = wr[n]; ax = fmaf(q, p.x, ax); ay = fmaf(q, p.y, ay); }
    ax = block_allreduce(ax, s_red, OpSum(), 0.f);
    ay = block_allreduce(ay, s_red, OpSum(), 0.f);
    if (threadIdx.x == 0) reinterpret_cast<float2*>(pred)[b] = make_float2(ax, ay);
}
__global__ void weighted_mean_bwd_kernel(const float* __restrict__ g_pred, const float* __restrict__ x, const float* __restrict__ w,
                                         size_t P, int N, float* __restrict__ d_x, float* __restrict__ d_w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float2 g = reinterpret_cast<const float2*>(g_pred)[i / N], p = reinterpret_cast<const float2*>(x)[i];
    const float q = w[i];
    if (d_x) reinterpret_cast<float2*>(d_x)[i] = make_float2(g.x * q, g.y * q);
    if (d_w) d_w[i] = fmaf(g.x, p.x, g.y * p.y);
}

__global__ void sum4_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                            const float* __restrict__ d, size_t n4, size_t n, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        float4 v = reinterpret_cast<const float4*>(a)[i];
        if (b) { const float4 u = reinterpret_cast<const float4*>(b)[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        if (c) { const float4 u = reinterpret_cast<const float4*>(c)[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        if (d) { const float4 u = reinterpret_cast<const float4*>(d)[i]; v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w; }
        reinterpret_cast<float4*>(out)[i] = v;
    }
    if (i == 0)
        for (size_t k = 4 * n4; k < n; ++k) out[k] = a[k] + (b ? b[k] : 0.f) + (c ? c[k] : 0.f) + (d ? d[k] : 0.f);
}

}  // namespace nfdpf

using namespace nfdpf;

extern "C" int nfdpf_sum4(const float* a, const float* b, const float* c, const float* d, int64_t n, float* out, void* stream) {
    NFDPF_REQUIRE(a && out && n > 0, "sum4: bad arguments");
    const size_t n4 = (size_t)n / 4;
    sum4_kernel<<<(unsigned)((n4 + 255) / 256 + (n4 == 0)), 256, 0, (cudaStream_t)stream>>>(a, b, c, d, n4, (size_t)n, out);
    return check_launch("sum4");
}

extern "C" int nfdpf_ess_gate(const float* ess_inv, int ess_stride, int B, int N, int force, int64_t* rng_state, int advance,
                              int32_t* gate_out, float* offsets_out, void* stream) {
    NFDPF_REQUIRE(gate_out && (force >= 0 || (ess_inv && ess_stride >= 1)) && B > 0 && N > 0, "ess_gate: bad arguments");
    NFDPF_REQUIRE(force >= -1 && force <= 1, "ess_gate: force must be -1 (the ESS rule), 0 or 1");
    ess_gate_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ess_inv, ess_stride, B, N, force, (long long*)rng_state, advance, gate_out, offsets_out);
    return check_launch("ess_gate");
}

extern "C" int nfdpf_motion_moments_rng(const float* particles, const float* vel, const int64_t* rng_state, float sigma, int B, int N, int d,
                                        float* out, float* noise_out, float* ctx, int ctx_stride, int ctx_off, void* stream) {
    NFDPF_REQUIRE(particles && vel && rng_state && out && noise_out, "motion_moments_rng: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && sigma >= 0.f, "motion_moments_rng: B and N must be positive");
    NFDPF_REQUIRE(!ctx || (ctx_off >= 0 && ctx_stride >= ctx_off + 4), "motion_moments_rng: context row too short");
    if (d != 2) { set_error("motion_moments_rng: built for state_dim 2, got %d", d); return NFDPF_ERR_UNSUPPORTED; }
    if (N <= 1024)
        motion_moments_rng_reg_kernel<4><<<B, 256, 0, (cudaStream_t)stream>>>(particles, vel, (const long long*)rng_state, sigma, N, out, noise_out,
                                                                             ctx, ctx_stride, ctx_off);
    else
        motion_moments_rng_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(particles, vel, (const long long*)rng_state, sigma, N, out, noise_out, ctx,
                                                                      ctx_stride, ctx_off);
    return check_launch("motion_moments_rng");
}

extern "C" int nfdpf_init_particles_rng(const float* start, int start_stride, const int64_t* rng_state, float width, int true_state, int B,
                                        int N, int d, float* out, void* stream) {
    NFDPF_REQUIRE(rng_state && out && B > 0 && N > 0 && (!true_state || (start && start_stride >= 2)), "init_particles_rng: bad arguments");
    if (d != 2) { set_error("init_particles_rng: built for state_dim 2, got %d", d); return NFDPF_ERR_UNSUPPORTED; }
    const size_t P = (size_t)B * N;
    init_particles_rng_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(start, start_stride, (const long long*)rng_state,
                                                                                           width, true_state, P, N, out);
    return check_launch("init_particles_rng");
}

extern "C" int nfdpf_gate_weights_fwd(const float* probs, const int32_t* gate, int B, int N, float* w_out, float* logw_out, void* stream) {
    NFDPF_REQUIRE(probs && gate && w_out && logw_out && B > 0 && N > 0, "gate_weights_fwd: bad arguments");
    const size_t P = (size_t)B * N;
    gate_weights_fwd_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(probs, gate, P, N, w_out, logw_out);
    return check_launch("gate_weights_fwd");
}

extern "C" int nfdpf_gate_weights_bwd(const float* g_w, const float* g_logw, const float* probs, const int32_t* gate, int B, int N,
                                      float* d_probs, void* stream) {
    NFDPF_REQUIRE(probs && gate && d_probs && B > 0 && N > 0, "gate_weights_bwd: bad arguments");
    const size_t P = (size_t)B * N;
    gate_weights_bwd_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_w, g_logw, probs, gate, P, d_probs);
    return check_launch("gate_weights_bwd");
}

extern "C" int nfdpf_weighted_mean_fwd(const float* particles, const float* probs, int B, int N, int d, float* pred, void* stream) {
    NFDPF_REQUIRE(particles && probs && pred && B > 0 && N > 0, "weighted_mean_fwd: bad arguments");
    if (d != 2) { set_error("weighted_mean_fwd: built for state_dim 2, got %d", d); return NFDPF_ERR_UNSUPPORTED; }
    weighted_mean_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(particles, probs, B, N, pred);
    return check_launch("weighted_mean_fwd");
}

extern "C" int nfdpf_weighted_mean_bwd(const float* g_pred, const float* particles, const float* probs, int B, int N, int d,
                                       float* d_particles, float* d_probs, void* stream) {
    NFDPF_REQUIRE(g_pred && particles && probs && (d_particles || d_probs) && B > 0 && N > 0, "weighted_mean_bwd: bad arguments");
    if (d != 2) { set_error("weighted_mean_bwd: built for state_dim 2, got %d", d); return NFDPF_ERR_UNSUPPORTED; }
    const size_t P = (size_t)B * N;
    weighted_mean_bwd_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_pred, particles, probs, P, N, d_particles, d_probs);
    return check_launch("weighted_mean_bwd");
}
