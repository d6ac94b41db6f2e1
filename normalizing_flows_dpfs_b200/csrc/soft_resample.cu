// (K3) Soft resampling -- one CTA per trajectory: mixture weights, ATen-order row sum, fp64 block scan,
// binary search of the caller's markers, gather, renormalise.  Replaces resamplers/resamplers.py:20-60, whose
// (B,N,N) bool compare (1 GiB at B=N=1024) is never materialised here.  HBM-bound: 32 B / particle forward.
#include "common.cuh"

namespace nfdpf {

// Bit-exact emulation of ATen's CPU `sum(dim=-1)` for a contiguous fp32 row (SumKernel.cpp cascade_sum ->
// vectorized_inner_sum -> row_sum -> multi_row_sum): 8 SIMD lanes x 4 ILP accumulators = 32 chains, chain c
// sums elements c, c+32, c+64, ... through a 4-level cascade (level step 2^max(4, ceil_log2(rows)/4)); then
// leftover whole vectors, ILP fold, scalar tail, lane fold.  One warp, lane = chain.  (oracle: cascade_row_sum)
__device__ float aten_row_sum_warp(const float* __restrict__ q, int n) {
    const int lane = threadIdx.x & 31;
    const int lanes = n >= 8 ? 8 : 1;       // rows shorter than a vector take ATen's scalar path
    const int chains = 4 * lanes;
    const int vec = n / lanes, rows = vec / 4;
    int lg = 0;
    while ((1 << lg) < rows) ++lg;          // ceil_log2(rows)
    const int p = max(4, lg / 4), step = 1 << p, mask = step - 1;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (lane < chains) {
        int i = 0;
        while (i + step <= rows) {
            for (int j = 0; j < step; ++j, ++i) a0 = __fadd_rn(a0, q[i * chains + lane]);
            a1 = __fadd_rn(a1, a0); a0 = 0.f;
            if ((i & (mask << p)) == 0) {
                a2 = __fadd_rn(a2, a1); a1 = 0.f;
                if ((i & (mask << (2 * p))) == 0) { a3 = __fadd_rn(a3, a2); a2 = 0.f; }
            }
        }
        for (; i < rows; ++i) a0 = __fadd_rn(a0, q[i * chains + lane]);
        a0 = __fadd_rn(a0, a1); a0 = __fadd_rn(a0, a2); a0 = __fadd_rn(a0, a3);
        if (lane < lanes)  // leftover whole vectors go to ILP accumulator 0
            for (int v = rows * 4; v < vec; ++v) a0 = __fadd_rn(q[v * lanes + lane], a0);
    }
    // ILP fold: ps[0] += ps[k], k = 1..3 (chain k*lanes + l lives in lane k*lanes + l)
    float t = a0;
    for (int k = 1; k < 4; ++k) {
        const float o = __shfl_sync(FULL, a0, (k * lanes + lane) & 31);
        t = __fadd_rn(t, o);
    }
    float out = 0.f;
    for (int k = vec * lanes; k < n; ++k) out = __fadd_rn(out, q[k]);  // scalar tail first
    for (int l = 0; l < lanes; ++l) out = __fadd_rn(out, __shfl_sync(FULL, t, l));
    return out;  // valid in every lane
}

// Block-wide inclusive scan in fp64 over s_q[0..n) (fp32 in, fp32-rounded prefixes out, in place).
// ATen's CPU cumsum accumulates a float row sequentially in double and rounds each prefix; with alpha < 1 every
// q_j >= (1-alpha)/(N*sum) so all partial sums are exactly representable in fp64 and the order is irrelevant.
__device__ void scan_fp64_inplace(float* s_q, int n, double* s_warp) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const int per = (n + nt - 1) / nt;
    const int lo = min(tid * per, n), hi = min(lo + per, n);
    double local = 0.0;
    for (int j = lo; j < hi; ++j) local += (double)s_q[j];
    double inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        double w = lane < nwarp ? s_warp[lane] : 0.0, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += v;
        }
        if (lane < nwarp) s_warp[lane] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    double run = s_warp[warp] + (inc - local);
    for (int j = lo; j < hi; ++j) {
        run += (double)s_q[j];
        s_q[j] = (float)run;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
soft_resample_fwd_kernel(const float* __restrict__ particles, const float* __restrict__ probs,
                         const float* __restrict__ offsets, const float* __restrict__ markers, float alpha_f,
                         float one_minus_alpha_f, int hard, int N, int d, float* __restrict__ particles_out,
                         float* __restrict__ probs_out, int64_t* __restrict__ idx_out, float* __restrict__ saved,
                         float* __restrict__ logprobs_out, const int* __restrict__ gate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (gate && *gate == 0) {   // device-side ESS gate closed (DPFs.py:168-170): particles / weights pass through, identity ancestors
        const int b = blockIdx.x;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const size_t o = (size_t)b * N + i;
            for (int k = 0; k < d; ++k) particles_out[o * d + k] = particles[o * d + k];
            const float w = probs[o];
            probs_out[o] = w;
            idx_out[o] = (int64_t)o;
            if (logprobs_out) logprobs_out[o] = logf(w);
        }
        if (threadIdx.x == 0) { saved[2 * b] = 1.f; saved[2 * b + 1] = 1.f; }
        return;
    }
    float* s_q = reinterpret_cast<float*>(smem_raw);  // q -> normalised q -> cum
    float* s_wis = s_q + N;                           // importance weights w / q
    __shared__ double s_warp[32];
    __shared__ float s_red[33];
    __shared__ float s_sum;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const float* w = probs + (size_t)b * N;
    const float unif = __fdiv_rn(1.0f, (float)N);                  // ones/N, resamplers.py:26
    const float uterm = __fmul_rn(unif, one_minus_alpha_f);        // uniform_probs*(1-alpha), :31
    for (int j = tid; j < N; j += nt)
        s_q[j] = hard ? w[j] : __fadd_rn(__fmul_rn(w[j], alpha_f), uterm);
    __syncthreads();
    float S = 1.0f;
    if (!hard) {
        if (tid < 32) {
            const float s = aten_row_sum_warp(s_q, N);
            if (tid == 0) s_sum = s;
        }
        __syncthreads();
        S = s_sum;
        for (int j = tid; j < N; j += nt) {
            const float qn = __fdiv_rn(s_q[j], S);                 // q / q.sum, :33
            s_q[j] = qn;
            s_wis[j] = __fdiv_rn(w[j], qn);                        // w / q, :34
        }
    } else {
        for (int j = tid; j < N; j += nt) s_wis[j] = unif;         // hard resampling, :36-38
    }
    __syncthreads();
    scan_fp64_inplace(s_q, N, s_warp);                             // cumsum, :45
    if (tid == 0) s_q[N - 1] = 1.0f;                               // cum[:, -1] = 1, :47
    __syncthreads();
    const float off = offsets[b];
    float part = 0.f;
    for (int i = tid; i < N; i += nt) {
        const float m = __fadd_rn(off, markers[i]);                // :44
        int lo = 0, hi = N - 1;  // count of cum[j] < m over the sorted prefix [0, N-1)  ('>' is strict, :49)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_q[mid] < m) lo = mid + 1; else hi = mid;
        }
        const int j = min(lo + (m > 1.0f ? 1 : 0), N - 1);         // forced last entry (markers beyond 1 only with out-of-range offsets)
        idx_out[(size_t)b * N + i] = (int64_t)j + (int64_t)N * b;  // :52
        const float* src = particles + ((size_t)b * N + j) * d;
        float* dst = particles_out + ((size_t)b * N + i) * d;
        if (d == 2) {
            *reinterpret_cast<float2*>(dst) = *reinterpret_cast<const float2*>(src);
        } else {
            for (int k = 0; k < d; ++k) dst[k] = src[k];
        }
        const float v = s_wis[j];
        probs_out[(size_t)b * N + i] = v;  // unnormalised for now (same thread rewrites it below)
        part += v;
    }
    const float S2 = block_allreduce(part, s_red, OpSum(), 0.f);
    for (int i = tid; i < N; i += nt) {
        const size_t o = (size_t)b * N + i;
        const float pn = __fdiv_rn(probs_out[o], S2);             // :56
        probs_out[o] = pn;
        if (logprobs_out) logprobs_out[o] = logf(pn);             // DPFs.py:167 (particle_probs_resampled.log())
    }
    if (tid == 0) { saved[2 * b] = S; saved[2 * b + 1] = S2; }
}

// Backward: indices carry no gradient; gradients flow through the gathered particles and through
// w/q (numerator and q), then the row renormalisation.  idx is monotone per row, so every source particle owns a
// contiguous run of destinations: the per-source sums are a SEGMENTED inclusive scan over the destinations
// (Hillis-Steele in shared memory; "same segment" == equal key because the keys are sorted), read off at the last
// element of each run.  fp32, fixed order, no atomics; work is balanced however peaked the weights are.
__global__ void __launch_bounds__(1024)
soft_resample_bwd_kernel(const float* __restrict__ g_particles, const float* __restrict__ g_probs,
                         const float* __restrict__ probs, const int64_t* __restrict__ idx,
                         const float* __restrict__ saved, float alpha_f, float one_minus_alpha_f, int hard, int N, int d,
                         float* __restrict__ d_particles, float* __restrict__ d_probs, const float* __restrict__ g_logprobs,
                         const int* __restrict__ gate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (gate && *gate == 0) {   // the forward passed everything through: so do the gradients (+ d log w = g / w)
        const int b = blockIdx.x;
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const size_t o = (size_t)b * N + i;
            for (int k = 0; k < d; ++k) d_particles[o * d + k] = g_particles ? g_particles[o * d + k] : 0.f;
            float g = g_probs ? g_probs[o] : 0.f;
            if (g_logprobs) g += g_logprobs[o] / probs[o];
            d_probs[o] = g;
        }
        return;
    }
    int* s_idx = reinterpret_cast<int*>(smem_raw);                    // [N] local source index per destination
    float4* s_a = reinterpret_cast<float4*>(s_idx + ((N + 3) & ~3));  // [N] (dL/dv, g_x, g_y, -), scanned in place
    __shared__ float s_red[33];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const float S = saved[2 * b], S2 = saved[2 * b + 1];
    const float* w = probs + (size_t)b * N;
    const size_t base = (size_t)b * N;
    const float unif = __fdiv_rn(1.0f, (float)N), uterm = __fmul_rn(unif, one_minus_alpha_f);
    auto qu = [&](int j) { return __fadd_rn(__fmul_rn(w[j], alpha_f), uterm); };
    float part = 0.f;
    for (int i = tid; i < N; i += nt) {
        const int j = (int)(idx[base + i] - (int64_t)N * b);
        s_idx[i] = j;
        float g = 0.f;   // total gradient reaching w'_i: direct + through log(w'_i)
        if (!hard && (g_probs || g_logprobs)) {
            const float wp = (w[j] * S / qu(j)) / S2;
            if (g_probs) g = g_probs[base + i];
            if (g_logprobs) g += g_logprobs[base + i] / wp;
            part += g * wp;                                  // sum_m g_m w'_m
        }
        float2 gp = make_float2(0.f, 0.f);
        if (g_particles && d == 2) gp = reinterpret_cast<const float2*>(g_particles)[base + i];
        s_a[i] = make_float4(g, gp.x, gp.y, 0.f);
    }
    const float c = block_allreduce(part, s_red, OpSum(), 0.f);
    for (int i = tid; i < N; i += nt) s_a[i].x = hard ? 0.f : (s_a[i].x - c) / S2;   // dL/dv_i, v_i = w_is[idx_i]
    __syncthreads();
    // Segmented inclusive scan of s_a over the runs of equal s_idx, three levels (a Hillis-Steele pass over all N elements cost
    // log2 N rounds of shared-memory traffic and barriers): (1) every thread scans its own chunk of consecutive destinations in
    // place; (2) the chunk tails (key of the last element, sum of its run inside the chunk) are scanned across threads with
    // warp shuffles and one cross-warp step -- a tail continues the previous thread's tail iff the chunk holds a single key
    // equal to it; (3) the resulting carry is added to the chunk's leading run.  fp32, fixed order.
    {
        __shared__ float s_wt[96];                            // cross-warp scratch: [32] x 3 tail sums / run totals
        __shared__ int s_wi[64];                              // [32] segment-start flags, [32] last keys
        const int lane = tid & 31, wid = tid >> 5;
        const int per = (N + nt - 1) / nt;
        const int lo = min(tid * per, N), hi = min(lo + per, N);
        float tx = 0.f, ty = 0.f, tz = 0.f;
        int kprev = -1;
        for (int i = lo; i < hi; ++i) {
            const int k = s_idx[i];
            float4 v = s_a[i];
            if (k == kprev) { v.x += tx; v.y += ty; v.z += tz; s_a[i] = v; }
            tx = v.x; ty = v.y; tz = v.z; kprev = k;
        }
        const bool has = hi > lo;
        const int kfirst = has ? s_idx[lo] : -2, klast = has ? kprev : -3;      // empty chunks (only at the end) never match anything
        if (lane == 31) s_wi[32 + wid] = klast;
        __syncthreads();
        int kleft = __shfl_up_sync(FULL, klast, 1);
        if (lane == 0) kleft = wid > 0 ? s_wi[32 + wid - 1] : -4;
        const bool cont = has && kfirst == kleft;                 // the chunk's leading run continues the previous chunk's tail
        int start = !(cont && kfirst == klast);                   // the tail starts a new segment unless the whole chunk continues it
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float ux = __shfl_up_sync(FULL, tx, o), uy = __shfl_up_sync(FULL, ty, o), uz = __shfl_up_sync(FULL, tz, o);
            const int us = __shfl_up_sync(FULL, start, o);
            if (lane >= o && !start) { tx += ux; ty += uy; tz += uz; start = us; }
        }
        if (lane == 31) { s_wt[wid] = tx; s_wt[32 + wid] = ty; s_wt[64 + wid] = tz; s_wi[wid] = start; }
        __syncthreads();
        if (!start && wid > 0) {                                  // segment reaches back beyond this warp: add the earlier warps' tails
            float cx = 0.f, cy = 0.f, cz = 0.f;
            for (int w = wid - 1; w >= 0; --w) {
                cx += s_wt[w]; cy += s_wt[32 + w]; cz += s_wt[64 + w];
                if (s_wi[w]) break;
            }
            tx += cx; ty += cy; tz += cz;
        }
        // (tx, ty, tz) = total of the run that ends with this chunk's tail; the next chunk's leading run continues it
        float px = __shfl_up_sync(FULL, tx, 1), py = __shfl_up_sync(FULL, ty, 1), pz = __shfl_up_sync(FULL, tz, 1);
        __syncthreads();                                          // s_wt reuse
        if (lane == 31) { s_wt[wid] = tx; s_wt[32 + wid] = ty; s_wt[64 + wid] = tz; }
        __syncthreads();
        if (lane == 0 && wid > 0) { px = s_wt[wid - 1]; py = s_wt[32 + wid - 1]; pz = s_wt[64 + wid - 1]; }
        if (cont) {
            for (int i = lo; i < hi && s_idx[i] == kfirst; ++i) { float4 v = s_a[i]; v.x += px; v.y += py; v.z += pz; s_a[i] = v; }
        }
        __syncthreads();
    }
    float4* src = s_a;
    // src[i] = inclusive segmented sums; a run ends at i when the next key differs
    for (int j = tid; j < N; j += nt) {
        d_probs[base + j] = 0.f;
        if (d == 2) reinterpret_cast<float2*>(d_particles)[base + j] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    float third = 0.f;
    for (int i = tid; i < N; i += nt) {
        const int j = s_idx[i];
        if (i + 1 < N && s_idx[i + 1] == j) continue;
        const float4 v = src[i];
        if (d == 2) reinterpret_cast<float2*>(d_particles)[base + j] = make_float2(v.y, v.z);
        if (!hard) {
            const float q = qu(j), wj = w[j], G = v.x;
            d_probs[base + j] = G * S / q - G * wj * S * alpha_f / (q * q);   // through w_is_j = w_j S / qu_j at fixed S
            third += G * wj / q;                                               // through S = sum_k qu_k
        }
    }
    if (d != 2) {  // generic state dimension: per-source ordered run sums (rare path)
        for (int j = tid; j < N; j += nt) {
            int lo = 0, hi = N;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (s_idx[mid] < j) lo = mid + 1; else hi = mid; }
            for (int k = 0; k < d; ++k) {
                float a = 0.f;
                if (g_particles)
                    for (int i = lo; i < N && s_idx[i] == j; ++i) a += g_particles[(base + i) * d + k];
                d_particles[(base + j) * d + k] = a;
            }
        }
    }
    const float t3 = block_allreduce(third, s_red, OpSum(), 0.f);   // (its barriers also order the zero-fill above)
    if (!hard)
        for (int j = tid; j < N; j += nt) d_probs[base + j] += alpha_f * t3;
}

static int pick_threads(int N) {
    int t = 128;
    while (t < 1024 && t * 4 < N) t <<= 1;
    return t;
}

}  // namespace nfdpf

using namespace nfdpf;

extern "C" int nfdpf_soft_resample_fwd(const float* particles, const float* probs, const float* offsets,
                                       const float* markers, double alpha, int B, int N, int d, float* particles_out,
                                       float* probs_out, int64_t* idx_out, float* saved, float* logprobs_out, const int32_t* gate,
                                       void* stream) {
    NFDPF_REQUIRE(particles && probs && offsets && markers && particles_out && probs_out && idx_out && saved,
                  "soft_resample_fwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && d > 0, "soft_resample_fwd: B, N, d must be positive (got %d, %d, %d)", B, N, d);
    NFDPF_REQUIRE(alpha > 0.0 && alpha <= 1.0, "soft_resample_fwd: need 0 < alpha <= 1 (resamplers.py:21), got %g", alpha);
    const size_t smem = (size_t)N * 2 * sizeof(float);
    // the backward keeps 20 B per particle in shared memory: accept only what it can also handle (no forward-only sizes)
    if ((size_t)((N + 3) & ~3) * sizeof(int) + (size_t)N * sizeof(float4) > 200 * 1024) {
        set_error("soft_resample_fwd: N=%d exceeds the shared-memory row limit (10200)", N); return NFDPF_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024)
        NFDPF_CUDA(cudaFuncSetAttribute(soft_resample_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    soft_resample_fwd_kernel<<<B, pick_threads(N), smem, (cudaStream_t)stream>>>(
        particles, probs, offsets, markers, (float)alpha, (float)(1.0 - alpha), alpha >= 1.0 ? 1 : 0, N, d, particles_out,
        probs_out, idx_out, saved, logprobs_out, gate);
    return check_launch("soft_resample_fwd");
}

extern "C" int nfdpf_soft_resample_bwd(const float* g_particles, const float* g_probs, const float* probs,
                                       const int64_t* idx, const float* saved, double alpha, int B, int N, int d,
                                       float* d_particles, float* d_probs, const float* g_logprobs, const int32_t* gate,
                                       void* stream) {
    NFDPF_REQUIRE(probs && idx && saved && d_particles && d_probs, "soft_resample_bwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && d > 0, "soft_resample_bwd: B, N, d must be positive");
    NFDPF_REQUIRE(alpha > 0.0 && alpha <= 1.0, "soft_resample_bwd: need 0 < alpha <= 1, got %g", alpha);
    const size_t smem = (size_t)((N + 3) & ~3) * sizeof(int) + (size_t)N * sizeof(float4);
    if (smem > 200 * 1024) { set_error("soft_resample_bwd: N=%d exceeds the shared-memory row limit (10200)", N); return NFDPF_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024)
        NFDPF_CUDA(cudaFuncSetAttribute(soft_resample_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    soft_resample_bwd_kernel<<<B, pick_threads(N), smem, (cudaStream_t)stream>>>(
        g_particles, g_probs, probs, idx, saved, (float)alpha, (float)(1.0 - alpha), alpha >= 1.0 ? 1 : 0, N, d, d_particles,
        d_probs, g_logprobs, gate);
    return check_launch("soft_resample_bwd");
}
