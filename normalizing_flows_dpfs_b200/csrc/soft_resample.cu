// (K3) Soft resampling -- one CTA per trajectory: mixture weights, ATen-order row sum, fp64 block scan,
// binary search of the caller's markers, gather, renormalise.  Replaces resamplers/resamplers.py:20-60, whose
// (B,N,N) bool compare (1 GiB at B=N=1024) is never materialised here.  HBM-bound: 32 B / particle forward.
#include <stdlib.h>

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

// =====================================================================================================================
// Fast path (round 2): N % 4 == 0, N <= 4096, d == 2.  The kernels above spend 334 instructions per particle (binary search:
// ~100; run-time-strided loops: addressing, compares and branches around every access) and were ISSUE-bound at 0.28 / 0.21 of
// the HBM roof.  Here a thread owns FOUR CONSECUTIVE particles for the whole kernel: every global access is a 128-bit load /
// store, the mixture weights / scan prefixes / ancestor keys stay in registers, and the marker search is INVERTED:
//   count_i = #{j : cum_j < m_i}  =  #{j : f_j <= i}   with   f_j = min{i : m_i > cum_j}
// (markers m_i = fl(off + mk_i) are non-decreasing).  f_j comes from one multiply (j's position among N equally spaced markers)
// plus an exact fix-up against the caller's marker values; a shared-memory histogram of the f_j and one integer block scan
// give every count -- no per-particle search loop.  Same arithmetic, same bits as the generic kernel and the reference.
constexpr int V4_MAX_N = 4096;

__device__ __forceinline__ float4 ld4f(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4f(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
soft_resample_fwd4_kernel(const float* __restrict__ particles, const float* __restrict__ probs,
                          const float* __restrict__ offsets, const float* __restrict__ markers, float alpha_f,
                          float one_minus_alpha_f, int hard, int N, float* __restrict__ particles_out,
                          float* __restrict__ probs_out, int64_t* __restrict__ idx_out, float* __restrict__ saved,
                          float* __restrict__ logprobs_out, const int* __restrict__ gate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_q = reinterpret_cast<float*>(smem_raw);    // [N] q for the ATen-order row sum; then the histogram of f_j
    int* s_hist = reinterpret_cast<int*>(smem_raw);
    float* s_wis = s_q + N;                             // [N] importance weights w / q
    float* s_mk = s_wis + N;                            // [N] the caller's markers
    __shared__ double s_wd[32];
    __shared__ int s_wi[32];
    __shared__ float s_red[33];
    __shared__ float s_sum;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    const int e0 = 4 * tid;
    const bool act = e0 < N;
    const size_t base = (size_t)b * N + (act ? e0 : 0);
    if (gate && *gate == 0) {   // ESS gate closed (DPFs.py:168-170): pass through, identity ancestors
        if (act) {
            const float4 w = ld4f(probs + base);
            const float4 x0 = ld4f(particles + base * 2), x1 = ld4f(particles + base * 2 + 4);
            *reinterpret_cast<float4*>(particles_out + base * 2) = x0;
            *reinterpret_cast<float4*>(particles_out + base * 2 + 4) = x1;
            *reinterpret_cast<float4*>(probs_out + base) = w;
            longlong2* io = reinterpret_cast<longlong2*>(idx_out + base);
            io[0] = make_longlong2((long long)base, (long long)base + 1);
            io[1] = make_longlong2((long long)base + 2, (long long)base + 3);
            if (logprobs_out) st4f(logprobs_out + base, logf(w.x), logf(w.y), logf(w.z), logf(w.w));
        }
        if (tid == 0) { saved[2 * b] = 1.f; saved[2 * b + 1] = 1.f; }
        return;
    }
    const float unif = __fdiv_rn(1.0f, (float)N);                  // ones/N, resamplers.py:26
    const float uterm = __fmul_rn(unif, one_minus_alpha_f);        // uniform_probs*(1-alpha), :31
    float w[4] = {0.f, 0.f, 0.f, 0.f}, q[4], mk[4] = {0.f, 0.f, 0.f, 0.f};
    if (act) {
        const float4 w4 = ld4f(probs + base), m4 = ld4f(markers + e0);
        w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
        mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = hard ? w[u] : __fadd_rn(__fmul_rn(w[u], alpha_f), uterm);
    if (act) { st4f(s_q + e0, q[0], q[1], q[2], q[3]); st4f(s_mk + e0, mk[0], mk[1], mk[2], mk[3]); }
    __syncthreads();
    float S = 1.0f, wis[4];
    if (!hard) {
        if (tid < 32) {
            const float s = aten_row_sum_warp(s_q, N);
            if (tid == 0) s_sum = s;
        }
        __syncthreads();
        S = s_sum;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            q[u] = __fdiv_rn(q[u], S);                             // q / q.sum, :33
            wis[u] = __fdiv_rn(w[u], q[u]);                        // w / q, :34
        }
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) wis[u] = unif;                 // hard resampling, :36-38
        __syncthreads();                                           // (s_q is about to become the histogram)
    }
    if (act) {
        st4f(s_wis + e0, wis[0], wis[1], wis[2], wis[3]);
        *reinterpret_cast<int4*>(s_hist + e0) = make_int4(0, 0, 0, 0);
    }
    // cumsum (:45): fp64 running sums, every prefix rounded to fp32 (exact in fp64, see scan_fp64_inplace)
    float cum[4];
    {
        double p[4];
        p[0] = act ? (double)q[0] : 0.0;
#pragma unroll
        for (int u = 1; u < 4; ++u) p[u] = p[u - 1] + (act ? (double)q[u] : 0.0);
        double inc = p[3];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double v = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_wd[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const double t = lane < nwarp ? s_wd[lane] : 0.0;
            double wi = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(FULL, wi, o);
                if (lane >= o) wi += v;
            }
            if (lane < nwarp) s_wd[lane] = wi - t;
        }
        __syncthreads();
        const double ex = s_wd[warp] + (inc - p[3]);
#pragma unroll
        for (int u = 0; u < 4; ++u) cum[u] = (float)(ex + p[u]);
    }
    // f_j = number of markers m_i <= cum_j, for the sorted prefix j < N - 1 (cum[N-1] = 1 is handled below, :47)
    const float off = offsets[b], Nf = (float)N;
    if (act) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (e0 + u >= N - 1) continue;
            const float c = cum[u];
            int g = (int)floorf((c - off) * Nf) + 1;
            g = max(0, min(g, N));
            while (g < N && !(__fadd_rn(off, s_mk[g]) > c)) ++g;           // m_g <= c: the first marker above c lies further right
            while (g > 0 && __fadd_rn(off, s_mk[g - 1]) > c) --g;
            if (g < N) atomicAdd(&s_hist[g], 1);
        }
    }
    __syncthreads();
    int cnt[4];
    {
        int4 h = act ? *reinterpret_cast<const int4*>(s_hist + e0) : make_int4(0, 0, 0, 0);
        cnt[0] = h.x; cnt[1] = cnt[0] + h.y; cnt[2] = cnt[1] + h.z; cnt[3] = cnt[2] + h.w;
        int inc = cnt[3];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_wi[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            const int t = lane < nwarp ? s_wi[lane] : 0;
            int wi = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, wi, o);
                if (lane >= o) wi += v;
            }
            if (lane < nwarp) s_wi[lane] = wi - t;
        }
        __syncthreads();
        const int ex = s_wi[warp] + (inc - cnt[3]);
#pragma unroll
        for (int u = 0; u < 4; ++u) cnt[u] += ex;
    }
    float v[4] = {0.f, 0.f, 0.f, 0.f}, part = 0.f;
    float2 x[4];
    int jj[4];
    if (act) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float m = __fadd_rn(off, mk[u]);                         // :44
            jj[u] = min(cnt[u] + (m > 1.0f ? 1 : 0), N - 1);               // + the forced last entry cum[N-1] = 1 ('>' is strict, :49)
            x[u] = __ldg(reinterpret_cast<const float2*>(particles) + (size_t)b * N + jj[u]);
            v[u] = s_wis[jj[u]];
        }
        part = (v[0] + v[1]) + (v[2] + v[3]);
    }
    const float S2 = block_allreduce(part, s_red, OpSum(), 0.f);
    if (act) {
        float pn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) pn[u] = __fdiv_rn(v[u], S2);           // :56
        *reinterpret_cast<float4*>(particles_out + base * 2) = make_float4(x[0].x, x[0].y, x[1].x, x[1].y);
        *reinterpret_cast<float4*>(particles_out + base * 2 + 4) = make_float4(x[2].x, x[2].y, x[3].x, x[3].y);
        st4f(probs_out + base, pn[0], pn[1], pn[2], pn[3]);
        const long long rb = (long long)N * b;
        longlong2* io = reinterpret_cast<longlong2*>(idx_out + base);
        io[0] = make_longlong2(rb + jj[0], rb + jj[1]);                    // :52
        io[1] = make_longlong2(rb + jj[2], rb + jj[3]);
        if (logprobs_out) st4f(logprobs_out + base, logf(pn[0]), logf(pn[1]), logf(pn[2]), logf(pn[3]));   // DPFs.py:167
    }
    if (tid == 0) { saved[2 * b] = S; saved[2 * b + 1] = S2; }
}

// Backward, same ownership: the sorted ancestor keys of a thread's four destinations and their (dL/dv, g_x, g_y) triples stay in
// registers; runs of equal keys are summed by a segmented scan (registers -> warp shuffles -> one cross-warp step, fixed order);
// a run's total lands in its SOURCE's shared-memory slot, and every thread then finishes the four sources it owns: all global
// traffic is 128-bit and coalesced, nothing is zero-filled and rewritten.
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
soft_resample_bwd4_kernel(const float* __restrict__ g_particles, const float* __restrict__ g_probs,
                          const float* __restrict__ probs, const int64_t* __restrict__ idx,
                          const float* __restrict__ saved, float alpha_f, float one_minus_alpha_f, int hard, int N,
                          float* __restrict__ d_particles, float* __restrict__ d_probs, const float* __restrict__ g_logprobs,
                          const int* __restrict__ gate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_w = reinterpret_cast<float*>(smem_raw);    // [N] the row's weights
    float* s_G = s_w + N;                               // [N] per-source dL/dv run totals
    float2* s_dp = reinterpret_cast<float2*>(s_G + N);  // [N] per-source particle-gradient run totals
    __shared__ float s_red[33];
    __shared__ float s_wt[96];
    __shared__ int s_wi[64];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int e0 = 4 * tid;
    const bool act = e0 < N;
    const size_t base = (size_t)b * N + (act ? e0 : 0);
    if (gate && *gate == 0) {   // the forward passed everything through: so do the gradients (+ d log w = g / w)
        if (act) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a, g = a;
            if (g_particles) { a = ld4f(g_particles + base * 2); c = ld4f(g_particles + base * 2 + 4); }
            if (g_probs) g = ld4f(g_probs + base);
            if (g_logprobs) {
                const float4 gl = ld4f(g_logprobs + base), w = ld4f(probs + base);
                g.x += gl.x / w.x; g.y += gl.y / w.y; g.z += gl.z / w.z; g.w += gl.w / w.w;
            }
            *reinterpret_cast<float4*>(d_particles + base * 2) = a;
            *reinterpret_cast<float4*>(d_particles + base * 2 + 4) = c;
            *reinterpret_cast<float4*>(d_probs + base) = g;
        }
        return;
    }
    const float S = saved[2 * b], S2 = saved[2 * b + 1];
    const float unif = __fdiv_rn(1.0f, (float)N), uterm = __fmul_rn(unif, one_minus_alpha_f);
    float wo[4] = {0.f, 0.f, 0.f, 0.f};                 // the weights of the four SOURCES this thread owns
    int key[4] = {-2, -2, -2, -3};                      // inactive threads never match anything
    float ax[4] = {0.f, 0.f, 0.f, 0.f}, gx[4] = {0.f, 0.f, 0.f, 0.f}, gy[4] = {0.f, 0.f, 0.f, 0.f};
    if (act) {
        const float4 w4 = ld4f(probs + base);
        wo[0] = w4.x; wo[1] = w4.y; wo[2] = w4.z; wo[3] = w4.w;
        *reinterpret_cast<float4*>(s_w + e0) = w4;
        *reinterpret_cast<float4*>(s_G + e0) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(s_dp + e0) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(s_dp + e0 + 2) = make_float4(0.f, 0.f, 0.f, 0.f);
        const longlong2* ip = reinterpret_cast<const longlong2*>(idx + base);
        const longlong2 i0 = ip[0], i1 = ip[1];
        const long long rb = (long long)N * b;
        key[0] = (int)(i0.x - rb); key[1] = (int)(i0.y - rb); key[2] = (int)(i1.x - rb); key[3] = (int)(i1.y - rb);
        if (g_particles) {
            const float4 a = ld4f(g_particles + base * 2), c = ld4f(g_particles + base * 2 + 4);
            gx[0] = a.x; gy[0] = a.y; gx[1] = a.z; gy[1] = a.w; gx[2] = c.x; gy[2] = c.y; gx[3] = c.z; gy[3] = c.w;
        }
        if (!hard && g_probs) { const float4 g = ld4f(g_probs + base); ax[0] = g.x; ax[1] = g.y; ax[2] = g.z; ax[3] = g.w; }
    }
    __syncthreads();
    float part = 0.f;
    if (act && !hard && (g_probs || g_logprobs)) {
        float gl[4] = {0.f, 0.f, 0.f, 0.f};
        if (g_logprobs) { const float4 g = ld4f(g_logprobs + base); gl[0] = g.x; gl[1] = g.y; gl[2] = g.z; gl[3] = g.w; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float wj = s_w[min(max(key[u], 0), N - 1)];
            const float wp = (wj * S / __fadd_rn(__fmul_rn(wj, alpha_f), uterm)) / S2;
            if (g_logprobs) ax[u] += gl[u] / wp;     // total gradient reaching w'_i: direct + through log(w'_i)
            part += ax[u] * wp;                      // sum_m g_m w'_m
        }
    }
    const float c = block_allreduce(part, s_red, OpSum(), 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) ax[u] = hard ? 0.f : (ax[u] - c) / S2;     // dL/dv_i, v_i = w_is[idx_i]
    // ---- segmented inclusive scan over the runs of equal keys: (1) inside the thread
#pragma unroll
    for (int u = 1; u < 4; ++u)
        if (key[u] == key[u - 1]) { ax[u] += ax[u - 1]; gx[u] += gx[u - 1]; gy[u] += gy[u - 1]; }
    const int kfirst = key[0], klast = key[3];
    float tx = ax[3], ty = gx[3], tz = gy[3];
    // (2) thread tails across the warp and the CTA: a tail continues the previous thread's tail iff the thread holds one key equal to it
    if (lane == 31) s_wi[32 + wid] = klast;
    __syncthreads();
    int kleft = __shfl_up_sync(FULL, klast, 1);
    if (lane == 0) kleft = wid > 0 ? s_wi[32 + wid - 1] : -4;
    const bool cont = act && kfirst == kleft;                     // the leading run continues the previous thread's tail
    int start = !(cont && kfirst == klast);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ux = __shfl_up_sync(FULL, tx, o), uy = __shfl_up_sync(FULL, ty, o), uz = __shfl_up_sync(FULL, tz, o);
        const int us = __shfl_up_sync(FULL, start, o);
        if (lane >= o && !start) { tx += ux; ty += uy; tz += uz; start = us; }
    }
    if (lane == 31) { s_wt[wid] = tx; s_wt[32 + wid] = ty; s_wt[64 + wid] = tz; s_wi[wid] = start; }
    __syncthreads();
    if (!start && wid > 0) {                                      // the segment reaches back beyond this warp
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int w = wid - 1; w >= 0; --w) {
            cx += s_wt[w]; cy += s_wt[32 + w]; cz += s_wt[64 + w];
            if (s_wi[w]) break;
        }
        tx += cx; ty += cy; tz += cz;
    }
    // (tx, ty, tz) = total of the run that ends with this thread's tail; (3) hand it to the next thread's leading run
    float px = __shfl_up_sync(FULL, tx, 1), py = __shfl_up_sync(FULL, ty, 1), pz = __shfl_up_sync(FULL, tz, 1);
    int knext = __shfl_down_sync(FULL, kfirst, 1);
    __syncthreads();                                              // s_wt / s_wi reuse
    if (lane == 31) { s_wt[wid] = tx; s_wt[32 + wid] = ty; s_wt[64 + wid] = tz; }
    if (lane == 0) s_wi[wid] = kfirst;
    __syncthreads();
    if (lane == 0 && wid > 0) { px = s_wt[wid - 1]; py = s_wt[32 + wid - 1]; pz = s_wt[64 + wid - 1]; }
    if (lane == 31) knext = (wid + 1) < (int)(blockDim.x >> 5) ? s_wi[wid + 1] : -5;
    if (cont) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (key[u] == kfirst) { ax[u] += px; gx[u] += py; gy[u] += pz; }
    }
    // a run ends where the next key differs: its total goes to its source's slot
    if (act) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int nk = u < 3 ? key[u + 1] : knext;
            if (nk != key[u]) {
                const int j = min(max(key[u], 0), N - 1);
                s_G[j] = ax[u];
                s_dp[j] = make_float2(gx[u], gy[u]);
            }
        }
    }
    __syncthreads();
    float third = 0.f, dq[4] = {0.f, 0.f, 0.f, 0.f};
    float4 da = make_float4(0.f, 0.f, 0.f, 0.f), dc = da;
    if (act) {
        const float4 G4 = *reinterpret_cast<const float4*>(s_G + e0);
        da = *reinterpret_cast<const float4*>(s_dp + e0);
        dc = *reinterpret_cast<const float4*>(s_dp + e0 + 2);
        const float G[4] = {G4.x, G4.y, G4.z, G4.w};
        if (!hard) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float wj = wo[u], q = __fadd_rn(__fmul_rn(wj, alpha_f), uterm);
                dq[u] = G[u] * S / q - G[u] * wj * S * alpha_f / (q * q);   // through w_is_j = w_j S / qu_j at fixed S
                third += G[u] * wj / q;                                      // through S = sum_k qu_k
            }
        }
    }
    const float t3 = block_allreduce(third, s_red, OpSum(), 0.f);
    if (act) {
        *reinterpret_cast<float4*>(d_particles + base * 2) = da;
        *reinterpret_cast<float4*>(d_particles + base * 2 + 4) = dc;
        const float a3 = hard ? 0.f : alpha_f * t3;
        st4f(d_probs + base, dq[0] + a3, dq[1] + a3, dq[2] + a3, dq[3] + a3);
    }
}

static int v4_threads(int N) { return ((N / 4) + 31) & ~31; }
static bool v4_ok(int N, int d) {
    static const bool generic = getenv("NFDPF_SOFT_GENERIC") != nullptr;    // A/B timing of the two paths (tools/time_soft.py)
    return !generic && d == 2 && N % 4 == 0 && N >= 4 && N <= V4_MAX_N;
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
    if (v4_ok(N, d)) {
        const size_t sm4 = (size_t)N * 3 * sizeof(float);
        if (N <= 1024) {     // <= 256 threads: seven rows resident per SM, so B = 1024 rows are ONE wave on 148 SMs
            soft_resample_fwd4_kernel<256, 7><<<B, v4_threads(N), sm4, (cudaStream_t)stream>>>(
                particles, probs, offsets, markers, (float)alpha, (float)(1.0 - alpha), alpha >= 1.0 ? 1 : 0, N, particles_out, probs_out,
                idx_out, saved, logprobs_out, gate);
        } else {
            if (sm4 > 40 * 1024)
                NFDPF_CUDA(cudaFuncSetAttribute(soft_resample_fwd4_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4));
            soft_resample_fwd4_kernel<1024, 1><<<B, v4_threads(N), sm4, (cudaStream_t)stream>>>(
                particles, probs, offsets, markers, (float)alpha, (float)(1.0 - alpha), alpha >= 1.0 ? 1 : 0, N, particles_out, probs_out,
                idx_out, saved, logprobs_out, gate);
        }
        return check_launch("soft_resample_fwd");
    }
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
    if (v4_ok(N, d)) {
        const size_t sm4 = (size_t)N * 4 * sizeof(float);
        if (N <= 1024) {     // 48 registers, no spills (a 32-register build spilled and ran 8 % slower)
            soft_resample_bwd4_kernel<256, 5><<<B, v4_threads(N), sm4, (cudaStream_t)stream>>>(
                g_particles, g_probs, probs, idx, saved, (float)alpha, (float)(1.0 - alpha), alpha >= 1.0 ? 1 : 0, N, d_particles, d_probs,
                g_logprobs, gate);
        } else {
            if (sm4 > 40 * 1024)
                NFDPF_CUDA(cudaFuncSetAttribute(soft_resample_bwd4_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm4));
            soft_resample_bwd4_kernel<1024, 1><<<B, v4_threads(N), sm4, (cudaStream_t)stream>>>(
                g_particles, g_probs, probs, idx, saved, (float)alpha, (float)(1.0 - alpha), alpha >= 1.0 ? 1 : 0, N, d_particles, d_probs,
                g_logprobs, gate);
        }
        return check_launch("soft_resample_bwd");
    }
    const size_t smem = (size_t)((N + 3) & ~3) * sizeof(int) + (size_t)N * sizeof(float4);
    if (smem > 200 * 1024) { set_error("soft_resample_bwd: N=%d exceeds the shared-memory row limit (10200)", N); return NFDPF_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024)
        NFDPF_CUDA(cudaFuncSetAttribute(soft_resample_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    soft_resample_bwd_kernel<<<B, pick_threads(N), smem, (cudaStream_t)stream>>>(
        g_particles, g_probs, probs, idx, saved, (float)alpha, (float)(1.0 - alpha), alpha >= 1.0 ? 1 : 0, N, d, d_particles,
        d_probs, g_logprobs, gate);
    return check_launch("soft_resample_bwd");
}
