// (K4) Entropy-regularised OT resampling: log-domain Sinkhorn with epsilon-scaling, then particles' = T x.
// Replaces resamplers/resamplers.py:62-277 (transport_function -> sinkhorn_potentials -> sinkhorn_loop ->
// transport_from_potentials -> apply_transport_matrix), which materialises four (B,N,N) fp64 cost matrices plus
// dozens of (B,N,N) temporaries per iteration (8.6 GB each at B=N=1024).  Here the N x N cost tile is recomputed
// from shared-memory-staged particles in every pass; HBM sees only O(B N) vectors.  SFU (ex2) / FMA bound.
//
// x == y in the reference's call (resamplers.py:223), so there is ONE symmetric cost matrix and only the two live
// potential chains a_y / b_x are evaluated (the a_x / b_y chains never reach the output).
//
// Stop rule (resamplers.py:126-129, 155-161): the loop runs while EVERY row still wants to continue, i.e. it stops
// at K = min_b k_b.  Each iteration is one launch; the last CTA to finish an iteration (ticket counter) evaluates
// the batch-wide rule on the device and arms / disarms the next launch -- no host synchronisation.
#include "common.cuh"

namespace nfdpf {

constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
constexpr int OT_T = 256;      // threads per CTA = rows i per CTA
constexpr int OT_J = 1024;     // columns j staged per shared-memory tile
constexpr int OT_CH = 16;      // columns per online-LSE rescale (one extra ex2 per chain per 16 pair evaluations)

struct OtCtrl { int iter; int go; unsigned ticket; int pad; };

__device__ __forceinline__ float ex2f(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2f(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// ---- prepare: centre / scale the cloud, epsilon_0 (resamplers.py:72-76, 87-91, 117, 218-222) ----------------
__global__ void __launch_bounds__(256)
ot_prepare_kernel(const float* __restrict__ x, int N, float2* __restrict__ sx, float* __restrict__ eps_run,
                  unsigned* __restrict__ diff, OtCtrl* __restrict__ ctrl, int max_iter, const int* __restrict__ gate) {
    __shared__ float s_red[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (gate && *gate == 0) {            // device-side ESS gate closed: disarm the whole Sinkhorn chain
        if (b == 0 && tid == 0) { ctrl->iter = 0; ctrl->go = 0; ctrl->ticket = 0u; ctrl->pad = 1; }
        return;
    }
    const float2* xr = reinterpret_cast<const float2*>(x) + (size_t)b * N;
    float sx0 = 0.f, sy0 = 0.f;
    for (int n = tid; n < N; n += 256) { const float2 v = xr[n]; sx0 += v.x; sy0 += v.y; }
    const float mx = block_allreduce(sx0, s_red, OpSum(), 0.f) / (float)N;
    const float my = block_allreduce(sy0, s_red, OpSum(), 0.f) / (float)N;
    float vx = 0.f, vy = 0.f;
    for (int n = tid; n < N; n += 256) { const float2 v = xr[n]; vx = fmaf(v.x - mx, v.x - mx, vx); vy = fmaf(v.y - my, v.y - my, vy); }
    vx = block_allreduce(vx, s_red, OpSum(), 0.f) / (float)N;   // population variance (unbiased=False)
    vy = block_allreduce(vy, s_red, OpSum(), 0.f) / (float)N;
    float diam = sqrtf(fmaxf(vx, vy));
    if (diam == 0.f) diam = 1.f;
    const float scale = diam * 1.41421356237309515f;            // * sqrt(d), d = 2
    float hi = -INFINITY, lo = INFINITY;
    for (int n = tid; n < N; n += 256) {
        const float2 v = xr[n];
        const float2 s = make_float2((v.x - mx) / scale, (v.y - my) / scale);
        sx[(size_t)b * N + n] = s;
        hi = fmaxf(hi, fmaxf(s.x, s.y));
        lo = fminf(lo, fminf(s.x, s.y));
    }
    hi = block_allreduce(hi, s_red, OpMax(), -INFINITY);
    lo = block_allreduce(lo, s_red, OpMin(), INFINITY);
    // max_min(x, x) = max_{n,d} x - min(min_d max_n x, min_{n,d} x) = max - min (the second argument always wins)
    if (tid == 0) {
        eps_run[b] = (hi - lo) * (hi - lo);
        diff[b] = 0u;
        if (b == 0) { ctrl->iter = 0; ctrl->go = max_iter - 1 > 0 ? 1 : 0; ctrl->ticket = 0u; ctrl->pad = 0; }
    }
}

// ---- one dual-chain softmin pass -----------------------------------------------------------------------------
// a'[i] = -eps LSE_j( logw_j    [+ b[j]/eps] - C_ij/eps ),  b'[i] = -eps LSE_j( -log N [+ a[j]/eps] - C_ij/eps )
// MODE 0 init (eps = eps_0[b], bracketed terms absent), 1 loop iteration (eps = eps_run[b], averaged with the old
// potentials, row-wise max |delta| recorded), 2 final (eps = target, no averaging).       (resamplers.py:94-178)
//
// Round 2: ONE exponential per (i, j) pair serves both chains.  With z1_j = (logw_j + b_j/eps) log2 e, z2_j likewise,
//   sum_j 2^(z1_j - c |xi - xj|^2) = 2^R1 sum_j K_ij E1_j,   K_ij = 2^(-c |xi - xj|^2),  E1_j = 2^(z1_j - R1),
// K_ij is shared by the two chains and the column factors E1_j, E2_j are formed once per COLUMN while the tile is staged
// (R = the running maximum of z over the columns seen so far, CTA-uniform; the sums are rescaled when a new tile raises it).
// Per pair: 2 FADD + FMUL + FFMA (the scaled squared distance, negated), one MUFU.EX2, one FFMA2 (both sums) -- against two
// exponentials and ~12 FP32 instructions of the per-chain online LSE it replaces.  A thread owns OT_R rows, so a staged column
// (one broadcast LDS.128 = 4 cycles of the SM's shared-memory return path) serves OT_R pairs: without the row blocking the
// LDS, not the MUFU pipe, would bound the loop.  Range: the j = i column always contributes K = 1, so a row's sum can only
// underflow if z_i lies ~100 below the tile maximum in log2 units (69 nats; the filter's weights span 28 nats, the potentials
// ~cost/eps: clouds with far outliers at the final eps do get there); such a row is recomputed with the per-row online LSE
// (ot_safe_row, two exponentials per pair straight from global memory: exact but ~10x slower for that row).
constexpr int OT_R = 4;        // rows per thread in the pass kernel
constexpr float OT_TINY = 1e-30f;   // ~2^-100: every term within 2^-24 of such a sum is still a normal fp32 number

__device__ __forceinline__ unsigned long long ot_pack2(float lo, float hi) {
    unsigned long long p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi));
    return p;
}

// per-row online log-sum-exp of both chains straight from global memory (2 exponentials per pair): the safe fallback
template <int MODE>
__device__ void ot_safe_row(const float2* __restrict__ sx, const float* __restrict__ logw, const float* __restrict__ a_old,
                            const float* __restrict__ b_old, size_t row, int N, int i, float inv, float c2, float log_beta,
                            float& lse1, float& lse2) {
    const float2 xi = sx[row + i];
    float m1 = -INFINITY, s1 = 0.f, m2 = -INFINITY, s2 = 0.f;
    for (int j = 0; j < N; ++j) {
        const float2 p = sx[row + j];
        float h1 = logw[row + j], h2 = log_beta;
        if (MODE != 0) { h1 = fmaf(b_old[j], inv, h1); h2 = fmaf(a_old[j], inv, h2); }
        const float dx = xi.x - p.x, dy = xi.y - p.y, d2 = fmaf(dy, dy, dx * dx);
        const float v1 = fmaf(-c2, d2, h1 * LOG2E), v2 = fmaf(-c2, d2, h2 * LOG2E);
        const float n1 = fmaxf(m1, v1), n2 = fmaxf(m2, v2);
        s1 = s1 * ex2f(m1 - n1) + ex2f(v1 - n1);
        s2 = s2 * ex2f(m2 - n2) + ex2f(v2 - n2);
        m1 = n1; m2 = n2;
    }
    lse1 = m1 + lg2f(s1); lse2 = m2 + lg2f(s2);
}

template <int MODE>
__global__ void __launch_bounds__(OT_T)
ot_pass_kernel(const float2* __restrict__ sx, const float* __restrict__ logw, const float* __restrict__ pot_a,
               const float* __restrict__ pot_b, float* __restrict__ out_a, float* __restrict__ out_b, int N,
               const float* eps_run_in, float* eps_run, unsigned* diff,
               OtCtrl* __restrict__ ctrl, float eps_target, float scaling2, float threshold, int max_iter, int B) {
    __shared__ float4 s_j[OT_J];
    __shared__ float s_red[33];
    __shared__ int s_last;
    const int b = blockIdx.y, tid = threadIdx.x, i0 = blockIdx.x * OT_T * OT_R + tid;
    int cur = 0;
    if (ctrl->pad) return;               // resampling gated off for this step (ot_prepare_kernel)
    if (MODE == 1) {
        if (!ctrl->go) return;           // the batch-wide stop rule fired in an earlier launch
        cur = ctrl->iter & 1;
    } else if (MODE == 2) {
        cur = ctrl->iter & 1;
    }
    const size_t row = (size_t)b * N, pstride = (size_t)B * N;
    const float* a_old = pot_a + cur * pstride + row;
    const float* b_old = pot_b + cur * pstride + row;
    const float eps = MODE == 2 ? eps_target : eps_run_in[b];
    const float inv = 1.0f / eps;
    const float c2 = 0.5f * LOG2E * inv;             // C_ij / eps in log2 units = c2 * |xi - xj|^2
    const float sc = sqrtf(c2);                      // coordinates are pre-multiplied: |sc xi - sc xj|^2 = c2 |xi - xj|^2
    const float log_beta = -logf((float)N);
    float xs[OT_R], ys[OT_R];
    unsigned long long acc[OT_R];                    // (s1, s2) as a packed pair: one FFMA2 per (row, column)
#pragma unroll
    for (int r = 0; r < OT_R; ++r) {
        const int i = i0 + r * OT_T;
        const float2 xi = sx[row + (i < N ? i : 0)];
        xs[r] = xi.x * sc; ys[r] = xi.y * sc;
        acc[r] = 0ull;
    }
    float R1 = -INFINITY, R2 = -INFINITY;            // running shifts (CTA-uniform)
    for (int j0 = 0; j0 < N; j0 += OT_J) {
        const int jn = min(OT_J, N - j0);
        // column exponents of this tile and their maxima
        float z1[OT_J / OT_T], z2[OT_J / OT_T], t1 = -INFINITY, t2 = -INFINITY;
#pragma unroll
        for (int u = 0; u < OT_J / OT_T; ++u) {
            const int j = tid + u * OT_T;
            z1[u] = z2[u] = -INFINITY;
            if (j < jn) {
                float h1 = logw[row + j0 + j], h2 = log_beta;
                if (MODE != 0) { h1 = fmaf(b_old[j0 + j], inv, h1); h2 = fmaf(a_old[j0 + j], inv, h2); }
                z1[u] = h1 * LOG2E; z2[u] = h2 * LOG2E;
            }
            t1 = fmaxf(t1, z1[u]); t2 = fmaxf(t2, z2[u]);
        }
        t1 = block_allreduce(t1, s_red, OpMax(), -INFINITY);      // (its barriers also protect s_j from the previous tile's readers)
        t2 = block_allreduce(t2, s_red, OpMax(), -INFINITY);
        const float n1 = fmaxf(R1, t1), n2 = fmaxf(R2, t2);        // finite: log-weights are finite (the filter adds 1e-12, DPFs.py:192)
        if (j0 > 0) {
            const unsigned long long f = ot_pack2(ex2f(R1 - n1), ex2f(R2 - n2));
#pragma unroll
            for (int r = 0; r < OT_R; ++r) asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(acc[r]) : "l"(f));
        }
        R1 = n1; R2 = n2;
#pragma unroll
        for (int u = 0; u < OT_J / OT_T; ++u) {
            const int j = tid + u * OT_T;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);                  // padding columns: E = 0
            if (j < jn) {
                const float2 p = sx[row + j0 + j];
                v = make_float4(p.x * sc, p.y * sc, ex2f(z1[u] - R1), ex2f(z2[u] - R2));
            }
            s_j[j] = v;
        }
        __syncthreads();
        const int jend = (jn + 3) & ~3;
#pragma unroll 4
        for (int j = 0; j < jend; ++j) {
            const float4 q = s_j[j];
            const unsigned long long e12 = ot_pack2(q.z, q.w);
#pragma unroll
            for (int r = 0; r < OT_R; ++r) {
                const float dx = xs[r] - q.x, dy = ys[r] - q.y;
                const float k = ex2f(fmaf(-dy, dy, -(dx * dx)));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[r]) : "l"(ot_pack2(k, k)), "l"(e12));
            }
        }
    }
    float d = 0.f;
#pragma unroll
    for (int r = 0; r < OT_R; ++r) {
        const int i = i0 + r * OT_T;
        const bool live = i < N;
        float s1, s2;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(s1), "=f"(s2) : "l"(acc[r]));
        float lse1 = R1 + lg2f(s1), lse2 = R2 + lg2f(s2);
        // A sum below OT_TINY has lost its leading terms to underflow (and lg2.approx.ftz maps a denormal sum to -inf): safe path
        if (live && !(s1 > OT_TINY && s2 > OT_TINY && s1 < INFINITY && s2 < INFINITY))
            ot_safe_row<MODE>(sx, logw, a_old, b_old, row, N, i, inv, c2, log_beta, lse1, lse2);
        const float sm1 = -eps * LN2 * lse1, sm2 = -eps * LN2 * lse2;               // softmin = -eps * LSE
        if (MODE != 1) {
            if (live) { out_a[row + i] = sm1; out_b[row + i] = sm2; }
        } else if (live) {
            const float ao = a_old[i], bo = b_old[i];
            const float an = 0.5f * (ao + sm1), bn = 0.5f * (bo + sm2);    // resamplers.py:147-148
            out_a[(cur ^ 1) * pstride + row + i] = an;
            out_b[(cur ^ 1) * pstride + row + i] = bn;
            d = fmaxf(d, fmaxf(fabsf(an - ao), fabsf(bn - bo)));
        }
    }
    if (MODE != 1) return;
    d = block_allreduce(d, s_red, OpMax(), 0.f);
    if (tid == 0) {
        atomicMax(diff + b, __float_as_uint(d));
        __threadfence();
        const unsigned t = atomicAdd(&ctrl->ticket, 1u);
        s_last = (t == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (!s_last) return;
    // last CTA of this iteration: batch-wide continue rule (resamplers.py:155-161, 126-129)
    __threadfence();
    int all_cont = 1;
    for (int r = tid; r < B; r += OT_T) {
        const float e_old = eps_run[r];
        const float e_new = fmaxf(e_old * scaling2, eps_target);
        const float dr = __uint_as_float(*reinterpret_cast<volatile unsigned*>(diff + r));
        const int cont = (e_new < e_old) || (dr > threshold);
        eps_run[r] = e_new;
        diff[r] = 0u;
        all_cont &= cont;
    }
    all_cont = __syncthreads_and(all_cont);
    if (tid == 0) {
        const int it = ctrl->iter + 1;
        ctrl->iter = it;
        ctrl->go = (it < max_iter - 1) && all_cont;
        ctrl->ticket = 0u;
        __threadfence();
    }
}

// ---- column normaliser of the transport plan (resamplers.py:199-207) -----------------------------------------
// U_j = -log2 sum_i 2^(f_i/eps - C_ij/eps) + log2(N w_j): same one-exponential-per-pair form and OT_R-row blocking as the
// softmin passes (sum_i K_ij E_i with E_i = 2^(f_i/eps log2 e - R), R the running maximum over the tiles).
__global__ void __launch_bounds__(OT_T)
ot_colnorm_kernel(const float2* __restrict__ sx, const float* __restrict__ logw, const float* __restrict__ f,
                  const float* __restrict__ g, int N, float eps_target, float4* __restrict__ saved, const int* __restrict__ gate) {
    __shared__ float4 s_j[OT_J];
    __shared__ float s_red[33];
    if (gate && *gate == 0) return;
    const int b = blockIdx.y, tid = threadIdx.x, c0 = blockIdx.x * OT_T * OT_R + tid;
    const size_t row = (size_t)b * N;
    const float inv = 1.0f / eps_target, c2 = 0.5f * LOG2E * inv, sc = sqrtf(c2);
    float xs[OT_R], ys[OT_R], acc[OT_R];
#pragma unroll
    for (int r = 0; r < OT_R; ++r) {
        const int j = c0 + r * OT_T;
        const float2 xj = sx[row + (j < N ? j : 0)];
        xs[r] = xj.x * sc; ys[r] = xj.y * sc; acc[r] = 0.f;
    }
    float R = -INFINITY;
    for (int i0 = 0; i0 < N; i0 += OT_J) {
        const int in = min(OT_J, N - i0);
        float z[OT_J / OT_T], t = -INFINITY;
#pragma unroll
        for (int u = 0; u < OT_J / OT_T; ++u) {
            const int i = tid + u * OT_T;
            z[u] = i < in ? f[row + i0 + i] * inv * LOG2E : -INFINITY;
            t = fmaxf(t, z[u]);
        }
        t = block_allreduce(t, s_red, OpMax(), -INFINITY);
        const float n = fmaxf(R, t);
        if (i0 > 0) {
            const float fr = ex2f(R - n);
#pragma unroll
            for (int r = 0; r < OT_R; ++r) acc[r] *= fr;
        }
        R = n;
#pragma unroll
        for (int u = 0; u < OT_J / OT_T; ++u) {
            const int i = tid + u * OT_T;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < in) { const float2 p = sx[row + i0 + i]; v = make_float4(p.x * sc, p.y * sc, ex2f(z[u] - R), 0.f); }
            s_j[i] = v;
        }
        __syncthreads();
        const int iend = (in + 3) & ~3;
#pragma unroll 4
        for (int i = 0; i < iend; ++i) {
            const float4 q = s_j[i];
#pragma unroll
            for (int r = 0; r < OT_R; ++r) {
                const float dx = xs[r] - q.x, dy = ys[r] - q.y;
                acc[r] = fmaf(ex2f(fmaf(-dy, dy, -(dx * dx))), q.z, acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < OT_R; ++r) {
        const int j = c0 + r * OT_T;
        if (j >= N) continue;
        float lse2 = R + lg2f(acc[r]);   // log2-domain LSE_i(f_i/eps - C_ij/eps); g_j/eps cancels inside U
        const float2 xj = sx[row + j];
        if (!(acc[r] > OT_TINY && acc[r] < INFINITY)) {      // sum (nearly) underflowed: per-column online LSE
            float m = -INFINITY, sacc = 0.f;
            for (int i = 0; i < N; ++i) {
                const float2 p = sx[row + i];
                const float dx = xj.x - p.x, dy = xj.y - p.y;
                const float v = fmaf(-c2, fmaf(dy, dy, dx * dx), f[row + i] * inv * LOG2E), nn = fmaxf(m, v);
                sacc = sacc * ex2f(m - nn) + ex2f(v - nn);
                m = nn;
            }
            lse2 = m + lg2f(sacc);
        }
        const float U = -lse2 + (logf((float)N) + logw[row + j]) * LOG2E;
        (void)g;
        saved[row + j] = make_float4(xj.x, xj.y, f[row + j] * inv * LOG2E, U);
    }
}

// ---- apply the plan: out_i = sum_j T_ij v_j (forward: v = particles) or out_j = sum_i T_ij v_i (backward: v = grad) ----
// T_ij = 2^(F_i + U_j - c |xi - xj|^2).  OT_R rows per thread share every staged column (LDS.128 + LDS.64 per column).
template <bool TRANSPOSED>
__global__ void __launch_bounds__(OT_T)
ot_apply_kernel(const float4* __restrict__ saved, const float* __restrict__ v, int N, float eps_target, float* __restrict__ out,
                const int* __restrict__ gate) {
    __shared__ float4 s_p[OT_J];
    __shared__ float2 s_v[OT_J];
    const int b = blockIdx.y, tid = threadIdx.x, m0 = blockIdx.x * OT_T * OT_R + tid;
    const size_t row = (size_t)b * N;
    if (gate && *gate == 0) {            // resampling gated off: the plan is the identity (forward and transposed)
#pragma unroll
        for (int r = 0; r < OT_R; ++r) {
            const int me = m0 + r * OT_T;
            if (me < N) reinterpret_cast<float2*>(out)[row + me] = reinterpret_cast<const float2*>(v)[row + me];
        }
        return;
    }
    const float c2 = 0.5f * LOG2E / eps_target;
    float mx[OT_R], my[OT_R], own[OT_R];
    unsigned long long acc[OT_R];
#pragma unroll
    for (int r = 0; r < OT_R; ++r) {
        const int me = m0 + r * OT_T;
        const float4 mine = saved[row + (me < N ? me : 0)];
        mx[r] = mine.x; my[r] = mine.y;
        own[r] = TRANSPOSED ? mine.w : mine.z;     // forward rows carry F_i, transposed rows carry U_j
        acc[r] = 0ull;
    }
    for (int k0 = 0; k0 < N; k0 += OT_J) {
        const int kn = min(OT_J, N - k0);
        __syncthreads();
        for (int k = tid; k < OT_J; k += OT_T) {
            float4 q = make_float4(0.f, 0.f, -INFINITY, -INFINITY);
            float2 w = make_float2(0.f, 0.f);
            if (k < kn) { q = saved[row + k0 + k]; w = reinterpret_cast<const float2*>(v)[row + k0 + k]; }
            s_p[k] = q; s_v[k] = w;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kn; ++k) {
            const float4 q = s_p[k];
            const float2 w = s_v[k];
            const unsigned long long w2 = ot_pack2(w.x, w.y);
            const float other = TRANSPOSED ? q.z : q.w;
#pragma unroll
            for (int r = 0; r < OT_R; ++r) {
                const float dx = mx[r] - q.x, dy = my[r] - q.y;
                const float t = ex2f(fmaf(-c2, fmaf(dy, dy, dx * dx), own[r] + other));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[r]) : "l"(ot_pack2(t, t)), "l"(w2));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < OT_R; ++r) {
        const int me = m0 + r * OT_T;
        float ax, ay;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(ax), "=f"(ay) : "l"(acc[r]));
        if (me < N) reinterpret_cast<float2*>(out)[row + me] = make_float2(ax, ay);
    }
}

__global__ void ot_iters_kernel(const OtCtrl* ctrl, int* iters_out) { *iters_out = ctrl->iter + 2; }

struct OtWs {
    float2* sx; float *a, *b, *f, *g, *eps_run; unsigned* diff; OtCtrl* ctrl;
    static size_t bytes(int B, int N) {
        const size_t P = (size_t)B * N;
        return P * 8 + P * 4 * 2 * 2 + P * 4 * 2 + (size_t)B * 8 + 256;
    }
    OtWs(void* p, int B, int N) {
        const size_t P = (size_t)B * N;
        char* c = (char*)p;
        sx = (float2*)c; c += P * 8;
        a = (float*)c; c += P * 8;
        b = (float*)c; c += P * 8;
        f = (float*)c; c += P * 4;
        g = (float*)c; c += P * 4;
        eps_run = (float*)c; c += (size_t)B * 4;
        diff = (unsigned*)c; c += (size_t)B * 4;
        c = (char*)(((uintptr_t)c + 15) & ~(uintptr_t)15);
        ctrl = (OtCtrl*)c;
    }
};

}  // namespace nfdpf

using namespace nfdpf;

extern "C" int64_t nfdpf_ot_workspace(int B, int N) { return B > 0 && N > 0 ? (int64_t)OtWs::bytes(B, N) : 0; }

extern "C" int nfdpf_ot_resample_fwd(const float* particles, const float* logw, float eps, float scaling, float threshold,
                                     int max_iter, int B, int N, int d, float* particles_out, float* saved, int32_t* iters_out,
                                     void* workspace, const int32_t* gate, void* stream) {
    NFDPF_REQUIRE(particles && logw && particles_out && saved && workspace, "ot_resample_fwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0, "ot_resample_fwd: B and N must be positive");
    NFDPF_REQUIRE(eps > 0.f && scaling > 0.f && scaling < 1.f && max_iter >= 1, "ot_resample_fwd: need eps > 0, 0 < scaling < 1, max_iter >= 1");
    if (d != 2) { set_error("ot_resample_fwd: kernels are built for state_dim 2 (DPFs.py:31), got %d", d); return NFDPF_ERR_UNSUPPORTED; }
    if (B > 65535) { set_error("ot_resample_fwd: B <= 65535 per call (got %d)", B); return NFDPF_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    OtWs w(workspace, B, N);
    const dim3 pgrid((N + OT_T * OT_R - 1) / (OT_T * OT_R), B);      // every N x N kernel: OT_R rows (columns) per thread
    const float s2 = scaling * scaling;
    ot_prepare_kernel<<<B, 256, 0, st>>>(particles, N, w.sx, w.eps_run, w.diff, w.ctrl, max_iter, gate);
    int rc = check_launch("ot_prepare");
    if (rc) return rc;
    ot_pass_kernel<0><<<pgrid, OT_T, 0, st>>>(w.sx, logw, w.a, w.b, w.a, w.b, N, w.eps_run, w.eps_run, w.diff, w.ctrl, eps, s2, threshold,
                                             max_iter, B);
    if ((rc = check_launch("ot_init"))) return rc;
    for (int it = 0; it < max_iter - 1; ++it) {
        ot_pass_kernel<1><<<pgrid, OT_T, 0, st>>>(w.sx, logw, w.a, w.b, w.a, w.b, N, w.eps_run, w.eps_run, w.diff, w.ctrl, eps, s2,
                                                 threshold, max_iter, B);
        if ((rc = check_launch("ot_iter"))) return rc;
    }
    ot_pass_kernel<2><<<pgrid, OT_T, 0, st>>>(w.sx, logw, w.a, w.b, w.f, w.g, N, w.eps_run, w.eps_run, w.diff, w.ctrl, eps, s2, threshold,
                                             max_iter, B);
    if ((rc = check_launch("ot_final"))) return rc;
    ot_colnorm_kernel<<<pgrid, OT_T, 0, st>>>(w.sx, logw, w.f, w.g, N, eps, (float4*)saved, gate);
    if ((rc = check_launch("ot_colnorm"))) return rc;
    ot_apply_kernel<false><<<pgrid, OT_T, 0, st>>>((const float4*)saved, particles, N, eps, particles_out, gate);
    if ((rc = check_launch("ot_apply"))) return rc;
    if (iters_out) {
        ot_iters_kernel<<<1, 1, 0, st>>>(w.ctrl, iters_out);
        rc = check_launch("ot_iters");
    }
    return rc;
}

extern "C" int nfdpf_ot_resample_bwd(const float* g_out, const float* saved, float eps, int B, int N, int d, float* d_particles,
                                     const int32_t* gate, void* stream) {
    NFDPF_REQUIRE(g_out && saved && d_particles, "ot_resample_bwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && eps > 0.f, "ot_resample_bwd: bad sizes");
    if (d != 2) { set_error("ot_resample_bwd: kernels are built for state_dim 2, got %d", d); return NFDPF_ERR_UNSUPPORTED; }
    const dim3 grid((N + OT_T * OT_R - 1) / (OT_T * OT_R), B);
    ot_apply_kernel<true><<<grid, OT_T, 0, (cudaStream_t)stream>>>((const float4*)saved, g_out, N, eps, d_particles, gate);
    return check_launch("ot_apply_T");
}
