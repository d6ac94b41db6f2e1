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
template <int MODE>
__global__ void __launch_bounds__(OT_T)
ot_pass_kernel(const float2* __restrict__ sx, const float* __restrict__ logw, const float* __restrict__ pot_a,
               const float* __restrict__ pot_b, float* __restrict__ out_a, float* __restrict__ out_b, int N,
               const float* eps_run_in, float* eps_run, unsigned* diff,
               OtCtrl* __restrict__ ctrl, float eps_target, float scaling2, float threshold, int max_iter, int B) {
    __shared__ float4 s_j[OT_J];
    __shared__ float s_red[33];
    __shared__ int s_last;
    const int b = blockIdx.y, tid = threadIdx.x, i = blockIdx.x * OT_T + tid;
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
    const float log_beta = -logf((float)N);
    const bool live = i < N;
    const float2 xi = sx[row + (live ? i : 0)];
    float m1 = -INFINITY, s1 = 0.f, m2 = -INFINITY, s2 = 0.f;
    for (int j0 = 0; j0 < N; j0 += OT_J) {
        const int jn = min(OT_J, N - j0);
        __syncthreads();
        for (int j = tid; j < OT_J; j += OT_T) {
            float4 v = make_float4(0.f, 0.f, -INFINITY, -INFINITY);      // padding columns contribute exp(-inf) = 0
            if (j < jn) {
                const float2 p = sx[row + j0 + j];
                float h1 = logw[row + j0 + j], h2 = log_beta;
                if (MODE != 0) { h1 = fmaf(b_old[j0 + j], inv, h1); h2 = fmaf(a_old[j0 + j], inv, h2); }
                v = make_float4(p.x, p.y, h1 * LOG2E, h2 * LOG2E);
            }
            s_j[j] = v;
        }
        __syncthreads();
        const int jend = (jn + OT_CH - 1) / OT_CH * OT_CH;
        for (int j = 0; j < jend; j += OT_CH) {
            float v1[OT_CH], v2[OT_CH];
            float c1 = -INFINITY, cm2 = -INFINITY;
#pragma unroll
            for (int u = 0; u < OT_CH; ++u) {
                const float4 q = s_j[j + u];
                const float dx = xi.x - q.x, dy = xi.y - q.y;
                const float d2 = fmaf(dy, dy, dx * dx);
                v1[u] = fmaf(-c2, d2, q.z);
                v2[u] = fmaf(-c2, d2, q.w);
                c1 = fmaxf(c1, v1[u]);
                cm2 = fmaxf(cm2, v2[u]);
            }
            // every chunk of a tile holds at least one real column (tiles start on real columns, padding is at the tail) and
            // log-weights are finite (the filter adds 1e-12 to every weight, DPFs.py:192), so n is finite and no
            // inf - inf can form: the first chunk sees s * 2^(-inf - n) = 0 * 0, padding terms are 2^(-inf) = 0.
            const float n1 = fmaxf(m1, c1), n2 = fmaxf(m2, cm2);
            s1 *= ex2f(m1 - n1);
            s2 *= ex2f(m2 - n2);
#pragma unroll
            for (int u = 0; u < OT_CH; ++u) {
                s1 += ex2f(v1[u] - n1);
                s2 += ex2f(v2[u] - n2);
            }
            m1 = n1; m2 = n2;
        }
    }
    const float sm1 = -eps * LN2 * (m1 + lg2f(s1));   // softmin = -eps * LSE
    const float sm2 = -eps * LN2 * (m2 + lg2f(s2));
    if (MODE != 1) {
        if (live) {
            float* oa = out_a + row;
            float* ob = out_b + row;
            oa[i] = sm1; ob[i] = sm2;
        }
        return;
    }
    float d = 0.f;
    if (live) {
        const float ao = a_old[i], bo = b_old[i];
        const float an = 0.5f * (ao + sm1), bn = 0.5f * (bo + sm2);    // resamplers.py:147-148
        out_a[(cur ^ 1) * pstride + row + i] = an;
        out_b[(cur ^ 1) * pstride + row + i] = bn;
        d = fmaxf(fabsf(an - ao), fabsf(bn - bo));
    }
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
// saved[j] = (sx_j, sy_j, F_j, U_j) in log2 units: F = f/eps, U = g/eps - LSE_i((f_i + g_j - C_ij)/eps) + log N + logw_j,
// so that T_ij = 2^(F_i + U_j - c2 |x_i - x_j|^2).
__global__ void __launch_bounds__(OT_T)
ot_colnorm_kernel(const float2* __restrict__ sx, const float* __restrict__ logw, const float* __restrict__ f,
                  const float* __restrict__ g, int N, float eps_target, float4* __restrict__ saved, const int* __restrict__ gate) {
    __shared__ float4 s_j[OT_J];
    if (gate && *gate == 0) return;
    const int b = blockIdx.y, tid = threadIdx.x, jcol = blockIdx.x * OT_T + tid;
    const size_t row = (size_t)b * N;
    const float inv = 1.0f / eps_target, c2 = 0.5f * LOG2E * inv;
    const bool live = jcol < N;
    const float2 xj = sx[row + (live ? jcol : 0)];
    float m = -INFINITY, s = 0.f;
    for (int i0 = 0; i0 < N; i0 += OT_J) {
        const int in = min(OT_J, N - i0);
        __syncthreads();
        for (int i = tid; i < OT_J; i += OT_T) {
            float4 v = make_float4(0.f, 0.f, -INFINITY, 0.f);
            if (i < in) { const float2 p = sx[row + i0 + i]; v = make_float4(p.x, p.y, f[row + i0 + i] * inv * LOG2E, 0.f); }
            s_j[i] = v;
        }
        __syncthreads();
        const int iend = (in + OT_CH - 1) / OT_CH * OT_CH;
        for (int i = 0; i < iend; i += OT_CH) {
            float v[OT_CH], cm = -INFINITY;
#pragma unroll
            for (int u = 0; u < OT_CH; ++u) {
                const float4 q = s_j[i + u];
                const float dx = xj.x - q.x, dy = xj.y - q.y;
                v[u] = fmaf(-c2, fmaf(dy, dy, dx * dx), q.z);
                cm = fmaxf(cm, v[u]);
            }
            const float n = fmaxf(m, cm);
            s *= ex2f(m - n);
#pragma unroll
            for (int u = 0; u < OT_CH; ++u) s += ex2f(v[u] - n);
            m = n;
        }
    }
    if (live) {
        const float lse2 = m + lg2f(s);   // log2-domain LSE_i(f_i/eps - C_ij/eps); g_j/eps cancels inside U
        const float U = -lse2 + (logf((float)N) + logw[row + jcol]) * LOG2E;
        (void)g;
        saved[row + jcol] = make_float4(xj.x, xj.y, f[row + jcol] * inv * LOG2E, U);
    }
}

// ---- apply the plan: out_i = sum_j T_ij v_j (forward: v = particles) or out_j = sum_i T_ij v_i (backward: v = grad) ----
template <bool TRANSPOSED>
__global__ void __launch_bounds__(OT_T)
ot_apply_kernel(const float4* __restrict__ saved, const float* __restrict__ v, int N, float eps_target, float* __restrict__ out,
                const int* __restrict__ gate) {
    __shared__ float4 s_p[OT_J];
    __shared__ float2 s_v[OT_J];
    const int b = blockIdx.y, tid = threadIdx.x, me = blockIdx.x * OT_T + tid;
    const size_t row = (size_t)b * N;
    if (gate && *gate == 0) {            // resampling gated off: the plan is the identity (forward and transposed)
        if (me < N) reinterpret_cast<float2*>(out)[row + me] = reinterpret_cast<const float2*>(v)[row + me];
        return;
    }
    const float c2 = 0.5f * LOG2E / eps_target;
    const bool live = me < N;
    const float4 mine = saved[row + (live ? me : 0)];
    const float own = TRANSPOSED ? mine.w : mine.z;     // forward rows carry F_i, transposed rows carry U_j
    float ax = 0.f, ay = 0.f;
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
            const float dx = mine.x - q.x, dy = mine.y - q.y;
            const float t = ex2f(fmaf(-c2, fmaf(dy, dy, dx * dx), own + (TRANSPOSED ? q.z : q.w)));
            const float2 w = s_v[k];
            ax = fmaf(t, w.x, ax); ay = fmaf(t, w.y, ay);
        }
    }
    if (live) reinterpret_cast<float2*>(out)[row + me] = make_float2(ax, ay);
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
    const dim3 grid((N + OT_T - 1) / OT_T, B);
    const float s2 = scaling * scaling;
    ot_prepare_kernel<<<B, 256, 0, st>>>(particles, N, w.sx, w.eps_run, w.diff, w.ctrl, max_iter, gate);
    int rc = check_launch("ot_prepare");
    if (rc) return rc;
    ot_pass_kernel<0><<<grid, OT_T, 0, st>>>(w.sx, logw, w.a, w.b, w.a, w.b, N, w.eps_run, w.eps_run, w.diff, w.ctrl, eps, s2, threshold,
                                             max_iter, B);
    if ((rc = check_launch("ot_init"))) return rc;
    for (int it = 0; it < max_iter - 1; ++it) {
        ot_pass_kernel<1><<<grid, OT_T, 0, st>>>(w.sx, logw, w.a, w.b, w.a, w.b, N, w.eps_run, w.eps_run, w.diff, w.ctrl, eps, s2,
                                                 threshold, max_iter, B);
        if ((rc = check_launch("ot_iter"))) return rc;
    }
    ot_pass_kernel<2><<<grid, OT_T, 0, st>>>(w.sx, logw, w.a, w.b, w.f, w.g, N, w.eps_run, w.eps_run, w.diff, w.ctrl, eps, s2, threshold,
                                             max_iter, B);
    if ((rc = check_launch("ot_final"))) return rc;
    ot_colnorm_kernel<<<grid, OT_T, 0, st>>>(w.sx, logw, w.f, w.g, N, eps, (float4*)saved, gate);
    if ((rc = check_launch("ot_colnorm"))) return rc;
    ot_apply_kernel<false><<<grid, OT_T, 0, st>>>((const float4*)saved, particles, N, eps, particles_out, gate);
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
    const dim3 grid((N + OT_T - 1) / OT_T, B);
    ot_apply_kernel<true><<<grid, OT_T, 0, (cudaStream_t)stream>>>((const float4*)saved, g_out, N, eps, d_particles, gate);
    return check_launch("ot_apply_T");
}
