// (K1, headline shape) Backward of the D = 2 coupling stack with row-constant context (nf_dyn: C = 4, cond_model: C = 36).
//
// The generic backward (coupling.cu) reduces weight gradients through a shared-memory tile and is LDS-bound
// (2 shared loads per FMA).  Here every weight-gradient product is accumulated in REGISTERS: a CTA owns a
// trajectory, each thread owns N/128 particles whose state (5 floats) sits in shared memory between stages, and for
// one FCNN at a time the thread sweeps its particles accumulating the 97 hoisted-layout gradients
//   dW1[:,x] (8), db1 = sum of layer-1 deltas (8, also drives the row-context columns), dW2 (64), db2 (8), dW3 (8), db3 (1)
// in registers, then the CTA reduces them once per (trajectory chunk, FCNN): warp butterfly -> per-warp slots -> owner thread.
// Activations are recomputed from the stage output (couplings are invertible), so nothing but y is read from HBM.
#include "coupling.cuh"

namespace nfdpf {

using L2_ = Lay<1, 0>;
constexpr int NACC = 97;          // == Rows<1,0>::NOUT, same ordering as out_entry<1,0>
constexpr int CHUNK_M = 8;        // particles per thread per chunk (chunk = 1024 particles)

struct D2Smem {
    static size_t bytes(int n_fcnn, int C_row) {
        size_t fl = (size_t)n_fcnn * L2_::SIZE + n_fcnn * H + (size_t)n_fcnn * H * C_row   // images, hb, w1r
                    + 6 * (size_t)CHUNK_M * TP                                              // lo, up, glo, gup, gld, ds
                    + 4 * 100                                                                // per-warp reduction slots
                    + (size_t)n_fcnn * NACC + (size_t)n_fcnn * H * C_row + n_fcnn * H + C_row + 4;
        return fl * sizeof(float);
    }
};

// Reduce the 97 per-thread accumulators over the CTA and add them to s_acc[f] (and the b1 block to s_d1row[f]).
__device__ __forceinline__ void reduce_flush(float (&acc)[NACC], float* s_part, float* s_acc_f, float* s_d1row_f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NACC; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) s_part[warp * 100 + k] = v;
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        const int k = threadIdx.x;
        const float v = (s_part[k] + s_part[100 + k]) + (s_part[200 + k] + s_part[300 + k]);
        s_acc_f[k] += v;
        if (k >= H && k < 2 * H) s_d1row_f[k - H] += v;
    }
    __syncthreads();
}

__device__ __forceinline__ void accumulate(float (&acc)[NACC], const float (&d1)[H], const float (&d2)[H], float dout, float c,
                                           const float (&h1)[H], const float (&h2)[H]) {
#pragma unroll
    for (int k = 0; k < H; ++k) { acc[k] = fmaf(d1[k], c, acc[k]); acc[H + k] += d1[k]; }
#pragma unroll
    for (int j = 0; j < H; ++j) {
#pragma unroll
        for (int k = 0; k < H; ++k) acc[2 * H + j * H + k] = fmaf(d2[j], h1[k], acc[2 * H + j * H + k]);
        acc[2 * H + H * H + j] += d2[j];
        acc[3 * H + H * H + j] = fmaf(dout, h2[j], acc[3 * H + H * H + j]);
    }
    acc[4 * H + H * H] += dout;
}

// One stage (nets t = f_t, s = f_t + 1) for all particles of the chunk.  c/gc: conditioning half and its gradient;
// v/gv: transformed half (output value on entry, input value on exit) and its gradient.
template <bool INV>
__device__ __forceinline__ void stage_bwd_d2(const float* img_t, const float* img_s, const float* hb_t, const float* hb_s, int f_t,
                                             int m_count, int n_live, const float* s_c, float* s_gc, float* s_v, float* s_gv,
                                             const float* s_gld, float* s_ds, float* s_part, float* s_acc, float* s_d1row) {
    const int tid = threadIdx.x;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
    const float (*dummy)[1] = nullptr; (void)dummy;
#pragma unroll 1
    for (int m = 0; m < m_count; ++m) {
        asm volatile("" ::: "memory");
        const int q = m * TP + tid;
        const bool live = q < n_live;
        const float c[1] = {s_c[q]};
        float h1[H], h2[H], t[1], s[1], hx[H], hy[H];
        fcnn_fwd<1, 0>(img_s, hb_s, c, nullptr, hx, hy, s);
        fcnn_fwd<1, 0>(img_t, hb_t, c, nullptr, h1, h2, t);
        const float v = s_v[q], gv = s_gv[q], gld = s_gld[q];
        const float es = expf(s[0]), ies = expf(-s[0]);
        float dt, ds, vin, gin;
        if (!INV) { vin = (v - t[0]) * ies; dt = gv; ds = fmaf(gv * vin, es, gld); gin = gv * es; }
        else      { gin = gv * ies; dt = -gin; ds = -fmaf(gv, v, gld); vin = fmaf(v, es, t[0]); }
        if (!live) { dt = 0.f; ds = 0.f; }
        s_v[q] = vin; s_gv[q] = gin; s_ds[q] = ds;
        float d1[H], d2[H], dc[1] = {0.f};
        const float dout[1] = {dt};
        fcnn_bwd<1, 0>(img_t, dout, h1, h2, d1, d2, dc, nullptr);
        s_gc[q] += dc[0];
        accumulate(acc, d1, d2, dt, c[0], h1, h2);
    }
    reduce_flush(acc, s_part, s_acc + f_t * NACC, s_d1row + f_t * H);
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
#pragma unroll 1
    for (int m = 0; m < m_count; ++m) {
        asm volatile("" ::: "memory");
        const int q = m * TP + tid;
        const float c[1] = {s_c[q]};
        float h1[H], h2[H], s[1];
        fcnn_fwd<1, 0>(img_s, hb_s, c, nullptr, h1, h2, s);
        const float ds = s_ds[q];
        float d1[H], d2[H], dc[1] = {0.f};
        const float dout[1] = {ds};
        fcnn_bwd<1, 0>(img_s, dout, h1, h2, d1, d2, dc, nullptr);
        s_gc[q] += dc[0];
        accumulate(acc, d1, d2, ds, c[0], h1, h2);
    }
    reduce_flush(acc, s_part, s_acc + (f_t + 1) * NACC, s_d1row + (f_t + 1) * H);
}

__global__ void __launch_bounds__(TP)
coupling_bwd_d2_kernel(const float* __restrict__ packed, int n_flows, int C_row, const float* __restrict__ y,
                       const float* __restrict__ row_ctx, int flags, int B, int N, const float* __restrict__ g_y,
                       const float* __restrict__ g_ld, float* __restrict__ d_x, float* __restrict__ d_row_ctx,
                       float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem[];
    const int n_fcnn = 4 * n_flows, tid = threadIdx.x, inverse = flags & 1;
    float* s_img = smem;
    float* s_hb = s_img + n_fcnn * L2_::SIZE;
    float* s_w1r = s_hb + n_fcnn * H;
    float* s_lo = s_w1r + (size_t)n_fcnn * H * C_row;
    float* s_up = s_lo + CHUNK_M * TP;
    float* s_glo = s_up + CHUNK_M * TP;
    float* s_gup = s_glo + CHUNK_M * TP;
    float* s_gld = s_gup + CHUNK_M * TP;
    float* s_ds = s_gld + CHUNK_M * TP;
    float* s_part = s_ds + CHUNK_M * TP;
    float* s_acc = s_part + 400;
    float* s_accR = s_acc + n_fcnn * NACC;
    float* s_d1row = s_accR + (size_t)n_fcnn * H * C_row;
    float* s_ctx = s_d1row + n_fcnn * H;
    const int pf = packed_fcnn_size(1, C_row);
    for (int f = 0; f < n_fcnn; ++f)
        load_fcnn_image<1, 0>(packed + (size_t)f * pf, C_row, s_img + f * L2_::SIZE, s_w1r + (size_t)f * H * C_row, tid, TP);
    for (int e = tid; e < n_fcnn * NACC; e += TP) s_acc[e] = 0.f;
    for (int e = tid; e < n_fcnn * H * C_row; e += TP) s_accR[e] = 0.f;
    __syncthreads();

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int e = tid; e < C_row; e += TP) s_ctx[e] = row_ctx[(size_t)b * C_row + e];
        for (int e = tid; e < n_fcnn * H; e += TP) s_d1row[e] = 0.f;
        __syncthreads();
        hoist_row_context<1, 0>(s_img, s_w1r, s_ctx, C_row, n_fcnn, s_hb, tid, TP);
        __syncthreads();
        for (int c0 = 0; c0 < N; c0 += CHUNK_M * TP) {
            const int n_live = min(CHUNK_M * TP, N - c0);
            const int m_count = (n_live + TP - 1) / TP;
            const size_t p0 = (size_t)b * N + c0;
            for (int q = tid; q < m_count * TP; q += TP) {   // each thread touches only its own slots q = m*TP + tid
                const bool live = q < n_live;
                const float2 yy = live ? reinterpret_cast<const float2*>(y)[p0 + q] : make_float2(0.f, 0.f);
                const float2 gg = live && g_y ? reinterpret_cast<const float2*>(g_y)[p0 + q] : make_float2(0.f, 0.f);
                s_lo[q] = yy.x; s_up[q] = yy.y; s_glo[q] = gg.x; s_gup[q] = gg.y;
                s_gld[q] = live && g_ld ? ((flags & 2) ? -g_ld[p0 + q] : g_ld[p0 + q]) : 0.f;
            }
            if (!inverse) {
#pragma unroll 1
                for (int f = n_flows - 1; f >= 0; --f) {
                    const float* im = s_img + 4 * f * L2_::SIZE;
                    const float* hb = s_hb + 4 * f * H;
                    stage_bwd_d2<false>(im + 2 * L2_::SIZE, im + 3 * L2_::SIZE, hb + 2 * H, hb + 3 * H, 4 * f + 2, m_count, n_live, s_up, s_gup,
                                        s_lo, s_glo, s_gld, s_ds, s_part, s_acc, s_d1row);
                    stage_bwd_d2<false>(im, im + L2_::SIZE, hb, hb + H, 4 * f, m_count, n_live, s_lo, s_glo, s_up, s_gup, s_gld, s_ds, s_part,
                                        s_acc, s_d1row);
                }
            } else {
#pragma unroll 1
                for (int f = 0; f < n_flows; ++f) {
                    const float* im = s_img + 4 * f * L2_::SIZE;
                    const float* hb = s_hb + 4 * f * H;
                    stage_bwd_d2<true>(im, im + L2_::SIZE, hb, hb + H, 4 * f, m_count, n_live, s_lo, s_glo, s_up, s_gup, s_gld, s_ds, s_part,
                                       s_acc, s_d1row);
                    stage_bwd_d2<true>(im + 2 * L2_::SIZE, im + 3 * L2_::SIZE, hb + 2 * H, hb + 3 * H, 4 * f + 2, m_count, n_live, s_up, s_gup,
                                       s_lo, s_glo, s_gld, s_ds, s_part, s_acc, s_d1row);
                }
            }
            for (int q = tid; q < n_live; q += TP) reinterpret_cast<float2*>(d_x)[p0 + q] = make_float2(s_glo[q], s_gup[q]);
        }
        for (int e = tid; e < n_fcnn * H * C_row; e += TP) s_accR[e] = fmaf(s_d1row[e / C_row], s_ctx[e % C_row], s_accR[e]);
        if (d_row_ctx)
            for (int cidx = tid; cidx < C_row; cidx += TP) {
                float a = 0.f;
                for (int fk = 0; fk < n_fcnn * H; ++fk) a = fmaf(s_w1r[(size_t)fk * C_row + cidx], s_d1row[fk], a);
                d_row_ctx[(size_t)b * C_row + cidx] = a;
            }
        __syncthreads();
    }
    float* out = partials + (size_t)blockIdx.x * n_fcnn * pf;
    const int fin = 1 + C_row;
    for (int e = tid; e < n_fcnn * NACC; e += TP) {
        int ra, rb, poff;
        out_entry<1, 0>(e % NACC, C_row, ra, rb, poff);
        out[(size_t)(e / NACC) * pf + poff] = s_acc[e];
    }
    for (int e = tid; e < n_fcnn * H * C_row; e += TP) {
        const int fk = e / C_row, cidx = e % C_row;
        out[(size_t)(fk / H) * pf + (fk % H) * fin + 1 + cidx] = s_accR[e];
    }
}

int launch_coupling_bwd_d2(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int inverse, int B, int N,
                           const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* d_packed, void* workspace,
                           cudaStream_t st) {
    const int n_fcnn = 4 * n_flows;
    const size_t smem = D2Smem::bytes(n_fcnn, C_row);
    if (smem > 48 * 1024) NFDPF_CUDA(cudaFuncSetAttribute(coupling_bwd_d2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = bwd_grid(B);
    const int n_params = n_fcnn * packed_fcnn_size(1, C_row);
    coupling_bwd_d2_kernel<<<grid, TP, smem, st>>>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx,
                                                    (float*)workspace);
    int rc = check_launch("coupling_bwd_d2");
    if (rc) return rc;
    return launch_reduce_partials((const float*)workspace, grid, n_params, d_packed, st);
}

}  // namespace nfdpf
