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
constexpr int NACC = 97;          // == Rows<1,0>::NOUT, same ordering as packed_offset<1,0>
constexpr int TPD = 256;          // threads per CTA (8 warps walking the same code: one CTA per SM, instruction-cache friendly)
constexpr int CHUNK_M = 4;        // particles per thread per chunk (chunk = 1024 particles)
constexpr int CHUNK = CHUNK_M * TPD;
constexpr int NWARP = TPD / 32;

struct D2Smem {
    static size_t bytes(int n_fcnn, int C_row) {
        size_t fl = (size_t)n_fcnn * L2_::SIZE + n_fcnn * H + (size_t)n_fcnn * H * C_row   // images, hb, w1r
                    + 6 * (size_t)CHUNK                                                     // lo, up, glo, gup, gld, ds
                    + 16 * (size_t)CHUNK                                                    // s-net activations (h1, h2) stash
                    + NWARP * 100                                                            // per-warp reduction slots
                    + (size_t)n_fcnn * NACC + n_fcnn * H + C_row + 4;
        return fl * sizeof(float);
    }
};

// Reduce the 97 per-thread accumulators over the CTA and add them to s_acc[f] (and the b1 block to s_d1row[f]).
// Warp level: transposed butterfly -- in round r the lanes with bit (16 >> r) set keep the upper half of the live
// values and send the lower half (and vice versa), so 96 values cost 93 shuffles instead of 480; lane L ends up
// owning the sums of entries 3L..3L+2.  Entry 96 (db3) takes a plain butterfly.
__device__ __forceinline__ void reduce_flush(float (&acc)[NACC], float* s_part, float* s_acc_f, float* s_d1row_f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#define NFDPF_ROUND(HALFN, OFF)                                                         \
    {                                                                                   \
        const bool up_ = (lane & OFF) != 0;                                             \
        _Pragma("unroll") for (int i = 0; i < HALFN; ++i) {                             \
            const float keep = up_ ? acc[i + HALFN] : acc[i];                           \
            const float send = up_ ? acc[i] : acc[i + HALFN];                           \
            acc[i] = keep + __shfl_xor_sync(FULL, send, OFF);                           \
        }                                                                               \
    }
    NFDPF_ROUND(48, 16) NFDPF_ROUND(24, 8) NFDPF_ROUND(12, 4) NFDPF_ROUND(6, 2) NFDPF_ROUND(3, 1)
#undef NFDPF_ROUND
    float last = acc[96];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) last += __shfl_xor_sync(FULL, last, o);
    float* slot = s_part + warp * 100;
    slot[3 * lane] = acc[0]; slot[3 * lane + 1] = acc[1]; slot[3 * lane + 2] = acc[2];
    if (lane == 0) slot[96] = last;
    __syncthreads();
    if (threadIdx.x < NACC) {
        const int k = threadIdx.x;
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) v += s_part[w * 100 + k];
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
// v/gv: transformed half (output value on entry, input value on exit) and its gradient.  inv: the stage was run as
// v_out = (v_in - t) e^{-s} (inverse direction) instead of v_out = t + v_in e^{s}.
// Sweep 1: both nets forward (s-net activations stashed in shared memory), invert, t-net backward + gradient products.
// Sweep 2: s-net backward from the stash.  Not inlined: ONE copy of this code serves all stages.
__device__ __noinline__ void stage_bwd_d2(const float* img_t, const float* img_s, const float* hb_t, const float* hb_s, int f_t, bool inv,
                                          int m_count, int n_live, const float* s_c, float* s_gc, float* s_v, float* s_gv,
                                          const float* s_gld, float* s_ds, float* s_stash, float* s_part, float* s_acc, float* s_d1row) {
    const int tid = threadIdx.x;
    float acc[NACC];
#pragma unroll
    for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
#pragma unroll 1
    for (int m = 0; m < m_count; ++m) {
        asm volatile("" ::: "memory");
        const int q = m * TPD + tid;
        const bool live = q < n_live;
        const float c[1] = {s_c[q]};
        float h1[H], h2[H], t[1], s[1];
        fcnn_fwd<1, 0>(img_s, hb_s, c, nullptr, h1, h2, s);
#pragma unroll
        for (int k = 0; k < H; ++k) { s_stash[k * CHUNK + q] = h1[k]; s_stash[(H + k) * CHUNK + q] = h2[k]; }
        fcnn_fwd<1, 0>(img_t, hb_t, c, nullptr, h1, h2, t);
        const float v = s_v[q], gv = s_gv[q], gld = s_gld[q];
        const float es = expf(s[0]), ies = expf(-s[0]);
        float dt, ds, vin, gin;
        if (!inv) { vin = (v - t[0]) * ies; dt = gv; ds = fmaf(gv * vin, es, gld); gin = gv * es; }
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
        const int q = m * TPD + tid;
        float h1[H], h2[H];
#pragma unroll
        for (int k = 0; k < H; ++k) { h1[k] = s_stash[k * CHUNK + q]; h2[k] = s_stash[(H + k) * CHUNK + q]; }
        const float ds = s_ds[q];
        float d1[H], d2[H], dc[1] = {0.f};
        const float dout[1] = {ds};
        fcnn_bwd<1, 0>(img_s, dout, h1, h2, d1, d2, dc, nullptr);
        s_gc[q] += dc[0];
        accumulate(acc, d1, d2, ds, s_c[q], h1, h2);
    }
    reduce_flush(acc, s_part, s_acc + (f_t + 1) * NACC, s_d1row + (f_t + 1) * H);
}

__global__ void __launch_bounds__(TPD)
coupling_bwd_d2_kernel(const float* __restrict__ packed, int n_flows, int C_row, const float* __restrict__ y,
                       const float* __restrict__ row_ctx, int flags, int B, int N, const float* __restrict__ g_y,
                       const float* __restrict__ g_ld, float* __restrict__ d_x, float* __restrict__ d_row_ctx,
                       float* __restrict__ partials, float* __restrict__ d1rows) {
    extern __shared__ __align__(16) float smem[];
    const int n_fcnn = 4 * n_flows, tid = threadIdx.x, inverse = flags & 1;
    float* s_img = smem;
    float* s_hb = s_img + n_fcnn * L2_::SIZE;
    float* s_w1r = s_hb + n_fcnn * H;
    float* s_lo = s_w1r + (size_t)n_fcnn * H * C_row;
    float* s_up = s_lo + CHUNK;
    float* s_glo = s_up + CHUNK;
    float* s_gup = s_glo + CHUNK;
    float* s_gld = s_gup + CHUNK;
    float* s_ds = s_gld + CHUNK;
    float* s_stash = s_ds + CHUNK;
    float* s_part = s_stash + 16 * CHUNK;
    float* s_acc = s_part + NWARP * 100;
    float* s_d1row = s_acc + n_fcnn * NACC;
    float* s_ctx = s_d1row + n_fcnn * H;
    const int pf = packed_fcnn_size(1, C_row);
    for (int f = 0; f < n_fcnn; ++f)
        load_fcnn_image<1, 0>(packed + (size_t)f * pf, C_row, s_img + f * L2_::SIZE, s_w1r + (size_t)f * H * C_row, tid, TPD);
    for (int e = tid; e < n_fcnn * NACC; e += TPD) s_acc[e] = 0.f;
    __syncthreads();

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int e = tid; e < C_row; e += TPD) s_ctx[e] = row_ctx[(size_t)b * C_row + e];
        for (int e = tid; e < n_fcnn * H; e += TPD) s_d1row[e] = 0.f;
        __syncthreads();
        hoist_row_context_par<1, 0>(s_img, s_w1r, s_ctx, C_row, n_fcnn, s_hb);
        __syncthreads();
        for (int c0 = 0; c0 < N; c0 += CHUNK) {
            const int n_live = min(CHUNK, N - c0);
            const int m_count = (n_live + TPD - 1) / TPD;
            const size_t p0 = (size_t)b * N + c0;
            for (int q = tid; q < m_count * TPD; q += TPD) {   // each thread touches only its own slots q = m*TPD + tid
                const bool live = q < n_live;
                const float2 yy = live ? reinterpret_cast<const float2*>(y)[p0 + q] : make_float2(0.f, 0.f);
                const float2 gg = live && g_y ? reinterpret_cast<const float2*>(g_y)[p0 + q] : make_float2(0.f, 0.f);
                s_lo[q] = yy.x; s_up[q] = yy.y; s_glo[q] = gg.x; s_gup[q] = gg.y;
                s_gld[q] = live && g_ld ? ((flags & 2) ? -g_ld[p0 + q] : g_ld[p0 + q]) : 0.f;
            }
            // forward pass ran flows 0..n-1 (t1/s1 then t2/s2): walk back n-1..0 (t2/s2 then t1/s1);
            // inverse pass ran flows n-1..0 (t2/s2 then t1/s1): walk back 0..n-1 (t1/s1 then t2/s2)
#pragma unroll 1
            for (int st = 0; st < 2 * n_flows; ++st) {
                const int f = inverse ? st / 2 : n_flows - 1 - st / 2;
                const int pair = inverse ? (st & 1) : 1 - (st & 1);       // 0: t1/s1 (c = lower), 1: t2/s2 (c = upper)
                const float* im = s_img + (4 * f + 2 * pair) * L2_::SIZE;
                const float* hb = s_hb + (4 * f + 2 * pair) * H;
                stage_bwd_d2(im, im + L2_::SIZE, hb, hb + H, 4 * f + 2 * pair, inverse != 0, m_count, n_live, pair ? s_up : s_lo,
                             pair ? s_gup : s_glo, pair ? s_lo : s_up, pair ? s_glo : s_gup, s_gld, s_ds, s_stash, s_part, s_acc, s_d1row);
            }
            for (int q = tid; q < n_live; q += TPD) reinterpret_cast<float2*>(d_x)[p0 + q] = make_float2(s_glo[q], s_gup[q]);
        }
        // per-trajectory layer-1 delta sums: the row-context columns of dW1 and d(row_ctx) are formed from them by
        // rowctx_grad_kernel after this kernel (keeps the 8 x C_row outer products out of the persistent loop)
        for (int e = tid; e < n_fcnn * H; e += TPD) d1rows[(size_t)e * B + b] = TANH_SCALE * s_d1row[e];   // true delta sums, [f*8+k][b]: coalesced for rowctx_grad
        __syncthreads();
    }
    float* out = partials + (size_t)blockIdx.x * n_fcnn * pf;
    for (int e = tid; e < n_fcnn * NACC; e += TPD) {
        out[(size_t)(e / NACC) * pf + packed_offset<1, 0>(e % NACC, C_row)] = s_acc[e] * grad_out_scale<1, 0>(e % NACC);
    }
    const int fin = 1 + C_row;
    for (int e = tid; e < n_fcnn * H * C_row; e += TPD)   // row-context columns are produced by rowctx_grad_kernel
        out[(size_t)(e / (H * C_row)) * pf + ((e / C_row) % H) * fin + 1 + (e % C_row)] = 0.f;
}

// dW1[f][k][1 + c] += sum_b D1[b][f][k] * ctx[b][c]   (one warp per output, fp64, fixed order)
// d_row_ctx[b][c]   = sum_{f,k} W1[f][k][1 + c] * D1[b][f][k]
__global__ void rowctx_grad_kernel(const float* __restrict__ packed, const float* __restrict__ d1rows, const float* __restrict__ row_ctx,
                                   int n_fcnn, int C_row, int B, float* __restrict__ d_packed, float* __restrict__ d_row_ctx) {
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int n_w = n_fcnn * H * C_row, pf = packed_fcnn_size(1, C_row), fin = 1 + C_row;
    if (warp < n_w) {
        const int fk = warp / C_row, c = warp % C_row;
        double a = 0.0;
        for (int b = lane; b < B; b += 32) a += (double)d1rows[(size_t)fk * B + b] * (double)row_ctx[(size_t)b * C_row + c];
        a = warp_sum(a);
        if (lane == 0) d_packed[(size_t)(fk / H) * pf + (fk % H) * fin + 1 + c] += (float)a;
    } else if (d_row_ctx) {
        const int e = warp - n_w;            // one warp per (b, c)
        if (e >= B * C_row) return;
        const int b = e / C_row, c = e % C_row;
        float a = 0.f;
        for (int fk = lane; fk < n_fcnn * H; fk += 32)
            a = fmaf(packed[(size_t)(fk / H) * pf + (fk % H) * fin + 1 + c], d1rows[(size_t)fk * B + b], a);
        a = warp_sum(a);
        if (lane == 0) d_row_ctx[(size_t)b * C_row + c] = a;
    }
}

int launch_coupling_bwd_d2(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int inverse, int B, int N,
                           const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* d_packed, void* workspace,
                           cudaStream_t st) {
    const int n_fcnn = 4 * n_flows;
    const size_t smem = D2Smem::bytes(n_fcnn, C_row);
    if (smem > 48 * 1024) NFDPF_CUDA(cudaFuncSetAttribute(coupling_bwd_d2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = min(B, sm_count());   // one 8-warp CTA per SM (registers), persistent over trajectories
    const int n_params = n_fcnn * packed_fcnn_size(1, C_row);
    float* d1rows = (float*)workspace + (size_t)bwd_grid(B) * n_params;
    coupling_bwd_d2_kernel<<<grid, TPD, smem, st>>>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx,
                                                    (float*)workspace, d1rows);
    int rc = check_launch("coupling_bwd_d2");
    if (rc) return rc;
    rc = launch_reduce_partials((const float*)workspace, grid, n_params, d_packed, st);
    if (rc || C_row == 0) return rc;
    const int warps = n_fcnn * H * C_row + (d_row_ctx ? B * C_row : 0);
    rowctx_grad_kernel<<<(warps + 7) / 8, 256, 0, st>>>(packed, d1rows, row_ctx, n_fcnn, C_row, B, d_packed, d_row_ctx);
    return check_launch("rowctx_grad");
}

}  // namespace nfdpf
