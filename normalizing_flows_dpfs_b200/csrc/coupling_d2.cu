// (K1, headline shape) Backward of the D = 2 coupling stack with row-constant context (nf_dyn: C = 4, cond_model: C = 36).
//
// Every weight-gradient product is accumulated in REGISTERS (the generic backward of coupling.cu reduces them through a
// shared-memory tile and is LDS-bound).  Layout of the work:
//   * a persistent CTA keeps the state (lo, up, g_lo, g_up, g_logdet: 5 floats) of up to E_MAX "entries" (<= 1024 particles
//     of one trajectory each) resident in shared memory and walks the stack STAGE-OUTER: all resident particles go through
//     stage st before anybody starts stage st - 1.  A thread therefore keeps the 97 gradient accumulators of ONE net
//       dW1[:,x] (8), db1 (8, per entry: it also drives the row-context columns), dW2 (64), db2 (8), dW3 (8), db3 (1)
//     live across ~56 particles and the CTA-wide reduction (warp butterfly -> per-warp slots -> owner thread) runs once per
//     (resident set, net) instead of once per (trajectory, net);
//   * the two nets of a stage are split over the two halves of the CTA: threads 0-127 own the t-net, threads 128-255 the
//     s-net, for the SAME particles.  Each half runs its net forward (activations recomputed from the stage output: couplings
//     are invertible, nothing but y is read from HBM), the halves exchange t and s through shared memory (one 64-thread
//     named barrier per warp pair and 64 particles), then each runs its net backward with the activations still in registers -- no activation stash;
//   * a thread evaluates its net for TWO particles at once: the pair is the packed operand of every FFMA2 (fma.rn.f32x2, the
//     weight being the broadcast scalar), so every weight is loaded once per two particles and the FMA issue slots halve.
#include "coupling.cuh"

namespace nfdpf {

using L2_ = Lay<1, 0>;
constexpr int NACC = 97;          // == Rows<1,0>::NOUT, same ordering as packed_offset<1,0>
constexpr int TPD = 256;          // threads per CTA: two net-groups of four warps, one CTA per SM
constexpr int GRP = 128;          // threads per net-group
constexpr int PPI = 2 * GRP;      // particles per iteration: two per thread (q, q + GRP)
constexpr int CHUNK = 1024;       // particles per entry
constexpr int NWARP = TPD / 32;
constexpr int E_CAP = 9;          // upper bound of resident entries (the launcher fits E_MAX <= E_CAP into shared memory)

struct D2Smem {
    static size_t fixed_floats(int n_fcnn, int C_row) {
        return (size_t)n_fcnn * L2_::SIZE + (size_t)n_fcnn * H * C_row        // images, w1r
               + 12 * GRP                                                      // exchange: t, s, dc_t (two particles per thread, double buffered)
               + NWARP * 100                                                   // per-warp reduction slots
               + (size_t)n_fcnn * NACC + 8;                                    // acc
    }
    static size_t entry_floats(int n_fcnn, int C_row) {
        return 5 * (size_t)CHUNK                                               // lo, up, glo, gup, gld
               + 2 * (size_t)n_fcnn * H                                        // hb, d1row
               + NWARP * H + C_row;                                            // per-warp b1 sums, ctx
    }
    static size_t bytes(int n_fcnn, int C_row, int e_max) { return (fixed_floats(n_fcnn, C_row) + e_max * entry_floats(n_fcnn, C_row)) * sizeof(float); }
};

// Warp-level transposed butterfly of the 97 per-thread accumulators: in round r the lanes with bit (16 >> r) set keep the
// upper half of the live values and send the lower half (and vice versa), so 96 values cost 93 shuffles instead of 480;
// lane L ends up owning the sums of entries 3L..3L+2.  Entry 96 (db3) takes a plain butterfly.  Result -> the warp's slot.
__device__ __forceinline__ void warp_reduce_to_slot(float (&acc)[NACC], float* s_part) {
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
}

// Warp sums of the 8 layer-1 delta accumulators of one entry -> d1part[warp][8]; the accumulators are cleared.
__device__ __forceinline__ void warp_reduce_b1(float (&acc)[NACC], float* s_d1part_e) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float v[H];
#pragma unroll
    for (int k = 0; k < H; ++k) { v[k] = acc[H + k]; acc[H + k] = 0.f; }
#define NFDPF_ROUND(HALFN, OFF)                                                         \
    {                                                                                   \
        const bool up_ = (lane & OFF) != 0;                                             \
        _Pragma("unroll") for (int i = 0; i < HALFN; ++i) {                             \
            const float keep = up_ ? v[i + HALFN] : v[i];                               \
            const float send = up_ ? v[i] : v[i + HALFN];                               \
            v[i] = keep + __shfl_xor_sync(FULL, send, OFF);                             \
        }                                                                               \
    }
    NFDPF_ROUND(4, 16) NFDPF_ROUND(2, 8) NFDPF_ROUND(1, 4)
#undef NFDPF_ROUND
    v[0] += __shfl_xor_sync(FULL, v[0], 2);
    v[0] += __shfl_xor_sync(FULL, v[0], 1);
    if ((lane & 3) == 0) s_d1part_e[warp * H + (lane >> 2)] = v[0];   // lane bits 4,3,2 select the entry index
}

// Two particles per thread: the pair (particle 0, particle 1) is the packed operand of every FMA, the weight its broadcast
// scalar -- every weight is loaded once per two particles and no value ever needs re-packing.  Arrays are [unit][particle].
__device__ __forceinline__ void fwd_d2x2(const float* __restrict__ img, const float* __restrict__ hb, float c0, float c1, float (&h1)[H][2],
                                         float (&h2)[H][2], float (&out)[2]) {
    using L = L2_;
    float hbv[8];
    ld8(hb, hbv);
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float a0 = hbv[k], a1 = hbv[k];
        ffma2_s(a0, a1, img[L::W1 + k * L::S1], c0, c1);
        tanh_prescaled_pair(a0, a1, h1[k][0], h1[k][1]);
    }
    float b2[8];
    ld8(img + L::B2, b2);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
        float a0 = b2[j], a1 = b2[j];
#pragma unroll
        for (int k = 0; k < H; ++k) ffma2_s(a0, a1, w[k], h1[k][0], h1[k][1]);
        tanh_prescaled_pair(a0, a1, h2[j][0], h2[j][1]);
    }
    float w3[8];
    ld8(img + L::W3, w3);
    float o0 = img[L::B3], o1 = o0;
#pragma unroll
    for (int j = 0; j < H; ++j) ffma2_s(o0, o1, w3[j], h2[j][0], h2[j][1]);
    out[0] = o0; out[1] = o1;
}

__device__ __forceinline__ void bwd_d2x2(const float* __restrict__ img, float dout0, float dout1, const float (&h1)[H][2],
                                         const float (&h2)[H][2], float (&d1)[H][2], float (&d2)[H][2], float (&dc)[2]) {
    using L = L2_;
    float w3[8];
    ld8(img + L::W3, w3);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float g0, g1, t0, t1;
        fmul2_p(t0, t1, -TANH_ISCALE, -TANH_ISCALE, h2[j][0], h2[j][1]);
        fma2_p(g0, g1, t0, t1, h2[j][0], h2[j][1], TANH_ISCALE, TANH_ISCALE);          // (1 - h2^2) / scale
        fmul2_p(t0, t1, w3[j], w3[j], dout0, dout1);
        fmul2_p(d2[j][0], d2[j][1], t0, t1, g0, g1);
    }
    float da[H][2];
#pragma unroll
    for (int k = 0; k < H; ++k) { da[k][0] = 0.f; da[k][1] = 0.f; }
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
#pragma unroll
        for (int k = 0; k < H; ++k) ffma2_s(da[k][0], da[k][1], w[k], d2[j][0], d2[j][1]);
    }
    float p0 = 0.f, p1 = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float g0, g1, t0, t1;
        fmul2_p(t0, t1, -TANH_ISCALE, -TANH_ISCALE, h1[k][0], h1[k][1]);
        fma2_p(g0, g1, t0, t1, h1[k][0], h1[k][1], TANH_ISCALE, TANH_ISCALE);
        fmul2_p(d1[k][0], d1[k][1], da[k][0], da[k][1], g0, g1);
        ffma2_s(p0, p1, img[L::W1 + k * L::S1], d1[k][0], d1[k][1]);
    }
    dc[0] = p0; dc[1] = p1;
}

// gradient products of particle P of the pair
template <int P>
__device__ __forceinline__ void accumulate_x2(float (&acc)[NACC], const float (&d1)[H][2], const float (&d2)[H][2], float dout, float c,
                                              const float (&h1)[H][2], const float (&h2)[H][2]) {
#pragma unroll
    for (int k = 0; k < H; k += 2) {
        ffma2_s(acc[k], acc[k + 1], c, d1[k][P], d1[k + 1][P]);
        ffma2_s(acc[H + k], acc[H + k + 1], 1.0f, d1[k][P], d1[k + 1][P]);
    }
    unsigned long long h1p[H / 2];
#pragma unroll
    for (int k = 0; k < H; k += 2) h1p[k / 2] = pack2(h1[k][P], h1[k + 1][P]);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const unsigned long long dj = pack2(d2[j][P], d2[j][P]);
#pragma unroll
        for (int k = 0; k < H; k += 2) ffma2(acc[2 * H + j * H + k], acc[2 * H + j * H + k + 1], dj, h1p[k / 2]);
    }
#pragma unroll
    for (int j = 0; j < H; j += 2) {
        ffma2_s(acc[2 * H + H * H + j], acc[2 * H + H * H + j + 1], 1.0f, d2[j][P], d2[j + 1][P]);
        ffma2_s(acc[3 * H + H * H + j], acc[3 * H + H * H + j + 1], dout, h2[j][P], h2[j + 1][P]);
    }
    acc[4 * H + H * H] += dout;
}

__global__ void __launch_bounds__(TPD)
coupling_bwd_d2_kernel(const float* __restrict__ packed, int n_flows, int C_row, const float* __restrict__ y,
                       const float* __restrict__ row_ctx, int flags, int B, int N, const float* __restrict__ g_y,
                       const float* __restrict__ g_ld, float* __restrict__ d_x, float* __restrict__ partials,
                       float* __restrict__ d1rows, int e_max) {
    extern __shared__ __align__(16) float smem[];
    const int n_fcnn = 4 * n_flows, tid = threadIdx.x, inverse = flags & 1;
    const int grp = tid >> 7, gi = tid & (GRP - 1), warp = tid >> 5;
    float* s_img = smem;
    float* s_w1r = s_img + n_fcnn * L2_::SIZE;
    float* s_xt = s_w1r + (size_t)n_fcnn * H * C_row;        // [2][PPI] t-net outputs
    float* s_xs = s_xt + 2 * PPI;                            // [2][PPI] s-net outputs
    float* s_dct = s_xs + 2 * PPI;                           // [2][PPI] t-net gradient wrt the conditioning half
    float* s_part = s_dct + 2 * PPI;                         // [NWARP][100]
    float* s_acc = s_part + NWARP * 100;                     // [n_fcnn][NACC]
    float* s_lo = s_acc + n_fcnn * NACC + 8;                 // state, [e_max][CHUNK] each
    float* s_up = s_lo + (size_t)e_max * CHUNK;
    float* s_glo = s_up + (size_t)e_max * CHUNK;
    float* s_gup = s_glo + (size_t)e_max * CHUNK;
    float* s_gld = s_gup + (size_t)e_max * CHUNK;
    float* s_hb = s_gld + (size_t)e_max * CHUNK;             // [e_max][n_fcnn][H] hoisted layer-1 biases
    float* s_d1row = s_hb + (size_t)e_max * n_fcnn * H;      // [e_max][n_fcnn][H] per-entry layer-1 delta sums
    float* s_d1part = s_d1row + (size_t)e_max * n_fcnn * H;  // [e_max][NWARP][H]
    float* s_ctx = s_d1part + (size_t)e_max * NWARP * H;     // [e_max][C_row]
    const int pf = packed_fcnn_size(1, C_row);
    for (int f = 0; f < n_fcnn; ++f)
        load_fcnn_image<1, 0>(packed + (size_t)f * pf, C_row, s_img + f * L2_::SIZE, s_w1r + (size_t)f * H * C_row, tid, TPD);
    for (int e = tid; e < n_fcnn * NACC; e += TPD) s_acc[e] = 0.f;
    __syncthreads();

    const int nc = (N + CHUNK - 1) / CHUNK;                                    // entries per trajectory
    const int n_traj = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_entries = n_traj * nc;
    for (int e0 = 0; e0 < n_entries; e0 += e_max) {
        const int ne = min(e_max, n_entries - e0);
        // ---- load the resident set
        for (int e = 0; e < ne; ++e) {
            const int j = e0 + e, b = blockIdx.x + (j / nc) * gridDim.x, c0 = (j % nc) * CHUNK;
            const int n_live = min(CHUNK, N - c0);
            const size_t p0 = (size_t)b * N + c0;
#pragma unroll                    // CHUNK / TPD = 4 independent iterations: all twelve loads in flight before the first store
            for (int q = tid; q < CHUNK; q += TPD) {
                const bool live = q < n_live;
                const float2 yy = live ? reinterpret_cast<const float2*>(y)[p0 + q] : make_float2(0.f, 0.f);
                const float2 gg = live && g_y ? reinterpret_cast<const float2*>(g_y)[p0 + q] : make_float2(0.f, 0.f);
                const int o = e * CHUNK + q;
                s_lo[o] = yy.x; s_up[o] = yy.y; s_glo[o] = gg.x; s_gup[o] = gg.y;
                s_gld[o] = live && g_ld ? ((flags & 2) ? -g_ld[p0 + q] : g_ld[p0 + q]) : 0.f;
            }
            for (int c = tid; c < C_row; c += TPD) s_ctx[e * C_row + c] = row_ctx[(size_t)b * C_row + c];
        }
        __syncthreads();
        for (int e = 0; e < ne; ++e) hoist_row_context_par<1, 0>(s_img, s_w1r, s_ctx + e * C_row, C_row, n_fcnn, s_hb + (size_t)e * n_fcnn * H);
        __syncthreads();
        // forward pass ran flows 0..n-1 (t1/s1 then t2/s2): walk back n-1..0 (t2/s2 then t1/s1);
        // inverse pass ran flows n-1..0 (t2/s2 then t1/s1): walk back 0..n-1 (t1/s1 then t2/s2)
#pragma unroll 1
        for (int st = 0; st < 2 * n_flows; ++st) {
            const int f = inverse ? st / 2 : n_flows - 1 - st / 2;
            const int pair = inverse ? (st & 1) : 1 - (st & 1);       // 0: t1/s1 (c = lower), 1: t2/s2 (c = upper)
            const int fm = 4 * f + 2 * pair + grp;                    // the net this half of the CTA owns
            const float* img = s_img + fm * L2_::SIZE;
            const float* s_c = pair ? s_up : s_lo;                    // conditioning half and its gradient
            float* s_gc = pair ? s_gup : s_glo;
            float* s_v = pair ? s_lo : s_up;                          // transformed half (output value -> input value) and its gradient
            float* s_gv = pair ? s_glo : s_gup;
            float acc[NACC];
#pragma unroll
            for (int k = 0; k < NACC; ++k) acc[k] = 0.f;
            int it = 0, prev_q = -1;
#pragma unroll 1
            for (int e = 0; e < ne; ++e) {
                const int j = e0 + e, c0 = (j % nc) * CHUNK;
                const int n_live = min(CHUNK, N - c0), iters = (n_live + PPI - 1) / PPI;
                const float* hb = s_hb + ((size_t)e * n_fcnn + fm) * H;
#pragma unroll 1
                for (int m = 0; m < iters; ++m, ++it) {
                    asm volatile("" ::: "memory");
                    const int q = e * CHUNK + m * PPI + gi, par = (it & 1) * PPI;     // this thread's particles: q and q + GRP
                    const bool live0 = m * PPI + gi < n_live, live1 = m * PPI + GRP + gi < n_live;
                    const float c0v = s_c[q], c1v = s_c[q + GRP];
                    float h1[H][2], h2[H][2], out[2];
                    fwd_d2x2(img, hb, c0v, c1v, h1, h2, out);
                    float gv0 = 0.f, gv1 = 0.f;
                    if (grp == 0) { s_xt[par + gi] = out[0]; s_xt[par + GRP + gi] = out[1]; gv0 = s_gv[q]; gv1 = s_gv[q + GRP]; }
                    else          { s_xs[par + gi] = out[0]; s_xs[par + GRP + gi] = out[1]; }
                    // t-warp w and s-warp w + 4 work on the same 64 particles and exchange only with each other: a 64-thread
                    // named barrier per warp pair instead of a CTA barrier (the four pairs drift independently within a stage)
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
                    float d1[H][2], d2[H][2], dc[2];
                    float do0, do1;
                    if (grp == 0) {         // t-net: d t = g_v (forward direction) or -g_v e^{-s} (inverse direction)
                        do0 = inverse ? -gv0 * exp_acc(-s_xs[par + gi]) : gv0;
                        do1 = inverse ? -gv1 * exp_acc(-s_xs[par + GRP + gi]) : gv1;
                    } else {                // s-net: inverts the stage, owns the state update
                        if (prev_q >= 0) { s_gc[prev_q] += s_dct[(PPI - par) + gi]; s_gc[prev_q + GRP] += s_dct[(PPI - par) + GRP + gi]; }
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            const int qq = q + p * GRP;
                            const float t = s_xt[par + p * GRP + gi], sv = out[p];
                            const float v = s_v[qq], gv = s_gv[qq], gld = s_gld[qq];
                            const float es = exp_acc(sv), ies = exp_acc(-sv);
                            float ds, vin, gin;
                            if (!inverse) { vin = (v - t) * ies; ds = fmaf(gv * vin, es, gld); gin = gv * es; }
                            else          { gin = gv * ies; ds = -fmaf(gv, v, gld); vin = fmaf(v, es, t); }
                            s_v[qq] = vin; s_gv[qq] = gin;
                            if (p == 0) do0 = ds; else do1 = ds;
                        }
                    }
                    if (!live0) do0 = 0.f;
                    if (!live1) do1 = 0.f;
                    bwd_d2x2(img, do0, do1, h1, h2, d1, d2, dc);
                    if (grp == 0) { s_dct[par + gi] = dc[0]; s_dct[par + GRP + gi] = dc[1]; }
                    else { s_gc[q] += dc[0]; s_gc[q + GRP] += dc[1]; prev_q = q; }
                    accumulate_x2<0>(acc, d1, d2, do0, c0v, h1, h2);
                    accumulate_x2<1>(acc, d1, d2, do1, c1v, h1, h2);
                }
                warp_reduce_b1(acc, s_d1part + (size_t)e * NWARP * H);
            }
            warp_reduce_to_slot(acc, s_part);
            __syncthreads();
            if (grp == 1 && prev_q >= 0) { s_gc[prev_q] += s_dct[((it - 1) & 1) * PPI + gi]; s_gc[prev_q + GRP] += s_dct[((it - 1) & 1) * PPI + GRP + gi]; }
            if (gi < NACC) {               // owner thread per parameter of the group's net: fixed-order sums
                const int k = gi;
                float v = 0.f;
                if (k >= H && k < 2 * H) {     // b1 block: per-entry sums (row-context hoist) and their total
                    for (int e = 0; e < ne; ++e) {
                        const float* dp = s_d1part + ((size_t)e * NWARP + 4 * grp) * H + (k - H);
                        const float r = (dp[0] + dp[H]) + (dp[2 * H] + dp[3 * H]);
                        s_d1row[((size_t)e * n_fcnn + fm) * H + (k - H)] = r;
                        v += r;
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < 4; ++w) v += s_part[(4 * grp + w) * 100 + k];
                }
                s_acc[fm * NACC + k] += v;
            }
            __syncthreads();
        }
        // ---- store d_x and the per-trajectory layer-1 delta sums (the row-context columns of dW1 and d(row_ctx) are formed
        // from them by rowctx_grad_kernel: keeps the 8 x C_row outer products out of the persistent loop)
        for (int e = 0; e < ne; ++e) {
            const int j = e0 + e, b = blockIdx.x + (j / nc) * gridDim.x, c0 = (j % nc) * CHUNK;
            const int n_live = min(CHUNK, N - c0);
            const size_t p0 = (size_t)b * N + c0;
            for (int q = tid; q < n_live; q += TPD) reinterpret_cast<float2*>(d_x)[p0 + q] = make_float2(s_glo[e * CHUNK + q], s_gup[e * CHUNK + q]);
        }
        if (tid < n_fcnn * H) {            // one thread per (net, k): entries in order (chunks of one trajectory add up)
            for (int e = 0; e < ne; ++e) {
                const int j = e0 + e, b = blockIdx.x + (j / nc) * gridDim.x, c0 = (j % nc) * CHUNK;
                const float val = TANH_SCALE * s_d1row[(size_t)e * n_fcnn * H + tid];   // true delta sums, [f*8+k][b]: coalesced for rowctx_grad
                float* dst = d1rows + (size_t)tid * B + b;
                *dst = c0 == 0 ? val : *dst + val;
            }
        }
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
    const int grid = min(B, sm_count());   // one 8-warp CTA per SM (registers), persistent over trajectories
    // resident entries: as many as the CTA's share of the work needs, bounded by shared memory
    const int nc = (N + CHUNK - 1) / CHUNK, need = ((B + grid - 1) / grid) * nc;
    int e_max = min(E_CAP, need);
    while (e_max > 1 && D2Smem::bytes(n_fcnn, C_row, e_max) > 220 * 1024) --e_max;
    const size_t smem = D2Smem::bytes(n_fcnn, C_row, e_max);
    if (smem > 220 * 1024) { set_error("coupling_bwd_d2: stack too large for shared memory (n_flows=%d, C_row=%d)", n_flows, C_row); return NFDPF_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024) NFDPF_CUDA(cudaFuncSetAttribute(coupling_bwd_d2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_params = n_fcnn * packed_fcnn_size(1, C_row);
    float* d1rows = (float*)workspace + (size_t)bwd_grid(B) * n_params;
    coupling_bwd_d2_kernel<<<grid, TPD, smem, st>>>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, (float*)workspace,
                                                    d1rows, e_max);
    int rc = check_launch("coupling_bwd_d2");
    if (rc) return rc;
    rc = launch_reduce_partials((const float*)workspace, grid, n_params, d_packed, st);
    if (rc || C_row == 0) return rc;
    const int warps = n_fcnn * H * C_row + (d_row_ctx ? B * C_row : 0);
    rowctx_grad_kernel<<<(warps + 7) / 8, 256, 0, st>>>(packed, d1rows, row_ctx, n_fcnn, C_row, B, d_packed, d_row_ctx);
    return check_launch("rowctx_grad");
}

}  // namespace nfdpf
