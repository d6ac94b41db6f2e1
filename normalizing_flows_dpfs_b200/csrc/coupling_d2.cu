// (K1, headline shape) Backward of the D = 2 coupling stack with row-constant context (nf_dyn: C = 4, cond_model: C = 36).
// Reference math: nf/flows.py:215-239 (RealNVP_cond forward / inverse), nf/flows.py:101-114 (FCNN), walked backwards from the
// stack OUTPUT (couplings are invertible: nothing but y is read from HBM, activations are recomputed).
//
// Round-2 design (the round-1 kernel: 8 warps per SM, 221 registers, 812 issue slots per particle pair and net, CTA halves
// exchanging t / s through named barriers):
//   * FFMA2 packs over adjacent UNITS of one particle, not over a particle pair.  Every activation / delta vector lives as four
//     (unit 2i, unit 2i+1) register pairs; the weights come from shared memory as ready-made pairs (both W2 and its transpose are
//     kept: the forward needs columns, the data-gradient rows); scalars enter as the FFMA2 broadcast operand (SASS `R.F32`).
//     The 97 weight-gradient products then are  acc[j][k:k+2] += d2[j] * h1[k:k+2]  -- natural pairs x broadcast scalar -- and
//     the ~110 register moves per iteration that re-paired values in the particle-pair form are gone (812 -> ~500 slots).
//   * tanh: (2^a0, 2^a1) -> FADD2 -> two MUFU.RCP -> FFMA2: three issue slots per activation (the shared-reciprocal pair of
//     the forward kernel costs six; this kernel is short of issue slots and FMA-pipe cycles, not of MUFU throughput).
//   * net-sequential passes instead of t-/s-net CTA halves: per stage a warp walks its particles twice, once per net.  One extra
//     float per particle carries what the second net needs from the first (forward direction: t; inverse direction: e^s).
//   * WARP-PRIVATE pipelines: a warp owns a contiguous range of tasks (32 x PPT particles) of the CTA's resident set for the whole
//     launch -- it loads them, runs all 4 n_flows passes over them and stores their d_x.  Particle state in shared memory is
//     touched by its owner lane only, so there is no barrier between passes and no exchange buffer; warps drift apart, which
//     spreads the MUFU-heavy forward halves and the FMA-heavy backward halves of the tasks over time.  After a pass the warp
//     folds its 97 accumulators with a transposed butterfly and writes ITS OWN partial-gradient row to global memory.
//   * PPT particles per thread share every weight load (PPT = 2, 8 warps: 49 LDS.128 per 64 particles and net; PPT = 1, 12 warps).
//   * the row-context columns of dW1 (and db1, and d(row_ctx)) are formed per CTA from the per-trajectory layer-1 delta sums;
//     d2_reduce_kernel sums the per-warp rows and per-CTA row-context partials in a fixed order (fp64): run-to-run deterministic.
// Measured dead ends of this round are listed in DESIGN.md (weights through constant memory / uniform registers, 16 warps).
#include <stdlib.h>

#include "coupling.cuh"

namespace nfdpf {

using u64 = unsigned long long;
constexpr int NACC = 97;          // == Rows<1,0>::NOUT, same ordering as packed_offset<1,0>: dW1 8 | db1 8 | dW2 64 | db2 8 | dW3 8 | db3 1
constexpr int CHUNK = 1024;       // particle slots per entry (an entry = up to 1024 particles of one trajectory)
constexpr int E_CAP = 9;          // upper bound of resident entries (the launcher fits E_MAX <= E_CAP into shared memory)
constexpr int NSTATE = 6;         // lo, up, g_lo, g_up, g_logdet, carry
constexpr int D2_MAX_WARPS = 12;  // sizes the partial-gradient workspace (nfdpf_coupling_bwd_workspace)

// Shared-memory image of one 1 -> 8 -> 8 -> 1 net (floats; every block 16-byte aligned).  "s" = multiplied by TANH_SCALE.
struct Img {
    static constexpr int W1 = 0;      // [8]      W1s[k]            (conditioning column)
    static constexpr int W2T = 8;     // [8][8]   W2s[j][k] at [k][j]   forward: a2[j:j+2] += W2T[k][j:j+2] * h1[k]
    static constexpr int B2 = 72;     // [8]      b2s
    static constexpr int W3 = 80;     // [8]
    static constexpr int W3I = 88;    // [8]      W3 * TANH_ISCALE      delta2 / scale = dout * (W3I - W3I h2^2)
    static constexpr int W2 = 96;     // [8][8]   W2s[j][k] at [j][k]   backward: da1[k:k+2] += W2[j][k:k+2] * d2[j]
    static constexpr int B3 = 160;    // [1] (+3 pad)
    static constexpr int SIZE = 164;
};

// ---- packed FP32 pairs ------------------------------------------------------------------------------------
__device__ __forceinline__ u64 P2(float lo, float hi) { u64 p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi)); return p; }
__device__ __forceinline__ void U2(u64 p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ u64 bc(float s) { return P2(s, s); }                      // folds into the broadcast-scalar operand form
__device__ __forceinline__ u64 neg2(u64 p) { float a, b; U2(p, a, b); return P2(-a, -b); }   // folds into the operand negation
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
// accumulate forms: the accumulator is tied to the destination register pair
__device__ __forceinline__ void acc_fma2(u64& acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ void acc_add2(u64& acc, u64 a) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(a)); }
template <int I>
__device__ __forceinline__ float half_of(u64 p) { float a, b; U2(p, a, b); return I ? b : a; }
__device__ __forceinline__ void ld4(const float* p, u64 (&w)[4]) {                   // eight floats = four pairs, two LDS.128
    const ulonglong2 a = reinterpret_cast<const ulonglong2*>(p)[0], b = reinterpret_cast<const ulonglong2*>(p)[1];
    w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}
// two tanh of pre-scaled arguments (a = 2 log2(e) x): 1 - 2 / (2^a + 1); ~2e-7 absolute error (common.cuh)
__device__ __forceinline__ u64 tanh2(u64 a) {
    float x, y, ex, ey, sx, sy, rx, ry;
    U2(a, x, y);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(y));
    U2(add2(P2(ex, ey), bc(1.0f)), sx, sy);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(sx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(sy));
    return fma2(bc(-2.0f), P2(rx, ry), bc(1.0f));
}

// Warp-level transposed butterfly of the 97 per-thread accumulators: in round r the lanes with bit (16 >> r) set keep the
// upper half of the live values and send the lower half (and vice versa), so 96 values cost 93 shuffles instead of 480;
// lane L ends up owning the sums of entries 3L..3L+2.  Entry 96 (db3) takes a plain butterfly.  Result -> the warp's slot.
__device__ __forceinline__ void warp_reduce_to_slot(float (&acc)[NACC], float* slot) {
    const int lane = threadIdx.x & 31;
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
    slot[3 * lane] = acc[0]; slot[3 * lane + 1] = acc[1]; slot[3 * lane + 2] = acc[2];
    if (lane == 0) slot[96] = last;
}

// Warp sums of the 8 layer-1 delta accumulators (pairs 4..7) of one entry -> dst[8]; the accumulators are cleared.
__device__ __forceinline__ void warp_flush_b1(u64 (&A)[48], float* dst) {
    const int lane = threadIdx.x & 31;
    float v[H];
#pragma unroll
    for (int k = 0; k < 4; ++k) { U2(A[4 + k], v[2 * k], v[2 * k + 1]); A[4 + k] = 0ull; }
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
    if ((lane & 3) == 0) dst[lane >> 2] = v[0];   // lane bits 4,3,2 select the entry index
}

// The four kinds of pass.  Forward-direction stage out = t + in e^s: the t-net goes first (its delta is g_v, no s needed) and
// leaves t in the carry slot; the s-net pass then inverts the stage.  Inverse-direction stage out = (in - t) e^-s: the s-net
// goes first (its delta -(g_v v + g_ld) needs no t), rescales g_v and leaves e^s; the t-net pass then restores the input.
enum { FWD_T = 0, FWD_S = 1, INV_T = 2, INV_S = 3 };

struct StatePtrs {
    float* c;     // conditioning half (read only in a stage)
    float* gc;    // its gradient (+= dc of both nets)
    float* v;     // transformed half: stage output -> stage input
    float* gv;    // its gradient
    float* gld;   // gradient of the log-det (constant along the walk)
    float* ex;    // carry between the two passes of a stage
};

// One task: PPT particles per lane (slots q0, q0 + 32, ...), one net.  n_rem = live particles from the lane's first slot on.
// Two diagnostics of this task body were run and removed again (they cost the default path 12 us through their run-time branches
// and the sixteen named barriers they reserved): (a) a PHASE LOCK -- two 64-thread named barriers per task that force warp w's
// backward half against warp w + 4's forward half (the two warps of a scheduler): 178 / 165 / 171 / 170 us against
// 171 / 165 / 163 / 170 us without it, no gain; (b) stopping the task after its forward half: 112-125 us of the 165-180 -- the forward
// half is a per-warp chain of MUFU and shared-memory latencies (~1200 cycles per task and net, of which the XU pipe needs 544), which
// is why neither the lock nor a second warp per scheduler (one warp per scheduler: 187-202 us) changes much.  DESIGN.md section 3.
template <int PPT, int KIND>
__device__ __forceinline__ void net_task(const float* __restrict__ img, const float* __restrict__ hb, const StatePtrs& S, int q0,
                                         int n_rem, u64 (&A)[48], float& ab3) {
    float c[PPT];
    bool live[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) { c[u] = S.c[q0 + 32 * u]; live[u] = 32 * u < n_rem; }
    u64 h1[PPT][4], h2[PPT][4];
    {   // layer 1: a1[k] = W1s[k] c + hb[k]
        u64 w1[4], hbp[4];
        ld4(img + Img::W1, w1);
        ld4(hb, hbp);
#pragma unroll
        for (int u = 0; u < PPT; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) h1[u][i] = tanh2(fma2(w1[i], bc(c[u]), hbp[i]));
    }
    {   // layer 2: a2[j] = b2s[j] + sum_k W2s[j][k] h1[k]
        u64 a2[PPT][4], b2[4];
        ld4(img + Img::B2, b2);
#pragma unroll
        for (int u = 0; u < PPT; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) a2[u][i] = b2[i];
#pragma unroll
        for (int k = 0; k < H; ++k) {
            u64 wr[4];
            ld4(img + Img::W2T + 8 * k, wr);
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const float hk = (k & 1) ? half_of<1>(h1[u][k >> 1]) : half_of<0>(h1[u][k >> 1]);
#pragma unroll
                for (int i = 0; i < 4; ++i) a2[u][i] = fma2(wr[i], bc(hk), a2[u][i]);
            }
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) h2[u][i] = tanh2(a2[u][i]);
    }
    float dout[PPT];
    {   // layer 3 and the stage algebra
        u64 w3[4];
        ld4(img + Img::W3, w3);
        const float b3 = img[Img::B3];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            u64 o = fma2(w3[0], h2[u][0], P2(b3, 0.f));
#pragma unroll
            for (int i = 1; i < 4; ++i) o = fma2(w3[i], h2[u][i], o);
            const float out = half_of<0>(o) + half_of<1>(o);
            const int q = q0 + 32 * u;
            float d;
            if (KIND == FWD_T) {            // t-net first: d t = g_v; carry t
                d = S.gv[q];
                if (live[u]) S.ex[q] = out;
            } else if (KIND == FWD_S) {     // s-net second: v_in = (v - t) e^-s, d s = g_v v_in e^s + g_ld, g_in = g_v e^s
                const float t = S.ex[q], v = S.v[q], gv = S.gv[q], gld = S.gld[q];
                const float es = exp_acc(out), ies = exp_acc(-out);
                const float vin = (v - t) * ies;
                d = fmaf(gv * vin, es, gld);
                if (live[u]) { S.v[q] = vin; S.gv[q] = gv * es; }
            } else if (KIND == INV_S) {     // s-net first: g_in = g_v e^-s, d s = -(g_v v + g_ld); carry e^s
                const float v = S.v[q], gv = S.gv[q], gld = S.gld[q];
                const float es = exp_acc(out), ies = exp_acc(-out);
                d = -fmaf(gv, v, gld);
                if (live[u]) { S.gv[q] = gv * ies; S.ex[q] = es; }
            } else {                        // INV_T, t-net second: d t = -g_in, v_in = v e^s + t
                d = -S.gv[q];
                if (live[u]) S.v[q] = fmaf(S.v[q], S.ex[q], out);
            }
            dout[u] = live[u] ? d : 0.f;
        }
    }
    u64 d2[PPT][4];
    {   // delta2 / scale = dout (W3I - W3I h2^2); dW3 += dout h2; db2 += delta2; db3 += dout
        u64 w3i[4];
        ld4(img + Img::W3I, w3i);
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc_fma2(A[44 + i], bc(dout[u]), h2[u][i]);
                const u64 g = fma2(neg2(mul2(h2[u][i], h2[u][i])), w3i[i], w3i[i]);
                d2[u][i] = mul2(g, bc(dout[u]));
                acc_add2(A[40 + i], d2[u][i]);
            }
            ab3 += dout[u];
        }
    }
    u64 d1[PPT][4];
    {   // da1[k] = sum_j W2s[j][k] d2[j]; dW2[j][k] += d2[j] h1[k]
#pragma unroll
        for (int j = 0; j < H; ++j) {
            u64 wr[4];
            ld4(img + Img::W2 + 8 * j, wr);
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const float dj = (j & 1) ? half_of<1>(d2[u][j >> 1]) : half_of<0>(d2[u][j >> 1]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    d1[u][i] = j == 0 ? mul2(wr[i], bc(dj)) : fma2(wr[i], bc(dj), d1[u][i]);
                    acc_fma2(A[8 + 4 * j + i], bc(dj), h1[u][i]);
                }
            }
        }
    }
    {   // delta1 / scale = da1 (ISCALE - ISCALE h1^2); dW1 += c delta1; db1 += delta1; dc = sum_k W1s[k] delta1[k]
        u64 w1[4];
        ld4(img + Img::W1, w1);
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            u64 s = 0ull;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u64 g = fma2(mul2(h1[u][i], h1[u][i]), bc(-TANH_ISCALE), bc(TANH_ISCALE));
                const u64 dd = mul2(d1[u][i], g);
                acc_fma2(A[i], bc(c[u]), dd);
                acc_add2(A[4 + i], dd);
                s = i == 0 ? mul2(w1[0], dd) : fma2(w1[i], dd, s);
            }
            const int q = q0 + 32 * u;
            if (live[u]) S.gc[q] += half_of<0>(s) + half_of<1>(s);
        }
    }
}


// One pass = one net over the warp's own tasks [t_lo, t_hi).  At an entry boundary the per-entry layer-1 delta sums are flushed
// (they drive the row-context gradient); at the end the 97 accumulators are folded over the lanes into `slot`.
template <int PPT, int NW, int KIND>
__device__ __forceinline__ void net_pass(const float* __restrict__ img, const float* __restrict__ s_hb_net, int hb_stride,
                                         const StatePtrs& S, const int* __restrict__ s_task, const int* __restrict__ s_nlive,
                                         int t_lo, int t_hi, int ne, float* __restrict__ s_d1part_net, float* __restrict__ slot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int TASK = 32 * PPT;
    u64 A[48];
    float ab3 = 0.f;
#pragma unroll
    for (int i = 0; i < 48; ++i) A[i] = 0ull;
    for (int e = 0; e < ne; ++e)
        if (lane < H) s_d1part_net[(e * NW + warp) * H + lane] = 0.f;
    __syncwarp();
    int t = t_lo;
#pragma unroll 1
    while (t < t_hi) {                                       // one segment = the warp's tasks inside one entry
        const int code = s_task[t], e = code >> 8, m0 = code & 255, n_live = s_nlive[e];
        const int seg_end = min(t_hi, t + (n_live + TASK - 1) / TASK - m0);
        const float* hb = s_hb_net + e * hb_stride;
        int q0 = e * CHUNK + m0 * TASK + lane, n_rem = n_live - m0 * TASK - lane;
#pragma unroll 1
        for (; t < seg_end; ++t, q0 += TASK, n_rem -= TASK) {
            asm volatile("" ::: "memory");                   // keep the weight loads inside the loop (registers)
            net_task<PPT, KIND>(img, hb, S, q0, n_rem, A, ab3);
        }
        warp_flush_b1(A, s_d1part_net + (e * NW + warp) * H);
    }
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < 48; ++i) U2(A[i], acc[2 * i], acc[2 * i + 1]);
    acc[96] = ab3;
    warp_reduce_to_slot(acc, slot);
}

template <int NT>
struct D2Smem {
    static constexpr int NW = NT / 32;
    static size_t fixed_floats(int n_fcnn, int C_row) {
        return (size_t)n_fcnn * Img::SIZE + (size_t)n_fcnn * H                // images, b1s
               + (size_t)n_fcnn * H * C_row + (size_t)n_fcnn * H * (C_row + 1) + 3   // w1r, accR (+ bias column, pad)
               + E_CAP * (CHUNK / 32) + 2 * 16 + 4 * 16;                     // task table, n_live / first-chunk flags, entry particle offsets (64-bit)
    }
    static size_t entry_floats(int n_fcnn, int C_row) {
        return NSTATE * (size_t)CHUNK + 2 * (size_t)n_fcnn * H + (size_t)n_fcnn * NW * H + ((C_row + 3) & ~3);   // state, hb, d1row, d1part, ctx
    }
    static size_t bytes(int n_fcnn, int C_row, int e_max) { return (fixed_floats(n_fcnn, C_row) + e_max * entry_floats(n_fcnn, C_row)) * sizeof(float); }
};

// element o of the image of net f (see Img)
__device__ __forceinline__ float img_value(const float* __restrict__ packed, int f, int o, int C_row) {
    const int fin = 1 + C_row;
    const float* pk = packed + (size_t)f * packed_fcnn_size(1, C_row);
    const float* tail = pk + H * fin;                      // b1 [8] | W2 [8][8] | b2 [8] | W3 [8] | b3
    if (o < Img::W2T) return TANH_SCALE * pk[o * fin];
    if (o < Img::B2) { const int k = (o - Img::W2T) >> 3, j = (o - Img::W2T) & 7; return TANH_SCALE * tail[H + j * H + k]; }
    if (o < Img::W3) return TANH_SCALE * tail[H + H * H + (o - Img::B2)];
    if (o < Img::W3I) return tail[2 * H + H * H + (o - Img::W3)];
    if (o < Img::W2) return TANH_ISCALE * tail[2 * H + H * H + (o - Img::W3I)];
    if (o < Img::B3) return TANH_SCALE * tail[H + (o - Img::W2)];
    return o == Img::B3 ? tail[3 * H + H * H] : 0.f;
}

// Workspace layout (floats): [grid * NW rows][n_fcnn * NACC] per-warp partial gradients (accumulator order, divided by the tanh
// scale where grad_out_scale says so), then [grid][n_fcnn * 8 * (C_row + 1)] per-CTA row-context partials (last column: db1),
// then [grid][n_fcnn * NACC]: the per-warp rows of each CTA folded in warp order (what the reduce kernel reads).
template <int PPT, int NT>
__global__ void __launch_bounds__(NT, 1)
coupling_bwd_d2_kernel(const float* __restrict__ packed, int n_flows, int C_row, const float* __restrict__ y,
                       const float* __restrict__ row_ctx, int flags, int B, int N, const float* __restrict__ g_y,
                       const float* __restrict__ g_ld, float* __restrict__ d_x, float* warp_rows,
                       float* __restrict__ ctx_rows, float* __restrict__ cta_rows, float* __restrict__ d_row_ctx, int e_max) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_slot[NT / 32][100];
    constexpr int NW = NT / 32, TASK = 32 * PPT;
    const int n_fcnn = 4 * n_flows, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, inverse = flags & 1;
    const int ctx_pad = (C_row + 3) & ~3, C1 = C_row + 1, nR = n_fcnn * H * C_row, nR1 = n_fcnn * H * C1;
    float* s_img = smem;                                       // [n_fcnn][Img::SIZE]
    float* s_b1 = s_img + n_fcnn * Img::SIZE;                  // [n_fcnn][8]   b1s
    float* s_w1r = s_b1 + n_fcnn * H;                          // [n_fcnn][8][C_row]      W1s row-context columns
    float* s_accR = s_w1r + nR;                                // [n_fcnn][8][C_row + 1]  their gradient (/ scale); last column: db1
    int* s_task = reinterpret_cast<int*>(s_accR + ((nR1 + 3) & ~3));   // [E_CAP * 32]  (entry << 8) | task-in-entry
    int* s_nlive = s_task + E_CAP * (CHUNK / 32);              // [16]
    int* s_first = s_nlive + 16;                               // [16] 1 = first chunk of its trajectory
    long long* s_p0 = reinterpret_cast<long long*>(s_first + 16);   // [16] first particle (global index), [16] trajectory
    float* s_state = reinterpret_cast<float*>(s_p0 + 32);      // [NSTATE][e_max][CHUNK]
    float* s_hb = s_state + (size_t)NSTATE * e_max * CHUNK;    // [e_max][n_fcnn][8]
    float* s_d1row = s_hb + (size_t)e_max * n_fcnn * H;        // [e_max][n_fcnn][8] per-entry layer-1 delta sums (/ scale)
    float* s_d1part = s_d1row + (size_t)e_max * n_fcnn * H;    // [n_fcnn][e_max][NW][8]
    float* s_ctx = s_d1part + (size_t)n_fcnn * e_max * NW * H; // [e_max][ctx_pad]
    const size_t plane = (size_t)e_max * CHUNK;
    float* s_lo = s_state, *s_up = s_lo + plane, *s_glo = s_up + plane, *s_gup = s_glo + plane, *s_gld = s_gup + plane, *s_ex = s_gld + plane;

    const int fin = 1 + C_row, pf = packed_fcnn_size(1, C_row);
    for (int d = tid; d < n_fcnn * Img::SIZE; d += NT) s_img[d] = img_value(packed, d / Img::SIZE, d % Img::SIZE, C_row);
    for (int e = tid; e < n_fcnn * H; e += NT) s_b1[e] = TANH_SCALE * packed[(size_t)(e / H) * pf + H * fin + (e % H)];
    for (int e = tid; e < nR; e += NT) {
        const int f = e / (H * C_row), r = e - f * H * C_row;
        s_w1r[e] = TANH_SCALE * packed[(size_t)f * pf + (r / C_row) * fin + 1 + (r % C_row)];
    }
    for (int e = tid; e < nR1; e += NT) s_accR[e] = 0.f;
    __syncthreads();

    float* my_rows = warp_rows + ((size_t)blockIdx.x * NW + warp) * n_fcnn * NACC;
    const int nc = (N + CHUNK - 1) / CHUNK;                                    // entries per trajectory
    const int n_traj = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_entries = n_traj * nc;
    for (int e0 = 0; e0 < n_entries; e0 += e_max) {
        const int ne = min(e_max, n_entries - e0);
        if (tid < ne) {
            const int j = e0 + tid, b = blockIdx.x + (j / nc) * gridDim.x, c0 = (j % nc) * CHUNK;
            s_nlive[tid] = min(CHUNK, N - c0);
            s_first[tid] = c0 == 0;
            s_p0[tid] = (long long)b * N + c0;
            s_p0[16 + tid] = b;
        }
        __syncthreads();
        for (int i = tid; i < ne * C_row; i += NT) {
            const int e = i / C_row, c = i - e * C_row;
            s_ctx[e * ctx_pad + c] = row_ctx[(size_t)s_p0[16 + e] * C_row + c];
        }
        // task table: entry e contributes ceil(n_live / TASK) tasks, entries in order
        int n_tasks = 0;
        for (int e = 0; e < ne; ++e) {
            const int te = (s_nlive[e] + TASK - 1) / TASK;
            for (int m = tid; m < te; m += NT) s_task[n_tasks + m] = (e << 8) | m;
            n_tasks += te;
        }
        __syncthreads();
        // hoisted layer-1 biases: hb[e][f][k] = b1s + sum_c W1s_r[f][k][c] ctx[e][c]   (four lanes per output)
        {
            const int sub = tid & 3, total = ne * n_fcnn * H;
            for (int o0 = 0; o0 < total; o0 += NT / 4) {
                const int o = o0 + (tid >> 2);
                float a = 0.f;
                if (o < total) {
                    const int e = o / (n_fcnn * H), fk = o - e * n_fcnn * H;
                    const float* w = s_w1r + (size_t)fk * C_row;
                    const float* cx = s_ctx + e * ctx_pad;
                    for (int c = sub; c < C_row; c += 4) a = fmaf(w[c], cx[c], a);
                }
                a += __shfl_xor_sync(FULL, a, 1);
                a += __shfl_xor_sync(FULL, a, 2);
                if (o < total && sub == 0) s_hb[o] = a + s_b1[o % (n_fcnn * H)];
            }
        }
        // ---- the warp's own tasks: load (each lane loads exactly the particle slots it will work on)
        const int t_lo = (warp * n_tasks) / NW, t_hi = ((warp + 1) * n_tasks) / NW;
#pragma unroll 2
        for (int t = t_lo; t < t_hi; ++t) {
            const int code = s_task[t], e = code >> 8, m = code & 255, n_live = s_nlive[e];
            const size_t p0 = (size_t)s_p0[e];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const int q = m * TASK + 32 * u + lane, i = e * CHUNK + q;
                const bool live = q < n_live;
                const float2 yy = live ? reinterpret_cast<const float2*>(y)[p0 + q] : make_float2(0.f, 0.f);
                const float2 gg = live && g_y ? reinterpret_cast<const float2*>(g_y)[p0 + q] : make_float2(0.f, 0.f);
                const float gl = live && g_ld ? g_ld[p0 + q] : 0.f;
                s_lo[i] = yy.x; s_up[i] = yy.y; s_glo[i] = gg.x; s_gup[i] = gg.y;
                s_gld[i] = (flags & 2) ? -gl : gl;
                s_ex[i] = 0.f;
            }
        }
        __syncthreads();          // hb complete (the particle state itself is lane-private)
        // forward pass ran flows 0..n-1 (t1/s1 then t2/s2): walk back n-1..0 (t2/s2 then t1/s1);
        // inverse pass ran flows n-1..0 (t2/s2 then t1/s1): walk back 0..n-1 (t1/s1 then t2/s2)
#pragma unroll 1
        for (int ps = 0; ps < 4 * n_flows; ++ps) {
            const int st = ps >> 1, second = ps & 1;
            const int f = inverse ? st / 2 : n_flows - 1 - st / 2;
            const int pair = inverse ? (st & 1) : 1 - (st & 1);       // 0: t1/s1 (c = lower), 1: t2/s2 (c = upper)
            const int net = inverse ? 1 - second : second;            // 0 = t-net, 1 = s-net
            const int fm = 4 * f + 2 * pair + net;
            const float* img = s_img + fm * Img::SIZE;
            StatePtrs S;
            S.c = pair ? s_up : s_lo;  S.gc = pair ? s_gup : s_glo;
            S.v = pair ? s_lo : s_up;  S.gv = pair ? s_glo : s_gup;
            S.gld = s_gld;             S.ex = s_ex;
            const float* hbn = s_hb + fm * H;
            const int hbs = n_fcnn * H;
            float* d1p = s_d1part + (size_t)fm * e_max * NW * H;
            float* slot = s_slot[warp];       // folded accumulators, then added to the warp's global row (resident sets accumulate)
            if (!inverse) {
                if (!net) net_pass<PPT, NW, FWD_T>(img, hbn, hbs, S, s_task, s_nlive, t_lo, t_hi, ne, d1p, slot);
                else      net_pass<PPT, NW, FWD_S>(img, hbn, hbs, S, s_task, s_nlive, t_lo, t_hi, ne, d1p, slot);
            } else {
                if (net)  net_pass<PPT, NW, INV_S>(img, hbn, hbs, S, s_task, s_nlive, t_lo, t_hi, ne, d1p, slot);
                else      net_pass<PPT, NW, INV_T>(img, hbn, hbs, S, s_task, s_nlive, t_lo, t_hi, ne, d1p, slot);
            }
            __syncwarp();
            float* row = my_rows + fm * NACC;
            for (int k = lane; k < NACC; k += 32) row[k] = e0 == 0 ? slot[k] : row[k] + slot[k];
            __syncwarp();
        }
        // ---- store the warp's d_x
#pragma unroll 2
        for (int t = t_lo; t < t_hi; ++t) {
            const int code = s_task[t], e = code >> 8, m = code & 255, n_live = s_nlive[e];
            const size_t p0 = (size_t)s_p0[e];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const int q = m * TASK + 32 * u + lane, i = e * CHUNK + q;
                if (q < n_live) reinterpret_cast<float2*>(d_x)[p0 + q] = make_float2(s_glo[i], s_gup[i]);
            }
        }
        __syncthreads();          // every warp's per-entry layer-1 delta sums are in place
        for (int o = tid; o < ne * n_fcnn * H; o += NT) {     // fixed-order sums over the warps
            const int e = o / (n_fcnn * H), fk = o - e * n_fcnn * H, f = fk / H, k = fk - f * H;
            const float* dp = s_d1part + (((size_t)f * e_max + e) * NW) * H + k;
            float r = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) r += dp[w * H];
            s_d1row[o] = r;
        }
        __syncthreads();
        // ---- row-context gradients from the per-entry layer-1 delta sums (scales cancel: W1s_r * (delta / s) = W1_r * delta)
        for (int o = tid; o < nR1; o += NT) {                     // dW1_r[f][k][c] (/ scale) += sum_e D1[e][f][k] ctx[e][c]; c = C_row: db1
            const int fk = o / C1, c = o - fk * C1;
            float a = s_accR[o];
            for (int e = 0; e < ne; ++e) a = fmaf(s_d1row[(size_t)e * n_fcnn * H + fk], c < C_row ? s_ctx[e * ctx_pad + c] : 1.0f, a);
            s_accR[o] = a;
        }
        if (d_row_ctx) {
            for (int c = tid; c < C_row; c += NT)                 // d ctx[b][c] = sum_{f,k} W1_r[f][k][c] delta1sum[b][f][k]
                for (int e = 0; e < ne; ++e) {                    // one thread per column, entries in order: the chunks of one
                    float a = 0.f;                                // trajectory (all in this CTA, consecutive) add up race-free
                    for (int fk = 0; fk < n_fcnn * H; ++fk) a = fmaf(s_w1r[(size_t)fk * C_row + c], s_d1row[(size_t)e * n_fcnn * H + fk], a);
                    float* dst = d_row_ctx + (size_t)s_p0[16 + e] * C_row + c;
                    *dst = s_first[e] ? a : *dst + a;
                }
        }
        __syncthreads();
    }
    float* out = ctx_rows + (size_t)blockIdx.x * nR1;
    for (int e = tid; e < nR1; e += NT) out[e] = s_accR[e];
    // fold the CTA's per-warp rows (still L2-hot) into one row, warps in order: the reduce kernel then reads 148 rows, not 1184
    const int n_acc = n_fcnn * NACC;
    const float* mine = warp_rows + (size_t)blockIdx.x * NW * n_acc;
    float* folded = cta_rows + (size_t)blockIdx.x * n_acc;
    for (int c = tid; c < n_acc; c += NT) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) v += mine[(size_t)w * n_acc + c];
        folded[c] = v;
    }
}

// d_packed[target] = fixed-order fp64 column sums of the per-CTA rows and the per-CTA row-context partials.  A block owns 32
// consecutive columns (coalesced 128-byte row segments); its 32 warps take the rows round-robin (<= 5 independent loads per
// thread: the kernel is one round trip to L2 long) and the partial sums are added in warp order -- the order depends only on
// the launch geometry: run-to-run deterministic.
constexpr int RED_G = 32;      // row groups per block of the reduce kernel (32 columns x 32 groups = 1024 threads)
__global__ void __launch_bounds__(32 * RED_G)
d2_reduce_kernel(const float* __restrict__ warp_rows, int n_rows, const float* __restrict__ ctx_rows, int n_cta,
                                 int n_fcnn, int C_row, float* __restrict__ d_packed, int n_calls, size_t call_stride) {
    __shared__ double s_sum[RED_G][32];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int fin = 1 + C_row, pf = packed_fcnn_size(1, C_row), C1 = C_row + 1;
    const int n_acc = n_fcnn * NACC, n_ctx = n_fcnn * H * C1;
    const int col = blockIdx.x * 32 + lane;                 // column of [warp rows | ctx rows]
    double a = 0.0;
    const bool is_acc = col < n_acc, is_ctx = !is_acc && col < n_acc + n_ctx;
    const float* src = is_acc ? warp_rows + col : ctx_rows + (col - n_acc);
    const int stride = is_acc ? n_acc : n_ctx, count = is_acc ? n_rows : (is_ctx ? n_cta : 0);
    if (n_calls == 1) {
#pragma unroll 4
        for (int r = grp; r < count; r += RED_G) a += (double)src[(size_t)r * stride];   // independent loads, few per thread
    } else {            // deferred reduction: the same rows of n_calls launches (blocks call_stride floats apart), calls in order
#pragma unroll 4
        for (int i = grp; i < count * n_calls; i += RED_G) {
            const int c = i / count, r = i - c * count;
            a += (double)src[(size_t)c * call_stride + (size_t)r * stride];
        }
    }
    s_sum[grp][lane] = a;
    __syncthreads();
    if (grp != 0 || col >= n_acc + n_ctx) return;
#pragma unroll
    for (int g = 1; g < RED_G; ++g) a += s_sum[g][lane];      // groups in order: deterministic
    int target;
    float scale;
    if (col < n_acc) {                                      // accumulator entry e of net f (b1 block: produced by the ctx partials)
        const int f = col / NACC, e = col - f * NACC;
        if (e >= H && e < 2 * H) return;
        target = f * pf + packed_offset<1, 0>(e, C_row);
        scale = grad_out_scale<1, 0>(e);
    } else {                                                // [f][k][c]: row-context column c of W1 row k, or (c = C_row) b1[k]
        const int o = col - n_acc, fk = o / C1, c = o - fk * C1, f = fk / H, k = fk - f * H;
        target = f * pf + (c < C_row ? k * fin + 1 + c : H * fin + k);
        scale = TANH_SCALE;
    }
    d_packed[target] = (float)(a * (double)scale);
}

// =====================================================================================================================
// Producer / consumer variant (round 2, second half).  The kernel above runs forward recompute (MUFU-heavy: 64 MUFU per 64-particle
// task and net) and backward (FMA-heavy: ~210 packed FMAs) back to back in the same warp, two warps per scheduler: the XU pipe is
// 38 % busy, the FMA pipe 46 %, and the two barely overlap (a warp needs 2860 cycles per task and net for 1250 pipe-cycles of work).
// Here a CTA has SIXTEEN warps in two roles, and the register file is split between them with setmaxnreg:
//   * forward warps (40 registers): per task and net they recompute h1, h2 and the net output, do the stage algebra (state
//     update, dout) and hand (h1, h2, dout) to their team's backward warp through a shared-memory ring of DEPTH slots;
//   * backward warps (160 registers: the 97 weight-gradient accumulators live only here): delta2, delta1, the 97 products per
//     particle, d(conditioning half); per-entry layer-1 delta sums and the per-pass fold of the accumulators as before.
// Team k = (forward warps 2k and 2k + 1, one half of every 64-particle task each, and backward warp 16 + k) owns the task range the
// single-role kernel gave warp k.  Per slot one `full` and one `empty`
// mbarrier; a per-pair progress counter orders the one cross-role dependency there is: the forward warp may start (pass p, task t)
// only after the backward warp has finished (pass p - 1, task t) -- the next stage reads (and overwrites) the gradient half the
// backward pass accumulated into.  A scheduler now holds two MUFU-heavy and two FMA-heavy warps at any time.
// MEASURED (B = N = 1024), first form -- eight forward warps (two particles per thread, 80 registers) + eight backward warps (176):
// 200 us with a ring of depth 2, 231 us with depth 1, against 161-168 us of the single-role kernel.  The form below -- SIXTEEN forward
// warps (one particle per thread, 40 registers; the forward half alone was measured latency-bound: 112-125 us at 8 warps per SM against
// 66 us for the same arithmetic in the forward kernel at 26) + eight backward warps (160: setmaxnreg.inc can only take what the CTA's
// own warps released, 512 x (80 - 40) = 256 x (160 - 80); asking for more deadlocks) -- 210-222 us / 238-252 us.  ncu of the first form
// (tools/prof_kernels.py coupling, NFDPF_D2_CFG=2): 116 M warp instructions instead of 81 M (ring / progress bookkeeping, the
// exchange stores and loads, spin loops), issue slots 54 % (44 %), FMA pipe 29 % of instruction peak (28 %), XU 33 % (38 %); the
// exchange ring costs two resident entries of shared memory, so a CTA's seven trajectories become two resident sets with their own
// prologue, drain and barrier (barrier stalls 6.5 % of the samples).  The pipes do overlap better, but not by more than the extra
// instructions cost.  NOT the default; kept selectable (NFDPF_D2_CFG=2) so the comparison can be repeated.
constexpr int WS_PAIRS = 8, WS_TASK = 64;                 // eight teams: two forward warps (one particle per thread) + one backward warp
constexpr int WS_FW = 2 * WS_PAIRS, WS_NT = 32 * (WS_FW + WS_PAIRS);   // 16 forward + 8 backward warps = 768 threads
constexpr int XSLOT = 4 * 2 * 32 * 4 + 2 * 32;          // floats per exchange slot: h1 | h2 as 16-byte chunks [chunk][u][lane], dout [u][lane]
constexpr int WS_REG_F = 40, WS_REG_B = 160;             // the backward warps can only take what the forward warps release: 512 x (80 - 40) = 256 x (160 - 80)

__device__ __forceinline__ void ws_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WSWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WSDONE_%=;\n\t"
        "bra WSWAIT_%=;\n\t"
        "WSDONE_%=:\n\t"
        "}" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ int ws_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void ws_st_release(int* p, int v) {
    asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// forward half of net_task for ONE particle per thread (slot half u of the team's 64-particle task): recompute, stage algebra,
// hand-over.  The forward warps are many and light (48 registers) because this half is a chain of MUFU latencies: measured alone it
// takes 112-125 us in the single-role kernel (8 warps per SM) against 66 us for the same arithmetic in the forward kernel (26 warps).
template <int KIND>
__device__ __forceinline__ void ws_f_task(const float* __restrict__ img, const float* __restrict__ hb, const StatePtrs& S, int q, bool live,
                                          float* __restrict__ slot, int u) {
    const int lane = threadIdx.x & 31;
    const float c = S.c[q];
    u64 h1[4], h2[4];
    {
        u64 w1[4], hbp[4];
        ld4(img + Img::W1, w1);
        ld4(hb, hbp);
#pragma unroll
        for (int i = 0; i < 4; ++i) h1[i] = tanh2(fma2(w1[i], bc(c), hbp[i]));
    }
    {
        ulonglong2* d = reinterpret_cast<ulonglong2*>(slot) + (0 * 2 + u) * 32 + lane;
        d[0] = make_ulonglong2(h1[0], h1[1]);
        d[2 * 32] = make_ulonglong2(h1[2], h1[3]);
    }
    {
        u64 a2[4];
        ld4(img + Img::B2, a2);
#pragma unroll
        for (int k = 0; k < H; ++k) {
            u64 wr[4];
            ld4(img + Img::W2T + 8 * k, wr);
            const float hk = (k & 1) ? half_of<1>(h1[k >> 1]) : half_of<0>(h1[k >> 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) a2[i] = fma2(wr[i], bc(hk), a2[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) h2[i] = tanh2(a2[i]);
    }
    {
        ulonglong2* d = reinterpret_cast<ulonglong2*>(slot) + (2 * 2 + u) * 32 + lane;
        d[0] = make_ulonglong2(h2[0], h2[1]);
        d[2 * 32] = make_ulonglong2(h2[2], h2[3]);
    }
    u64 w3[4];
    ld4(img + Img::W3, w3);
    u64 o = fma2(w3[0], h2[0], P2(img[Img::B3], 0.f));
#pragma unroll
    for (int i = 1; i < 4; ++i) o = fma2(w3[i], h2[i], o);
    const float out = half_of<0>(o) + half_of<1>(o);
    float d;
    if (KIND == FWD_T) {
        d = S.gv[q];
        if (live) S.ex[q] = out;
    } else if (KIND == FWD_S) {
        const float t = S.ex[q], v = S.v[q], gv = S.gv[q], gld = S.gld[q];
        const float es = exp_acc(out), ies = exp_acc(-out);
        const float vin = (v - t) * ies;
        d = fmaf(gv * vin, es, gld);
        if (live) { S.v[q] = vin; S.gv[q] = gv * es; }
    } else if (KIND == INV_S) {
        const float v = S.v[q], gv = S.gv[q], gld = S.gld[q];
        const float es = exp_acc(out), ies = exp_acc(-out);
        d = -fmaf(gv, v, gld);
        if (live) { S.gv[q] = gv * ies; S.ex[q] = es; }
    } else {
        d = -S.gv[q];
        if (live) S.v[q] = fmaf(S.v[q], S.ex[q], out);
    }
    slot[4 * 2 * 32 * 4 + 32 * u + lane] = live ? d : 0.f;
}

// backward half of net_task (the same for all four kinds of pass): everything downstream of dout
__device__ __forceinline__ void ws_b_task(const float* __restrict__ img, const StatePtrs& S, int q0, int n_rem,
                                          const float* __restrict__ slot, u64 (&A)[48], float& ab3) {
    const int lane = threadIdx.x & 31;
    float c[2], dout[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) { c[u] = S.c[q0 + 32 * u]; live[u] = 32 * u < n_rem; dout[u] = slot[4 * 2 * 32 * 4 + 32 * u + lane]; }
    u64 d2[2][4];
    {
        u64 w3i[4];
        ld4(img + Img::W3I, w3i);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(slot) + (2 * 2 + u) * 32 + lane;
            const ulonglong2 x0 = sp[0], x1 = sp[2 * 32];
            const u64 h2[4] = {x0.x, x0.y, x1.x, x1.y};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc_fma2(A[44 + i], bc(dout[u]), h2[i]);
                const u64 g = fma2(neg2(mul2(h2[i], h2[i])), w3i[i], w3i[i]);
                d2[u][i] = mul2(g, bc(dout[u]));
                acc_add2(A[40 + i], d2[u][i]);
            }
            ab3 += dout[u];
        }
    }
    u64 h1[2][4], d1[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(slot) + (0 * 2 + u) * 32 + lane;
        const ulonglong2 x0 = sp[0], x1 = sp[2 * 32];
        h1[u][0] = x0.x; h1[u][1] = x0.y; h1[u][2] = x1.x; h1[u][3] = x1.y;
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
        u64 wr[4];
        ld4(img + Img::W2 + 8 * j, wr);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float dj = (j & 1) ? half_of<1>(d2[u][j >> 1]) : half_of<0>(d2[u][j >> 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                d1[u][i] = j == 0 ? mul2(wr[i], bc(dj)) : fma2(wr[i], bc(dj), d1[u][i]);
                acc_fma2(A[8 + 4 * j + i], bc(dj), h1[u][i]);
            }
        }
    }
    {
        u64 w1[4];
        ld4(img + Img::W1, w1);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            u64 s = 0ull;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const u64 g = fma2(mul2(h1[u][i], h1[u][i]), bc(-TANH_ISCALE), bc(TANH_ISCALE));
                const u64 dd = mul2(d1[u][i], g);
                acc_fma2(A[i], bc(c[u]), dd);
                acc_add2(A[4 + i], dd);
                s = i == 0 ? mul2(w1[0], dd) : fma2(w1[i], dd, s);
            }
            const int q = q0 + 32 * u;
            if (live[u]) S.gc[q] += half_of<0>(s) + half_of<1>(s);
        }
    }
}

struct WsCtx {                      // what both roles need to walk the task list of one pass
    const int* s_task;
    const int* s_nlive;
    int t_lo, t_hi;
    uint64_t* full;                 // [DEPTH] of this pair
    uint64_t* empty;                // [DEPTH]
    int* done;                      // the pair's progress counter (tasks the backward warp finished in this resident set)
    float* xch;                     // the pair's exchange slots
};

template <int DEPTH, int KIND>
__device__ __forceinline__ void ws_f_pass(const float* __restrict__ img, const float* __restrict__ s_hb_net, int hb_stride, const StatePtrs& S,
                                          const WsCtx& W, int need0, unsigned& fill) {
    const int lane = threadIdx.x & 31, u = (threadIdx.x >> 5) & 1;     // the team's two forward warps take the two halves of a task
    int t = W.t_lo;
#pragma unroll 1
    while (t < W.t_hi) {
        const int code = W.s_task[t], e = code >> 8, m0 = code & 255, n_live = W.s_nlive[e];
        const int seg_end = min(W.t_hi, t + (n_live + WS_TASK - 1) / WS_TASK - m0);
        const float* hb = s_hb_net + e * hb_stride;
        int q = e * CHUNK + m0 * WS_TASK + 32 * u + lane, n_rem = n_live - m0 * WS_TASK - 32 * u - lane;
#pragma unroll 1
        for (; t < seg_end; ++t, q += WS_TASK, n_rem -= WS_TASK) {
            asm volatile("" ::: "memory");
            if (need0 >= 0) {                                  // the backward warp has finished this task in the previous pass
                const int need = need0 + (t - W.t_lo) + 1;
                while (ws_ld_acquire(W.done) < need) { }
            }
            const unsigned sl = fill % DEPTH;
            ws_mbar_wait(W.empty + sl, ((fill / DEPTH) & 1u) ^ 1u);
            ws_f_task<KIND>(img, hb, S, q, n_rem > 0, W.xch + sl * XSLOT, u);
            ws_mbar_arrive(W.full + sl);
            ++fill;
        }
    }
}

template <int DEPTH>
__device__ __forceinline__ void ws_b_pass(const float* __restrict__ img, const StatePtrs& S, const WsCtx& W, int ne, int pair,
                                          float* __restrict__ s_d1part_net, float* __restrict__ slot_row, unsigned& cons, int& done_local) {
    const int lane = threadIdx.x & 31;
    u64 A[48];
    float ab3 = 0.f;
#pragma unroll
    for (int i = 0; i < 48; ++i) A[i] = 0ull;
    for (int e = 0; e < ne; ++e)
        if (lane < H) s_d1part_net[(e * WS_PAIRS + pair) * H + lane] = 0.f;
    __syncwarp();
    int t = W.t_lo;
#pragma unroll 1
    while (t < W.t_hi) {
        const int code = W.s_task[t], e = code >> 8, m0 = code & 255, n_live = W.s_nlive[e];
        const int seg_end = min(W.t_hi, t + (n_live + WS_TASK - 1) / WS_TASK - m0);
        int q0 = e * CHUNK + m0 * WS_TASK + lane, n_rem = n_live - m0 * WS_TASK - lane;
#pragma unroll 1
        for (; t < seg_end; ++t, q0 += WS_TASK, n_rem -= WS_TASK) {
            asm volatile("" ::: "memory");
            const unsigned sl = cons % DEPTH;
            ws_mbar_wait(W.full + sl, (cons / DEPTH) & 1u);
            ws_b_task(img, S, q0, n_rem, W.xch + sl * XSLOT, A, ab3);
            ws_mbar_arrive(W.empty + sl);
            ++cons;
            __syncwarp();                                      // every lane's state update precedes the published count
            ++done_local;
            if (lane == 0) ws_st_release(W.done, done_local);
        }
        warp_flush_b1(A, s_d1part_net + (e * WS_PAIRS + pair) * H);
    }
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < 48; ++i) U2(A[i], acc[2 * i], acc[2 * i + 1]);
    acc[96] = ab3;
    warp_reduce_to_slot(acc, slot_row);
}

struct WsArgs {
    const float* packed; int n_flows, C_row; const float* y; const float* row_ctx; int flags, B, N; const float* g_y; const float* g_ld;
    float* d_x; float* warp_rows; float* ctx_rows; float* cta_rows; float* d_row_ctx; int e_max;
};
struct WsShared {
    float (*s_slot)[100];
    uint64_t* s_full; uint64_t* s_empty;     // [WS_PAIRS][DEPTH]
    int* s_done;
};

// One role's whole life after the register split.  IS_F is a template parameter on purpose: the two roles must not share a single
// instruction after setmaxnreg (ptxas allocates a region that both roles can reach for the SMALLER budget).
template <int DEPTH, bool IS_F>
__device__ __forceinline__ void ws_role(const WsArgs& a, float* smem, const WsShared& sh) {
    const float* __restrict__ packed = a.packed; const float* __restrict__ y = a.y; const float* __restrict__ row_ctx = a.row_ctx;
    const float* __restrict__ g_y = a.g_y; const float* __restrict__ g_ld = a.g_ld;
    float* __restrict__ d_x = a.d_x; float* warp_rows = a.warp_rows; float* __restrict__ ctx_rows = a.ctx_rows;
    float* __restrict__ cta_rows = a.cta_rows; float* __restrict__ d_row_ctx = a.d_row_ctx;
    const int n_flows = a.n_flows, C_row = a.C_row, flags = a.flags, B = a.B, N = a.N, e_max = a.e_max;
    (void)packed; (void)y; (void)row_ctx; (void)g_y; (void)g_ld; (void)d_x; (void)ctx_rows; (void)cta_rows; (void)d_row_ctx; (void)warp_rows;
    constexpr bool is_f = IS_F;
    float (*s_slot)[100] = sh.s_slot;
    int* s_done = sh.s_done;
    constexpr int NW = WS_PAIRS, TASK = WS_TASK, NTF = IS_F ? 32 * WS_FW : 32 * WS_PAIRS;    // NTF: threads of this role
    const int n_fcnn = 4 * n_flows, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, inverse = flags & 1;
    const int pair = IS_F ? warp >> 1 : warp - WS_FW;                    // team: forward warps 2k, 2k + 1 and backward warp 16 + k
    const int rt = IS_F ? tid : tid - 32 * WS_FW;                        // thread index inside the role
    const int ctx_pad = (C_row + 3) & ~3, C1 = C_row + 1, nR = n_fcnn * H * C_row, nR1 = n_fcnn * H * C1;
    float* s_img = smem;
    float* s_b1 = s_img + n_fcnn * Img::SIZE;
    float* s_w1r = s_b1 + n_fcnn * H;
    float* s_accR = s_w1r + nR;
    int* s_task = reinterpret_cast<int*>(s_accR + ((nR1 + 3) & ~3));
    int* s_nlive = s_task + E_CAP * (CHUNK / 32);
    int* s_first = s_nlive + 16;
    long long* s_p0 = reinterpret_cast<long long*>(s_first + 16);
    float* s_state = reinterpret_cast<float*>(s_p0 + 32);
    float* s_hb = s_state + (size_t)NSTATE * e_max * CHUNK;
    float* s_d1row = s_hb + (size_t)e_max * n_fcnn * H;
    float* s_d1part = s_d1row + (size_t)e_max * n_fcnn * H;
    float* s_ctx = s_d1part + (size_t)n_fcnn * e_max * NW * H;
    float* s_xch = s_ctx + (((size_t)e_max * ctx_pad + 3) & ~(size_t)3);   // [WS_PAIRS][DEPTH][XSLOT]
    const size_t plane = (size_t)e_max * CHUNK;
    float* s_lo = s_state, *s_up = s_lo + plane, *s_glo = s_up + plane, *s_gup = s_glo + plane, *s_gld = s_gup + plane, *s_ex = s_gld + plane;

    float* my_rows = warp_rows + ((size_t)blockIdx.x * NW + pair) * n_fcnn * NACC;
    const int nc = (N + CHUNK - 1) / CHUNK;
    const int n_traj = (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_entries = n_traj * nc;
    WsCtx W;
    W.s_task = s_task; W.s_nlive = s_nlive; W.full = sh.s_full + pair * DEPTH; W.empty = sh.s_empty + pair * DEPTH; W.done = &s_done[pair];
    W.xch = s_xch + (size_t)pair * DEPTH * XSLOT;
    unsigned ring = 0;                // tasks produced (forward role) / consumed (backward role) so far: slot and parity of the ring
    for (int e0 = 0; e0 < n_entries; e0 += e_max) {
        const int ne = min(e_max, n_entries - e0);
        if (is_f) {
            if (rt < ne) {
                const int j = e0 + rt, b = blockIdx.x + (j / nc) * gridDim.x, c0 = (j % nc) * CHUNK;
                s_nlive[rt] = min(CHUNK, N - c0);
                s_first[rt] = c0 == 0;
                s_p0[rt] = (long long)b * N + c0;
                s_p0[16 + rt] = b;
            }
        } else if (rt < WS_PAIRS) {
            s_done[rt] = 0;
        }
        __syncthreads();                                                           // (1)
        int n_tasks = 0;
        for (int e = 0; e < ne; ++e) n_tasks += (s_nlive[e] + TASK - 1) / TASK;
        if (is_f) {
            for (int i = rt; i < ne * C_row; i += NTF) {
                const int e = i / C_row, c = i - e * C_row;
                s_ctx[e * ctx_pad + c] = row_ctx[(size_t)s_p0[16 + e] * C_row + c];
            }
            int nt_ = 0;
            for (int e = 0; e < ne; ++e) {
                const int te = (s_nlive[e] + TASK - 1) / TASK;
                for (int m = rt; m < te; m += NTF) s_task[nt_ + m] = (e << 8) | m;
                nt_ += te;
            }
        }
        __syncthreads();                                                           // (2)
        W.t_lo = (pair * n_tasks) / NW; W.t_hi = ((pair + 1) * n_tasks) / NW;
        const int Tn = W.t_hi - W.t_lo;
        if (is_f) {
            {   // hoisted layer-1 biases
                const int sub = rt & 3, total = ne * n_fcnn * H;
                for (int o0 = 0; o0 < total; o0 += NTF / 4) {
                    const int o = o0 + (rt >> 2);
                    float a = 0.f;
                    if (o < total) {
                        const int e = o / (n_fcnn * H), fk = o - e * n_fcnn * H;
                        const float* w = s_w1r + (size_t)fk * C_row;
                        const float* cx = s_ctx + e * ctx_pad;
                        for (int c = sub; c < C_row; c += 4) a = fmaf(w[c], cx[c], a);
                    }
                    a += __shfl_xor_sync(FULL, a, 1);
                    a += __shfl_xor_sync(FULL, a, 2);
                    if (o < total && sub == 0) s_hb[o] = a + s_b1[o % (n_fcnn * H)];
                }
            }
            // particle state of the team's own tasks (each forward warp loads the half it will work on)
#pragma unroll 2
            for (int t = W.t_lo; t < W.t_hi; ++t) {
                const int code = s_task[t], e = code >> 8, m = code & 255, n_live = s_nlive[e];
                const size_t p0 = (size_t)s_p0[e];
                {
                    const int u = warp & 1;
                    const int q = m * TASK + 32 * u + lane, i = e * CHUNK + q;
                    const bool live = q < n_live;
                    const float2 yy = live ? reinterpret_cast<const float2*>(y)[p0 + q] : make_float2(0.f, 0.f);
                    const float2 gg = live && g_y ? reinterpret_cast<const float2*>(g_y)[p0 + q] : make_float2(0.f, 0.f);
                    const float gl = live && g_ld ? g_ld[p0 + q] : 0.f;
                    s_lo[i] = yy.x; s_up[i] = yy.y; s_glo[i] = gg.x; s_gup[i] = gg.y;
                    s_gld[i] = (flags & 2) ? -gl : gl;
                    s_ex[i] = 0.f;
                }
            }
        }
        __syncthreads();                                                           // (3) hb complete, state loaded
        int done_local = 0;
#pragma unroll 1
        for (int ps = 0; ps < 4 * n_flows; ++ps) {
            const int st = ps >> 1, second = ps & 1;
            const int f = inverse ? st / 2 : n_flows - 1 - st / 2;
            const int pr = inverse ? (st & 1) : 1 - (st & 1);
            const int net = inverse ? 1 - second : second;
            const int fm = 4 * f + 2 * pr + net;
            const float* img = s_img + fm * Img::SIZE;
            StatePtrs S;
            S.c = pr ? s_up : s_lo;  S.gc = pr ? s_gup : s_glo;
            S.v = pr ? s_lo : s_up;  S.gv = pr ? s_glo : s_gup;
            S.gld = s_gld;           S.ex = s_ex;
            if (is_f) {
                const float* hbn = s_hb + fm * H;
                const int hbs = n_fcnn * H, need0 = ps ? (ps - 1) * Tn : -1;
                if (!inverse) {
                    if (!net) ws_f_pass<DEPTH, FWD_T>(img, hbn, hbs, S, W, need0, ring);
                    else      ws_f_pass<DEPTH, FWD_S>(img, hbn, hbs, S, W, need0, ring);
                } else {
                    if (net)  ws_f_pass<DEPTH, INV_S>(img, hbn, hbs, S, W, need0, ring);
                    else      ws_f_pass<DEPTH, INV_T>(img, hbn, hbs, S, W, need0, ring);
                }
            } else {
                float* d1p = s_d1part + (size_t)fm * e_max * NW * H;
                float* slot = s_slot[pair];
                ws_b_pass<DEPTH>(img, S, W, ne, pair, d1p, slot, ring, done_local);
                __syncwarp();
                float* row = my_rows + fm * NACC;
                for (int k = lane; k < NACC; k += 32) row[k] = e0 == 0 ? slot[k] : row[k] + slot[k];
                __syncwarp();
            }
        }
        __syncthreads();                                                           // (4) every pass of every pair is finished
        if (!is_f) {
            // ---- store the pair's d_x
#pragma unroll 2
            for (int t = W.t_lo; t < W.t_hi; ++t) {
                const int code = s_task[t], e = code >> 8, m = code & 255, n_live = s_nlive[e];
                const size_t p0 = (size_t)s_p0[e];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int q = m * TASK + 32 * u + lane, i = e * CHUNK + q;
                    if (q < n_live) reinterpret_cast<float2*>(d_x)[p0 + q] = make_float2(s_glo[i], s_gup[i]);
                }
            }
            for (int o = rt; o < ne * n_fcnn * H; o += NTF) {     // fixed-order sums over the pairs
                const int e = o / (n_fcnn * H), fk = o - e * n_fcnn * H, f = fk / H, k = fk - f * H;
                const float* dp = s_d1part + (((size_t)f * e_max + e) * NW) * H + k;
                float r = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) r += dp[w * H];
                s_d1row[o] = r;
            }
        }
        __syncthreads();                                                           // (5)
        if (!is_f) {
            for (int o = rt; o < nR1; o += NTF) {
                const int fk = o / C1, c = o - fk * C1;
                float a = s_accR[o];
                for (int e = 0; e < ne; ++e) a = fmaf(s_d1row[(size_t)e * n_fcnn * H + fk], c < C_row ? s_ctx[e * ctx_pad + c] : 1.0f, a);
                s_accR[o] = a;
            }
            if (d_row_ctx) {
                for (int c = rt; c < C_row; c += NTF)
                    for (int e = 0; e < ne; ++e) {
                        float a = 0.f;
                        for (int fk = 0; fk < n_fcnn * H; ++fk) a = fmaf(s_w1r[(size_t)fk * C_row + c], s_d1row[(size_t)e * n_fcnn * H + fk], a);
                        float* dst = d_row_ctx + (size_t)s_p0[16 + e] * C_row + c;
                        *dst = s_first[e] ? a : *dst + a;
                    }
            }
        }
        __syncthreads();                                                           // (6)
    }
    if (!is_f) {
        float* out = ctx_rows + (size_t)blockIdx.x * nR1;
        for (int e = rt; e < nR1; e += NTF) out[e] = s_accR[e];
        const int n_acc = n_fcnn * NACC;
        const float* mine = warp_rows + (size_t)blockIdx.x * NW * n_acc;
        float* folded = cta_rows + (size_t)blockIdx.x * n_acc;
        __syncwarp();
        asm volatile("bar.sync 2, 256;" ::: "memory");            // every backward warp's global row is written (same-CTA visibility)
        for (int c = rt; c < n_acc; c += NTF) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += mine[(size_t)w * n_acc + c];
            folded[c] = v;
        }
    }
}

size_t coupling_bwd_d2_workspace_floats(int n_flows, int C_row, int B) {
    const int n_fcnn = 4 * n_flows, grid = min(B, sm_count());
    return (size_t)grid * ((size_t)(D2_MAX_WARPS + 1) * n_fcnn * NACC + (size_t)n_fcnn * H * (C_row + 1));
}

template <int PPT, int NT>
static int launch_cfg(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int flags, int B, int N,
                      const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* d_packed, void* workspace,
                      cudaStream_t st, float* block = nullptr) {
    static_assert(NT / 32 <= D2_MAX_WARPS, "workspace is sized for D2_MAX_WARPS warps per CTA");
    const int n_fcnn = 4 * n_flows;
    const int grid = min(B, sm_count());   // one CTA per SM, persistent over trajectories
    // resident entries: as many as the CTA's share of the work needs, bounded by shared memory
    const int nc = (N + CHUNK - 1) / CHUNK, need = ((B + grid - 1) / grid) * nc;
    int e_max = min(E_CAP, need);
    using SM = D2Smem<NT>;
    constexpr size_t STATIC = (NT / 32) * 100 * sizeof(float);
    while (e_max > 1 && SM::bytes(n_fcnn, C_row, e_max) + STATIC > 226 * 1024) --e_max;
    const size_t smem = SM::bytes(n_fcnn, C_row, e_max);
    if (smem + STATIC > 226 * 1024) { set_error("coupling_bwd_d2: stack too large for shared memory (n_flows=%d, C_row=%d)", n_flows, C_row); return NFDPF_ERR_UNSUPPORTED; }
    auto kern = coupling_bwd_d2_kernel<PPT, NT>;
    NFDPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float* warp_rows = (float*)workspace;
    // the rows the reduction reads ([row-context partials | folded CTA rows]): behind the per-warp rows, or -- deferred reduction --
    // in the caller's block, where they wait for ONE reduce launch over all the calls of a training step
    float* ctx_rows = block ? block : warp_rows + (size_t)grid * (NT / 32) * n_fcnn * NACC;
    float* cta_rows = ctx_rows + (size_t)grid * n_fcnn * H * (C_row + 1);
    kern<<<grid, NT, smem, st>>>(packed, n_flows, C_row, y, row_ctx, flags, B, N, g_y, g_ld, d_x, warp_rows, ctx_rows, cta_rows, d_row_ctx, e_max);
    int rc = check_launch("coupling_bwd_d2");
    if (rc || block) return rc;
    const int n_cols = n_fcnn * NACC + n_fcnn * H * (C_row + 1);
    d2_reduce_kernel<<<(n_cols + 31) / 32, 32 * RED_G, 0, st>>>(cta_rows, grid, ctx_rows, grid, n_fcnn, C_row, d_packed, 1, 0);
    return check_launch("d2_reduce");
}


template <int DEPTH>
__global__ void __launch_bounds__(WS_NT, 1)
coupling_bwd_d2_ws_kernel(WsArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_slot[WS_PAIRS][100];
    __shared__ uint64_t s_full[WS_PAIRS][DEPTH], s_empty[WS_PAIRS][DEPTH];
    __shared__ int s_done[WS_PAIRS];
    {
        const float* __restrict__ packed = a.packed;
        const int n_flows = a.n_flows, C_row = a.C_row;
        constexpr int NT = WS_NT;
        const int n_fcnn = 4 * n_flows, tid = threadIdx.x;
        const int C1 = C_row + 1, nR = n_fcnn * H * C_row, nR1 = n_fcnn * H * C1;
        float* s_img = smem;
        float* s_b1 = s_img + n_fcnn * Img::SIZE;
        float* s_w1r = s_b1 + n_fcnn * H;
        float* s_accR = s_w1r + nR;
    const int fin = 1 + C_row, pf = packed_fcnn_size(1, C_row);
    for (int d = tid; d < n_fcnn * Img::SIZE; d += NT) s_img[d] = img_value(packed, d / Img::SIZE, d % Img::SIZE, C_row);
    for (int e = tid; e < n_fcnn * H; e += NT) s_b1[e] = TANH_SCALE * packed[(size_t)(e / H) * pf + H * fin + (e % H)];
    for (int e = tid; e < nR; e += NT) {
        const int f = e / (H * C_row), r = e - f * H * C_row;
        s_w1r[e] = TANH_SCALE * packed[(size_t)f * pf + (r / C_row) * fin + 1 + (r % C_row)];
    }
    for (int e = tid; e < nR1; e += NT) s_accR[e] = 0.f;
    if (tid < WS_PAIRS * DEPTH) { ws_mbar_init(&s_full[0][0] + tid, 64); ws_mbar_init(&s_empty[0][0] + tid, 32); }   // two forward warps fill a slot
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    }
    const WsShared sh{s_slot, &s_full[0][0], &s_empty[0][0], s_done};
    // the register file is re-split between the roles; from here on the two roles share no code (only barrier counts)
    if ((threadIdx.x >> 5) < WS_FW) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_REG_F));
        ws_role<DEPTH, true>(a, smem, sh);
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_REG_B));
        ws_role<DEPTH, false>(a, smem, sh);
    }
}

template <int DEPTH>
static int launch_ws(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int flags, int B, int N,
                     const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* d_packed, void* workspace,
                     cudaStream_t st, bool* fits) {
    const int n_fcnn = 4 * n_flows;
    const int grid = min(B, sm_count());
    const int nc = (N + CHUNK - 1) / CHUNK, need = ((B + grid - 1) / grid) * nc;
    using SM = D2Smem<256>;                                   // same layout as the single-role kernel (eight accumulating warps) ...
    const size_t xch = (size_t)WS_PAIRS * DEPTH * XSLOT * sizeof(float) + 16;   // ... plus the exchange ring
    constexpr size_t STATIC = WS_PAIRS * 100 * sizeof(float) + 2 * WS_PAIRS * DEPTH * 8 + 64;
    int e_max = min(E_CAP, need);
    while (e_max > 1 && SM::bytes(n_fcnn, C_row, e_max) + xch + STATIC > 226 * 1024) --e_max;
    const size_t smem = SM::bytes(n_fcnn, C_row, e_max) + xch;
    *fits = smem + STATIC <= 226 * 1024;
    if (!*fits) return NFDPF_OK;
    auto kern = coupling_bwd_d2_ws_kernel<DEPTH>;
    NFDPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float* warp_rows = (float*)workspace;
    float* ctx_rows = warp_rows + (size_t)grid * WS_PAIRS * n_fcnn * NACC;
    float* cta_rows = ctx_rows + (size_t)grid * n_fcnn * H * (C_row + 1);
    const WsArgs args{packed, n_flows, C_row, y, row_ctx, flags, B, N, g_y, g_ld, d_x, warp_rows, ctx_rows, cta_rows, d_row_ctx, e_max};
    kern<<<grid, WS_NT, smem, st>>>(args);
    int rc = check_launch("coupling_bwd_d2_ws");
    if (rc) return rc;
    const int n_cols = n_fcnn * NACC + n_fcnn * H * (C_row + 1);
    d2_reduce_kernel<<<(n_cols + 31) / 32, 32 * RED_G, 0, st>>>(cta_rows, grid, ctx_rows, grid, n_fcnn, C_row, d_packed, 1, 0);
    return check_launch("d2_reduce");
}

// ---- deferred reduction (one reduce launch per training step instead of one per call) ------------------------------------------
size_t coupling_bwd_d2_block_floats(int n_flows, int C_row, int B) {
    const int n_fcnn = 4 * n_flows, grid = min(B, sm_count());
    return (size_t)grid * ((size_t)n_fcnn * H * (C_row + 1) + (size_t)n_fcnn * NACC);
}
int launch_coupling_bwd_d2_deferred(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int inverse, int B,
                                    int N, const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* block, void* workspace,
                                    cudaStream_t st) {
    return launch_cfg<2, 256>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, nullptr, workspace, st, block);
}
int launch_coupling_bwd_d2_reduce(int n_flows, int C_row, int B, const float* blocks, int n_calls, float* d_packed, cudaStream_t st) {
    const int n_fcnn = 4 * n_flows, grid = min(B, sm_count());
    const int n_cols = n_fcnn * NACC + n_fcnn * H * (C_row + 1);
    const float* ctx_rows = blocks;
    const float* cta_rows = blocks + (size_t)grid * n_fcnn * H * (C_row + 1);
    d2_reduce_kernel<<<(n_cols + 31) / 32, 32 * RED_G, 0, st>>>(cta_rows, grid, ctx_rows, grid, n_fcnn, C_row, d_packed, n_calls,
                                                                coupling_bwd_d2_block_floats(n_flows, C_row, B));
    return check_launch("d2_reduce (deferred)");
}

int launch_coupling_bwd_d2(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int inverse, int B, int N,
                           const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* d_packed, void* workspace,
                           cudaStream_t st) {
    // NFDPF_D2_CFG=1 selects the one-particle-per-thread / 12-warp geometry (A/B measurements); default: two particles, 8 warps
    // NFDPF_D2_CFG: 0 = single-role kernel (two particles per thread, 8 warps), 1 = its one-particle / 12-warp geometry,
    // 2 / 3 = producer / consumer kernel with an exchange ring of depth 2 / 1.  Default 0: the producer / consumer kernel is parity-green
    // but was measured SLOWER (200 / 231 us against 161-168 us at B = N = 1024) -- it stays selectable for A/B runs.
    static const int cfg = [] { const char* s = getenv("NFDPF_D2_CFG"); return s ? atoi(s) : 0; }();
    if (cfg >= 2) {
        bool fits = false;
        const int rc = cfg == 3 ? launch_ws<1>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_packed, workspace, st, &fits)
                                : launch_ws<2>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_packed, workspace, st, &fits);
        if (rc || fits) return rc;
    }
    if (cfg == 4)      // diagnostic: ONE warp per scheduler (how long does a warp take when it has the pipes to itself?)
        return launch_cfg<2, 128>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_packed, workspace, st);
    if (cfg == 1)
        return launch_cfg<1, 384>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_packed, workspace, st);
    return launch_cfg<2, 256>(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_packed, workspace, st);
}

}  // namespace nfdpf
