// Device building blocks of the fused (conditional) RealNVP coupling stack -- shared by the stand-alone
// coupling kernels (coupling.cu), the measurement kernels (measure.cu) and the fused filter step.
//
// Reference semantics: nf/flows.py:101-114 (FCNN = Linear-Tanh-Linear-Tanh-Linear, hidden 8),
// nf/flows.py:155-179 / 215-239 (RealNVP / RealNVP_cond forward + inverse), nf/models.py:45-61 (stack order).
//
// Layout decisions
//   * one thread = one particle; all 8 FCNNs of a 2-flow stack are evaluated in registers;
//   * weights live in shared memory as a 16-byte aligned "image" per FCNN (Lay<>), read with broadcast LDS.128;
//   * the row-constant part of the context (per-trajectory mean/std/observation encoding, which the reference
//     materialises as a (P,C) tensor and concatenates 8x, model/models.py:309-315, 338-346) is folded once per
//     trajectory into the layer-1 bias: hb = b1 + W1[:, row cols] . ctx_row  ("hoisting");
//   * backward walks the stack from its OUTPUT (couplings are invertible), recomputing activations, so nothing
//     but y is saved; weight gradients are reduced per CTA through a transposed shared-memory tile (one owner
//     thread per parameter => no atomics, fixed order => deterministic) and across CTAs by a second kernel.
#pragma once
#include "common.cuh"
#include "mma_tile.cuh"

namespace nfdpf {

constexpr int H = NFDPF_HIDDEN;  // 8

int bwd_grid(int B);  // persistent grid of the backward kernels (also sizes the partial-gradient workspace)
// d_packed[i] += sum_c partials[c][i], fixed order, fp64 accumulate (coupling.cu)
int launch_reduce_partials(const float* partials, int n_parts, int n_params, float* d_packed, cudaStream_t st);
// register-accumulating backward for D = 2 stacks with row-constant context only (coupling_d2.cu) and its workspace size
size_t coupling_bwd_d2_workspace_floats(int n_flows, int C_row, int B);
int launch_coupling_bwd_d2(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int inverse, int B, int N,
                           const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* d_packed, void* workspace,
                           cudaStream_t st);

// deferred reduction of the D = 2 backward: a call leaves [row-context partials | folded CTA rows] in the caller's block
// (coupling_bwd_d2_block_floats); one reduce launch sums the blocks of n_calls calls (same n_flows, C_row, B) into d_packed
size_t coupling_bwd_d2_block_floats(int n_flows, int C_row, int B);
int launch_coupling_bwd_d2_deferred(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, int inverse, int B,
                                    int N, const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx, float* block, void* workspace,
                                    cudaStream_t st);
int launch_coupling_bwd_d2_reduce(int n_flows, int C_row, int B, const float* blocks, int n_calls, float* d_packed, cudaStream_t st);

// Aligned shared-memory image of one FCNN (row-context columns of W1 excluded; they live in s_w1r).
template <int HALF, int CP>
struct Lay {
    static constexpr int IN1 = HALF + CP;
    static constexpr int S1 = (IN1 + 3) & ~3;   // padded W1 row stride
    static constexpr int W1 = 0;                // [8][S1]  columns: conditioning half | per-particle context
    static constexpr int B1 = W1 + H * S1;      // [8]
    static constexpr int W2 = B1 + H;           // [8][8]
    static constexpr int B2 = W2 + H * H;       // [8]
    static constexpr int W3 = B2 + H;           // [HALF][8]
    static constexpr int B3 = W3 + HALF * H;    // [HALF]
    static constexpr int SIZE = (B3 + HALF + 3) & ~3;
};

// size of one FCNN in the packed (state_dict order) parameter vector
__host__ __device__ inline int packed_fcnn_size(int half, int C) { return H * (half + C) + H + H * H + H + half * H + half; }

// Copy one packed FCNN into its aligned image (and its row-context W1 columns into w1r[8][C_row]).
template <int HALF, int CP>
__device__ void load_fcnn_image(const float* __restrict__ pk, int C_row, float* __restrict__ img, float* __restrict__ w1r,
                                int tid, int nt) {
    using L = Lay<HALF, CP>;
    const int fin = HALF + C_row + CP;
    for (int e = tid; e < H * L::S1; e += nt) {
        const int k = e / L::S1, i = e % L::S1;
        float v = 0.f;
        if (i < HALF) v = pk[k * fin + i];
        else if (i < L::IN1) v = pk[k * fin + C_row + i];
        img[L::W1 + e] = TANH_SCALE * v;
    }
    // Everything that feeds a tanh (layers 1 and 2: weights, biases, row-context columns) is stored multiplied by
    // 2 log2(e), so the activation is tanh_prescaled(pre-activation): 4 instructions instead of 7.  The backward
    // carries delta / scale (free: folded into the 1 - h^2 factor) and rescales the weight gradients once on output.
    for (int e = tid; e < H * C_row; e += nt) w1r[e] = TANH_SCALE * pk[(e / C_row) * fin + HALF + (e % C_row)];
    const float* p = pk + H * fin;
    for (int e = tid; e < H; e += nt) img[L::B1 + e] = TANH_SCALE * p[e];
    p += H;
    for (int e = tid; e < H * H; e += nt) img[L::W2 + e] = TANH_SCALE * p[e];
    p += H * H;
    for (int e = tid; e < H; e += nt) img[L::B2 + e] = TANH_SCALE * p[e];
    p += H;
    for (int e = tid; e < HALF * H; e += nt) img[L::W3 + e] = p[e];
    p += HALF * H;
    for (int e = tid; e < HALF; e += nt) img[L::B3 + e] = p[e];
}

// All FCNN images of a stack in two flat passes (image words, then row-context columns): every iteration is an independent
// global load, so a CTA's loads overlap instead of queueing behind seven short loops per net (the one-CTA-per-trajectory
// forward kernels spend a visible share of their life in this prologue).
template <int HALF, int CP>
__device__ void load_stack_images(const float* __restrict__ packed, int n_fcnn, int C_row, float* __restrict__ imgs, float* __restrict__ w1r,
                                  int tid, int nt) {
    using L = Lay<HALF, CP>;
    const int fin = HALF + C_row + CP, pf = packed_fcnn_size(HALF, C_row + CP);
#pragma unroll 4
    for (int d = tid; d < n_fcnn * L::SIZE; d += nt) {
        const int f = d / L::SIZE, o = d % L::SIZE;
        const float* pk = packed + (size_t)f * pf;
        float v = 0.f;
        if (o < L::B1) {                                   // W1 [8][S1]: conditioning half | per-particle context, zero padded
            const int k = o / L::S1, i = o % L::S1;
            if (i < HALF) v = TANH_SCALE * pk[k * fin + i];
            else if (i < L::IN1) v = TANH_SCALE * pk[k * fin + C_row + i];
        } else if (o < L::W3) v = TANH_SCALE * pk[H * fin + (o - L::B1)];                       // b1, W2, b2: contiguous in both layouts
        else if (o < L::B3 + HALF) v = pk[H * fin + H + H * H + H + (o - L::W3)];               // W3, b3
        imgs[d] = v;
    }
    const int per = H * C_row;
#pragma unroll 4
    for (int e = tid; e < n_fcnn * per; e += nt) {
        const int f = e / per, r = e - f * per;
        w1r[e] = TANH_SCALE * packed[(size_t)f * pf + (r / C_row) * fin + HALF + (r % C_row)];
    }
}

__device__ __forceinline__ void ld8(const float* p, float (&w)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// FCNN forward.  hb = hoisted layer-1 bias (b1 + row-context contribution).  Keeps h1/h2 for the backward.
// PAIRED: activations through tanh_prescaled_pair (3 MUFU per two tanh; the forward-only kernels are SFU-bound).
template <int HALF, int CP, bool PAIRED = false>
__device__ __forceinline__ void fcnn_fwd(const float* __restrict__ img, const float* __restrict__ hb, const float (&c)[HALF],
                                         const float* pc, float (&h1)[H], float (&h2)[H], float (&out)[HALF]) {
    using L = Lay<HALF, CP>;
    // dot products on packed FP32 pairs (FFMA2): (even, odd) partial sums over the input index, one horizontal add at the end
    float hbv[8];
    ld8(hb, hbv);
    float in[L::S1];                       // zero padded like the W1 rows
#pragma unroll
    for (int i = 0; i < HALF; ++i) in[i] = c[i];
#pragma unroll
    for (int i = 0; i < CP; ++i) in[HALF + i] = pc[i];
#pragma unroll
    for (int i = L::IN1; i < L::S1; ++i) in[i] = 0.f;
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float p0 = hbv[k], p1 = 0.f;
        const float* w = img + L::W1 + k * L::S1;
#pragma unroll
        for (int i = 0; i < L::S1; i += 4) {  // rows are zero-padded to a multiple of 4
            const float4 w4 = *reinterpret_cast<const float4*>(w + i);
            if (i < L::IN1) ffma2_p(p0, p1, w4.x, w4.y, in[i], in[i + 1]);
            if (i + 2 < L::IN1) ffma2_p(p0, p1, w4.z, w4.w, in[i + 2], in[i + 3]);
        }
        const float a = p0 + p1;
        h1[k] = PAIRED ? a : tanh_prescaled(a);
    }
    if (PAIRED) {
#pragma unroll
        for (int k = 0; k < H; k += 2) tanh_prescaled_pair(h1[k], h1[k + 1], h1[k], h1[k + 1]);
    }
    float b2[8];
    ld8(img + L::B2, b2);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
        float p0 = b2[j], p1 = 0.f;
#pragma unroll
        for (int k = 0; k < H; k += 2) ffma2_p(p0, p1, w[k], w[k + 1], h1[k], h1[k + 1]);
        const float a = p0 + p1;
        h2[j] = PAIRED ? a : tanh_prescaled(a);
    }
    if (PAIRED) {
#pragma unroll
        for (int j = 0; j < H; j += 2) tanh_prescaled_pair(h2[j], h2[j + 1], h2[j], h2[j + 1]);
    }
#pragma unroll
    for (int o = 0; o < HALF; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
        float p0 = img[L::B3 + o], p1 = 0.f;
#pragma unroll
        for (int j = 0; j < H; j += 2) ffma2_p(p0, p1, w[j], w[j + 1], h2[j], h2[j + 1]);
        out[o] = p0 + p1;
    }
}

// FCNN backward (data path): dout -> d2, d1 = pre-activation grads DIVIDED BY TANH_SCALE (the layer-1/2 images hold
// scaled weights, so W_s^T (delta / s) = W^T delta exactly); dc += W1c^T delta1, dpc += W1p^T delta1.
template <int HALF, int CP>
__device__ __forceinline__ void fcnn_bwd(const float* __restrict__ img, const float (&dout)[HALF], const float (&h1)[H],
                                         const float (&h2)[H], float (&d1)[H], float (&d2)[H], float (&dc)[HALF], float* dpc) {
    using L = Lay<HALF, CP>;
    // matrix-transpose products on packed FP32 pairs (FFMA2): the delta is the broadcast operand, consecutive weights the pair
    float da2[H];
#pragma unroll
    for (int j = 0; j < H; ++j) da2[j] = 0.f;
#pragma unroll
    for (int o = 0; o < HALF; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
#pragma unroll
        for (int j = 0; j < H; j += 2) ffma2_s(da2[j], da2[j + 1], dout[o], w[j], w[j + 1]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) d2[j] = da2[j] * fmaf(-TANH_ISCALE * h2[j], h2[j], TANH_ISCALE);   // delta2 / scale
    float da1[H];
#pragma unroll
    for (int k = 0; k < H; ++k) da1[k] = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
#pragma unroll
        for (int k = 0; k < H; k += 2) ffma2_s(da1[k], da1[k + 1], d2[j], w[k], w[k + 1]);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) d1[k] = da1[k] * fmaf(-TANH_ISCALE * h1[k], h1[k], TANH_ISCALE);   // delta1 / scale (W2 image is scaled: W2s^T (d2/s) = W2^T d2)
    if constexpr (HALF % 2 == 0 && CP % 2 == 0) {   // d in = W1^T d1 straight into the (even-sized) gradient arrays, two inputs per FFMA2
#pragma unroll
        for (int k = 0; k < H; ++k) {
            const float* w = img + L::W1 + k * L::S1;
#pragma unroll
            for (int i = 0; i < L::S1; i += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + i);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int u = 0; u < 4; u += 2) {
                    if (i + u < HALF) ffma2_s(dc[i + u < HALF ? i + u : 0], dc[i + u < HALF ? i + u + 1 : 1], d1[k], wv[u], wv[u + 1]);
                    else if (i + u < HALF + CP)
                        ffma2_s(dpc[i + u >= HALF ? i + u - HALF : 0], dpc[i + u >= HALF ? i + u - HALF + 1 : 1], d1[k], wv[u], wv[u + 1]);
                }
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < H; ++k) {   // d in = W1^T d1, 128-bit weight loads (rows are zero-padded to a multiple of 4)
            const float* w = img + L::W1 + k * L::S1;
#pragma unroll
            for (int i = 0; i < L::S1; i += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w + i);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (i + u < HALF) dc[i + u < HALF ? i + u : 0] = fmaf(wv[u], d1[k], dc[i + u < HALF ? i + u : 0]);
                    else if (i + u < HALF + CP) dpc[i + u >= HALF ? i + u - HALF : 0] = fmaf(wv[u], d1[k], dpc[i + u >= HALF ? i + u - HALF : 0]);
                }
            }
        }
    }
}

// hoisted layer-1 biases for every FCNN of a stack: hb[f][k] = b1[f][k] + sum_c W1r[f][k][c] * ctx[c]
template <int HALF, int CP>
__device__ void hoist_row_context(const float* __restrict__ imgs, const float* __restrict__ w1r, const float* __restrict__ ctx,
                                  int C_row, int n_fcnn, float* __restrict__ hb, int tid, int nt) {
    using L = Lay<HALF, CP>;
    for (int e = tid; e < n_fcnn * H; e += nt) {
        const int f = e / H, k = e % H;
        float a = imgs[f * L::SIZE + L::B1 + k];
        const float* w = w1r + (size_t)(f * H + k) * C_row;
        for (int c = 0; c < C_row; ++c) a = fmaf(w[c], ctx[c], a);
        hb[e] = a;
    }
}

// Same, with every thread of the CTA participating: (f,k) pairs are spread over groups of 4 lanes.
template <int HALF, int CP>
__device__ void hoist_row_context_par(const float* __restrict__ imgs, const float* __restrict__ w1r, const float* __restrict__ ctx,
                                      int C_row, int n_fcnn, float* __restrict__ hb) {
    using L = Lay<HALF, CP>;
    const int sub = threadIdx.x & 3, nt = blockDim.x >> 2;
    for (int e = threadIdx.x >> 2; e < ((n_fcnn * H + nt - 1) / nt) * nt; e += nt) {
        float a = 0.f;
        if (e < n_fcnn * H) {
            const float* w = w1r + (size_t)e * C_row;
            for (int c = sub; c < C_row; c += 4) a = fmaf(w[c], ctx[c], a);
        }
        a += __shfl_xor_sync(FULL, a, 1);
        a += __shfl_xor_sync(FULL, a, 2);
        if (e < n_fcnn * H && sub == 0) hb[e] = a + imgs[(e / H) * L::SIZE + L::B1 + (e % H)];
    }
}

// One coupling stage, forward evaluation.  inv = false: v = t(c) + v*exp(s(c)); inv = true: v = (v - t(c))*exp(-s(c)).
template <int HALF, int CP, bool PAIRED = false>
__device__ __forceinline__ void stage_fwd(const float* img_t, const float* img_s, const float* hb_t, const float* hb_s, bool inv,
                                          const float (&c)[HALF], const float* pc, float (&v)[HALF], float& ld) {
    float h1[H], h2[H], t[HALF], s[HALF];
    fcnn_fwd<HALF, CP, PAIRED>(img_t, hb_t, c, pc, h1, h2, t);
    fcnn_fwd<HALF, CP, PAIRED>(img_s, hb_s, c, pc, h1, h2, s);
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        if (!inv) { v[i] = fmaf(v[i], expf(s[i]), t[i]); ld += s[i]; }
        else      { v[i] = (v[i] - t[i]) * expf(-s[i]); ld -= s[i]; }
    }
}

// FCNN forward for TWO particles at once (narrow stacks, HALF <= 2): every weight is loaded from shared memory once and
// used twice, which halves the LDS traffic (LDS and MUFU share the MIO queue the forward kernel stalls on) and doubles the
// independent work per thread.  PAIRED pairs the two particles' activations in tanh_prescaled_pair.
template <int HALF, int CP, bool PAIRED>
__device__ __forceinline__ void fcnn_fwd_x2(const float* __restrict__ img, const float* __restrict__ hb, const float (&c)[2][HALF],
                                            const float (&pc)[2][CP > 0 ? CP : 1], float (&out)[2][HALF]) {
    // the two particles form the packed FP32 pair of every FMA: (a_p0, a_p1) += w * (x_p0, x_p1) is ONE FFMA2 with the weight
    // as its broadcast operand -- half the FMA issue slots of the (issue-bound) forward kernel
    using L = Lay<HALF, CP>;
    float hbv[8];
    ld8(hb, hbv);
    float in[L::IN1][2], h1[H][2], h2[H][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int i = 0; i < HALF; ++i) in[i][q] = c[q][i];
#pragma unroll
        for (int i = 0; i < CP; ++i) in[HALF + i][q] = pc[q][i];
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float a0 = hbv[k], a1 = hbv[k];
        const float* w = img + L::W1 + k * L::S1;
#pragma unroll
        for (int i = 0; i < L::S1; i += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(w + i);
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u < L::IN1) ffma2_s(a0, a1, wv[u], in[i + u < L::IN1 ? i + u : 0][0], in[i + u < L::IN1 ? i + u : 0][1]);
        }
        if (PAIRED) tanh_prescaled_pair(a0, a1, h1[k][0], h1[k][1]);
        else { h1[k][0] = tanh_prescaled(a0); h1[k][1] = tanh_prescaled(a1); }
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
        if (PAIRED) tanh_prescaled_pair(a0, a1, h2[j][0], h2[j][1]);
        else { h2[j][0] = tanh_prescaled(a0); h2[j][1] = tanh_prescaled(a1); }
    }
#pragma unroll
    for (int o = 0; o < HALF; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
        float a0 = img[L::B3 + o], a1 = a0;
#pragma unroll
        for (int j = 0; j < H; ++j) ffma2_s(a0, a1, w[j], h2[j][0], h2[j][1]);
        out[0][o] = a0; out[1][o] = a1;
    }
}

template <int HALF, int CP, bool PAIRED>
__device__ __forceinline__ void stage_fwd_x2(const float* img_t, const float* img_s, const float* hb_t, const float* hb_s, bool inv,
                                             const float (&c)[2][HALF], const float (&pc)[2][CP > 0 ? CP : 1], float (&v)[2][HALF],
                                             float (&ld)[2]) {
    float t[2][HALF], s[2][HALF];
    fcnn_fwd_x2<HALF, CP, PAIRED>(img_t, hb_t, c, pc, t);
    fcnn_fwd_x2<HALF, CP, PAIRED>(img_s, hb_s, c, pc, s);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int i = 0; i < HALF; ++i) {
            if (!inv) { v[q][i] = fmaf(v[q][i], exp_acc(s[q][i]), t[q][i]); ld[q] += s[q][i]; }
            else      { v[q][i] = (v[q][i] - t[q][i]) * exp_acc(-s[q][i]); ld[q] -= s[q][i]; }
        }
    }
}

// Exchange the roles of the two halves between stages so that ONE inlined copy of the stage code serves every stage.
template <int HALF>
__device__ __forceinline__ void swap_halves(float (&a)[HALF], float (&b)[HALF]) {
#pragma unroll
    for (int i = 0; i < HALF; ++i) { const float t = a[i]; a[i] = b[i]; b[i] = t; }
}


// ---- CTA batch geometry of the backward kernels (deltas / activations are staged transposed in a shared-memory tile,
// row = feature, column = particle; stride TSM, see mma_tile.cuh) -------------------------------------------------
constexpr int TP = 128;           // particles per CTA batch (= threads per CTA)


// ---- backward of one coupling stage ---------------------------------------------------------------------
// Tile rows (one column per particle of the batch):
template <int HALF, int CP>
struct Rows {
    static constexpr int ONE = 0, C = 1, PC = C + HALF, H1 = PC + CP, H2 = H1 + H, D1 = H2 + H, D2 = D1 + H, DO = D2 + H,
                         COUNT = DO + HALF,
                         ZERO = DO + 16,       // all-zero row (padding columns of the mma B tiles); DO's m16 tile may read up to here
                         TROWS = DO + 17;      // rows to allocate (stride TSM)
    // per-FCNN gradient outputs, in packed order minus the row-context columns
    static constexpr int NOUT = H * (HALF + CP) + H + H * H + H + HALF * H + HALF;
};

// Weight gradients of ONE FCNN on the tensor path (3xTF32 mma.sync, see mma_tile.cuh).  Each warp contracts over the
// 32 particles its own threads staged (no CTA barrier) and adds the result into ITS OWN accumulator copy acc_f[NOUT]
// (packed order minus row-context columns) and, for the b1 block, d1row_f (row-context hoist); the copies are summed
// in a fixed order when the CTA writes its partial gradient.
//   A1 = [delta1 (8 rows); delta2 (8 rows)]   x  B tiles: inputs [c | pc] (NIN tiles), ONE, h1
//        rows 0-7 x inputs -> dW1, rows 0-7 x ONE -> db1, rows 8-15 x ONE -> db2, rows 8-15 x h1 -> dW2
//   A2 = [dout (HALF rows, padded)]            x  B tiles: h2 -> dW3, ONE -> db3
template <int HALF, int CP, int T0, int NTG>
__device__ __forceinline__ void a1_tile_group(const float* __restrict__ s_tile, float* __restrict__ acc_f, float* __restrict__ d1row_f,
                                              int k0) {
    using R = Rows<HALF, CP>;
    constexpr int IN = HALF + CP, NIN = (IN + 7) / 8;
    constexpr int B1 = H * IN, W2 = B1 + H, B2 = W2 + H * H;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, col = 2 * t;
    int rowB[NTG];
#pragma unroll
    for (int n = 0; n < NTG; ++n) {
        const int ti = T0 + n;
        rowB[n] = ti < NIN ? ((8 * ti + g < IN) ? R::C + 8 * ti + g : R::ZERO) : (ti == NIN ? (g == 0 ? R::ONE : R::ZERO) : R::H1 + g);
    }
    float c[NTG][4] = {};
    mma_outer<NTG>(s_tile, R::D1, rowB, k0, k0 + 32, c);
#pragma unroll
    for (int n = 0; n < NTG; ++n) {
        const int ti = T0 + n;
        if (ti < NIN) {                                     // rows g (< 8): dW1[k = g][i]
            const int i = 8 * ti + col;
            if (i < IN) acc_f[g * IN + i] += c[n][0];
            if (i + 1 < IN) acc_f[g * IN + i + 1] += c[n][1];
        } else if (ti == NIN) {                             // ONE: col 0 -> db1 (rows 0-7), db2 (rows 8-15)
            if (t == 0) {
                acc_f[B1 + g] += c[n][0];
                if (d1row_f) d1row_f[g] += c[n][0];
                acc_f[B2 + g] += c[n][2];
            }
        } else {                                            // h1: rows 8-15 -> dW2[j = g][k]
            acc_f[W2 + g * H + col] += c[n][2];
            acc_f[W2 + g * H + col + 1] += c[n][3];
        }
    }
}

template <int HALF, int CP, int T0>
__device__ __forceinline__ void a1_tiles(const float* __restrict__ s_tile, float* __restrict__ acc_f, float* __restrict__ d1row_f, int k0) {
    constexpr int NT1 = (HALF + CP + 7) / 8 + 2;            // input tiles + ONE + h1
    if constexpr (T0 < NT1) {
        constexpr int NTG = NT1 - T0 < 4 ? NT1 - T0 : 4;    // at most four accumulator fragments live at a time
        a1_tile_group<HALF, CP, T0, NTG>(s_tile, acc_f, d1row_f, k0);
        a1_tiles<HALF, CP, T0 + NTG>(s_tile, acc_f, d1row_f, k0);
    }
}

template <int HALF, int CP>
__device__ __forceinline__ void stage_weight_grads_mma(const float* __restrict__ s_tile, float* __restrict__ acc_f,
                                                       float* __restrict__ d1row_f) {
    using R = Rows<HALF, CP>;
    constexpr int IN = HALF + CP;
    constexpr int W3 = H * IN + H + H * H + H, B3 = W3 + HALF * H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3, col = 2 * t;
    const int k0 = 32 * warp;
    a1_tiles<HALF, CP, 0>(s_tile, acc_f, d1row_f, k0);
    float c[2][4] = {};
    const int rowB[2] = {R::H2 + g, g == 0 ? R::ONE : R::ZERO};
    mma_outer<2>(s_tile, R::DO, rowB, k0, k0 + 32, c);
    if (g < HALF) { acc_f[W3 + g * H + col] += c[0][0]; acc_f[W3 + g * H + col + 1] += c[0][1]; }                       // dW3[o][j]
    if (g + 8 < HALF) { acc_f[W3 + (g + 8) * H + col] += c[0][2]; acc_f[W3 + (g + 8) * H + col + 1] += c[0][3]; }
    if (t == 0) {                                                                                                         // db3[o]
        if (g < HALF) acc_f[B3 + g] += c[1][0];
        if (g + 8 < HALF) acc_f[B3 + g + 8] += c[1][2];
    }
}

// gradient entries of layers 1 and 2 were accumulated from delta / TANH_SCALE: factor to apply when they are written out
template <int HALF, int CP>
__device__ __forceinline__ float grad_out_scale(int e) { return e < H * (HALF + CP) + H + H * H + H ? TANH_SCALE : 1.0f; }

// packed-FCNN offset (row-context columns skipped) of gradient entry e, e in [0, NOUT): the order of acc_f above
template <int HALF, int CP>
__device__ int packed_offset(int e, int C_row) {
    const int fin = HALF + C_row + CP;
    if (e < H * (HALF + CP)) {
        const int k = e / (HALF + CP), i = e % (HALF + CP);
        return k * fin + (i < HALF ? i : C_row + i);
    }
    return e - H * (HALF + CP) + H * fin;   // b1, W2, b2, W3, b3 follow contiguously in both layouts
}

template <int HALF, int CP>
struct BwdSmem {
    using L = Lay<HALF, CP>;
    using R = Rows<HALF, CP>;
    static size_t bytes(int n_fcnn, int C_row) {
        size_t fl = (size_t)n_fcnn * L::SIZE + n_fcnn * H + (size_t)n_fcnn * H * C_row  // images, hb, w1r
                    + (size_t)R::TROWS * TSM                                             // tile
                    + (size_t)(TP / 32) * n_fcnn * R::NOUT                               // acc (one copy per warp)
                    + (size_t)n_fcnn * H * C_row                                         // accR
                    + (TP / 32) * n_fcnn * H                                             // d1row (one copy per warp)
                    + C_row + 4;                                                         // ctx
        return fl * sizeof(float);
    }
};

// Backward of one stage for the particle held by this thread, plus the CTA-wide weight-gradient accumulation.
// On entry (c, v) are the stage's OUTPUT values with gradients (gc, gv); on exit v / gv are the stage's input.
// Where a stage's weight-gradient products go: per-warp shared-memory accumulator copies (the generic backward) ...
struct SmemGradSink {
    float* acc;      // this warp's copy [n_fcnn][NOUT]
    float* d1row;    // this warp's layer-1 delta sums [n_fcnn][8]
    template <int HALF, int CP>
    __device__ __forceinline__ void accumulate(const float* __restrict__ s_tile, int f) {
        stage_weight_grads_mma<HALF, CP>(s_tile, acc + f * Rows<HALF, CP>::NOUT, d1row + f * H);
    }
};
// ... or any type with the same accumulate<HALF, CP>(tile, net) member (measure.cu keeps the fragments in tensor memory).

template <int HALF, int CP, class Sink>
__device__ __forceinline__ void stage_bwd(const float* img_t, const float* img_s, const float* hb_t, const float* hb_s, int f_t,
                                          bool inv, bool live, const float (&c)[HALF], float (&gc)[HALF], const float* pc, float* gpc,
                                          float (&v)[HALF], float (&gv)[HALF], float gld, float* s_tile, Sink& sink) {
    using R = Rows<HALF, CP>;
    using L = Lay<HALF, CP>;
    const int tid = threadIdx.x;
    float h1t[H], h2t[H], h1s[H], h2s[H], t[HALF], s[HALF], dt[HALF], ds[HALF];
    fcnn_fwd<HALF, CP>(img_t, hb_t, c, pc, h1t, h2t, t);
    fcnn_fwd<HALF, CP>(img_s, hb_s, c, pc, h1s, h2s, s);
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const float es = expf(s[i]), ies = expf(-s[i]);
        if (!inv) {  // out = t + in*e^s
            const float vin = (v[i] - t[i]) * ies;
            dt[i] = gv[i];
            ds[i] = fmaf(gv[i] * vin, es, gld);
            gv[i] = gv[i] * es;
            v[i] = vin;
        } else {     // out = (in - t) e^{-s}
            const float gin = gv[i] * ies;
            dt[i] = -gin;
            ds[i] = -fmaf(gv[i], v[i], gld);
            v[i] = fmaf(v[i], es, t[i]);
            gv[i] = gin;
        }
        if (!live) { dt[i] = 0.f; ds[i] = 0.f; }
    }
    // conditioning half is shared by both nets of the stage: stage it once
#pragma unroll
    for (int i = 0; i < HALF; ++i) s_tile[(R::C + i) * TSM + tid] = c[i];
    // the two nets share ONE copy of the backward code (rolled loop, operands selected): for the wide D = 32 stack the
    // unrolled version is > 18 k SASS instructions and the kernel stalls on instruction fetch
#pragma unroll 1
    for (int net = 0; net < 2; ++net) {
        float d1[H], d2[H], ha[H], hb2[H], dout[HALF];
#pragma unroll
        for (int k = 0; k < H; ++k) { ha[k] = net ? h1s[k] : h1t[k]; hb2[k] = net ? h2s[k] : h2t[k]; }
#pragma unroll
        for (int i = 0; i < HALF; ++i) dout[i] = net ? ds[i] : dt[i];
        fcnn_bwd<HALF, CP>(net ? img_s : img_t, dout, ha, hb2, d1, d2, gc, gpc);
#pragma unroll
        for (int k = 0; k < H; ++k) {
            s_tile[(R::H1 + k) * TSM + tid] = ha[k];
            s_tile[(R::H2 + k) * TSM + tid] = hb2[k];
            s_tile[(R::D1 + k) * TSM + tid] = d1[k];
            s_tile[(R::D2 + k) * TSM + tid] = d2[k];
        }
#pragma unroll
        for (int i = 0; i < HALF; ++i) s_tile[(R::DO + i) * TSM + tid] = dout[i];
        __syncwarp();   // the warp contracts over its own 32 columns only: no CTA barrier in the gradient phase
        const int f = f_t + net;
        // b1 slots (first after the W1 block) double as the per-trajectory layer-1 delta sums (row-context hoist)
        sink.template accumulate<HALF, CP>(s_tile, f);
        __syncwarp();
    }
    (void)sizeof(L);
}


}  // namespace nfdpf
