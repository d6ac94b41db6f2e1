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

namespace nfdpf {

constexpr int H = NFDPF_HIDDEN;  // 8

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
        img[L::W1 + e] = v;
    }
    for (int e = tid; e < H * C_row; e += nt) w1r[e] = pk[(e / C_row) * fin + HALF + (e % C_row)];
    const float* p = pk + H * fin;
    for (int e = tid; e < H; e += nt) img[L::B1 + e] = p[e];
    p += H;
    for (int e = tid; e < H * H; e += nt) img[L::W2 + e] = p[e];
    p += H * H;
    for (int e = tid; e < H; e += nt) img[L::B2 + e] = p[e];
    p += H;
    for (int e = tid; e < HALF * H; e += nt) img[L::W3 + e] = p[e];
    p += HALF * H;
    for (int e = tid; e < HALF; e += nt) img[L::B3 + e] = p[e];
}

__device__ __forceinline__ void ld8(const float* p, float (&w)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// FCNN forward.  hb = hoisted layer-1 bias (b1 + row-context contribution).  Keeps h1/h2 for the backward.
template <int HALF, int CP>
__device__ __forceinline__ void fcnn_fwd(const float* __restrict__ img, const float* __restrict__ hb, const float (&c)[HALF],
                                         const float* pc, float (&h1)[H], float (&h2)[H], float (&out)[HALF]) {
    using L = Lay<HALF, CP>;
    float hbv[8];
    ld8(hb, hbv);
    float in[L::IN1];
#pragma unroll
    for (int i = 0; i < HALF; ++i) in[i] = c[i];
#pragma unroll
    for (int i = 0; i < CP; ++i) in[HALF + i] = pc[i];
#pragma unroll
    for (int k = 0; k < H; ++k) {
        float a = hbv[k];
        const float* w = img + L::W1 + k * L::S1;
#pragma unroll
        for (int i = 0; i < L::S1; i += 4) {  // rows are zero-padded to a multiple of 4
            const float4 w4 = *reinterpret_cast<const float4*>(w + i);
            if (i + 0 < L::IN1) a = fmaf(w4.x, in[i + 0 < L::IN1 ? i + 0 : 0], a);
            if (i + 1 < L::IN1) a = fmaf(w4.y, in[i + 1 < L::IN1 ? i + 1 : 0], a);
            if (i + 2 < L::IN1) a = fmaf(w4.z, in[i + 2 < L::IN1 ? i + 2 : 0], a);
            if (i + 3 < L::IN1) a = fmaf(w4.w, in[i + 3 < L::IN1 ? i + 3 : 0], a);
        }
        h1[k] = tanh_acc(a);
    }
    float b2[8];
    ld8(img + L::B2, b2);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
        float a = b2[j];
#pragma unroll
        for (int k = 0; k < H; ++k) a = fmaf(w[k], h1[k], a);
        h2[j] = tanh_acc(a);
    }
#pragma unroll
    for (int o = 0; o < HALF; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
        float a = img[L::B3 + o];
#pragma unroll
        for (int j = 0; j < H; ++j) a = fmaf(w[j], h2[j], a);
        out[o] = a;
    }
}

// FCNN backward (data path): dout -> d2, d1 (pre-activation grads), dc += W1c^T d1, dpc += W1p^T d1.
template <int HALF, int CP>
__device__ __forceinline__ void fcnn_bwd(const float* __restrict__ img, const float (&dout)[HALF], const float (&h1)[H],
                                         const float (&h2)[H], float (&d1)[H], float (&d2)[H], float (&dc)[HALF], float* dpc) {
    using L = Lay<HALF, CP>;
    float da2[H];
#pragma unroll
    for (int j = 0; j < H; ++j) da2[j] = 0.f;
#pragma unroll
    for (int o = 0; o < HALF; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
#pragma unroll
        for (int j = 0; j < H; ++j) da2[j] = fmaf(w[j], dout[o], da2[j]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) d2[j] = da2[j] * fmaf(-h2[j], h2[j], 1.0f);
    float da1[H];
#pragma unroll
    for (int k = 0; k < H; ++k) da1[k] = 0.f;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
#pragma unroll
        for (int k = 0; k < H; ++k) da1[k] = fmaf(w[k], d2[j], da1[k]);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) d1[k] = da1[k] * fmaf(-h1[k], h1[k], 1.0f);
#pragma unroll
    for (int k = 0; k < H; ++k) {
        const float* w = img + L::W1 + k * L::S1;
#pragma unroll
        for (int i = 0; i < HALF; ++i) dc[i] = fmaf(w[i], d1[k], dc[i]);
        if constexpr (CP > 0) {
#pragma unroll
            for (int i = 0; i < CP; ++i) dpc[i] = fmaf(w[HALF + i], d1[k], dpc[i]);
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

// One coupling stage, forward evaluation.  INV=false: v = t(c) + v*exp(s(c)); INV=true: v = (v - t(c))*exp(-s(c)).
template <int HALF, int CP, bool INV>
__device__ __forceinline__ void stage_fwd(const float* img_t, const float* img_s, const float* hb_t, const float* hb_s,
                                          const float (&c)[HALF], const float* pc, float (&v)[HALF], float& ld) {
    float h1[H], h2[H], t[HALF], s[HALF];
    fcnn_fwd<HALF, CP>(img_t, hb_t, c, pc, h1, h2, t);
    fcnn_fwd<HALF, CP>(img_s, hb_s, c, pc, h1, h2, s);
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        if (!INV) { v[i] = fmaf(v[i], expf(s[i]), t[i]); ld += s[i]; }
        else      { v[i] = (v[i] - t[i]) * expf(-s[i]); ld -= s[i]; }
    }
}

}  // namespace nfdpf
