// (K2) Measurement log-likelihood fused with the log-weight update and the softmax normalisation.
// One CTA per trajectory, one thread per particle:
//   particle encoder MLP 2-16-32-32 (model/models.py:130-139)
//   -> Gaussian (models.py:237-254) | cosine (206-219) | conditional-RealNVP (256-278, D=32, C=32) log-likelihood
//   -> row max shift -> logw = logw_prev + lki + prior - propose (DPFs.py:187) -> softmax + eps, ESS term, sum logw.
// The reference materialises the (B,N,32) encodings, the repeated (B,N,32) observation encodings and, for CRNVP,
// eight (P,48) concatenations; here the only HBM traffic is particles in, (B,N) vectors out.
// Tensor cores (umma.cuh): encoder layers 2-3 (forward and data gradient) and layer 1 of the CRNVP stages (forward, and in the
// warp-specialised backward also its recompute and W1^T delta1) run as 3xTF32 tcgen05 products over batches of 128 particles
// (thread = particle = tensor-memory lane); the backward's weight gradients are mma.sync 3xTF32 contractions whose fragments
// accumulate in tensor memory.  Layer 1 of the encoder, the 8-wide layers of the stack, activations and the likelihood algebra
// stay in registers.
#include "coupling.cuh"
#include "mma_tile.cuh"
#include "umma.cuh"

#include <stdlib.h>

#include <type_traits>

namespace nfdpf {

enum { MODE_GAUSS = 0, MODE_COS = 1, MODE_CNF = 2, MODE_NN = 3 };
constexpr int FWD_TMEM_COLS = 128;   // accumulator 32 + activation hi 32 + lo 32 columns, rounded up to a power of two
constexpr int FWD_TMEM_COLS_NN = 256;   // NN likelihood head: accumulator 64 + activation hi 64 + lo 64 columns
// packed likelihood head (build_likelihood, model/models.py:119-128; state_dict order 0.weight,0.bias,2.weight,2.bias,4.weight,4.bias)
constexpr int NH = 64, NH_W1 = 0, NH_B1 = NH_W1 + NH * NH, NH_W2 = NH_B1 + NH, NH_B2 = NH_W2 + NH * NH, NH_W3 = NH_B2 + NH, NH_B3 = NH_W3 + NH,
              NH_SIZE = NH_B3 + 1;
constexpr int HID = 32;                      // encoding width (args.hiddensize)
// packed particle encoder (state_dict order 0.weight,0.bias,2.weight,2.bias,4.weight,4.bias)
constexpr int PE_W1 = 0, PE_B1 = 32, PE_W2 = 48, PE_B2 = 560, PE_W3 = 592, PE_B3 = 1616, PE_SIZE = 1648;
using LC = Lay<16, 32>;
using RC = Rows<16, 32>;

// ---- particle encoder layers 2 and 3 on the tcgen05 tensor cores (umma.cuh) ------------------------------------
// M = the CTA's 128 particles (TMEM lane = particle = thread), N / K = 16 / 32 features, 3xTF32.  Layer 1 (2 -> 16) and
// the bias / ReLU epilogues stay in registers.  The activation operand lives in TENSOR MEMORY: a thread splits its row into
// TF32 hi / lo and writes both with tcgen05.st into its own lane (columns [32,64) hi, [64,96) lo; the accumulator is
// columns [0,32)), so activations never touch shared memory; only the weight tiles (hi | lo, both orientations in the
// backward) are shared-memory operands.
struct PeTc {
    using W2 = umma::Operand<32, 16>;    // rows j (out), K = k (in)          a2 = W2 a1
    using W3 = umma::Operand<32, 32>;    // rows o,       K = j               e  = W3 a2
    using W3T = umma::Operand<32, 32>;   // rows j,       K = o               d a2 = W3^T delta3
    using W2T = umma::Operand<16, 32>;   // rows k,       K = j               d a1 = W2^T delta2
    static constexpr int W2_HI = 0, W2_LO = W2_HI + W2::FLOATS, W3_HI = W2_LO + W2::FLOATS, W3_LO = W3_HI + W3::FLOATS,  // offsets from w
                         WFWD_FLOATS = W3_LO + W3::FLOATS, W3T_HI = WFWD_FLOATS, W3T_LO = W3T_HI + W3T::FLOATS,
                         W2T_HI = W3T_LO + W3T::FLOATS, W2T_LO = W2T_HI + W2T::FLOATS, WBWD_FLOATS = W2T_LO + W2T::FLOATS;
    static constexpr int COL_D = 0, COL_AHI = 32, COL_ALO = 64, COLS = 96;     // tensor-memory columns
    float* w;            // weight tiles (128-byte aligned)
    uint64_t* bar;       // mbarrier: completion of the issued MMAs
    uint32_t tmem;       // base of the CTA's tensor-memory allocation
    uint32_t parity;

    // split the layer-2 / layer-3 weights into the hi / lo operand tiles (all threads; followed by a CTA barrier in the caller)
    __device__ void load_weights(const float* __restrict__ pe, bool bwd) {
        for (int e = threadIdx.x; e < 32 * 16; e += blockDim.x) {
            const int j = e >> 4, k = e & 15;
            const float wv = pe[PE_W2 + e];
            W2::store_elem(w + W2_HI, w + W2_LO, j, k, wv);
            if (bwd) W2T::store_elem(w + W2T_HI, w + W2T_LO, k, j, wv);
        }
        for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
            const int o = e >> 5, j = e & 31;
            const float wv = pe[PE_W3 + e];
            W3::store_elem(w + W3_HI, w + W3_LO, o, j, wv);
            if (bwd) W3T::store_elem(w + W3T_HI, w + W3T_LO, j, o, wv);
        }
        umma::fence_smem_to_async();     // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
    }
    __device__ __forceinline__ uint32_t lane_addr() const { return tmem + ((uint32_t)(threadIdx.x & ~31) << 16); }
    // this thread's activation row -> its TMEM lane (hi and lo copies)
    template <int K>
    __device__ __forceinline__ void store_row(const float (&v)[K]) {
        float hi[K], lo[K];
#pragma unroll
        for (int k = 0; k < K; ++k) umma::split(v[k], hi[k], lo[k]);
        umma::st_frag<K>(lane_addr() + COL_AHI, hi);
        umma::st_frag<K>(lane_addr() + COL_ALO, lo);
    }
    // One product round, called by ALL 128 threads after store_row: one thread issues the 3xTF32 MMAs, everybody waits for them.
    template <int N, int K>
    __device__ __forceinline__ void round(int w_hi, int w_lo) {
        umma::wait_st();                // this thread's activation row has landed in tensor memory
        umma::fence_before_sync();      // ... and is ordered, like its earlier tcgen05.ld of the accumulator, before the barrier
        __syncthreads();
        if (threadIdx.x == 0) {
            umma::fence_after_sync();
            umma::gemm3_ts<N, K>(tmem + COL_D, tmem + COL_AHI, tmem + COL_ALO, w + w_hi, w + w_lo);
            umma::commit(bar);
        }
        wait();
    }
    __device__ __forceinline__ void wait() {
        umma::mbar_wait(bar, parity);
        parity ^= 1;
        umma::fence_after_sync();
    }
};

// Twin of PeTc with the activation operand in SHARED memory (SS form: 32 tensor-memory columns instead of 96).  The CRNVP
// backward keeps 212 columns of gradient fragments per CTA in tensor memory and still wants two CTAs per SM (512 columns).
struct PeSs {
    using A32 = umma::Operand<128, 32>;
    static constexpr int A_HI = 0, A_LO = A32::FLOATS, A_FLOATS = 2 * A32::FLOATS;
    static constexpr int COL_D = 0, COLS = 32;
    float* a;            // activation tile hi | lo (128-byte aligned); chunk stride is 128 rows x 16 B for every K
    float* w;
    uint64_t* bar;
    uint32_t tmem;
    uint32_t parity;
    __device__ void load_weights(const float* __restrict__ pe, bool bwd) {
        PeTc t{w, bar, tmem, parity};
        t.load_weights(pe, bwd);
    }
    __device__ __forceinline__ uint32_t lane_addr() const { return tmem + ((uint32_t)(threadIdx.x & ~31) << 16); }
    template <int K>
    __device__ __forceinline__ void store_row(const float (&v)[K]) { umma::Operand<128, K>::store_row(a + A_HI, a + A_LO, threadIdx.x, v); }
    template <int N, int K>
    __device__ __forceinline__ void round(int w_hi, int w_lo) {
        umma::fence_smem_to_async();
        umma::fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            umma::fence_after_sync();
            umma::gemm3<N, K>(tmem + COL_D, a + A_HI, a + A_LO, w + w_hi, w + w_lo);
            umma::commit(bar);
        }
        wait();
    }
    __device__ __forceinline__ void wait() {
        umma::mbar_wait(bar, parity);
        parity ^= 1;
        umma::fence_after_sync();
    }
};

__device__ __forceinline__ void pe_l1(const float* __restrict__ w, float x0, float x1, float (&a1)[16]) {
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
        const float4 q = *reinterpret_cast<const float4*>(w + PE_W1 + 2 * k);
        const float2 b = *reinterpret_cast<const float2*>(w + PE_B1 + k);
        a1[k] = fmaxf(fmaf(q.y, x1, fmaf(q.x, x0, b.x)), 0.f);
        a1[k + 1] = fmaxf(fmaf(q.w, x1, fmaf(q.z, x0, b.y)), 0.f);
    }
}

// encoder forward for the particle of this thread; collective over the CTA (contains barriers)
template <class Tc>
__device__ __forceinline__ void pe_fwd_tc(Tc& tc, const float* __restrict__ w, float x0, float x1, float (&a1)[16], float (&a2)[32],
                                          float (&e)[32]) {
    pe_l1(w, x0, x1, a1);
    tc.template store_row<16>(a1);
    tc.template round<32, 16>(PeTc::W2_HI, PeTc::W2_LO);
    umma::ld32(tc.lane_addr(), a2);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(w + PE_B2 + j);
        a2[j] = fmaxf(a2[j] + b.x, 0.f); a2[j + 1] = fmaxf(a2[j + 1] + b.y, 0.f);
        a2[j + 2] = fmaxf(a2[j + 2] + b.z, 0.f); a2[j + 3] = fmaxf(a2[j + 3] + b.w, 0.f);
    }
    tc.template store_row<32>(a2);
    tc.template round<32, 32>(PeTc::W3_HI, PeTc::W3_LO);
    umma::ld32(tc.lane_addr(), e);
#pragma unroll
    for (int o = 0; o < 32; o += 4) {
        const float4 b = *reinterpret_cast<const float4*>(w + PE_B3 + o);
        e[o] += b.x; e[o + 1] += b.y; e[o + 2] += b.z; e[o + 3] += b.w;
    }
}

__device__ __forceinline__ void pe_fwd(const float* __restrict__ w, float x0, float x1, float (&a1)[16], float (&a2)[32],
                                       float (&e)[32]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a1[k] = fmaxf(fmaf(w[PE_W1 + 2 * k + 1], x1, fmaf(w[PE_W1 + 2 * k], x0, w[PE_B1 + k])), 0.f);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float a = w[PE_B2 + j];
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            const float4 q = *reinterpret_cast<const float4*>(w + PE_W2 + j * 16 + k);
            a = fmaf(q.x, a1[k], a); a = fmaf(q.y, a1[k + 1], a); a = fmaf(q.z, a1[k + 2], a); a = fmaf(q.w, a1[k + 3], a);
        }
        a2[j] = fmaxf(a, 0.f);
    }
#pragma unroll
    for (int o = 0; o < 32; ++o) {
        float a = w[PE_B3 + o];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 q = *reinterpret_cast<const float4*>(w + PE_W3 + o * 32 + j);
            a = fmaf(q.x, a2[j], a); a = fmaf(q.y, a2[j + 1], a); a = fmaf(q.z, a2[j + 2], a); a = fmaf(q.w, a2[j + 3], a);
        }
        e[o] = a;
    }
}

// layers 1-2 only (the Gaussian backward fuses layer 3 forward with its transposed backward in one rolled loop)
__device__ __forceinline__ void pe_fwd_l12(const float* __restrict__ w, float x0, float x1, float (&a1)[16], float (&a2)[32]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a1[k] = fmaxf(fmaf(w[PE_W1 + 2 * k + 1], x1, fmaf(w[PE_W1 + 2 * k], x0, w[PE_B1 + k])), 0.f);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float a = w[PE_B2 + j];
#pragma unroll
        for (int k = 0; k < 16; k += 4) {
            const float4 q = *reinterpret_cast<const float4*>(w + PE_W2 + j * 16 + k);
            a = fmaf(q.x, a1[k], a); a = fmaf(q.y, a1[k + 1], a); a = fmaf(q.z, a1[k + 2], a); a = fmaf(q.w, a1[k + 3], a);
        }
        a2[j] = fmaxf(a, 0.f);
    }
}

// log-likelihood of one particle given its encoding e; for MODE_CNF also returns z (= [lo|up]) for the backward.
template <int MODE>
__device__ __forceinline__ float loglik(const float (&e)[32], const float* __restrict__ s_enc, float p0, float p1,
                                        const float* s_img, const float* s_hb, int n_flows, float (&lo)[16], float (&up)[16]) {
    if (MODE == MODE_GAUSS) {   // MVN(loc = p0, cov = p1^2 I).log_prob(enc - e)
        float m = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) { const float u = s_enc[k] - e[k] - p0; m = fmaf(u, u, m); }
        return -0.5f * m / (p1 * p1) - 32.0f * (logf(p1) + 0.91893853320467274f);
    } else if (MODE == MODE_COS) {  // log(1 / (1e-7 + 1 - cos(enc, e))), F.normalize eps 1e-12
        float ne = 0.f, dot = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) { ne = fmaf(e[k], e[k], ne); dot = fmaf(s_enc[k], e[k], dot); }
        const float cosd = 1.0f - dot / (fmaxf(sqrtf(ne), 1e-12f) * s_enc[32]);
        return -logf(1e-7f + cosd);
    } else {                    // conditional RealNVP: x = enc (row constant), context = e; prior N(p0, p1^2 I)
        float ld = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { lo[i] = s_enc[i]; up[i] = s_enc[16 + i]; }
#pragma unroll 1
        for (int st = 0; st < 2 * n_flows; ++st) {        // one inlined stage body; the halves swap roles after every stage
            const float* im = s_img + 2 * st * LC::SIZE;
            const float* hb = s_hb + 2 * st * H;
            stage_fwd<16, 32>(im, im + LC::SIZE, hb, hb + H, false, lo, e, up, ld);
            swap_halves<16>(lo, up);
        }
        float m = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float a = lo[i] - p0, b = up[i] - p0; m = fmaf(a, a, m); m = fmaf(b, b, m); }
        return -0.5f * m / (p1 * p1) - 32.0f * (logf(p1) + 0.91893853320467274f) + ld;
    }
}

// ---- CRNVP stack, forward: layer 1 of a stage's t- and s-net on tcgen05 ----------------------------------------------
// The two nets of a stage read the same 48 inputs [c (16) | particle encoding e (32)]; their layer-1 weights side by side are a
// [16 x 48] operand, i.e. an M = 128 particles, N = 16, K = 48 product per stage: 2/3 of the stack's FMAs.  The activation row
// lives in tensor memory like the encoder's (TS form): e is written once per particle (columns [48,80) hi, [96,128) lo), c every
// stage ([32,48) hi, [80,96) lo); the 16 pre-activations come back in columns [0,16).  Layers 2 and 3 (8 x 8, 8 -> 16) stay in
// registers.  CnfL1 = the per-stage weight tiles in shared memory (hi | lo, chunk-major K-major, umma.cuh).
struct CnfL1 {
    using W = umma::Operand<16, 48>;
    static constexpr int STAGE_FLOATS = 2 * W::FLOATS;     // hi | lo
    static constexpr int COL_D = 0, COL_CHI = 32, COL_EHI = 48, COL_CLO = 80, COL_ELO = 96;
    // The forward keeps only the TAIL of every FCNN image in shared memory (b1 | W2 | b2 | W3 | b3, laid out exactly as Lay<16,32>
    // from its B1 offset on): the layer-1 weights live in the operand tiles.  tail_image() returns the pointer that makes the LC
    // offsets valid (its W1 range is never dereferenced).
    static constexpr int TAIL = LC::SIZE - LC::B1;         // 224 floats per net instead of 608
    __device__ static __forceinline__ const float* tail_image(const float* s_tail, int f) { return s_tail + f * TAIL - LC::B1; }
    // build the tail images and the operand tiles (rows 0-7 t-net units, 8-15 s-net units) from the packed stack; layers 1-2
    // carry the tanh pre-scale (coupling.cuh)
    __device__ static void load(const float* __restrict__ cnf, int n_fcnn, float* __restrict__ s_tail, float* __restrict__ s_hb,
                                float* __restrict__ s_w) {
        const int pf = packed_fcnn_size(16, 32);
        for (int e = threadIdx.x; e < n_fcnn * TAIL; e += blockDim.x) {
            const int f = e / TAIL, o = e % TAIL;
            float v = 0.f;
            if (o < pf - H * 48) v = cnf[(size_t)f * pf + H * 48 + o];        // b1 (8) W2 (64) b2 (8) W3 (128) b3 (16): contiguous in both layouts
            if (o < H + H * H + H) v *= TANH_SCALE;
            s_tail[e] = v;
            if (o < H) s_hb[f * H + o] = v;                                       // no row context here: hoisted bias = scaled b1
        }
        for (int e = threadIdx.x; e < (n_fcnn / 2) * 16 * 48; e += blockDim.x) {
            const int st = e / (16 * 48), r = (e / 48) % 16, k = e % 48;
            const float w = TANH_SCALE * cnf[(size_t)(2 * st + (r >> 3)) * pf + (r & 7) * 48 + k];
            W::store_elem(s_w + st * STAGE_FLOATS, s_w + st * STAGE_FLOATS + W::FLOATS, r, k, w);
        }
        umma::fence_smem_to_async();
    }
    // one round: this thread's c half -> tensor memory, 18 MMAs (3xTF32, K = 48), pre-activations back; CTA-collective
    __device__ static __forceinline__ void round(PeTc& tc, const float* __restrict__ s_w_stage, const float (&c)[16], float (&pre)[16]) {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) umma::split(c[i], hi[i], lo[i]);
        umma::st_frag<16>(tc.lane_addr() + COL_CHI, hi);
        umma::st_frag<16>(tc.lane_addr() + COL_CLO, lo);
        umma::wait_st();
        umma::fence_before_sync();
        __syncthreads();
        if (threadIdx.x == 0) {
            umma::fence_after_sync();
            constexpr uint32_t idesc = umma::idesc_tf32(128, 16);
            const float* w_hi = s_w_stage;
            const float* w_lo = s_w_stage + W::FLOATS;
            uint32_t acc = 0;
#pragma unroll
            for (int k0 = 0; k0 < 48; k0 += 8) { umma::mma_tf32_ts(tc.tmem + COL_D, tc.tmem + COL_CLO + k0, W::desc(w_hi, k0), idesc, acc); acc = 1; }
#pragma unroll
            for (int k0 = 0; k0 < 48; k0 += 8) umma::mma_tf32_ts(tc.tmem + COL_D, tc.tmem + COL_CHI + k0, W::desc(w_lo, k0), idesc, 1);
#pragma unroll
            for (int k0 = 0; k0 < 48; k0 += 8) umma::mma_tf32_ts(tc.tmem + COL_D, tc.tmem + COL_CHI + k0, W::desc(w_hi, k0), idesc, 1);
            umma::commit(tc.bar);
        }
        tc.wait();
        umma::ld16(tc.lane_addr() + COL_D, pre);
    }
};

// layers 2 and 3 of one net from its layer-1 PRE-activations (in place: a1 -> h1); same arithmetic as fcnn_fwd<16, 32, true>
__device__ __forceinline__ void fcnn_tail16(const float* __restrict__ img, float (&a1)[H], float (&out)[16]) {
    using L = LC;
#pragma unroll
    for (int k = 0; k < H; k += 2) tanh_prescaled_pair(a1[k], a1[k + 1], a1[k], a1[k + 1]);
    float b2[8], h2[H];
    ld8(img + L::B2, b2);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
        float p0 = b2[j], p1 = 0.f;
#pragma unroll
        for (int k = 0; k < H; k += 2) ffma2_p(p0, p1, w[k], w[k + 1], a1[k], a1[k + 1]);
        h2[j] = p0 + p1;
    }
#pragma unroll
    for (int j = 0; j < H; j += 2) tanh_prescaled_pair(h2[j], h2[j + 1], h2[j], h2[j + 1]);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
        float p0 = img[L::B3 + o], p1 = 0.f;
#pragma unroll
        for (int j = 0; j < H; j += 2) ffma2_p(p0, p1, w[j], w[j + 1], h2[j], h2[j + 1]);
        out[o] = p0 + p1;
    }
}

// CRNVP log-likelihood of the particle of this thread (x = observation encoding, context = e), layer 1 on the tensor cores.
// CTA-collective: every thread of the CTA must call it (dead threads with any finite e).
__device__ __forceinline__ float loglik_cnf_tc(PeTc& tc, const float (&e)[32], const float* __restrict__ s_enc, float p0, float p1,
                                               const float* __restrict__ s_img /* tail images */, const float* __restrict__ s_hb,
                                               const float* __restrict__ s_l1w, int n_flows, float (&lo)[16], float (&up)[16]) {
    {
        float hi[32], l[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) umma::split(e[i], hi[i], l[i]);
        umma::st_frag<32>(tc.lane_addr() + CnfL1::COL_EHI, hi);
        umma::st_frag<32>(tc.lane_addr() + CnfL1::COL_ELO, l);
    }
    float ld = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo[i] = s_enc[i]; up[i] = s_enc[16 + i]; }
#pragma unroll 1
    for (int st = 0; st < 2 * n_flows; ++st) {        // one stage body; the halves swap roles after every stage
        float pre[16];
        CnfL1::round(tc, s_l1w + st * CnfL1::STAGE_FLOATS, lo, pre);
        const float* hb = s_hb + 2 * st * H;
        float a_t[H], a_s[H], t[16], sc[16];
#pragma unroll
        for (int k = 0; k < H; ++k) { a_t[k] = pre[k] + hb[k]; a_s[k] = pre[H + k] + hb[H + k]; }
        fcnn_tail16(CnfL1::tail_image(s_img, 2 * st), a_t, t);
        fcnn_tail16(CnfL1::tail_image(s_img, 2 * st + 1), a_s, sc);
#pragma unroll
        for (int i = 0; i < 16; ++i) { up[i] = fmaf(up[i], exp_acc(sc[i]), t[i]); ld += sc[i]; }
        swap_halves<16>(lo, up);
    }
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const float a = lo[i] - p0, b = up[i] - p0; m = fmaf(a, a, m); m = fmaf(b, b, m); }
    return -0.5f * m / (p1 * p1) - 32.0f * (logf(p1) + 0.91893853320467274f) + ld;
}

// ---- NN likelihood (measurement_model_NN, model/models.py:221-235): sigmoid MLP 64 -> 64 -> 64 -> 1 on [enc | e], then log ----------
// The observation half of layer 1 is row-constant: hb1 = b1 + W1[:, :32] enc is formed once per trajectory.  The particle half
// (M = 128 particles, N = 64, K = 32) and layer 2 (N = 64, K = 64) are tcgen05 TS rounds like the encoder's (3xTF32): the thread
// writes its activation row (hi | lo) into its own tensor-memory lane, columns [64,128) / [128,192), the accumulator is [0,64).
struct NnHead {
    using W1E = umma::Operand<NH, 32>;   // rows j (out), K = k over the particle-encoding half of the input
    using W2 = umma::Operand<NH, NH>;    // rows j, K = k
    static constexpr int W1E_HI = 0, W1E_LO = W1E_HI + W1E::FLOATS, W2_HI = W1E_LO + W1E::FLOATS, W2_LO = W2_HI + W2::FLOATS,
                         TILE_FLOATS = W2_LO + W2::FLOATS;
    static constexpr int HB1 = TILE_FLOATS, B2 = HB1 + NH, W3 = B2 + NH, B3 = W3 + NH, FLOATS = B3 + 4;   // vectors behind the tiles
    static constexpr int COL_D = 0, COL_AHI = 64, COL_ALO = 128;
    // all threads; followed by a CTA barrier in the caller.  enc_row: the trajectory's observation encoding (global)
    __device__ static void load(const float* __restrict__ head, const float* __restrict__ enc_row, float* s) {
        for (int e = threadIdx.x; e < NH * 32; e += blockDim.x) {
            const int j = e >> 5, k = e & 31;
            W1E::store_elem(s + W1E_HI, s + W1E_LO, j, k, head[NH_W1 + j * NH + 32 + k]);
        }
        for (int e = threadIdx.x; e < NH * NH; e += blockDim.x) W2::store_elem(s + W2_HI, s + W2_LO, e >> 6, e & 63, head[NH_W2 + e]);
        for (int j = threadIdx.x; j < NH; j += blockDim.x) {
            float a = head[NH_B1 + j];
            for (int c = 0; c < 32; ++c) a = fmaf(head[NH_W1 + j * NH + c], enc_row[c], a);
            s[HB1 + j] = a;
            s[B2 + j] = head[NH_B2 + j];
            s[W3 + j] = head[NH_W3 + j];
        }
        if (threadIdx.x == 0) s[B3] = head[NH_B3];
        umma::fence_smem_to_async();
    }
};

// One TS round of the head on the CTA's 128 lanes: A = the K values every thread just wrote, D[0,64) = A W^T (3xTF32).
template <int K>
__device__ __forceinline__ void nn_round(PeTc& tc, const float (&a)[K], const float* w_hi, const float* w_lo, float (&d)[NH]) {
    {
        float hi[K], lo[K];
#pragma unroll
        for (int k = 0; k < K; ++k) umma::split(a[k], hi[k], lo[k]);
        umma::st_frag<K>(tc.lane_addr() + NnHead::COL_AHI, hi);
        umma::st_frag<K>(tc.lane_addr() + NnHead::COL_ALO, lo);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        umma::gemm3_ts<NH, K>(tc.tmem + NnHead::COL_D, tc.tmem + NnHead::COL_AHI, tc.tmem + NnHead::COL_ALO, w_hi, w_lo);
        umma::commit(tc.bar);
    }
    tc.wait();
    float lo32[32], hi32[32];
    umma::ld32(tc.lane_addr() + NnHead::COL_D, lo32);
    umma::ld32(tc.lane_addr() + NnHead::COL_D + 32, hi32);
#pragma unroll
    for (int j = 0; j < 32; ++j) { d[j] = lo32[j]; d[32 + j] = hi32[j]; }
}

// CTA-collective (every thread calls it, dead threads with any finite e).  Returns log sigmoid(z) as the reference forms it.
__device__ __forceinline__ float loglik_nn_tc(PeTc& tc, const float (&e)[32], const float* __restrict__ s_nn) {
    float h[NH];
    nn_round<32>(tc, e, s_nn + NnHead::W1E_HI, s_nn + NnHead::W1E_LO, h);
#pragma unroll
    for (int j = 0; j < NH; ++j) h[j] = fmaxf(h[j] + s_nn[NnHead::HB1 + j], 0.f);
    float h2[NH];
    nn_round<NH>(tc, h, s_nn + NnHead::W2_HI, s_nn + NnHead::W2_LO, h2);
    float z = s_nn[NnHead::B3];
#pragma unroll
    for (int j = 0; j < NH; ++j) z = fmaf(s_nn[NnHead::W3 + j], fmaxf(h2[j] + s_nn[NnHead::B2 + j], 0.f), z);
    return logf(1.0f / (1.0f + expf(-z)));        // likelihood[..., 0].log() of a Sigmoid output, models.py:233-235
}

__device__ void load_enc(const float* __restrict__ enc_row, float* s_enc) {  // s_enc[32] = ||enc|| clamped (cos mode)
    if (threadIdx.x < 32) {
        const float v = enc_row[threadIdx.x];
        s_enc[threadIdx.x] = v;
        const float n2 = warp_sum(v * v);
        if (threadIdx.x == 0) s_enc[32] = fmaxf(sqrtf(n2), 1e-12f);
    }
}

template <int MODE>
__global__ void __launch_bounds__(TP)
measure_fwd_kernel(const float* __restrict__ pe, const float* __restrict__ cnf, int n_flows, float p0, float p1,
                   const float* __restrict__ enc, const float* __restrict__ particles, int N, const float* __restrict__ lw0,
                   const float* __restrict__ prior, const float* __restrict__ propose, float add_eps, float* __restrict__ lki,
                   int* __restrict__ argmax, float* __restrict__ logw_out, float* __restrict__ probs_out,
                   float* __restrict__ row_stats, float* __restrict__ z_out, float* __restrict__ pred_out) {
    extern __shared__ __align__(128) float smem[];
    __shared__ float s_red[33];
    __shared__ int s_redi[33];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tslot;
    const int tid = threadIdx.x, b = blockIdx.x, n_fcnn = MODE == MODE_CNF ? 4 * n_flows : 0;
    float* s_tc = smem;                       // tensor-core weight tiles (the activation operand lives in tensor memory)
    float* s_pe = s_tc + PeTc::WFWD_FLOATS;   // [1648]
    float* s_enc = s_pe + PE_SIZE;            // [36]
    float* s_img = s_enc + 36;                // [n_fcnn][CnfL1::TAIL]  CRNVP tail images (b1 | W2 | b2 | W3 | b3)
    float* s_hb = s_img + n_fcnn * CnfL1::TAIL;   // [n_fcnn][8]
    float* s_l1w = s_hb + n_fcnn * H;         // [n_fcnn / 2][CnfL1::STAGE_FLOATS]  layer-1 operand tiles of the CRNVP stages
    float* s_nn = s_l1w + (n_fcnn / 2) * CnfL1::STAGE_FLOATS;   // [NnHead::FLOATS]  NN likelihood head (mode 3)
    float* s_ll = s_nn + (MODE == MODE_NN ? NnHead::FLOATS : 0);   // [N]
    constexpr int TCOLS = MODE == MODE_NN ? FWD_TMEM_COLS_NN : FWD_TMEM_COLS;
    constexpr bool SHIFT = MODE == MODE_GAUSS || MODE == MODE_CNF;     // likelihood - likelihood.max(dim=-1), models.py:252, 276
    if (tid < 32) umma::tmem_alloc<TCOLS>(&s_tslot);
    if (tid == 0) umma::mbar_init(&s_bar, 1);
    PeTc tc{s_tc, &s_bar, 0u, 0u};
    tc.load_weights(pe, false);
    for (int e = tid; e < PE_SIZE; e += TP) s_pe[e] = pe[e];
    if (MODE == MODE_CNF) CnfL1::load(cnf, n_fcnn, s_img, s_hb, s_l1w);
    if (MODE == MODE_NN) NnHead::load(cnf, enc + (size_t)b * HID, s_nn);      // (the `cnf` slot carries the packed head)
    load_enc(enc + (size_t)b * HID, s_enc);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    tc.tmem = s_tslot;
    float mx = -INFINITY;
    float2 x_next = *reinterpret_cast<const float2*>(particles + ((size_t)b * N + (tid < N ? tid : 0)) * 2);
    for (int n0 = 0; n0 < N; n0 += TP) {      // uniform trip count: the tensor-core rounds are CTA-collective
        asm volatile("" ::: "memory");  // keep the shared-memory weight loads inside the loop
        const int n = n0 + tid;
        const bool live = n < N;
        const float2 x = x_next;
        x_next = *reinterpret_cast<const float2*>(particles + ((size_t)b * N + (n + TP < N ? n + TP : 0)) * 2);   // next batch: latency hidden
        float a1[16], a2[32], e[32], lo[16], up[16];
        pe_fwd_tc(tc, s_pe, x.x, x.y, a1, a2, e);
        float ll;
        if constexpr (MODE == MODE_CNF) ll = loglik_cnf_tc(tc, e, s_enc, p0, p1, s_img, s_hb, s_l1w, n_flows, lo, up);
        else if constexpr (MODE == MODE_NN) ll = loglik_nn_tc(tc, e, s_nn);
        else ll = loglik<MODE>(e, s_enc, p0, p1, s_img, s_hb, n_flows, lo, up);
        if (MODE == MODE_CNF && z_out && live) {   // the flow output: lets the backward walk the stack from z without re-running it
            float4* zo = reinterpret_cast<float4*>(z_out + ((size_t)b * N + n) * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                zo[i] = make_float4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
                zo[4 + i] = make_float4(up[4 * i], up[4 * i + 1], up[4 * i + 2], up[4 * i + 3]);
            }
        }
        if (live) { s_ll[n] = ll; mx = fmaxf(mx, ll); }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_free<TCOLS>(tc.tmem);
    float shift = 0.f;
    if (SHIFT) {
        shift = block_allreduce(mx, s_red, OpMax(), -INFINITY);
        int am = 0x7fffffff;
        for (int n = tid; n < N; n += TP) if (s_ll[n] == shift) am = min(am, n);
        am = block_allreduce(am, s_redi, [] __device__(int a, int c) { return min(a, c); }, 0x7fffffff);
        if (tid == 0 && argmax) argmax[b] = am;
    } else {
        __syncthreads();
        if (tid == 0 && argmax) argmax[b] = -1;
    }
    const size_t base = (size_t)b * N;
    if (!lw0) {
        for (int n = tid; n < N; n += TP) lki[base + n] = s_ll[n] - shift;
        return;
    }
    // fused weight update + normalisation (DPFs.py:187-192)
    float vmx = -INFINITY, sl = 0.f;
    for (int n = tid; n < N; n += TP) {
        const float l = s_ll[n] - shift;
        lki[base + n] = l;
        float v = lw0[base + n] + l;
        if (prior) v += prior[base + n];
        if (propose) v -= propose[base + n];
        if (logw_out) logw_out[base + n] = v;
        s_ll[n] = v;
        vmx = fmaxf(vmx, v);
        sl += v;
    }
    vmx = block_allreduce(vmx, s_red, OpMax(), -INFINITY);
    sl = block_allreduce(sl, s_red, OpSum(), 0.f);
    float se = 0.f;
    for (int n = tid; n < N; n += TP) { const float ex = expf(s_ll[n] - vmx); s_ll[n] = ex; se += ex; }
    se = block_allreduce(se, s_red, OpSum(), 0.f);
    float s2 = 0.f;
    for (int n = tid; n < N; n += TP) { const float p = __fdiv_rn(s_ll[n], se) + add_eps; probs_out[base + n] = p; s2 = fmaf(p, p, s2); }
    s2 = block_allreduce(s2, s_red, OpSum(), 0.f);
    if (tid == 0 && row_stats) { row_stats[2 * b] = sl; row_stats[2 * b + 1] = 1.0f / s2; }
    if (pred_out) {      // the prediction of the supervised loss, sum_n probs[n] particles[n] (losses.py:22): the row is still L2-hot
        float px = 0.f, py = 0.f;
        for (int n = tid; n < N; n += TP) {
            const float pr = probs_out[base + n];           // written by this thread above
            const float2 x = *reinterpret_cast<const float2*>(particles + (base + n) * 2);
            px = fmaf(pr, x.x, px); py = fmaf(pr, x.y, py);
        }
        px = block_allreduce(px, s_red, OpSum(), 0.f);
        py = block_allreduce(py, s_red, OpSum(), 0.f);
        if (tid == 0) { pred_out[2 * b] = px; pred_out[2 * b + 1] = py; }
    }
}

// ----------------------------------------------------------------------------------------------- backward
// particle-encoder tile rows (row = feature, column = particle of the CTA batch; stride TSM, see mma_tile.cuh).
// The weight gradients are contracted in two phases that reuse the same rows (half the tile: two CTAs per SM):
//   phase A: delta3 x [a2 | 1]                      -> dW3, db3
//   phase B: delta2 x [a1 | 1], delta1 x [x0 x1 1]  -> dW2, db2, dW1, db1
// per-warp accumulator copy: row strides 40 / 24 (instead of 32 / 16) spread the eight fragment rows a warp adds into at
// once over all banks (LDS.64 / STS.64 at the two-wavefront minimum)
struct AC {
    static constexpr int W3S = 40, W2S = 24;
    static constexpr int W3 = 0, B3 = W3 + 32 * W3S, W2 = B3 + 32, B2 = W2 + 32 * W2S, W1 = B2 + 32, B1 = W1 + 32, SIZE = B1 + 16;
    // accumulator index of packed particle-encoder parameter e
    __device__ static int of_packed(int e) {
        if (e < PE_B1) return W1 + e;
        if (e < PE_W2) return B1 + (e - PE_B1);
        if (e < PE_B2) return W2 + ((e - PE_W2) >> 4) * W2S + ((e - PE_W2) & 15);
        if (e < PE_W3) return B2 + (e - PE_B2);
        if (e < PE_B3) return W3 + ((e - PE_W3) >> 5) * W3S + ((e - PE_W3) & 31);
        return B3 + (e - PE_B3);
    }
};
struct PR {
    static constexpr int ONE = 0, ZERO = 1;
    static constexpr int D3 = 2, A2 = 34;                      // phase A
    static constexpr int X = 2, A1 = 4, D1 = 20, D2 = 36;      // phase B
    static constexpr int COUNT = 68;
};

// Particle-encoder weight gradients on the warp-level tensor path (3xTF32 mma.sync): every warp contracts over the 32
// particles its own threads staged.  The accumulator fragments are LANE-PRIVATE and live in tensor memory between batches
// (umma::ld_frag -> mma -> umma::st_frag on the warp's own 32 TMEM lanes): no per-warp shared-memory copies, no CTA barrier.
// Column map (per lane, relative to tacc = the first column after the data-path ones): phase A m-tile mt: [20 mt, 20 mt + 20) = c[5][4];
// phase B m-tile mt: 40 + [12 mt, 12 mt + 12) = c[3][4]; delta1 tile: 64 + [0, 4).
constexpr int TA_A = 0, TA_B = TA_A + 40, TA_D1 = TA_B + 24, TA_END = TA_D1 + 4;   // 68 columns after the data-path columns
constexpr int TA_CNF = TA_END, CNF_COLS = 18;                                      // CRNVP nets: 18 columns each behind them
constexpr int BWD_TMEM_COLS = 256;   // gaussian / cos: 96 + 68; CRNVP (SS data path): 32 + 68 + 8 x 18 = 244
__device__ __forceinline__ void pe_weight_grads_a(const float* __restrict__ s_tile, uint32_t tacc, int warp) {
    const int lane = threadIdx.x & 31, g = lane >> 2;
    const int k0 = 32 * warp;
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
        // the MMA chain of a batch starts from zero and is ADDED to the running sum with an IEEE add: the tensor core's own
        // fp32 accumulate is not round-to-nearest, and chaining hundreds of batches through it biased the gradients (~6e-4)
        float c[5][4] = {}, r[20];
        const int rowB[5] = {PR::A2 + g, PR::A2 + 8 + g, PR::A2 + 16 + g, PR::A2 + 24 + g, g == 0 ? PR::ONE : PR::ZERO};
        mma_outer<5, 1u << 4>(s_tile, PR::D3 + 16 * mt, rowB, k0, k0 + 32, c);
        umma::ld_frag<20>(tacc + TA_A + 20 * mt, r);
#pragma unroll
        for (int i = 0; i < 20; ++i) r[i] += c[i >> 2][i & 3];
        umma::st_frag<20>(tacc + TA_A + 20 * mt, r);
    }
}
__device__ __forceinline__ void pe_weight_grads_a(const float* __restrict__ s_tile, uint32_t tacc) {
    pe_weight_grads_a(s_tile, tacc, threadIdx.x >> 5);
}
__device__ __forceinline__ void pe_weight_grads_b(const float* __restrict__ s_tile, uint32_t tacc, int warp) {
    const int lane = threadIdx.x & 31, g = lane >> 2;
    const int k0 = 32 * warp;
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
        float c[3][4] = {}, r[12];
        const int rowB[3] = {PR::A1 + g, PR::A1 + 8 + g, g == 0 ? PR::ONE : PR::ZERO};
        mma_outer<3, 1u << 2>(s_tile, PR::D2 + 16 * mt, rowB, k0, k0 + 32, c);
        umma::ld_frag<12>(tacc + TA_B + 12 * mt, r);
#pragma unroll
        for (int i = 0; i < 12; ++i) r[i] += c[i >> 2][i & 3];
        umma::st_frag<12>(tacc + TA_B + 12 * mt, r);
    }
    float c1[1][4] = {}, r1[4];
    const int rowX[1] = {g < 2 ? PR::X + g : (g == 2 ? PR::ONE : PR::ZERO)};
    mma_outer<1>(s_tile, PR::D1, rowX, k0, k0 + 32, c1);
    umma::ld_frag<4>(tacc + TA_D1, r1);
#pragma unroll
    for (int i = 0; i < 4; ++i) r1[i] += c1[0][i];
    umma::st_frag<4>(tacc + TA_D1, r1);
    umma::wait_st();     // the next batch (or the read-out) loads these columns again
}
__device__ __forceinline__ void pe_weight_grads_b(const float* __restrict__ s_tile, uint32_t tacc) {
    pe_weight_grads_b(s_tile, tacc, threadIdx.x >> 5);
}
// The same two phases for the warp-specialised kernels' gradient warps: both 16-row delta tiles of a phase in ONE pass over the
// particles, so the activation fragments are loaded and split once (bit-identical sums; a third fewer instructions per batch).
__device__ __forceinline__ void pe_weight_grads_a_x2(const float* __restrict__ s_tile, uint32_t tacc, int warp) {
    const int lane = threadIdx.x & 31, g = lane >> 2;
    const int k0 = 32 * warp;
    float c0[5][4] = {}, c1[5][4] = {}, r[20];
    const int rowB[5] = {PR::A2 + g, PR::A2 + 8 + g, PR::A2 + 16 + g, PR::A2 + 24 + g, g == 0 ? PR::ONE : PR::ZERO};
    mma_outer2<5, 1u << 4>(s_tile, PR::D3, PR::D3 + 16, rowB, k0, k0 + 32, c0, c1);
    umma::ld_frag<20>(tacc + TA_A, r);
#pragma unroll
    for (int i = 0; i < 20; ++i) r[i] += c0[i >> 2][i & 3];
    umma::st_frag<20>(tacc + TA_A, r);
    umma::ld_frag<20>(tacc + TA_A + 20, r);
#pragma unroll
    for (int i = 0; i < 20; ++i) r[i] += c1[i >> 2][i & 3];
    umma::st_frag<20>(tacc + TA_A + 20, r);
}
__device__ __forceinline__ void pe_weight_grads_b_x2(const float* __restrict__ s_tile, uint32_t tacc, int warp) {
    const int lane = threadIdx.x & 31, g = lane >> 2;
    const int k0 = 32 * warp;
    {
        float c0[3][4] = {}, c1[3][4] = {}, r[12];
        const int rowB[3] = {PR::A1 + g, PR::A1 + 8 + g, g == 0 ? PR::ONE : PR::ZERO};
        mma_outer2<3, 1u << 2>(s_tile, PR::D2, PR::D2 + 16, rowB, k0, k0 + 32, c0, c1);
        umma::ld_frag<12>(tacc + TA_B, r);
#pragma unroll
        for (int i = 0; i < 12; ++i) r[i] += c0[i >> 2][i & 3];
        umma::st_frag<12>(tacc + TA_B, r);
        umma::ld_frag<12>(tacc + TA_B + 12, r);
#pragma unroll
        for (int i = 0; i < 12; ++i) r[i] += c1[i >> 2][i & 3];
        umma::st_frag<12>(tacc + TA_B + 12, r);
    }
    float c1[1][4] = {}, r1[4];
    const int rowX[1] = {g < 2 ? PR::X + g : (g == 2 ? PR::ONE : PR::ZERO)};
    mma_outer<1>(s_tile, PR::D1, rowX, k0, k0 + 32, c1);
    umma::ld_frag<4>(tacc + TA_D1, r1);
#pragma unroll
    for (int i = 0; i < 4; ++i) r1[i] += c1[0][i];
    umma::st_frag<4>(tacc + TA_D1, r1);
    umma::wait_st();     // the next batch (or the read-out) loads these columns again
}
// Read this warp's accumulated fragments back and scatter them into its accumulator copy accpe[AC::SIZE] (every entry is
// owned by exactly one lane; bias columns by the t == 0 / t == 1 lanes).
__device__ __forceinline__ void pe_weight_grads_readout(uint32_t tacc, float* __restrict__ accpe) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int mt = 0; mt < 2; ++mt) {
        float c[5][4];
        umma::ld_frag<20>(tacc + TA_A + 20 * mt, &c[0][0]);
        const int o = 16 * mt + g;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            float* w = accpe + AC::W3 + o * AC::W3S + 8 * n + 2 * t;
            w[0] = c[n][0]; w[1] = c[n][1]; w[8 * AC::W3S] = c[n][2]; w[8 * AC::W3S + 1] = c[n][3];
        }
        if (t == 0) { accpe[AC::B3 + o] = c[4][0]; accpe[AC::B3 + o + 8] = c[4][2]; }
    }
    for (int mt = 0; mt < 2; ++mt) {
        float c[3][4];
        umma::ld_frag<12>(tacc + TA_B + 12 * mt, &c[0][0]);
        const int o = 16 * mt + g;
#pragma unroll
        for (int n = 0; n < 2; ++n) {
            float* w = accpe + AC::W2 + o * AC::W2S + 8 * n + 2 * t;
            w[0] = c[n][0]; w[1] = c[n][1]; w[8 * AC::W2S] = c[n][2]; w[8 * AC::W2S + 1] = c[n][3];
        }
        if (t == 0) { accpe[AC::B2 + o] = c[2][0]; accpe[AC::B2 + o + 8] = c[2][2]; }
    }
    float c1[4];
    umma::ld_frag<4>(tacc + TA_D1, c1);
    if (t == 0) {          // columns 0,1 = dW1[o][0..1]
        accpe[AC::W1 + 2 * g] = c1[0]; accpe[AC::W1 + 2 * g + 1] = c1[1];
        accpe[AC::W1 + 2 * (g + 8)] = c1[2]; accpe[AC::W1 + 2 * (g + 8) + 1] = c1[3];
    } else if (t == 1) {   // column 2 = db1[o]
        accpe[AC::B1 + g] = c1[0]; accpe[AC::B1 + g + 8] = c1[2];
    }
}

template <class Tc> __device__ __forceinline__ Tc make_tc(float* a_tile, float* w, uint64_t* bar);
template <> __device__ __forceinline__ PeTc make_tc<PeTc>(float*, float* w, uint64_t* bar) { return PeTc{w, bar, 0u, 0u}; }
template <> __device__ __forceinline__ PeSs make_tc<PeSs>(float* a_tile, float* w, uint64_t* bar) { return PeSs{a_tile, w, bar, 0u, 0u}; }

// CRNVP stack (nets 48 -> 8 -> 8 -> 16): weight-gradient sink of stage_bwd that keeps the fragments in tensor memory.
// Per net and lane 18 values: dW1 (six input tiles x 2), dW2 (2), dW3 (4); the bias fragments, which only the t == 0 lanes
// carry, go to a small per-warp shared-memory block [n_fcnn][32] = b1 (8) | b2 (8) | b3 (16).  Replaces four per-warp
// shared-memory copies of all eight nets (78 KB), which pinned the kernel to one 4-warp CTA per SM.
struct CnfTmemSink {
    uint32_t tcol;     // this warp's lanes, first CRNVP column
    float* bias;       // this warp's bias block
    template <int HALF, int CP>
    __device__ __forceinline__ void accumulate(const float* __restrict__ s_tile, int f) {
        static_assert(HALF == 16 && CP == 32, "CRNVP measurement stack only");
        using R = RC;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
        const int k0 = 32 * warp;
        const uint32_t col = tcol + CNF_COLS * f;
        float a[CNF_COLS];
        umma::ld_frag<16>(col, a);
        umma::ld2w(col + 16, a + 16);
        {   // A = [delta1; delta2] x input tiles 0..3 -> rows 0-7: dW1[g][8 n + 2 t, +1]
            float c[4][4];
            int rowB[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) { rowB[n] = R::C + 8 * n + g; c[n][0] = 0.f; c[n][1] = 0.f; c[n][2] = 0.f; c[n][3] = 0.f; }
            mma_outer<4>(s_tile, R::D1, rowB, k0, k0 + 32, c);       // chains start from zero; IEEE adds into the running sums (see pe_weight_grads_a)
#pragma unroll
            for (int n = 0; n < 4; ++n) { a[2 * n] += c[n][0]; a[2 * n + 1] += c[n][1]; }
        }
        {   // input tiles 4, 5 | ONE (db1 rows 0-7, db2 rows 8-15) | h1 (rows 8-15: dW2[g][2 t, +1])
            float c[4][4] = {};
            const int rowB[4] = {R::C + 32 + g, R::C + 40 + g, g == 0 ? R::ONE : R::ZERO, R::H1 + g};
            mma_outer<4>(s_tile, R::D1, rowB, k0, k0 + 32, c);
            a[8] += c[0][0]; a[9] += c[0][1]; a[10] += c[1][0]; a[11] += c[1][1]; a[12] += c[3][2]; a[13] += c[3][3];
            if (t == 0) { bias[f * 32 + g] += c[2][0]; bias[f * 32 + 8 + g] += c[2][2]; }
        }
        {   // A = d out (16 rows) x [h2 | ONE] -> dW3[g][2 t, +1], dW3[g + 8][..], db3
            float c[2][4] = {};
            const int rowB[2] = {R::H2 + g, g == 0 ? R::ONE : R::ZERO};
            mma_outer<2>(s_tile, R::DO, rowB, k0, k0 + 32, c);
            a[14] += c[0][0]; a[15] += c[0][1]; a[16] += c[0][2]; a[17] += c[0][3];
            if (t == 0) { bias[f * 32 + 16 + g] += c[1][0]; bias[f * 32 + 24 + g] += c[1][2]; }
        }
        umma::st_frag<16>(col, a);
        umma::st2(col + 16, a + 16);
        umma::wait_st();
    }
    // write this warp's accumulated gradients of all nets into its own partial (packed order; C_row = 0: identity layout)
    __device__ void readout(int n_fcnn, float* __restrict__ out) const {
        constexpr int IN = 48, B1 = H * IN, W2 = B1 + H, B2 = W2 + H * H, W3 = B2 + H, B3 = W3 + 16 * H, PF = B3 + 16;
        const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
        for (int f = 0; f < n_fcnn; ++f) {
            float a[CNF_COLS];
            umma::ld_frag<16>(tcol + CNF_COLS * f, a);
            umma::ld2w(tcol + CNF_COLS * f + 16, a + 16);
            float* o = out + (size_t)f * PF;
#pragma unroll
            for (int n = 0; n < 6; ++n) { o[g * IN + 8 * n + 2 * t] = TANH_SCALE * a[2 * n]; o[g * IN + 8 * n + 2 * t + 1] = TANH_SCALE * a[2 * n + 1]; }
            o[W2 + g * H + 2 * t] = TANH_SCALE * a[12]; o[W2 + g * H + 2 * t + 1] = TANH_SCALE * a[13];
            o[W3 + g * H + 2 * t] = a[14]; o[W3 + g * H + 2 * t + 1] = a[15];
            o[W3 + (g + 8) * H + 2 * t] = a[16]; o[W3 + (g + 8) * H + 2 * t + 1] = a[17];
            if (t == 0) {
                o[B1 + g] = TANH_SCALE * bias[f * 32 + g]; o[B2 + g] = TANH_SCALE * bias[f * 32 + 8 + g];
                o[B3 + g] = bias[f * 32 + 16 + g]; o[B3 + 8 + g] = bias[f * 32 + 24 + g];
            }
        }
    }
};

template <int MODE>
__global__ void __launch_bounds__(TP, 2)   // two CTAs per SM (256 tensor-memory columns each)
measure_bwd_kernel(const float* __restrict__ pe, const float* __restrict__ cnf, int n_flows, float p0, float p1,
                   const float* __restrict__ enc, const float* __restrict__ particles, int B, int N,
                   const float* __restrict__ g_lki, const int* __restrict__ argmax, float* __restrict__ d_particles,
                   float* __restrict__ d_enc, float* __restrict__ part_pe, float* __restrict__ part_cnf, const float* __restrict__ z_saved,
                   const float* __restrict__ g_pred, const float* __restrict__ probs) {
    extern __shared__ __align__(128) float smem[];
    __shared__ float s_red[33];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tslot;
    const int tid = threadIdx.x, n_fcnn = MODE == MODE_CNF ? 4 * n_flows : 0;
    using Tc = typename std::conditional<MODE == MODE_CNF, PeSs, PeTc>::type;   // CRNVP: SS data path (tensor-memory budget)
    constexpr int TILE_ROWS = MODE == MODE_CNF && RC::TROWS > PR::COUNT ? RC::TROWS : PR::COUNT;
    constexpr int TILE_FLOATS = (TILE_ROWS * TSM + 31) & ~31;   // keeps the weight tiles 128-byte aligned
    // The gradient tile: every thread owns one column (its particle), every warp contracts over its own 32 columns; the CRNVP
    // tile and the encoder tile alias (warp-private columns, __syncwarp).  In the CRNVP kernel the SS activation tile (hi | lo,
    // 32 KB) aliases it too: there a CTA barrier separates a batch's rounds from the tile uses before and after them.
    float* s_tile = smem;                                    // [TILE_FLOATS]
    float* s_tcw = s_tile + TILE_FLOATS;                     // tensor-core weight tiles
    float* s_pe = s_tcw + PeTc::WBWD_FLOATS;
    float* s_enc = s_pe + PE_SIZE;
    float* s_img = s_enc + 36;
    float* s_hb = s_img + n_fcnn * LC::SIZE;
    constexpr int NW = TP / 32;
    const int warp = tid >> 5;
    float* s_accpe = s_tile;                                 // [NW][AC::SIZE]   read-out staging at the very end (the tile is free then)
    float* s_cnfbias = s_hb + n_fcnn * H;                    // [NW][n_fcnn][32]  CRNVP bias gradients (t == 0 lanes)
    float* s_denc = s_cnfbias + NW * n_fcnn * 32 + 4;        // [NW][32]
    static_assert(TILE_FLOATS % 32 == 0, "weight tiles must stay 128-byte aligned");
    static_assert(NW * AC::SIZE <= TILE_FLOATS, "read-out staging must fit in the tile");
    static_assert(MODE != MODE_CNF || PeSs::A_FLOATS <= TILE_FLOATS, "SS activation tile must fit in the tile");
    if (tid < 32) umma::tmem_alloc<BWD_TMEM_COLS>(&s_tslot);
    if (tid == 0) umma::mbar_init(&s_bar, 1);
    Tc tc = make_tc<Tc>(s_tile, s_tcw, &s_bar);
    tc.load_weights(pe, true);
    for (int e = tid; e < PE_SIZE; e += TP) {
        s_pe[e] = pe[e];
    }
    if (MODE == MODE_CNF) {
        const int pf = packed_fcnn_size(16, 32);
        for (int f = 0; f < n_fcnn; ++f) load_fcnn_image<16, 32>(cnf + (size_t)f * pf, 0, s_img + f * LC::SIZE, nullptr, tid, TP);
        for (int e = tid; e < NW * n_fcnn * 32; e += TP) s_cnfbias[e] = 0.f;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    tc.tmem = s_tslot;
    const uint32_t tacc = tc.lane_addr() + Tc::COLS;   // this warp's 32 lanes: gradient fragments behind the data-path columns
    {
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0.f;
        const int ncols = TA_END + (MODE == MODE_CNF ? CNF_COLS * n_fcnn : 0);
        for (int c0 = 0; c0 < ncols; c0 += 4) umma::st4(tacc + c0, z);
        umma::wait_st();
    }
    CnfTmemSink cnf_sink{tacc + TA_CNF, s_cnfbias + warp * n_fcnn * 32};
    float rs3_prev[4] = {0.f, 0.f, 0.f, 0.f};    // Gaussian mode, t == 0 lanes: running delta3 sums of earlier trajectories
    if (MODE == MODE_CNF) hoist_row_context<16, 32>(s_img, nullptr, nullptr, 0, n_fcnn, s_hb, tid, TP);

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        load_enc(enc + (size_t)b * HID, s_enc);
        const size_t base = (size_t)b * N;
        // row-max shift: d ll[n] = g[n] - [n == argmax] * sum_m g[m]
        float gs = 0.f;
        if (MODE != MODE_COS) {
            for (int n = tid; n < N; n += TP) gs += g_lki[base + n];
            gs = block_allreduce(gs, s_red, OpSum(), 0.f);
        }
        const int am = MODE != MODE_COS ? argmax[b] : -1;
        float denc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) denc[k] = 0.f;
        __syncthreads();
        float2 x_next = *reinterpret_cast<const float2*>(particles + (base + (tid < N ? tid : 0)) * 2);
        float g_next = g_lki[base + (tid < N ? tid : 0)];
        for (int n0 = 0; n0 < N; n0 += TP) {
            asm volatile("" ::: "memory");  // no LICM of shared-memory weight loads across particles
            if (MODE == MODE_CNF) __syncthreads();   // SS activation tile aliases the gradient tile the previous batch contracted over
            const int n = n0 + tid;
            const bool live = n < N;
            const size_t p = base + (live ? n : 0);
            const float2 x = x_next;
            float g = live ? g_next : 0.f;
            if (live && n == am) g -= gs;
            {   // prefetch the next batch's particle and incoming gradient: their latency hides behind this batch
                const size_t pn = base + (n + TP < N ? n + TP : 0);
                x_next = *reinterpret_cast<const float2*>(particles + pn * 2);
                g_next = g_lki[pn];
            }
            float a1[16], a2[32], e[32], de[32];
            pe_fwd_tc(tc, s_pe, x.x, x.y, a1, a2, e);        // tensor-core rounds 1-2: layers 2 and 3 forward
            if (MODE == MODE_GAUSS) {
                const float c = live ? g / (p1 * p1) : 0.f;
#pragma unroll
                for (int o = 0; o < 32; ++o) de[o] = c * (s_enc[o] - e[o] - p0);
            } else if (MODE == MODE_COS) {
                float ne = 0.f, dot = 0.f;
#pragma unroll
                for (int k = 0; k < 32; ++k) { ne = fmaf(e[k], e[k], ne); dot = fmaf(s_enc[k], e[k], dot); }
                const float nrm = fmaxf(sqrtf(ne), 1e-12f), nenc = s_enc[32];
                const float ab = dot / (nrm * nenc);
                const float c = g / (1e-7f + 1.0f - ab);     // d lki / d(ab)
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const float ah = s_enc[k] / nenc, bh = e[k] / nrm;
                    de[k] = live ? c * (ah - ab * bh) / nrm : 0.f;
                    denc[k] += c * (bh - ab * ah) / nenc;
                }
            } else {
                float lo[16], up[16], glo[16], gup[16];
                if (z_saved) {
                    const float4* zi = reinterpret_cast<const float4*>(z_saved + p * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = zi[i], c4 = zi[4 + i];
                        lo[4 * i] = a.x; lo[4 * i + 1] = a.y; lo[4 * i + 2] = a.z; lo[4 * i + 3] = a.w;
                        up[4 * i] = c4.x; up[4 * i + 1] = c4.y; up[4 * i + 2] = c4.z; up[4 * i + 3] = c4.w;
                    }
                } else {
                    loglik<MODE_CNF>(e, s_enc, p0, p1, s_img, s_hb, n_flows, lo, up);
                }
                const float c = -g / (p1 * p1);
#pragma unroll
                for (int i = 0; i < 16; ++i) { glo[i] = c * (lo[i] - p0); gup[i] = c * (up[i] - p0); }
#pragma unroll
                for (int k = 0; k < 32; ++k) { de[k] = 0.f; s_tile[(RC::PC + k) * TSM + tid] = e[k]; }
                s_tile[RC::ONE * TSM + tid] = 1.0f;
                for (int r = RC::COUNT; r < RC::TROWS; ++r) s_tile[r * TSM + tid] = 0.0f;
                swap_halves<16>(lo, up); swap_halves<16>(glo, gup);     // the last forward stage had c = upper
#pragma unroll 1
                for (int st = 2 * n_flows - 1; st >= 0; --st) {         // walk the forward stages back; one inlined stage body
                    const float* im = s_img + 2 * st * LC::SIZE;
                    const float* hb = s_hb + 2 * st * H;
                    stage_bwd<16, 32>(im, im + LC::SIZE, hb, hb + H, 2 * st, false, live, lo, glo, e, de, up, gup, g, s_tile, cnf_sink);
                    swap_halves<16>(lo, up); swap_halves<16>(glo, gup);
                }
                swap_halves<16>(lo, up); swap_halves<16>(glo, gup);
                if (live) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) { denc[i] += glo[i]; denc[16 + i] += gup[i]; }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k) de[k] = 0.f;
                }
                __syncthreads();            // every warp is done with the CRNVP tile before the (aliasing) SS activation tile is rewritten
            }
            // encoder backward, tensor-core rounds 3-4: d a2 = W3^T delta3, d a1 = W2^T delta2
            float d2[32], d1[16];
            tc.template store_row<32>(de);
            tc.template round<32, 32>(PeTc::W3T_HI, PeTc::W3T_LO);
            umma::ld32(tc.lane_addr(), d2);
#pragma unroll
            for (int j = 0; j < 32; ++j) d2[j] = a2[j] > 0.f ? d2[j] : 0.f;
            tc.template store_row<32>(d2);
            tc.template round<16, 32>(PeTc::W2T_HI, PeTc::W2T_LO);
            umma::ld16(tc.lane_addr(), d1);
            float dx0 = 0.f, dx1 = 0.f;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const float4 q = *reinterpret_cast<const float4*>(s_pe + PE_W1 + 2 * k);
                d1[k] = a1[k] > 0.f ? d1[k] : 0.f;
                d1[k + 1] = a1[k + 1] > 0.f ? d1[k + 1] : 0.f;
                dx0 = fmaf(q.x, d1[k], dx0); dx1 = fmaf(q.y, d1[k], dx1);
                dx0 = fmaf(q.z, d1[k + 1], dx0); dx1 = fmaf(q.w, d1[k + 1], dx1);
            }
            if (live) {
                if (g_pred) {      // fused prediction (losses.py:22): d pred / d x_n = probs[n]
                    const float pr = probs[p];
                    dx0 = fmaf(g_pred[2 * b], pr, dx0); dx1 = fmaf(g_pred[2 * b + 1], pr, dx1);
                }
                *reinterpret_cast<float2*>(d_particles + p * 2) = make_float2(dx0, dx1);
            }
            // weight gradients: the warp contracts over its own 32 tile columns, no CTA barrier between the two phases
            s_tile[PR::ONE * TSM + tid] = 1.0f;
            s_tile[PR::ZERO * TSM + tid] = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { s_tile[(PR::D3 + j) * TSM + tid] = de[j]; s_tile[(PR::A2 + j) * TSM + tid] = a2[j]; }
            __syncwarp();
            pe_weight_grads_a(s_tile, tacc);
            __syncwarp();
            s_tile[(PR::X + 0) * TSM + tid] = x.x;
            s_tile[(PR::X + 1) * TSM + tid] = x.y;
#pragma unroll
            for (int k = 0; k < 16; ++k) { s_tile[(PR::A1 + k) * TSM + tid] = a1[k]; s_tile[(PR::D1 + k) * TSM + tid] = d1[k]; }
#pragma unroll
            for (int j = 0; j < 32; ++j) s_tile[(PR::D2 + j) * TSM + tid] = d2[j];
            __syncwarp();
            pe_weight_grads_b(s_tile, tacc);
            __syncwarp();
        }
        // d_enc[b][k] = sum over the row's particles (deterministic: through the tile)
        __syncthreads();   // all warps are done with this trajectory (tile columns, per-warp sums)
        if (MODE == MODE_GAUSS) {
            // d_enc = -sum_p delta3: the bias column of phase A holds the warp's RUNNING sum over the trajectories done so far
            // (t == 0 lanes: entries c[4][0], c[4][2] of both m-tiles); this trajectory's share is the increment
            float cur[4], c4[4];
            umma::ld_frag<4>(tacc + TA_A + 16, c4); cur[0] = c4[0]; cur[1] = c4[2];
            umma::ld_frag<4>(tacc + TA_A + 36, c4); cur[2] = c4[0]; cur[3] = c4[2];
            if ((tid & 3) == 0) {
                const int g = (tid & 31) >> 2;
                s_denc[warp * 32 + g] = cur[0] - rs3_prev[0]; s_denc[warp * 32 + g + 8] = cur[1] - rs3_prev[1];
                s_denc[warp * 32 + 16 + g] = cur[2] - rs3_prev[2]; s_denc[warp * 32 + 24 + g] = cur[3] - rs3_prev[3];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) rs3_prev[i] = cur[i];
            __syncthreads();
            if (d_enc && tid < 32) d_enc[(size_t)b * HID + tid] = -((s_denc[tid] + s_denc[32 + tid]) + (s_denc[64 + tid] + s_denc[96 + tid]));
        } else if (d_enc) {
#pragma unroll
            for (int k = 0; k < 32; ++k) s_tile[k * TSM + tid] = denc[k];
            __syncthreads();
            if (tid < 32) {
                float a = 0.f;
                for (int p = 0; p < TP; ++p) a += s_tile[tid * TSM + p];
                d_enc[(size_t)b * HID + tid] = a;
            }
        }
    }
    __syncthreads();                                   // every warp is done with the tile: it becomes the read-out staging area
    for (int e = tid; e < NW * AC::SIZE; e += TP) s_accpe[e] = 0.f;
    __syncthreads();
    pe_weight_grads_readout(tacc, s_accpe + warp * AC::SIZE);
    if (MODE == MODE_CNF)     // every warp writes its own partial: NW partial gradients per CTA
        cnf_sink.readout(n_fcnn, part_cnf + ((size_t)blockIdx.x * NW + warp) * n_fcnn * packed_fcnn_size(16, 32));
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_free<BWD_TMEM_COLS>(tc.tmem);
    for (int e = tid; e < PE_SIZE; e += TP) {
        const float* a = s_accpe + AC::of_packed(e);
        part_pe[(size_t)blockIdx.x * PE_SIZE + e] = (a[0] + a[AC::SIZE]) + (a[2 * AC::SIZE] + a[3 * AC::SIZE]);
    }
}

// ---- warp-specialised backward (gaussian / cos) -----------------------------------------------------------------------
// The kernel above walks a batch through four tcgen05 rounds and then through the two mma.sync weight-gradient phases, all on
// the same four warps: while a CTA waits for a round nobody contracts, while it contracts the tensor cores' data path idles
// (ncu: issue slots 35 %, tensor pipe 36 %, 17 % of the samples in the mbarrier spin).  Here a CTA has EIGHT warps: warps 0-3 own
// the particles (thread = particle = tensor-memory lane: forward rounds, likelihood gradient, data-gradient rounds, d_particles),
// warps 4-7 own the weight gradients (warp 4 + w contracts over the 32 particles warp w staged; its lane-private fragments live in
// the same tensor-memory lanes, behind the data-path columns).  The two halves meet only in the gradient tile: per warp pair one
// `full` and one `empty` mbarrier, alternating between the tile's two (aliased) phases --
//   data warp:      ... round 2 | wait empty | stage A (delta3, a2) | arrive full | rounds 3-4 | wait empty | stage B | arrive full | next batch
//   gradient warp:  wait full | phase A (dW3, db3) | arrive empty | wait full | phase B (dW2, db2, dW1, db1) | arrive empty
// so phase A of a batch overlaps its data-gradient rounds and phase B the next batch's forward rounds.  Two CTAs per SM (16 warps).
constexpr int WS_THREADS = 256;
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void data_group_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// PeTc whose rounds synchronise the 128 data threads only (named barrier 1)
struct PeTcWs : PeTc {
    template <int N, int K>
    __device__ __forceinline__ void round(int w_hi, int w_lo) {
        umma::wait_st();
        umma::fence_before_sync();
        data_group_sync();
        if (threadIdx.x == 0) {
            umma::fence_after_sync();
            umma::gemm3_ts<N, K>(tmem + COL_D, tmem + COL_AHI, tmem + COL_ALO, w + w_hi, w + w_lo);
            umma::commit(bar);
        }
        wait();
    }
};

template <int MODE>
__global__ void __launch_bounds__(WS_THREADS, 2)
measure_bwd_ws_kernel(const float* __restrict__ pe, float p0, float p1, const float* __restrict__ enc,
                      const float* __restrict__ particles, int B, int N, const float* __restrict__ g_lki,
                      const int* __restrict__ argmax, float* __restrict__ d_particles, float* __restrict__ d_enc,
                      float* __restrict__ part_pe, const float* __restrict__ g_pred, const float* __restrict__ probs) {
    static_assert(MODE == MODE_GAUSS || MODE == MODE_COS, "CRNVP keeps the single-role kernel");
    extern __shared__ __align__(128) float smem[];
    __shared__ float s_red[8];
    __shared__ uint64_t s_bar;
    __shared__ uint64_t s_full[4], s_empty[4];
    __shared__ uint32_t s_tslot;
    constexpr int TILE_FLOATS = (PR::COUNT * TSM + 31) & ~31;
    constexpr int NWD = 4;                                   // data warps == gradient warps
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_data = warp < NWD;
    const int pw = warp & 3;                                 // warp pair = tensor-memory lane quarter
    float* s_tile = smem;                                    // [TILE_FLOATS]
    float* s_tcw = s_tile + TILE_FLOATS;                     // tensor-core weight tiles
    float* s_pe = s_tcw + PeTc::WBWD_FLOATS;
    float* s_enc = s_pe + PE_SIZE;
    float* s_denc = s_enc + 36;                              // [4][32]
    float* s_accpe = s_tile;                                 // read-out staging at the very end
    static_assert(NWD * AC::SIZE <= TILE_FLOATS, "read-out staging must fit in the tile");
    // (a pre-split hi | lo tile pair -- the data warps splitting once, the gradient warps only loading -- was measured: 279 us
    //  against 194: twice the shared-memory stores and fragment loads cost more than the 860 split instructions per batch saved)
    auto stage = [&](int row, float v) { s_tile[row * TSM + tid] = v; };
    if (is_data) {   // the constant rows of the bias columns, once (no phase aliases them)
        s_tile[PR::ONE * TSM + tid] = 1.0f;
        s_tile[PR::ZERO * TSM + tid] = 0.0f;
    }
    if (tid < 32) umma::tmem_alloc<BWD_TMEM_COLS>(&s_tslot);
    if (tid == 0) {
        umma::mbar_init(&s_bar, 1);
        for (int w = 0; w < NWD; ++w) { umma::mbar_init(&s_full[w], 32); umma::mbar_init(&s_empty[w], 32); }
    }
    PeTcWs tc;
    tc.w = s_tcw; tc.bar = &s_bar; tc.tmem = 0u; tc.parity = 0u;
    tc.load_weights(pe, true);
    for (int e = tid; e < PE_SIZE; e += WS_THREADS) s_pe[e] = pe[e];
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    tc.tmem = s_tslot;
    const uint32_t tacc = tc.tmem + ((uint32_t)(32 * pw) << 16) + PeTc::COLS;   // the pair's 32 lanes, behind the data-path columns
    if (is_data) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = 0; c0 < TA_END; c0 += 4) umma::st4(tacc + c0, z);
        umma::wait_st();
    }
    uint32_t ph_full = 0, ph_empty = 1;                      // the tile starts out empty: the first `empty` wait passes
    float rs3_prev[4] = {0.f, 0.f, 0.f, 0.f};

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        umma::fence_before_sync();
        __syncthreads();                                     // S1: previous trajectory is finished everywhere
        umma::fence_after_sync();
        const size_t base = (size_t)b * N;
        if (!is_data) {
            // ---------------------------------------------------------------- gradient warps
            for (int n0 = 0; n0 < N; n0 += 128) {
                umma::mbar_wait(&s_full[pw], ph_full); ph_full ^= 1;
                pe_weight_grads_a_x2(s_tile, tacc, pw);
                mbar_arrive(&s_empty[pw]);
                umma::mbar_wait(&s_full[pw], ph_full); ph_full ^= 1;
                pe_weight_grads_b_x2(s_tile, tacc, pw);
                mbar_arrive(&s_empty[pw]);
            }
            umma::wait_st();
        } else {
            // ---------------------------------------------------------------- data warps
            load_enc(enc + (size_t)b * HID, s_enc);
            float gs = 0.f;
            if (MODE != MODE_COS) {      // row-max shift: d ll[n] = g[n] - [n == argmax] * sum_m g[m]
                for (int n = tid; n < N; n += 128) gs += g_lki[base + n];
                gs = warp_sum(gs);
                if (lane == 0) s_red[warp] = gs;
            }
            data_group_sync();
            if (MODE != MODE_COS) gs = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
            const int am = MODE != MODE_COS ? argmax[b] : -1;
            float denc[MODE == MODE_COS ? 32 : 1];
#pragma unroll
            for (int k = 0; k < (MODE == MODE_COS ? 32 : 1); ++k) denc[k] = 0.f;
            float2 x_next = *reinterpret_cast<const float2*>(particles + (base + (tid < N ? tid : 0)) * 2);
            float g_next = g_lki[base + (tid < N ? tid : 0)];
            for (int n0 = 0; n0 < N; n0 += 128) {
                asm volatile("" ::: "memory");
                const int n = n0 + tid;
                const bool live = n < N;
                const size_t p = base + (live ? n : 0);
                const float2 x = x_next;
                float g = live ? g_next : 0.f;
                if (live && n == am) g -= gs;
                {
                    const size_t pn = base + (n + 128 < N ? n + 128 : 0);
                    x_next = *reinterpret_cast<const float2*>(particles + pn * 2);
                    g_next = g_lki[pn];
                }
                float a1[16], a2[32], e[32];
                pe_fwd_tc(tc, s_pe, x.x, x.y, a1, a2, e);        // rounds 1-2: layers 2 and 3 forward
                if (MODE == MODE_GAUSS) {
                    const float c = live ? g / (p1 * p1) : 0.f;
#pragma unroll
                    for (int o = 0; o < 32; ++o) e[o] = c * (s_enc[o] - e[o] - p0);          // e := delta3
                } else {
                    float ne = 0.f, dot = 0.f;
#pragma unroll
                    for (int k = 0; k < 32; ++k) { ne = fmaf(e[k], e[k], ne); dot = fmaf(s_enc[k], e[k], dot); }
                    const float nrm = fmaxf(sqrtf(ne), 1e-12f), nenc = s_enc[32];
                    const float ab = dot / (nrm * nenc);
                    const float c = g / (1e-7f + 1.0f - ab);     // d lki / d(ab)
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const float ah = s_enc[k] / nenc, bh = e[k] / nrm;
                        e[k] = live ? c * (ah - ab * bh) / nrm : 0.f;
                        denc[k] += c * (bh - ab * ah) / nenc;
                    }
                }
                // stage phase A (the gradient warp has finished phase B of the previous batch)
                umma::mbar_wait(&s_empty[pw], ph_empty); ph_empty ^= 1;
                uint32_t m2 = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    stage(PR::D3 + j, e[j]);
                    stage(PR::A2 + j, a2[j]);
                    m2 |= (a2[j] > 0.f ? 1u : 0u) << j;
                }
                mbar_arrive(&s_full[pw]);
                // rounds 3-4: d a2 = W3^T delta3, d a1 = W2^T delta2
                float d2[32], d1[16];
                tc.template store_row<32>(e);
                tc.template round<32, 32>(PeTc::W3T_HI, PeTc::W3T_LO);
                umma::ld32(tc.lane_addr(), d2);
#pragma unroll
                for (int j = 0; j < 32; ++j) d2[j] = (m2 >> j) & 1u ? d2[j] : 0.f;
                tc.template store_row<32>(d2);
                tc.template round<16, 32>(PeTc::W2T_HI, PeTc::W2T_LO);
                umma::ld16(tc.lane_addr(), d1);
                float dx0 = 0.f, dx1 = 0.f;
#pragma unroll
                for (int k = 0; k < 16; k += 2) {
                    const float4 q = *reinterpret_cast<const float4*>(s_pe + PE_W1 + 2 * k);
                    d1[k] = a1[k] > 0.f ? d1[k] : 0.f;
                    d1[k + 1] = a1[k + 1] > 0.f ? d1[k + 1] : 0.f;
                    dx0 = fmaf(q.x, d1[k], dx0); dx1 = fmaf(q.y, d1[k], dx1);
                    dx0 = fmaf(q.z, d1[k + 1], dx0); dx1 = fmaf(q.w, d1[k + 1], dx1);
                }
                if (live) {
                    if (g_pred) {      // fused prediction (losses.py:22): d pred / d x_n = probs[n]
                        const float pr = probs[p];
                        dx0 = fmaf(g_pred[2 * b], pr, dx0); dx1 = fmaf(g_pred[2 * b + 1], pr, dx1);
                    }
                    *reinterpret_cast<float2*>(d_particles + p * 2) = make_float2(dx0, dx1);
                }
                // stage phase B (its rows alias phase A's: the gradient warp has finished phase A)
                umma::mbar_wait(&s_empty[pw], ph_empty); ph_empty ^= 1;
                stage(PR::X + 0, x.x);
                stage(PR::X + 1, x.y);
#pragma unroll
                for (int k = 0; k < 16; ++k) { stage(PR::A1 + k, a1[k]); stage(PR::D1 + k, d1[k]); }
#pragma unroll
                for (int j = 0; j < 32; ++j) stage(PR::D2 + j, d2[j]);
                mbar_arrive(&s_full[pw]);
            }
            if (MODE == MODE_COS) {
                // d_enc[b][k] = sum over the row's particles: through the tile once the gradient warps are done with it
                umma::mbar_wait(&s_empty[pw], ph_empty);         // (no toggle: phase B of the last batch; the next stage-A wait re-observes it)
#pragma unroll
                for (int k = 0; k < 32; ++k) s_tile[(2 + k) * TSM + tid] = denc[k];     // (rows 0-1 hold the constant rows)
            }
        }
        umma::fence_before_sync();
        __syncthreads();                                     // S2: every fragment of this trajectory is in tensor memory
        umma::fence_after_sync();
        if (MODE == MODE_GAUSS) {
            // d_enc = -sum_p delta3: the bias column of phase A holds the RUNNING sum over the trajectories done so far
            if (is_data) {
                float cur[4], c4[4];
                umma::ld_frag<4>(tacc + TA_A + 16, c4); cur[0] = c4[0]; cur[1] = c4[2];
                umma::ld_frag<4>(tacc + TA_A + 36, c4); cur[2] = c4[0]; cur[3] = c4[2];
                if ((tid & 3) == 0) {
                    const int g = lane >> 2;
                    s_denc[warp * 32 + g] = cur[0] - rs3_prev[0]; s_denc[warp * 32 + g + 8] = cur[1] - rs3_prev[1];
                    s_denc[warp * 32 + 16 + g] = cur[2] - rs3_prev[2]; s_denc[warp * 32 + 24 + g] = cur[3] - rs3_prev[3];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) rs3_prev[i] = cur[i];
            }
            __syncthreads();
            if (d_enc && tid < 32) d_enc[(size_t)b * HID + tid] = -((s_denc[tid] + s_denc[32 + tid]) + (s_denc[64 + tid] + s_denc[96 + tid]));
        } else if (d_enc) {
            if (tid < 32) {
                float a = 0.f;
                for (int q = 0; q < 128; ++q) a += s_tile[(2 + tid) * TSM + q];
                d_enc[(size_t)b * HID + tid] = a;
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();                                   // every warp is done with the tile: it becomes the read-out staging area
    umma::fence_after_sync();
    for (int e = tid; e < NWD * AC::SIZE; e += WS_THREADS) s_accpe[e] = 0.f;
    __syncthreads();
    if (is_data) pe_weight_grads_readout(tacc, s_accpe + warp * AC::SIZE);
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_free<BWD_TMEM_COLS>(tc.tmem);
    for (int e = tid; e < PE_SIZE; e += WS_THREADS) {
        const float* a = s_accpe + AC::of_packed(e);
        part_pe[(size_t)blockIdx.x * PE_SIZE + e] = (a[0] + a[AC::SIZE]) + (a[2 * AC::SIZE] + a[3 * AC::SIZE]);
    }
}

// ---- CRNVP backward, warp-specialised (round 2) ----------------------------------------------------------------------
// The single-role kernel above spends 19 k warp instructions per 128-particle batch at 255 registers with 16.7 M local loads /
// stores per launch: the 48-wide layer 1 of the eight nets runs on the CUDA cores twice (forward recompute and W1^T delta1) with the
// particle encoding e and its gradient (64 registers) live across the whole stage loop, and the same four warps then contract the
// weight gradients.  Here a CTA owns a whole SM (all 512 tensor-memory columns, 170 KB of shared memory) and has EIGHT warps:
//   warps 0-3 (data; thread = particle = tensor-memory lane): per stage two tcgen05 products in the TS form --
//       forward   pre[16]  = [c | e] (K = 48) x W1cat^T   (CnfL1's operand tiles, as in the forward kernel), and
//       backward  [d c | d e] += delta1[t | s] (K = 16) x W1cat   (N = 48; the thread zeroes its d c columns every stage, d e is ONE
//                 tensor-memory accumulator over all stages and nets),
//     so e lives in tensor memory (hi | lo), d e is read once per batch, and only the 8-wide tails (layers 2-3, 384 of the 1152
//     FMAs of a stage) stay in registers; the encoder activations a1 / a2 and the thread's d_enc sums wait in spare tensor-memory
//     columns.  The NEXT stage's conditioning half is this stage's inverted half, known before the tails' backward: the backward
//     product of stage st and the forward product of stage st - 1 are issued together -- ONE issue -> commit -> wait round per stage
//     (nine per batch with the encoder's four) instead of two;
//   warps 4-7 (gradient): warp 4 + w contracts over the 32 particles warp w staged (mma.sync 3xTF32, fragments of all eight nets +
//     the encoder lane-private in tensor memory: 212 columns), one stage (both nets) per hand-over.
// The halves meet in the gradient tile through one full / one empty mbarrier per warp pair: six hand-overs per batch (four stages,
// encoder phase A, encoder phase B); the tile is staged while the stage's backward round is in flight.
struct CnfB {
    using WB = umma::Operand<48, 16>;                        // rows = layer-1 input i (c 16 | e 32), K = unit (t-net 8 | s-net 8)
    static constexpr int L1B_STAGE_FLOATS = 2 * WB::FLOATS;  // hi | lo
    // gradient tile rows (one column per particle): constants, conditioning half, particle encoding, two net blocks
    static constexpr int ONE = 0, ZERO = 1, C = 2, E = 18, NET = 50, NET_ROWS = 48, ROWS = NET + 2 * NET_ROWS;
    static constexpr int H1 = 0, H2 = 8, D1 = 16, D2 = 24, DO = 32;   // inside a net block ([D1; D2] and DO are m16 A tiles)
    // tensor-memory columns: pre-activation accumulator | layer-1 activation row [c | e] hi, lo (K = 48 each) | delta1 hi, lo (K = 16
    // each) | d c (fresh every stage) and d e (ONE accumulator over the stack), contiguous: one N = 48 product | parked encoder
    // activations a1, a2 | the thread's running d_enc sums | gradient fragments.  The encoder's rounds (PeTc's map) alias [0,96)
    // before and after the stage loop.
    static constexpr int COL_PRE = 0, COL_AHI = 16, COL_ALO = 64, COL_D1HI = 112, COL_D1LO = 128, COL_DC = 144, COL_DE = 160,
                         COL_A1 = 192, COL_A2 = 208, COL_DENC = 240, COL_FRAG = 272, TMEM_COLS = 512;
    static_assert(ROWS >= PR::COUNT, "the encoder phases alias the CRNVP rows");
    static_assert(COL_FRAG + TA_END + 8 * CNF_COLS <= TMEM_COLS, "fragments of two flows must fit");
};

// layers 2 and 3 of one net from its layer-1 PRE-activations (in place: a1 -> h1), keeping h2: fcnn_tail16 for the backward
__device__ __forceinline__ void cnf_tail_fwd(const float* __restrict__ img, float (&a1)[H], float (&h2)[H], float (&out)[16]) {
    using L = LC;
#pragma unroll
    for (int k = 0; k < H; k += 2) tanh_prescaled_pair(a1[k], a1[k + 1], a1[k], a1[k + 1]);
    float b2[8];
    ld8(img + L::B2, b2);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
        float p0 = b2[j], p1 = 0.f;
#pragma unroll
        for (int k = 0; k < H; k += 2) ffma2_p(p0, p1, w[k], w[k + 1], a1[k], a1[k + 1]);
        h2[j] = p0 + p1;
    }
#pragma unroll
    for (int j = 0; j < H; j += 2) tanh_prescaled_pair(h2[j], h2[j + 1], h2[j], h2[j + 1]);
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
        float p0 = img[L::B3 + o], p1 = 0.f;
#pragma unroll
        for (int j = 0; j < H; j += 2) ffma2_p(p0, p1, w[j], w[j + 1], h2[j], h2[j + 1]);
        out[o] = p0 + p1;
    }
}
// data gradient of the tail: dout -> delta2, delta1 (both divided by TANH_SCALE, as in fcnn_bwd)
__device__ __forceinline__ void cnf_tail_bwd(const float* __restrict__ img, const float (&dout)[16], const float (&h1)[H],
                                             const float (&h2)[H], float (&d1)[H], float (&d2)[H]) {
    using L = LC;
    float da2[H], da1[H];
#pragma unroll
    for (int j = 0; j < H; ++j) { da2[j] = 0.f; da1[j] = 0.f; }
#pragma unroll
    for (int o = 0; o < 16; ++o) {
        float w[8];
        ld8(img + L::W3 + o * H, w);
#pragma unroll
        for (int j = 0; j < H; j += 2) ffma2_s(da2[j], da2[j + 1], dout[o], w[j], w[j + 1]);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) d2[j] = da2[j] * fmaf(-TANH_ISCALE * h2[j], h2[j], TANH_ISCALE);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float w[8];
        ld8(img + L::W2 + j * H, w);
#pragma unroll
        for (int k = 0; k < H; k += 2) ffma2_s(da1[k], da1[k + 1], d2[j], w[k], w[k + 1]);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) d1[k] = da1[k] * fmaf(-TANH_ISCALE * h1[k], h1[k], TANH_ISCALE);
}

// weight gradients of net f from the tile rows of its block (gradient warp; contracts over the 32 columns of warp pair pw):
// CnfTmemSink::accumulate with the row map of CnfB
__device__ __forceinline__ void cnf_accumulate_ws(const float* __restrict__ s_tile, int pw, uint32_t col, float* __restrict__ bias_f,
                                                  int nb) {
    using G = CnfB;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int k0 = 32 * pw;
    float a[CNF_COLS];
    umma::ld_frag<16>(col, a);
    umma::ld2w(col + 16, a + 16);
    {   // A = [delta1; delta2] x input tiles 0..3 -> rows 0-7: dW1[g][8 n + 2 t, +1]
        float c[4][4];
        int rowB[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) { rowB[n] = G::C + 8 * n + g; c[n][0] = 0.f; c[n][1] = 0.f; c[n][2] = 0.f; c[n][3] = 0.f; }
        mma_outer<4>(s_tile, nb + G::D1, rowB, k0, k0 + 32, c);
#pragma unroll
        for (int n = 0; n < 4; ++n) { a[2 * n] += c[n][0]; a[2 * n + 1] += c[n][1]; }
    }
    {   // input tiles 4, 5 | ONE (db1 rows 0-7, db2 rows 8-15) | h1 (rows 8-15: dW2[g][2 t, +1])
        float c[4][4] = {};
        const int rowB[4] = {G::C + 32 + g, G::C + 40 + g, g == 0 ? G::ONE : G::ZERO, nb + G::H1 + g};
        mma_outer<4, 1u << 2>(s_tile, nb + G::D1, rowB, k0, k0 + 32, c);
        a[8] += c[0][0]; a[9] += c[0][1]; a[10] += c[1][0]; a[11] += c[1][1]; a[12] += c[3][2]; a[13] += c[3][3];
        if (t == 0) { bias_f[g] += c[2][0]; bias_f[8 + g] += c[2][2]; }
    }
    {   // A = d out (16 rows) x [h2 | ONE] -> dW3[g][2 t, +1], dW3[g + 8][..], db3
        float c[2][4] = {};
        const int rowB[2] = {nb + G::H2 + g, g == 0 ? G::ONE : G::ZERO};
        mma_outer<2, 1u << 1>(s_tile, nb + G::DO, rowB, k0, k0 + 32, c);
        a[14] += c[0][0]; a[15] += c[0][1]; a[16] += c[0][2]; a[17] += c[0][3];
        if (t == 0) { bias_f[16 + g] += c[1][0]; bias_f[24 + g] += c[1][2]; }
    }
    umma::st_frag<16>(col, a);
    umma::st2(col + 16, a + 16);
    umma::wait_st();
}

__global__ void __launch_bounds__(WS_THREADS, 1)
measure_bwd_cnf_ws_kernel(const float* __restrict__ pe, const float* __restrict__ cnf, int n_flows, float p0, float p1,
                          const float* __restrict__ enc, const float* __restrict__ particles, int B, int N,
                          const float* __restrict__ g_lki, const int* __restrict__ argmax, float* __restrict__ d_particles,
                          float* __restrict__ d_enc, float* __restrict__ part_pe, float* __restrict__ part_cnf,
                          const float* __restrict__ z_saved, const float* __restrict__ g_pred, const float* __restrict__ probs) {
    using G = CnfB;
    extern __shared__ __align__(128) float smem[];
    __shared__ float s_red[4];
    __shared__ uint64_t s_bar;
    __shared__ uint64_t s_full[4], s_empty[4];
    __shared__ uint32_t s_tslot;
    constexpr int TILE_FLOATS = (G::ROWS * TSM + 31) & ~31;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_data = warp < 4;
    const int pw = warp & 3;                                 // warp pair = tensor-memory lane quarter
    const int n_fcnn = 4 * n_flows, n_st = 2 * n_flows;
    float* s_tile = smem;                                    // [TILE_FLOATS]
    float* s_tcw = s_tile + TILE_FLOATS;                     // encoder operand tiles
    float* s_l1f = s_tcw + PeTc::WBWD_FLOATS;                // [n_st] layer-1 operand tiles, forward  ([16 units][48 inputs], hi | lo)
    float* s_l1b = s_l1f + n_st * CnfL1::STAGE_FLOATS;       // [n_st] layer-1 operand tiles, backward ([48 inputs][16 units], hi | lo)
    float* s_tail = s_l1b + n_st * G::L1B_STAGE_FLOATS;      // [n_fcnn][CnfL1::TAIL]
    float* s_hb = s_tail + n_fcnn * CnfL1::TAIL;             // [n_fcnn][8]
    float* s_pe = s_hb + n_fcnn * H;
    float* s_enc = s_pe + PE_SIZE;
    float* s_cnfbias = s_enc + 36;                           // [4][n_fcnn][32]  bias gradients of the gradient warps (t == 0 lanes)
    float* s_denc = s_cnfbias + 4 * n_fcnn * 32;             // [4][32]
    float* s_accpe = s_tile;                                 // read-out staging at the very end
    static_assert(4 * AC::SIZE <= TILE_FLOATS, "read-out staging must fit in the tile");
    if (is_data) {   // the constant rows of the bias columns, once (nothing aliases them)
        s_tile[G::ONE * TSM + tid] = 1.0f;
        s_tile[G::ZERO * TSM + tid] = 0.0f;
    }
    if (tid < 32) umma::tmem_alloc<G::TMEM_COLS>(&s_tslot);
    if (tid == 0) {
        umma::mbar_init(&s_bar, 1);
        for (int w = 0; w < 4; ++w) { umma::mbar_init(&s_full[w], 32); umma::mbar_init(&s_empty[w], 32); }
    }
    PeTcWs tc;
    tc.w = s_tcw; tc.bar = &s_bar; tc.tmem = 0u; tc.parity = 0u;
    tc.load_weights(pe, true);
    for (int e = tid; e < PE_SIZE; e += WS_THREADS) s_pe[e] = pe[e];
    CnfL1::load(cnf, n_fcnn, s_tail, s_hb, s_l1f);
    {
        const int pf = packed_fcnn_size(16, 32);
        for (int e = tid; e < n_st * 48 * 16; e += WS_THREADS) {
            const int st = e / (48 * 16), i = (e >> 4) % 48, u = e & 15;
            const float w = TANH_SCALE * cnf[(size_t)(2 * st + (u >> 3)) * pf + (u & 7) * 48 + i];
            G::WB::store_elem(s_l1b + st * G::L1B_STAGE_FLOATS, s_l1b + st * G::L1B_STAGE_FLOATS + G::WB::FLOATS, i, u, w);
        }
        for (int e = tid; e < 4 * n_fcnn * 32; e += WS_THREADS) s_cnfbias[e] = 0.f;
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    tc.tmem = s_tslot;
    const uint32_t lane_base = tc.tmem + ((uint32_t)(32 * pw) << 16);          // this warp's tensor-memory lanes
    const uint32_t tacc = lane_base + G::COL_FRAG;                             // the pair's gradient fragments
    CnfTmemSink sink{tacc + TA_CNF, s_cnfbias + pw * n_fcnn * 32};
    uint32_t ph_full = 0, ph_empty = 1;                      // the tile starts out empty: the first `empty` wait passes

    if (!is_data) {
        // ------------------------------------------------------------------------------------ gradient warps
        {
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            for (int c0 = 0; c0 < TA_END + CNF_COLS * n_fcnn; c0 += 4) umma::st4(tacc + c0, z);
            umma::wait_st();
        }
        for (int b = blockIdx.x; b < B; b += gridDim.x) {
            for (int n0 = 0; n0 < N; n0 += 128) {
#pragma unroll 1
                for (int st = n_st - 1; st >= 0; --st) {
                    umma::mbar_wait(&s_full[pw], ph_full); ph_full ^= 1;
#pragma unroll 1
                    for (int net = 0; net < 2; ++net)
                        cnf_accumulate_ws(s_tile, pw, sink.tcol + CNF_COLS * (2 * st + net), sink.bias + (2 * st + net) * 32,
                                          G::NET + net * G::NET_ROWS);
                    mbar_arrive(&s_empty[pw]);
                }
                umma::mbar_wait(&s_full[pw], ph_full); ph_full ^= 1;
                pe_weight_grads_a_x2(s_tile, tacc, pw);
                umma::wait_st();
                mbar_arrive(&s_empty[pw]);
                umma::mbar_wait(&s_full[pw], ph_full); ph_full ^= 1;
                pe_weight_grads_b_x2(s_tile, tacc, pw);
                mbar_arrive(&s_empty[pw]);
            }
        }
    } else {
        // ------------------------------------------------------------------------------------ data warps
        auto stage = [&](int row, float v) { s_tile[row * TSM + tid] = v; };
        for (int b = blockIdx.x; b < B; b += gridDim.x) {
            data_group_sync();                               // the previous trajectory's readers of s_enc / s_red / s_denc are done
            load_enc(enc + (size_t)b * HID, s_enc);
            const size_t base = (size_t)b * N;
            float gs = 0.f;                                  // row-max shift: d ll[n] = g[n] - [n == argmax] * sum_m g[m]
            for (int n = tid; n < N; n += 128) gs += g_lki[base + n];
            gs = warp_sum(gs);
            if (lane == 0) s_red[warp] = gs;
            data_group_sync();
            gs = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
            const int am = argmax[b];
            {   // this thread's running d_enc sums of the trajectory wait in tensor memory (32 registers less in the stage loop)
                float z[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) z[k] = 0.f;
                umma::st_frag<32>(lane_base + G::COL_DENC, z);
            }
            float2 x_next = *reinterpret_cast<const float2*>(particles + (base + (tid < N ? tid : 0)) * 2);
            float g_next = g_lki[base + (tid < N ? tid : 0)];
            for (int n0 = 0; n0 < N; n0 += 128) {
                asm volatile("" ::: "memory");
                const int n = n0 + tid;
                const bool live = n < N;
                const size_t p = base + (live ? n : 0);
                const float2 x = x_next;
                float g = live ? g_next : 0.f;
                if (live && n == am) g -= gs;
                {
                    const size_t pn = base + (n + 128 < N ? n + 128 : 0);
                    x_next = *reinterpret_cast<const float2*>(particles + pn * 2);
                    g_next = g_lki[pn];
                }
                float lo[16], up[16], glo[16], gup[16];
                // the flow output of the forward (one 128-byte line per thread): on its way while the encoder's rounds run
                asm volatile("prefetch.global.L1 [%0];" ::"l"(z_saved + p * 32));
                {
                    float a1[16], a2[32], e[32];
                    pe_fwd_tc(tc, s_pe, x.x, x.y, a1, a2, e);        // rounds 1-2: encoder layers 2 and 3 forward
                    umma::st_frag<16>(lane_base + G::COL_A1, a1);    // parked until the encoder's backward
                    umma::st_frag<32>(lane_base + G::COL_A2, a2);
                    {
                        float hi[32], l[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) umma::split(e[i], hi[i], l[i]);
                        umma::st_frag<32>(lane_base + G::COL_AHI + 16, hi);
                        umma::st_frag<32>(lane_base + G::COL_ALO + 16, l);
#pragma unroll
                        for (int i = 0; i < 32; ++i) hi[i] = 0.f;
                        umma::st_frag<32>(lane_base + G::COL_DE, hi);    // d e accumulates over the stack: starts from zero
                    }
                    // the gradient warp has finished the previous batch (its encoder phase B): the tile is this warp's until the
                    // first stage is handed over
                    umma::mbar_wait(&s_empty[pw], ph_empty); ph_empty ^= 1;
#pragma unroll
                    for (int k = 0; k < 32; ++k) stage(G::E + k, e[k]);
                }
                {
                    const float4* zi = reinterpret_cast<const float4*>(z_saved + p * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 a = zi[i], c4 = zi[4 + i];
                        lo[4 * i] = a.x; lo[4 * i + 1] = a.y; lo[4 * i + 2] = a.z; lo[4 * i + 3] = a.w;
                        up[4 * i] = c4.x; up[4 * i + 1] = c4.y; up[4 * i + 2] = c4.z; up[4 * i + 3] = c4.w;
                    }
                    const float c = -g / (p1 * p1);          // dead threads: g = 0, so glo / gup and every delta below stay exactly 0
#pragma unroll
                    for (int i = 0; i < 16; ++i) { glo[i] = c * (lo[i] - p0); gup[i] = c * (up[i] - p0); }
                }
                swap_halves<16>(lo, up); swap_halves<16>(glo, gup);     // the last forward stage had c = upper
                // this thread's conditioning half (hi | lo) -> the first 16 columns of the layer-1 activation row
                auto put_c = [&](const float (&c)[16]) {
                    float hi[16], l[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) umma::split(c[i], hi[i], l[i]);
                    umma::st_frag<16>(lane_base + G::COL_AHI, hi);
                    umma::st_frag<16>(lane_base + G::COL_ALO, l);
                };
                // leader: pre = [c | e] W1cat^T of stage st (18 MMAs)
                auto issue_fwd = [&](int st) {
                    constexpr uint32_t idesc = umma::idesc_tf32(128, 16);
                    const float* w_hi = s_l1f + st * CnfL1::STAGE_FLOATS;
                    const float* w_lo = w_hi + CnfL1::W::FLOATS;
                    const uint32_t d = tc.tmem + G::COL_PRE, a_hi = tc.tmem + G::COL_AHI, a_lo = tc.tmem + G::COL_ALO;
                    uint32_t acc = 0;
#pragma unroll
                    for (int k0 = 0; k0 < 48; k0 += 8) { umma::mma_tf32_ts(d, a_lo + k0, CnfL1::W::desc(w_hi, k0), idesc, acc); acc = 1; }
#pragma unroll
                    for (int k0 = 0; k0 < 48; k0 += 8) umma::mma_tf32_ts(d, a_hi + k0, CnfL1::W::desc(w_lo, k0), idesc, 1);
#pragma unroll
                    for (int k0 = 0; k0 < 48; k0 += 8) umma::mma_tf32_ts(d, a_hi + k0, CnfL1::W::desc(w_hi, k0), idesc, 1);
                };
                put_c(lo);
                umma::wait_st();
                umma::fence_before_sync();
                data_group_sync();
                if (tid == 0) {
                    umma::fence_after_sync();
                    issue_fwd(n_st - 1);
                    umma::commit(tc.bar);
                }
                tc.wait();
                float pre[16];
                umma::ld16(lane_base + G::COL_PRE, pre);
#pragma unroll 1
                for (int st = n_st - 1; st >= 0; --st) {                // walk the forward stages back
                    float h1t[H], h2t[H], h1s[H], h2s[H], ds[16], d1t[H], d1s[H], d2s[H];
                    {
                        const float* hb = s_hb + 2 * st * H;
#pragma unroll
                        for (int k = 0; k < H; ++k) { h1t[k] = pre[k] + hb[k]; h1s[k] = pre[H + k] + hb[H + k]; }
                    }
                    const float* im_t = CnfL1::tail_image(s_tail, 2 * st);
                    const float* im_s = CnfL1::tail_image(s_tail, 2 * st + 1);
                    {
                        float t[16], s[16];
                        cnf_tail_fwd(im_t, h1t, h2t, t);
                        cnf_tail_fwd(im_s, h1s, h2s, s);
                        {   // t-net: its d out is the incoming gradient of the transformed half (out = t + in e^s); staged at once
                            float d2t[H];
                            cnf_tail_bwd(im_t, gup, h1t, h2t, d1t, d2t);
                            if (st != n_st - 1) { umma::mbar_wait(&s_empty[pw], ph_empty); ph_empty ^= 1; }
#pragma unroll
                            for (int i = 0; i < 16; ++i) { stage(G::C + i, lo[i]); stage(G::NET + G::DO + i, gup[i]); }
#pragma unroll
                            for (int k = 0; k < H; ++k) {
                                stage(G::NET + G::H1 + k, h1t[k]); stage(G::NET + G::H2 + k, h2t[k]);
                                stage(G::NET + G::D1 + k, d1t[k]); stage(G::NET + G::D2 + k, d2t[k]);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i) {                  // coupling.cuh stage_bwd, forward direction
                            const float es = exp_acc(s[i]), ies = exp_acc(-s[i]);
                            const float vin = (up[i] - t[i]) * ies;
                            ds[i] = fmaf(gup[i] * vin, es, g);
                            gup[i] = gup[i] * es;
                            up[i] = vin;
                        }
                    }
                    if (st > 0) put_c(up);          // the next stage conditions on the half just inverted: its forward product rides along
                    cnf_tail_bwd(im_s, ds, h1s, h2s, d1s, d2s);
                    {   // delta1 of both nets (K = 16) and a zeroed d c accumulator -> tensor memory
                        float hi[16], l[16];
#pragma unroll
                        for (int k = 0; k < H; ++k) { umma::split(d1t[k], hi[k], l[k]); umma::split(d1s[k], hi[H + k], l[H + k]); }
                        umma::st_frag<16>(lane_base + G::COL_D1HI, hi);
                        umma::st_frag<16>(lane_base + G::COL_D1LO, l);
#pragma unroll
                        for (int k = 0; k < 16; ++k) hi[k] = 0.f;
                        umma::st_frag<16>(lane_base + G::COL_DC, hi);
                    }
                    umma::wait_st();
                    umma::fence_before_sync();
                    data_group_sync();
                    if (tid == 0) {
                        umma::fence_after_sync();
                        constexpr uint32_t id48 = umma::idesc_tf32(128, 48);
                        const float* b_hi = s_l1b + st * G::L1B_STAGE_FLOATS;
                        const float* b_lo = b_hi + G::WB::FLOATS;
                        const uint32_t a_hi = tc.tmem + G::COL_D1HI, a_lo = tc.tmem + G::COL_D1LO, d = tc.tmem + G::COL_DC;
                        // [d c | d e] += delta1 W1cat
                        umma::mma_tf32_ts(d, a_lo, G::WB::desc(b_hi, 0), id48, 1);
                        umma::mma_tf32_ts(d, a_lo + 8, G::WB::desc(b_hi, 8), id48, 1);
                        umma::mma_tf32_ts(d, a_hi, G::WB::desc(b_lo, 0), id48, 1);
                        umma::mma_tf32_ts(d, a_hi + 8, G::WB::desc(b_lo, 8), id48, 1);
                        umma::mma_tf32_ts(d, a_hi, G::WB::desc(b_hi, 0), id48, 1);
                        umma::mma_tf32_ts(d, a_hi + 8, G::WB::desc(b_hi, 8), id48, 1);
                        if (st > 0) issue_fwd(st - 1);
                        umma::commit(tc.bar);
                    }
                    // hand the stage to the gradient warp while the products are in flight
#pragma unroll
                    for (int i = 0; i < 16; ++i) stage(G::NET + G::NET_ROWS + G::DO + i, ds[i]);
#pragma unroll
                    for (int k = 0; k < H; ++k) {
                        stage(G::NET + G::NET_ROWS + G::H1 + k, h1s[k]); stage(G::NET + G::NET_ROWS + G::H2 + k, h2s[k]);
                        stage(G::NET + G::NET_ROWS + G::D1 + k, d1s[k]); stage(G::NET + G::NET_ROWS + G::D2 + k, d2s[k]);
                    }
                    mbar_arrive(&s_full[pw]);
                    tc.wait();
                    {
                        float dc[16];
                        umma::ld16(lane_base + G::COL_DC, dc);
#pragma unroll
                        for (int i = 0; i < 16; ++i) glo[i] += dc[i];
                    }
                    if (st > 0) umma::ld16(lane_base + G::COL_PRE, pre);
                    swap_halves<16>(lo, up); swap_halves<16>(glo, gup);
                }
                swap_halves<16>(lo, up); swap_halves<16>(glo, gup);
                {
                    float denc[32];
                    umma::ld32(lane_base + G::COL_DENC, denc);
#pragma unroll
                    for (int i = 0; i < 16; ++i) { denc[i] += live ? glo[i] : 0.f; denc[16 + i] += live ? gup[i] : 0.f; }
                    umma::st_frag<32>(lane_base + G::COL_DENC, denc);
                }
                // encoder backward: delta3 = d e (dead threads: every delta above was zeroed, so the accumulator row is exactly 0)
                float de[32], a2[32];
                umma::ld32(lane_base + G::COL_DE, de);
                umma::ld32(lane_base + G::COL_A2, a2);
                umma::mbar_wait(&s_empty[pw], ph_empty); ph_empty ^= 1;
                uint32_t m2 = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    stage(PR::D3 + j, de[j]);
                    stage(PR::A2 + j, a2[j]);
                    m2 |= (a2[j] > 0.f ? 1u : 0u) << j;
                }
                mbar_arrive(&s_full[pw]);
                float d2[32], d1[16], a1[16];
                tc.template store_row<32>(de);
                tc.template round<32, 32>(PeTc::W3T_HI, PeTc::W3T_LO);
                umma::ld32(lane_base, d2);
#pragma unroll
                for (int j = 0; j < 32; ++j) d2[j] = (m2 >> j) & 1u ? d2[j] : 0.f;
                tc.template store_row<32>(d2);
                tc.template round<16, 32>(PeTc::W2T_HI, PeTc::W2T_LO);
                umma::ld16(lane_base, d1);
                umma::ld16(lane_base + G::COL_A1, a1);
                float dx0 = 0.f, dx1 = 0.f;
#pragma unroll
                for (int k = 0; k < 16; k += 2) {
                    const float4 q = *reinterpret_cast<const float4*>(s_pe + PE_W1 + 2 * k);
                    d1[k] = a1[k] > 0.f ? d1[k] : 0.f;
                    d1[k + 1] = a1[k + 1] > 0.f ? d1[k + 1] : 0.f;
                    dx0 = fmaf(q.x, d1[k], dx0); dx1 = fmaf(q.y, d1[k], dx1);
                    dx0 = fmaf(q.z, d1[k + 1], dx0); dx1 = fmaf(q.w, d1[k + 1], dx1);
                }
                if (live) {
                    if (g_pred) {      // fused prediction (losses.py:22): d pred / d x_n = probs[n]
                        const float pr = probs[p];
                        dx0 = fmaf(g_pred[2 * b], pr, dx0); dx1 = fmaf(g_pred[2 * b + 1], pr, dx1);
                    }
                    *reinterpret_cast<float2*>(d_particles + p * 2) = make_float2(dx0, dx1);
                }
                umma::mbar_wait(&s_empty[pw], ph_empty); ph_empty ^= 1;
                stage(PR::X + 0, x.x);
                stage(PR::X + 1, x.y);
#pragma unroll
                for (int k = 0; k < 16; ++k) { stage(PR::A1 + k, a1[k]); stage(PR::D1 + k, d1[k]); }
#pragma unroll
                for (int j = 0; j < 32; ++j) stage(PR::D2 + j, d2[j]);
                mbar_arrive(&s_full[pw]);
            }
            // d_enc[b][k] = sum over the row's particles of the stack's input gradient (fixed order: butterfly, then warps)
            if (d_enc) {
                float denc[32];
                umma::wait_st();
                umma::ld32(lane_base + G::COL_DENC, denc);
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    const float v = warp_sum(denc[k]);
                    if (lane == 0) s_denc[warp * 32 + k] = v;
                }
                data_group_sync();
                if (tid < 32) d_enc[(size_t)b * HID + tid] = (s_denc[tid] + s_denc[32 + tid]) + (s_denc[64 + tid] + s_denc[96 + tid]);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();                                   // every contraction is done: the tile becomes the read-out staging area
    umma::fence_after_sync();
    for (int e = tid; e < 4 * AC::SIZE; e += WS_THREADS) s_accpe[e] = 0.f;
    __syncthreads();
    if (!is_data) {
        pe_weight_grads_readout(tacc, s_accpe + pw * AC::SIZE);
        sink.readout(n_fcnn, part_cnf + ((size_t)blockIdx.x * 4 + pw) * n_fcnn * packed_fcnn_size(16, 32));
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_free<G::TMEM_COLS>(tc.tmem);
    for (int e = tid; e < PE_SIZE; e += WS_THREADS) {
        const float* a = s_accpe + AC::of_packed(e);
        part_pe[(size_t)blockIdx.x * PE_SIZE + e] = (a[0] + a[AC::SIZE]) + (a[2 * AC::SIZE] + a[3 * AC::SIZE]);
    }
}

// ---- NN likelihood, backward (mode 3) --------------------------------------------------------------------------------
// One 128-thread CTA per SM (the head's four operand tiles alone are 96 KB), persistent over trajectories.  Per 128-particle batch:
//   data path (thread = particle = tensor-memory lane): encoder forward (2 rounds), head forward (2 rounds: N = 64, K = 32 / 64),
//     dz = g (1 - sigmoid z), delta2, round W2^T -> delta1, round W1e^T -> d e, then the encoder's backward rounds as in the other modes;
//   head weight gradients on the CUDA cores from the transposed tile (row = feature, column = particle): a thread owns a 4 x 8
//     block of dW2 and a 4 x 4 block of dW1[:, 32:] IN REGISTERS for the CTA's whole life and walks the batch's 128 particles with
//     128-bit loads (rows interleaved by 16 / 8 so that a warp's loads are conflict-free); db1 / db2 are row sums of the same tile,
//     dW3 a transposed warp butterfly of dz h2; the observation half dW1[:, :32] = (sum_p delta1) (x) enc and d_enc = W1[:, :32]^T
//     (sum_p delta1) are formed once per trajectory.  IEEE adds throughout (no tensor-core accumulation across batches);
//   encoder weight gradients: the mma.sync phases A / B of the other modes, fragments in tensor memory (columns 192..259: the CTA
//     owns all 512 columns).
struct NnB {
    using W1E = umma::Operand<NH, 32>;
    using W2 = umma::Operand<NH, NH>;
    using W2T = umma::Operand<NH, NH>;    // rows k (layer-2 input), K = j:   d h1[k] = sum_j delta2[j] W2[j][k]
    using W1ET = umma::Operand<32, NH>;   // rows c (encoding index), K = j:  d e[c]  = sum_j delta1[j] W1[j][32 + c]
    static constexpr int W1E_HI = 0, W1E_LO = W1E_HI + W1E::FLOATS, W2_HI = W1E_LO + W1E::FLOATS, W2_LO = W2_HI + W2::FLOATS,
                         W2T_HI = W2_LO + W2::FLOATS, W2T_LO = W2T_HI + W2T::FLOATS, W1ET_HI = W2T_LO + W2T::FLOATS,
                         W1ET_LO = W1ET_HI + W1ET::FLOATS, TILE_FLOATS = W1ET_LO + W1ET::FLOATS;
    static constexpr int HB1 = TILE_FLOATS, B2 = HB1 + NH, W3 = B2 + NH, B3 = W3 + NH, D1SUM = B3 + 4, ENC = D1SUM + NH, W3ACC = ENC + 32,
                         FLOATS = W3ACC + 4 * NH;
    // gradient-tile rows: the encoder phases (PR, rows 0..67) alias rows R_H1.. once the head phases are done
    static constexpr int R_H1 = 2, R_D = R_H1 + NH, R_E = R_D + NH, ROWS = R_E + 32;
    static constexpr int TMEM_COLS = 512, TACC = 192;
};

// acc[r][c] += sum_p tile[rowA0 + jg + 16 r][p] * tile[rowB0 + kg + 8 c][p] over the batch's 128 particles
template <int NB>
__device__ __forceinline__ void nn_contract(const float* __restrict__ tile, int rowA0, int rowB0, int jg, int kg, float (&acc)[4][NB]) {
#pragma unroll 2
    for (int p = 0; p < 128; p += 4) {
        float4 a[4], b[NB];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(tile + (rowA0 + jg + 16 * r) * TSM + p);
#pragma unroll
        for (int c = 0; c < NB; ++c) b[c] = *reinterpret_cast<const float4*>(tile + (rowB0 + kg + 8 * c) * TSM + p);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < NB; ++c)
                acc[r][c] += fmaf(a[r].x, b[c].x, fmaf(a[r].y, b[c].y, fmaf(a[r].z, b[c].z, a[r].w * b[c].w)));
    }
}
__device__ __forceinline__ float nn_rowsum(const float* __restrict__ row) {
    float s = 0.f;
#pragma unroll 4
    for (int p = 0; p < 128; p += 4) { const float4 v = *reinterpret_cast<const float4*>(row + p); s += (v.x + v.y) + (v.z + v.w); }
    return s;
}
// one TS round of the head's backward: A = the K values every thread just wrote, D[0, N) = A W^T
template <int N, int K>
__device__ __forceinline__ void nn_round_nk(PeTc& tc, const float (&a)[K], const float* w_hi, const float* w_lo, float (&d)[N]) {
    {
        float hi[K], lo[K];
#pragma unroll
        for (int k = 0; k < K; ++k) umma::split(a[k], hi[k], lo[k]);
        umma::st_frag<K>(tc.lane_addr() + NnHead::COL_AHI, hi);
        umma::st_frag<K>(tc.lane_addr() + NnHead::COL_ALO, lo);
    }
    umma::wait_st();
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0) {
        umma::fence_after_sync();
        umma::gemm3_ts<N, K>(tc.tmem + NnHead::COL_D, tc.tmem + NnHead::COL_AHI, tc.tmem + NnHead::COL_ALO, w_hi, w_lo);
        umma::commit(tc.bar);
    }
    tc.wait();
    float lo32[32];
    umma::ld32(tc.lane_addr() + NnHead::COL_D, lo32);
#pragma unroll
    for (int j = 0; j < 32; ++j) d[j] = lo32[j];
    if (N == 64) {
        float hi32[32];
        umma::ld32(tc.lane_addr() + NnHead::COL_D + 32, hi32);
#pragma unroll
        for (int j = 0; j < 32; ++j) d[(N == 64 ? 32 : 0) + j] = hi32[j];
    }
}

__global__ void __launch_bounds__(TP, 1)
measure_bwd_nn_kernel(const float* __restrict__ pe, const float* __restrict__ head, const float* __restrict__ enc,
                      const float* __restrict__ particles, int B, int N, const float* __restrict__ g_lki,
                      float* __restrict__ d_particles, float* __restrict__ d_enc, float* __restrict__ part_pe,
                      float* __restrict__ part_head, const float* __restrict__ g_pred, const float* __restrict__ probs) {
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_tslot;
    constexpr int TILE_FLOATS = (NnB::ROWS * TSM + 31) & ~31;
    constexpr int NW = TP / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* s_tile = smem;                                    // [TILE_FLOATS]
    float* s_tcw = s_tile + TILE_FLOATS;                     // encoder tensor-core weight tiles
    float* s_pe = s_tcw + PeTc::WBWD_FLOATS;
    float* s_nn = s_pe + PE_SIZE;                            // head operand tiles + vectors (NnB)
    float* s_accpe = s_tile;                                 // read-out staging at the very end
    static_assert(NW * AC::SIZE <= TILE_FLOATS, "read-out staging must fit in the tile");
    static_assert((TILE_FLOATS + PeTc::WBWD_FLOATS + PE_SIZE) % 4 == 0, "head tiles must stay 16-byte aligned");
    if (tid < 32) umma::tmem_alloc<NnB::TMEM_COLS>(&s_tslot);
    if (tid == 0) umma::mbar_init(&s_bar, 1);
    PeTc tc{s_tcw, &s_bar, 0u, 0u};
    tc.load_weights(pe, true);
    for (int e = tid; e < PE_SIZE; e += TP) s_pe[e] = pe[e];
    for (int e = tid; e < NH * 32; e += TP) {
        const int j = e >> 5, c = e & 31;
        const float wv = head[NH_W1 + j * NH + 32 + c];
        NnB::W1E::store_elem(s_nn + NnB::W1E_HI, s_nn + NnB::W1E_LO, j, c, wv);
        NnB::W1ET::store_elem(s_nn + NnB::W1ET_HI, s_nn + NnB::W1ET_LO, c, j, wv);
    }
    for (int e = tid; e < NH * NH; e += TP) {
        const int j = e >> 6, k = e & 63;
        const float wv = head[NH_W2 + e];
        NnB::W2::store_elem(s_nn + NnB::W2_HI, s_nn + NnB::W2_LO, j, k, wv);
        NnB::W2T::store_elem(s_nn + NnB::W2T_HI, s_nn + NnB::W2T_LO, k, j, wv);
    }
    for (int j = tid; j < NH; j += TP) { s_nn[NnB::B2 + j] = head[NH_B2 + j]; s_nn[NnB::W3 + j] = head[NH_W3 + j]; }
    if (tid == 0) s_nn[NnB::B3] = head[NH_B3];
    s_tile[PR::ONE * TSM + tid] = 1.0f;
    s_tile[PR::ZERO * TSM + tid] = 0.0f;
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    tc.tmem = s_tslot;
    const uint32_t tacc = tc.lane_addr() + NnB::TACC;
    {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c0 = 0; c0 < TA_END; c0 += 4) umma::st4(tacc + c0, z);
        umma::wait_st();
    }
    // ---- register tiles of the head's weight gradients (this thread's outputs, summed over everything the CTA sees)
    const int jg = tid >> 3, kg = tid & 7;                   // dW2[jg + 16 r][kg + 8 c], dW1[jg + 16 r][32 + kg + 8 c]
    const int oj = tid >> 1, oc = (tid & 1) * 16;            // dW1[oj][oc .. oc + 16): the observation half
    float acc2[4][8], acc1[4][4], acc1o[16], accb1 = 0.f, accb2 = 0.f, acc3a = 0.f, acc3b = 0.f, accb3 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc2[r][c] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc1[r][c] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) acc1o[i] = 0.f;

    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        __syncthreads();
        if (tid < 32) s_nn[NnB::ENC + tid] = enc[(size_t)b * HID + tid];
        __syncthreads();
        if (tid < NH) {     // observation half of layer 1, once per trajectory
            float a = head[NH_B1 + tid];
            for (int c = 0; c < 32; ++c) a = fmaf(head[NH_W1 + tid * NH + c], s_nn[NnB::ENC + c], a);
            s_nn[NnB::HB1 + tid] = a;
        }
        float d1traj = 0.f;                                  // threads < 64: sum_p delta1[tid] over this trajectory
        const size_t base = (size_t)b * N;
        __syncthreads();
        for (int n0 = 0; n0 < N; n0 += TP) {
            asm volatile("" ::: "memory");
            const int n = n0 + tid;
            const bool live = n < N;
            const size_t p = base + (live ? n : 0);
            const float2 x = *reinterpret_cast<const float2*>(particles + p * 2);
            const float g = live ? g_lki[p] : 0.f;
            float a1[16], a2[32], e[32];
            __syncthreads();                                 // the previous batch's encoder phases are done with the tile
            pe_fwd_tc(tc, s_pe, x.x, x.y, a1, a2, e);
#pragma unroll
            for (int k = 0; k < 32; ++k) s_tile[(NnB::R_E + k) * TSM + tid] = e[k];
            uint64_t m1 = 0ull;
            float dz;
            float d2[NH];
            {
                float h1[NH], h2[NH];
                nn_round<32>(tc, e, s_nn + NnB::W1E_HI, s_nn + NnB::W1E_LO, h1);
#pragma unroll
                for (int j = 0; j < NH; ++j) {
                    h1[j] = fmaxf(h1[j] + s_nn[NnB::HB1 + j], 0.f);
                    m1 |= (uint64_t)(h1[j] > 0.f) << j;
                    s_tile[(NnB::R_H1 + j) * TSM + tid] = h1[j];
                }
                nn_round<NH>(tc, h1, s_nn + NnB::W2_HI, s_nn + NnB::W2_LO, h2);
                float z = s_nn[NnB::B3];
#pragma unroll
                for (int j = 0; j < NH; ++j) { h2[j] = fmaxf(h2[j] + s_nn[NnB::B2 + j], 0.f); z = fmaf(s_nn[NnB::W3 + j], h2[j], z); }
                const float sg = 1.0f / (1.0f + expf(-z));
                dz = g * (1.0f - sg);                        // d log(sigmoid z) / dz
#pragma unroll
                for (int j = 0; j < NH; ++j) { d2[j] = h2[j] > 0.f ? dz * s_nn[NnB::W3 + j] : 0.f; h2[j] *= dz; }
                // dW3[j] += sum_p dz h2[j]: transposed butterfly over the warp, lane l ends up with entries 2l, 2l + 1
#define NFDPF_NN_ROUND(HALFN, OFF)                                                     \
                {                                                                       \
                    const bool up_ = (lane & OFF) != 0;                                 \
                    _Pragma("unroll") for (int i = 0; i < HALFN; ++i) {                 \
                        const float keep = up_ ? h2[i + HALFN] : h2[i];                 \
                        const float send = up_ ? h2[i] : h2[i + HALFN];                 \
                        h2[i] = keep + __shfl_xor_sync(FULL, send, OFF);                \
                    }                                                                   \
                }
                NFDPF_NN_ROUND(32, 16) NFDPF_NN_ROUND(16, 8) NFDPF_NN_ROUND(8, 4) NFDPF_NN_ROUND(4, 2) NFDPF_NN_ROUND(2, 1)
#undef NFDPF_NN_ROUND
                acc3a += h2[0]; acc3b += h2[1];
                accb3 += warp_sum(dz);
            }
#pragma unroll
            for (int j = 0; j < NH; ++j) s_tile[(NnB::R_D + j) * TSM + tid] = d2[j];
            float d1[NH];
            nn_round_nk<NH, NH>(tc, d2, s_nn + NnB::W2T_HI, s_nn + NnB::W2T_LO, d1);     // (its barrier also publishes the tile rows)
#pragma unroll
            for (int j = 0; j < NH; ++j) d1[j] = (m1 >> j) & 1ull ? d1[j] : 0.f;
            // head weight gradients, phase H2: dW2 += delta2 (x) h1, db2 += sum_p delta2
            nn_contract<8>(s_tile, NnB::R_D, NnB::R_H1, jg, kg, acc2);
            if (tid < NH) accb2 += nn_rowsum(s_tile + (NnB::R_D + tid) * TSM);
            __syncthreads();                                 // everybody has read delta2: the rows take delta1
#pragma unroll
            for (int j = 0; j < NH; ++j) s_tile[(NnB::R_D + j) * TSM + tid] = d1[j];
            float de[32];
            nn_round_nk<32, NH>(tc, d1, s_nn + NnB::W1ET_HI, s_nn + NnB::W1ET_LO, de);
            // phase H1: dW1[:, 32:] += delta1 (x) e, db1 += sum_p delta1 (also per trajectory: the observation half)
            nn_contract<4>(s_tile, NnB::R_D, NnB::R_E, jg, kg, acc1);
            if (tid < NH) { const float rs = nn_rowsum(s_tile + (NnB::R_D + tid) * TSM); accb1 += rs; d1traj += rs; }
            __syncthreads();                                 // the head phases are done with the tile: the encoder phases alias it
            // ---- encoder backward, as in the other modes
            float ed2[32], ed1[16];
            tc.template store_row<32>(de);
            tc.template round<32, 32>(PeTc::W3T_HI, PeTc::W3T_LO);
            umma::ld32(tc.lane_addr(), ed2);
#pragma unroll
            for (int j = 0; j < 32; ++j) ed2[j] = a2[j] > 0.f ? ed2[j] : 0.f;
            tc.template store_row<32>(ed2);
            tc.template round<16, 32>(PeTc::W2T_HI, PeTc::W2T_LO);
            umma::ld16(tc.lane_addr(), ed1);
            float dx0 = 0.f, dx1 = 0.f;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const float4 q = *reinterpret_cast<const float4*>(s_pe + PE_W1 + 2 * k);
                ed1[k] = a1[k] > 0.f ? ed1[k] : 0.f;
                ed1[k + 1] = a1[k + 1] > 0.f ? ed1[k + 1] : 0.f;
                dx0 = fmaf(q.x, ed1[k], dx0); dx1 = fmaf(q.y, ed1[k], dx1);
                dx0 = fmaf(q.z, ed1[k + 1], dx0); dx1 = fmaf(q.w, ed1[k + 1], dx1);
            }
            if (live) {
                if (g_pred) {
                    const float pr = probs[p];
                    dx0 = fmaf(g_pred[2 * b], pr, dx0); dx1 = fmaf(g_pred[2 * b + 1], pr, dx1);
                }
                *reinterpret_cast<float2*>(d_particles + p * 2) = make_float2(dx0, dx1);
            }
            s_tile[PR::ONE * TSM + tid] = 1.0f;              // (rows 0-1 are not aliased by the head phases; cheap to keep right)
            s_tile[PR::ZERO * TSM + tid] = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { s_tile[(PR::D3 + j) * TSM + tid] = de[j]; s_tile[(PR::A2 + j) * TSM + tid] = a2[j]; }
            __syncwarp();
            pe_weight_grads_a(s_tile, tacc);
            __syncwarp();
            s_tile[(PR::X + 0) * TSM + tid] = x.x;
            s_tile[(PR::X + 1) * TSM + tid] = x.y;
#pragma unroll
            for (int k = 0; k < 16; ++k) { s_tile[(PR::A1 + k) * TSM + tid] = a1[k]; s_tile[(PR::D1 + k) * TSM + tid] = ed1[k]; }
#pragma unroll
            for (int j = 0; j < 32; ++j) s_tile[(PR::D2 + j) * TSM + tid] = ed2[j];
            __syncwarp();
            pe_weight_grads_b(s_tile, tacc);
            __syncwarp();
        }
        // ---- per trajectory: the observation half of layer 1
        __syncthreads();
        if (tid < NH) s_nn[NnB::D1SUM + tid] = d1traj;
        __syncthreads();
        {
            const float dj = s_nn[NnB::D1SUM + oj];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc1o[i] = fmaf(dj, s_nn[NnB::ENC + oc + i], acc1o[i]);
        }
        if (d_enc && tid < 32) {
            float a = 0.f;
            for (int j = 0; j < NH; ++j) a = fmaf(head[NH_W1 + j * NH + tid], s_nn[NnB::D1SUM + j], a);
            d_enc[(size_t)b * HID + tid] = a;
        }
    }
    // ---- write this CTA's partial gradients
    float* ph = part_head + (size_t)blockIdx.x * NH_SIZE;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 8; ++c) ph[NH_W2 + (jg + 16 * r) * NH + kg + 8 * c] = acc2[r][c];
#pragma unroll
        for (int c = 0; c < 4; ++c) ph[NH_W1 + (jg + 16 * r) * NH + 32 + kg + 8 * c] = acc1[r][c];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) ph[NH_W1 + oj * NH + oc + i] = acc1o[i];
    if (tid < NH) { ph[NH_B1 + tid] = accb1; ph[NH_B2 + tid] = accb2; }
    __syncthreads();
    s_nn[NnB::W3ACC + warp * NH + 2 * lane] = acc3a;
    s_nn[NnB::W3ACC + warp * NH + 2 * lane + 1] = acc3b;
    if (lane == 0) s_nn[NnB::D1SUM + warp] = accb3;
    __syncthreads();
    if (tid < NH) ph[NH_W3 + tid] = (s_nn[NnB::W3ACC + tid] + s_nn[NnB::W3ACC + NH + tid]) + (s_nn[NnB::W3ACC + 2 * NH + tid] + s_nn[NnB::W3ACC + 3 * NH + tid]);
    if (tid == 0) ph[NH_B3] = (s_nn[NnB::D1SUM] + s_nn[NnB::D1SUM + 1]) + (s_nn[NnB::D1SUM + 2] + s_nn[NnB::D1SUM + 3]);
    __syncthreads();                                   // every warp is done with the tile: it becomes the read-out staging area
    for (int e = tid; e < NW * AC::SIZE; e += TP) s_accpe[e] = 0.f;
    __syncthreads();
    pe_weight_grads_readout(tacc, s_accpe + warp * AC::SIZE);
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_free<NnB::TMEM_COLS>(tc.tmem);
    for (int e = tid; e < PE_SIZE; e += TP) {
        const float* a = s_accpe + AC::of_packed(e);
        part_pe[(size_t)blockIdx.x * PE_SIZE + e] = (a[0] + a[AC::SIZE]) + (a[2 * AC::SIZE] + a[3 * AC::SIZE]);
    }
}

static int launch_measure_bwd_nn(const float* pe, const float* head, const float* enc, const float* particles, int B, int N,
                                 const float* g_lki, float* d_particles, float* d_enc, float* d_pe, float* d_head, void* workspace,
                                 const float* g_pred, const float* probs, cudaStream_t st) {
    const int grid = min(B, sm_count());
    const size_t smem = ((size_t)((NnB::ROWS * TSM + 31) & ~31) + PeTc::WBWD_FLOATS + PE_SIZE + NnB::FLOATS) * sizeof(float);
    NFDPF_CUDA(cudaFuncSetAttribute(measure_bwd_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float* part_pe = (float*)workspace;
    float* part_head = part_pe + (size_t)grid * PE_SIZE;
    measure_bwd_nn_kernel<<<grid, TP, smem, st>>>(pe, head, enc, particles, B, N, g_lki, d_particles, d_enc, part_pe, part_head, g_pred, probs);
    int rc = check_launch("measure_bwd_nn");
    if (rc) return rc;
    rc = launch_reduce_partials(part_pe, grid, PE_SIZE, d_pe, st);
    if (rc) return rc;
    return launch_reduce_partials(part_head, grid, NH_SIZE, d_head, st);
}

static size_t fwd_smem(int mode, int n_flows, int N) {
    const int n_fcnn = mode == MODE_CNF ? 4 * n_flows : 0;
    return ((size_t)PeTc::WFWD_FLOATS + PE_SIZE + 36 + (size_t)n_fcnn * CnfL1::TAIL + n_fcnn * H + (size_t)(n_fcnn / 2) * CnfL1::STAGE_FLOATS +
            (mode == MODE_NN ? NnHead::FLOATS : 0) + N) * sizeof(float);
}
static size_t bwd_smem(int mode, int n_flows) {
    const int n_fcnn = mode == MODE_CNF ? 4 * n_flows : 0;
    size_t tile = (size_t)PR::COUNT * TSM;
    if (mode == MODE_CNF && (size_t)RC::TROWS * TSM > tile) tile = (size_t)RC::TROWS * TSM;
    tile = (tile + 31) & ~(size_t)31;
    const int nw = TP / 32;
    size_t fl = tile + PeTc::WBWD_FLOATS + (size_t)PE_SIZE + 36 + (size_t)n_fcnn * LC::SIZE + n_fcnn * H + (size_t)nw * n_fcnn * 32 + 4 + nw * 32;
    return fl * sizeof(float);
}

static size_t bwd_cnf_ws_smem(int n_flows) {
    const int n_fcnn = 4 * n_flows, n_st = 2 * n_flows;
    const size_t tile = ((size_t)CnfB::ROWS * TSM + 31) & ~(size_t)31;
    size_t fl = tile + PeTc::WBWD_FLOATS + (size_t)n_st * (CnfL1::STAGE_FLOATS + CnfB::L1B_STAGE_FLOATS) + (size_t)n_fcnn * CnfL1::TAIL + n_fcnn * H +
                PE_SIZE + 36 + (size_t)4 * n_fcnn * 32 + 4 * 32;
    return fl * sizeof(float);
}

template <int MODE>
static int launch_measure_fwd(const float* pe, const float* cnf, int n_flows, float p0, float p1, const float* enc, const float* particles,
                              int B, int N, const float* lw0, const float* prior, const float* propose, float add_eps, float* lki,
                              int* argmax, float* logw_out, float* probs_out, float* row_stats, float* z_out, float* pred_out, cudaStream_t st) {
    const size_t smem = fwd_smem(MODE, n_flows, N);
    if (smem > 220 * 1024) { set_error("measure_fwd: N=%d too large for the shared-memory row buffer", N); return NFDPF_ERR_UNSUPPORTED; }
    auto kern = measure_fwd_kernel<MODE>;
    if (smem > 48 * 1024) NFDPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, TP, smem, st>>>(pe, cnf, n_flows, p0, p1, enc, particles, N, lw0, prior, propose, add_eps, lki, argmax, logw_out, probs_out,
                              row_stats, z_out, pred_out);
    return check_launch("measure_fwd");
}

// persistent grid of the backward: two resident CTAs per SM for the gaussian / cos kernels, two waves of one for CRNVP
static int measure_bwd_grid(int mode, int B) { (void)mode; return min(B, 2 * sm_count()); }

template <int MODE>
static void launch_measure_bwd_ws(const float* pe, float p0, float p1, const float* enc, const float* particles, int B, int N,
                                  const float* g_lki, const int* argmax, float* d_particles, float* d_enc, float* part_pe, int grid,
                                  size_t smem, const float* g_pred, const float* probs, cudaStream_t st) {
    auto kern = measure_bwd_ws_kernel<MODE>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, WS_THREADS, smem, st>>>(pe, p0, p1, enc, particles, B, N, g_lki, argmax, d_particles, d_enc, part_pe, g_pred, probs);
}

template <int MODE>
static int launch_measure_bwd(const float* pe, const float* cnf, int n_flows, float p0, float p1, const float* enc, const float* particles,
                              int B, int N, const float* g_lki, const int* argmax, float* d_particles, float* d_enc, float* d_pe,
                              float* d_cnf, void* workspace, const float* z_saved, const float* g_pred, const float* probs, cudaStream_t st) {
    const size_t smem = bwd_smem(MODE, n_flows);
    const int grid = measure_bwd_grid(MODE, B);
    float* part_pe = (float*)workspace;
    float* part_cnf = part_pe + (size_t)grid * PE_SIZE;
    static const bool single_role = getenv("NFDPF_MEASURE_BWD_V1") != nullptr;     // A/B timing against the round-1 kernel
    if (MODE == MODE_CNF && !single_role && z_saved && n_flows <= 2) {
        // warp-specialised CRNVP kernel: one CTA per SM (all of its tensor memory); walks the stack back from the saved flow output
        const int g1 = min(B, sm_count());
        const size_t s1 = bwd_cnf_ws_smem(n_flows);
        NFDPF_CUDA(cudaFuncSetAttribute(measure_bwd_cnf_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1));
        measure_bwd_cnf_ws_kernel<<<g1, WS_THREADS, s1, st>>>(pe, cnf, n_flows, p0, p1, enc, particles, B, N, g_lki, argmax, d_particles, d_enc,
                                                              part_pe, part_cnf, z_saved, g_pred, probs);
        int rc1 = check_launch("measure_bwd (CRNVP, warp-specialised)");
        if (rc1) return rc1;
        rc1 = launch_reduce_partials(part_pe, g1, PE_SIZE, d_pe, st);
        if (rc1) return rc1;
        return launch_reduce_partials(part_cnf, g1 * 4, 4 * n_flows * packed_fcnn_size(16, 32), d_cnf, st);
    }
    if (MODE == MODE_GAUSS && !single_role) {     // (cos: its 32 running d_enc sums per thread spill at 128 registers -- measured slower)
        const size_t ws_smem = smem;
        launch_measure_bwd_ws<MODE == MODE_CNF ? MODE_GAUSS : MODE>(pe, p0, p1, enc, particles, B, N, g_lki, argmax, d_particles, d_enc, part_pe,
                                                                     grid, ws_smem, g_pred, probs, st);
    } else {
        auto kern = measure_bwd_kernel<MODE>;
        if (smem > 48 * 1024) NFDPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, TP, smem, st>>>(pe, cnf, n_flows, p0, p1, enc, particles, B, N, g_lki, argmax, d_particles, d_enc, part_pe, part_cnf, z_saved,
                                     g_pred, probs);
    }
    int rc = check_launch("measure_bwd");
    if (rc) return rc;
    rc = launch_reduce_partials(part_pe, grid, PE_SIZE, d_pe, st);
    if (rc) return rc;
    if (MODE == MODE_CNF) rc = launch_reduce_partials(part_cnf, grid * (TP / 32), 4 * n_flows * packed_fcnn_size(16, 32), d_cnf, st);
    return rc;
}

}  // namespace nfdpf

using namespace nfdpf;

extern "C" int nfdpf_measure_fwd(int mode, const float* pe_packed, const float* cnf_packed, int n_flows, float p0, float p1,
                                 const float* enc, const float* particles, int B, int N, int hidden, const float* logw_prev,
                                 const float* prior, const float* propose, float add_eps, float* lki, int32_t* argmax,
                                 float* logw_out, float* probs_out, float* row_stats, float* z_out, float* pred_out, void* stream) {
    NFDPF_REQUIRE(pe_packed && enc && particles && lki, "measure_fwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0, "measure_fwd: B and N must be positive");
    NFDPF_REQUIRE(mode >= 0 && mode <= 3, "measure_fwd: mode must be 0 (gaussian), 1 (cos), 2 (CRNVP) or 3 (NN), got %d", mode);
    NFDPF_REQUIRE(mode != MODE_NN || cnf_packed, "measure_fwd: the NN likelihood needs the packed head (8385 floats) in cnf_packed");
    NFDPF_REQUIRE(mode != MODE_CNF || (cnf_packed && n_flows >= 1 && n_flows <= 4), "measure_fwd: CRNVP needs a packed stack, 1..4 flows");
    NFDPF_REQUIRE(!logw_prev || probs_out, "measure_fwd: fused update needs probs_out");
    NFDPF_REQUIRE(!pred_out || logw_prev, "measure_fwd: the fused prediction needs the fused weight update");
    if (hidden != HID) { set_error("measure_fwd: kernels are built for hiddensize 32 (got %d)", hidden); return NFDPF_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
#define ARGS pe_packed, cnf_packed, n_flows, p0, p1, enc, particles, B, N, logw_prev, prior, propose, add_eps, lki, argmax, logw_out, probs_out, row_stats, z_out, pred_out, st
    if (mode == MODE_GAUSS) return launch_measure_fwd<MODE_GAUSS>(ARGS);
    if (mode == MODE_COS) return launch_measure_fwd<MODE_COS>(ARGS);
    if (mode == MODE_NN) return launch_measure_fwd<MODE_NN>(ARGS);
    return launch_measure_fwd<MODE_CNF>(ARGS);
#undef ARGS
}

extern "C" int64_t nfdpf_measure_bwd_workspace(int mode, int n_flows, int B, int N) {
    (void)N;
    if (B < 1) return 0;
    if (mode == MODE_NN) return (int64_t)min(B, sm_count()) * (PE_SIZE + NH_SIZE) * (int64_t)sizeof(float);
    int64_t per = PE_SIZE + (mode == MODE_CNF ? (TP / 32) * 4 * n_flows * packed_fcnn_size(16, 32) : 0);   // CRNVP: one partial per warp
    return (int64_t)measure_bwd_grid(mode, B) * per * (int64_t)sizeof(float);
}

extern "C" int nfdpf_measure_bwd(int mode, const float* pe_packed, const float* cnf_packed, int n_flows, float p0, float p1,
                                 const float* enc, const float* particles, int B, int N, int hidden, const float* g_lki,
                                 const int32_t* argmax, float* d_particles, float* d_enc, float* d_pe, float* d_cnf, void* workspace,
                                 const float* z_saved, const float* g_pred, const float* probs, void* stream) {
    NFDPF_REQUIRE(pe_packed && enc && particles && g_lki && d_particles && d_pe && workspace, "measure_bwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0, "measure_bwd: B and N must be positive");
    NFDPF_REQUIRE(mode >= 0 && mode <= 3, "measure_bwd: bad mode %d", mode);
    NFDPF_REQUIRE(mode == MODE_COS || mode == MODE_NN || argmax, "measure_bwd: argmax required for max-shifted likelihoods");
    NFDPF_REQUIRE(mode != MODE_NN || (cnf_packed && d_cnf), "measure_bwd: the NN likelihood needs the packed head + its gradient buffer");
    NFDPF_REQUIRE(!g_pred || probs, "measure_bwd: the fused prediction gradient needs the forward's probs");
    NFDPF_REQUIRE(mode != MODE_CNF || (cnf_packed && d_cnf && n_flows >= 1 && n_flows <= 4), "measure_bwd: CRNVP needs packed stack + gradient buffer");
    if (hidden != HID) { set_error("measure_bwd: kernels are built for hiddensize 32 (got %d)", hidden); return NFDPF_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
#define ARGS pe_packed, cnf_packed, n_flows, p0, p1, enc, particles, B, N, g_lki, argmax, d_particles, d_enc, d_pe, d_cnf, workspace, z_saved, g_pred, probs, st
    if (mode == MODE_NN)
        return launch_measure_bwd_nn(pe_packed, cnf_packed, enc, particles, B, N, g_lki, d_particles, d_enc, d_pe, d_cnf, workspace, g_pred, probs, st);
    if (mode == MODE_GAUSS) return launch_measure_bwd<MODE_GAUSS>(ARGS);
    if (mode == MODE_COS) return launch_measure_bwd<MODE_COS>(ARGS);
    return launch_measure_bwd<MODE_CNF>(ARGS);
#undef ARGS
}
