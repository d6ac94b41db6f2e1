// tcgen05 (5th-generation tensor core) building blocks for the measurement kernels.
//
// The particle-encoder layers (16->32, 32->32, model/models.py:130-139) and layer 1 of the CRNVP stack's t-/s-net pairs
// (48 -> 8 + 8) are the products on the path whose widths make a dense tile: M = 128 particles of a CTA batch, N = 16 / 32
// features, K = 16 / 32 / 48.  The operands are produced by the CTA's own threads (one thread = one particle), so there is no
// TMA: every thread splits its activations into a TF32 hi / lo pair and a single thread issues the 3xTF32 product
//     D = A_lo B_hi + A_hi B_lo + A_hi B_hi          (fp32 accumulate in tensor memory)
// which keeps ~1e-7 relative accuracy (plain TF32, 2^-11, would not hold the rtol-1e-4 parity bar).  The accumulator
// comes back with tcgen05.ld (lane = particle row, one register per feature).
//
// Two forms of the activation operand A:
//   TS (default): A lives in tensor memory -- the owning thread writes its row with tcgen05.st (lane = its particle,
//       K consecutive columns); only the weights are shared-memory operands.  gemm3_ts().
//   SS: A in shared memory in the same layout as the weights (below).  Kept for the CRNVP backward, whose tensor memory is
//       full of gradient fragments.  gemm3().  Measured 1.65x slower per round: every K = 8 step re-reads 128 x 32 B of A.
// Tensor memory also serves as lane-private accumulator storage for mma.sync gradient fragments (ld_frag / st_frag).
//
// Shared-memory operand layout (no swizzle, K-major; cute "INTERLEAVE" canonical form ((8,m),(4,2)):((4,SBO),(1,LBO))
// in fp32 elements): element (row r, k) lives at byte  (k / 4) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 4) * 4.
// We use SBO = 128 (8-row core matrices back to back) and LBO = rows * 16, i.e. "chunk-major": chunk c = k / 4 is a
// [rows][4] slab, row r of chunk c at byte c * rows * 16 + r * 16 -- consecutive threads write consecutive 16-byte
// words (conflict-free).  16-byte alignment suffices; MN-major TF32 operands in this layout read as zeros (tools/tc_probe.cu).
#pragma once
#include "common.cuh"

namespace nfdpf {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- tensor memory allocation (one warp allocates / frees; columns: power of two >= 32) -------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier (completion of the issued MMAs) -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// arrives on the mbarrier when every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle: start address [0,14), LBO [16,30), SBO [32,46) (all >> 4), version 1 at [46,48)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3ffff) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// instruction descriptor, kind::tf32, fp32 accumulate: c_format F32 [4,6), a/b format TF32 (2) at [7,10) / [10,13),
// a/b major (0 = K-major) at 15 / 16, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K = 8 step (tf32); issued by ONE thread
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 8 step (tf32): A rows are TMEM lanes, its K elements 8 consecutive columns
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- accumulator read-back: lane = row (the warp's 32 TMEM lanes), one register per column -----------------
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- tensor memory as lane-private accumulator storage --------------------------------------------------------
// A warp's 32 TMEM lanes x N columns hold N floats per thread: the mma.sync gradient fragments live there between batches
// instead of in per-warp shared-memory copies (tensor memory is 256 KB per SM and otherwise nearly empty on this path).
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ld4w(uint32_t taddr, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
    wait_ld();   // the destination registers are undefined until the wait: nothing may read them earlier
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld8w(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    wait_ld();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld16w(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    wait_ld();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld2w(uint32_t taddr, float* v) {
    uint32_t r[2];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
    wait_ld();
    v[0] = __uint_as_float(r[0]); v[1] = __uint_as_float(r[1]);
}
__device__ __forceinline__ void st2(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])) : "memory");
}
__device__ __forceinline__ void st4(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])),
                 "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])),
                 "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}
// N floats per thread (N a multiple of 4, compile time), starting at column address taddr
template <int N>
__device__ __forceinline__ void ld_frag(uint32_t taddr, float* v) {
    static_assert(N % 4 == 0, "fragment sizes are multiples of 4");
    constexpr int N16 = N / 16 * 16, R8 = (N - N16) >= 8 ? 8 : 0, R4 = (N - N16 - R8) >= 4 ? 4 : 0;
#pragma unroll
    for (int o = 0; o < N16; o += 16) ld16w(taddr + o, v + o);
    if (R8) ld8w(taddr + N16, v + N16);
    if (R4) ld4w(taddr + N16 + R8, v + N16 + R8);
}
template <int N>
__device__ __forceinline__ void st_frag(uint32_t taddr, const float* v) {
    static_assert(N % 4 == 0, "fragment sizes are multiples of 4");
    constexpr int N16 = N / 16 * 16, R8 = (N - N16) >= 8 ? 8 : 0, R4 = (N - N16 - R8) >= 4 ? 4 : 0;
#pragma unroll
    for (int o = 0; o < N16; o += 16) st16(taddr + o, v + o);
    if (R8) st8(taddr + N16, v + N16);
    if (R4) st4(taddr + N16 + R8, v + N16 + R8);
}

// ---- TF32 split ---------------------------------------------------------------------------------------------
// hi = x ROUNDED to TF32 (add half an ulp of the 10-bit mantissa, clear the 13 low bits: three instructions with the
// subtraction), lo = x - hi (exact, either sign, |lo| <= 2^-11 |x|; the tensor core reads its top 19 bits).  A truncating
// split (hi = x & mask) is one instruction cheaper but leaves lo with the sign of x: the dropped lo*lo products then add up
// coherently (~4e-7 relative error of a K = 32 product instead of ~1e-7), and ReLU masks that sit within that error of zero
// flip ten times more often than in an fp32 evaluation.
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// Operand tile in the chunk-major K-major layout: ROWS rows, K features (multiple of 4), hi and lo copies.
template <int ROWS, int K>
struct Operand {
    static constexpr int CHUNK_BYTES = ROWS * 16;
    static constexpr int BYTES = (K / 4) * CHUNK_BYTES;       // one copy (hi or lo)
    static constexpr int FLOATS = BYTES / 4;
    // write the K features of row r (thread-private registers) into the hi / lo copies
    __device__ static __forceinline__ void store_row(float* hi_base, float* lo_base, int r, const float (&v)[K]) {
#pragma unroll
        for (int c = 0; c < K / 4; ++c) {
            float4 h, l;
            split(v[4 * c + 0], h.x, l.x); split(v[4 * c + 1], h.y, l.y); split(v[4 * c + 2], h.z, l.z); split(v[4 * c + 3], h.w, l.w);
            *reinterpret_cast<float4*>(hi_base + c * (CHUNK_BYTES / 4) + r * 4) = h;
            *reinterpret_cast<float4*>(lo_base + c * (CHUNK_BYTES / 4) + r * 4) = l;
        }
    }
    __device__ static __forceinline__ void store_elem(float* hi_base, float* lo_base, int r, int k, float x) {
        float h, l;
        split(x, h, l);
        const int o = (k >> 2) * (CHUNK_BYTES / 4) + r * 4 + (k & 3);
        hi_base[o] = h; lo_base[o] = l;
    }
    // descriptor of the K = 8 step starting at feature k0 (multiple of 8)
    __device__ static __forceinline__ uint64_t desc(const float* base, int k0) {
        return smem_desc(smem_u32(base) + (k0 / 4) * CHUNK_BYTES, CHUNK_BYTES, 128);
    }
};

// 3xTF32 product of a [128][K] activation tile with an [N][K] weight tile into tmem_d (overwrite); ONE thread calls this.
template <int N, int K>
__device__ __forceinline__ void gemm3(uint32_t tmem_d, const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo) {
    using A = Operand<128, K>;
    using Bm = Operand<N, K>;
    constexpr uint32_t idesc = idesc_tf32(128, N);
    uint32_t acc = 0;
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) { mma_tf32_ss(tmem_d, A::desc(a_lo, k0), Bm::desc(b_hi, k0), idesc, acc); acc = 1; }
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) mma_tf32_ss(tmem_d, A::desc(a_hi, k0), Bm::desc(b_lo, k0), idesc, 1);
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) mma_tf32_ss(tmem_d, A::desc(a_hi, k0), Bm::desc(b_hi, k0), idesc, 1);
}

// Same product with the activation operand in tensor memory (TS form): A = 128 lanes x K columns, hi copy at tmem_a_hi, lo at
// tmem_a_lo (each thread wrote its own row with tcgen05.st).  Per K = 8 step the tensor core then reads only the N x 32 B weight
// slab from shared memory instead of 128 x 32 B of activations as well: measured 1.65x the round throughput of the SS form.
template <int N, int K>
__device__ __forceinline__ void gemm3_ts(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo, const float* b_hi, const float* b_lo) {
    using Bm = Operand<N, K>;
    constexpr uint32_t idesc = idesc_tf32(128, N);
    uint32_t acc = 0;
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) { mma_tf32_ts(tmem_d, tmem_a_lo + k0, Bm::desc(b_hi, k0), idesc, acc); acc = 1; }
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) mma_tf32_ts(tmem_d, tmem_a_hi + k0, Bm::desc(b_lo, k0), idesc, 1);
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += 8) mma_tf32_ts(tmem_d, tmem_a_hi + k0, Bm::desc(b_hi, k0), idesc, 1);
}

}  // namespace umma
}  // namespace nfdpf
