// Weight-gradient contractions on the (legacy, warp-level) tensor path: dW[o][i] = sum_p delta[p][o] * act[p][i].
// K = particles is the long dimension; the deltas / activations of the CTA's batch sit transposed in a shared-memory
// tile (row = feature, column = particle).  mma.sync.m16n8k8 TF32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi) keeps
// ~2^-21 relative accuracy, which the rtol 1e-4 gradient bar needs (plain TF32 is 2^-11).  These contractions have M, N <= 48
// with K = particles: on tcgen05 they were measured slower (M is 128 whatever the live rows, the SS operands are re-read from
// shared memory every K = 8 step, and the TS form needs a thread to write other lanes' rows) -- DESIGN.md 3.  The forward /
// data-gradient products with M = particles run on tcgen05 instead (umma.cuh).
#pragma once
#include "common.cuh"

namespace nfdpf {

constexpr int TSM = 132;  // tile row stride in floats: == 4 (mod 32) makes the m16n8k8 fragment loads conflict-free

// hi = x rounded to TF32 (half-ulp add + mask), lo = the (exact, signed) remainder cut to TF32: both are valid TF32 bit
// patterns, |lo| <= 2^-11 |x|, what is dropped is < 2^-21 |x| and unbiased.  Four instructions (IADD, LOP3, FADD, LOP3);
// cvt.rna.tf32.f32 is emulated on sm_100a (VIADD + FSETP + SEL + LOP3 per conversion: nine for the pair) and the split is
// done once per fragment element, i.e. it used to be almost half of the instructions of the gradient kernels.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One warp: c[n] (16 x 8 accumulator fragments) += A(16 delta rows starting at tile row rowA) x B_n over the particles
// (tile columns) [k_begin, k_end) (multiples of 8).  rowB[n] is THIS LANE's tile row for column (lane >> 2) of n-tile n --
// callers build it so that bias / padding columns point at the ONE / ZERO rows of the tile.  Every warp contracts over
// ITS OWN 32 particles (the columns its threads staged), so the weight-gradient phase needs no CTA barrier.
// EXACT_B: bit n set = the B operand of n-tile n is exactly representable in TF32 (the ONE / ZERO bias rows): its lo part is
// zero, so the hi x lo product (a third of that tile's MMAs) is skipped.
template <int NT, unsigned EXACT_B = 0u>
__device__ __forceinline__ void mma_outer(const float* __restrict__ tile, int rowA, const int (&rowB)[NT], int k_begin, int k_end,
                                          float (&c)[NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* a_lo = tile + (rowA + g) * TSM + t;
    const float* a_hi = tile + (rowA + g + 8) * TSM + t;
#pragma unroll
    for (int k0 = k_begin; k0 < k_end; k0 += 8) {
        uint32_t ah[4], al[4], bh[NT][2], bl[NT][2];
        split_tf32(a_lo[k0], ah[0], al[0]);
        split_tf32(a_hi[k0], ah[1], al[1]);
        split_tf32(a_lo[k0 + 4], ah[2], al[2]);
        split_tf32(a_hi[k0 + 4], ah[3], al[3]);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const float* b = tile + rowB[n] * TSM + k0 + t;
            split_tf32(b[0], bh[n][0], bl[n][0]);
            split_tf32(b[4], bh[n][1], bl[n][1]);
        }
        // pass-outer: consecutive MMAs hit different accumulator fragments (the asm statements keep their order, and three
        // back-to-back MMAs into the same fragment would serialise on the tensor-pipe latency)
#pragma unroll
        for (int n = 0; n < NT; ++n) mma_tf32(c[n], al, bh[n]);
#pragma unroll
        for (int n = 0; n < NT; ++n)
            if (!((EXACT_B >> n) & 1u)) mma_tf32(c[n], ah, bl[n]);
#pragma unroll
        for (int n = 0; n < NT; ++n) mma_tf32(c[n], ah, bh[n]);
    }
}

// Two A tiles (rowA0, rowA1) against the SAME B tiles: the B fragments are loaded and split once for both (the weight-gradient
// phases of the particle encoder contract two 16-row delta tiles with the same activation rows).  Every accumulator sees exactly the
// MMA sequence mma_outer gives it, so the results are bit-identical; twice as many independent accumulators sit between two
// dependent MMAs.
template <int NT, unsigned EXACT_B = 0u>
__device__ __forceinline__ void mma_outer2(const float* __restrict__ tile, int rowA0, int rowA1, const int (&rowB)[NT], int k_begin,
                                           int k_end, float (&c0)[NT][4], float (&c1)[NT][4]) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const float* a0_lo = tile + (rowA0 + g) * TSM + t;
    const float* a0_hi = tile + (rowA0 + g + 8) * TSM + t;
    const float* a1_lo = tile + (rowA1 + g) * TSM + t;
    const float* a1_hi = tile + (rowA1 + g + 8) * TSM + t;
#pragma unroll
    for (int k0 = k_begin; k0 < k_end; k0 += 8) {
        uint32_t ah0[4], al0[4], ah1[4], al1[4], bh[NT][2], bl[NT][2];
        split_tf32(a0_lo[k0], ah0[0], al0[0]);
        split_tf32(a0_hi[k0], ah0[1], al0[1]);
        split_tf32(a0_lo[k0 + 4], ah0[2], al0[2]);
        split_tf32(a0_hi[k0 + 4], ah0[3], al0[3]);
        split_tf32(a1_lo[k0], ah1[0], al1[0]);
        split_tf32(a1_hi[k0], ah1[1], al1[1]);
        split_tf32(a1_lo[k0 + 4], ah1[2], al1[2]);
        split_tf32(a1_hi[k0 + 4], ah1[3], al1[3]);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
            const float* b = tile + rowB[n] * TSM + k0 + t;
            split_tf32(b[0], bh[n][0], bl[n][0]);
            split_tf32(b[4], bh[n][1], bl[n][1]);
        }
#pragma unroll
        for (int n = 0; n < NT; ++n) { mma_tf32(c0[n], al0, bh[n]); mma_tf32(c1[n], al1, bh[n]); }
#pragma unroll
        for (int n = 0; n < NT; ++n)
            if (!((EXACT_B >> n) & 1u)) { mma_tf32(c0[n], ah0, bl[n]); mma_tf32(c1[n], ah1, bl[n]); }
#pragma unroll
        for (int n = 0; n < NT; ++n) { mma_tf32(c0[n], ah0, bh[n]); mma_tf32(c1[n], ah1, bh[n]); }
    }
}

}  // namespace nfdpf
