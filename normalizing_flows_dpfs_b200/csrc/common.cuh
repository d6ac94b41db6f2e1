// Shared device/host helpers for libnfdpf (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nfdpf.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libnfdpf is written for sm_100a (B200) only"
#endif

namespace nfdpf {

// ---- error plumbing (thread-local message, C-ABI returns negative codes) -------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> NFDPF_ERR_CUDA
int sm_count();

#define NFDPF_REQUIRE(cond, ...)              \
    do {                                      \
        if (!(cond)) {                        \
            nfdpf::set_error(__VA_ARGS__);    \
            return NFDPF_ERR_INVALID;         \
        }                                     \
    } while (0)

#define NFDPF_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            nfdpf::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
            return NFDPF_ERR_CUDA;                                                    \
        }                                                                             \
    } while (0)

// ---- warp / block primitives ----------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// Block-wide all-reduce through a 33-slot shared scratch (fixed order => deterministic).
// All threads of the block must call; result is broadcast to every thread.
template <typename T, typename Op>
__device__ __forceinline__ T block_allreduce(T v, T* scratch, Op op, T identity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(FULL, v, o));
    __syncthreads();  // protect scratch from a previous use
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T t = lane < nwarp ? scratch[lane] : identity;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = op(t, __shfl_xor_sync(FULL, t, o));
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}
// Two sums at once over a 256-thread block (eight warps), fixed order, ONE barrier; s = 16 floats that no other phase of the kernel
// reuses (so no trailing barrier).  The glue kernels that run one CTA per trajectory spend their time in barrier chains, not in
// memory traffic: block_allreduce costs three barriers and a serial second stage per scalar.
__device__ __forceinline__ void block_sum2_256(float& a, float& b, float* __restrict__ s) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(FULL, a, o); b += __shfl_xor_sync(FULL, b, o); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s[warp] = a; s[8 + warp] = b; }
    __syncthreads();
    a = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    b = ((s[8] + s[9]) + (s[10] + s[11])) + ((s[12] + s[13]) + (s[14] + s[15]));
}
struct OpSum {
    template <typename T>
    __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMax {
    __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); }
};
struct OpMin {
    __device__ __forceinline__ float operator()(float a, float b) const { return fminf(a, b); }
};

// ---- transcendental helpers ------------------------------------------------------------------------------
// tanh with ~2e-7 ABSOLUTE error from two MUFU ops (ex2 + rcp).  The parity bar is rtol 1e-4 / atol 1e-5 against fp32
// torch; tanh.approx.f32 (2^-11 relative) would not hold it.  The argument arrives PRE-SCALED, a = 2 log2(e) x (the scale
// is folded into the weights that feed the tanh):
// tanh(x) = 1 - 2 / (2^a + 1).  Four instructions (MUFU.EX2, FADD, MUFU.RCP, FFMA); 2^a = inf gives rcp = 0 -> 1, 2^a = 0 -> -1.
constexpr float TANH_SCALE = 2.885390081777927f, TANH_ISCALE = 0.34657359027997264f;
__device__ __forceinline__ float tanh_prescaled(float a) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(a));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
    return fmaf(-2.0f, r, 1.0f);
}
// Two tanh for three MUFU ops instead of four: one reciprocal of the product (t_a + 1)(t_b + 1) serves both,
// 1 / u = v r, 1 / v = u r.  The arguments are clamped at 60 (tanh is exactly 1.0f from 2^a ~ 2^25 on) so the product
// stays below 2^121.  Costs four more FP32/ALU instructions per pair: used by the SFU-bound forward kernels only.
__device__ __forceinline__ void tanh_prescaled_pair(float a, float b, float& ta, float& tb) {
    float ea, eb, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(fminf(a, 60.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(fminf(b, 60.0f)));
    const float u = ea + 1.0f, v = eb + 1.0f;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(u * v));
    ta = fmaf(-2.0f * v, r, 1.0f);
    tb = fmaf(-2.0f * u, r, 1.0f);
}
// ex2.approx(x * log2e): 2 ulp + |x| 2^-23 relative -- ~5e-7 for the |s| < 5 log-scales of the coupling stages (bar: rtol 1e-4)
// Packed FP32 pair FMA (Blackwell FFMA2): (a0, a1) += (x0, x1) * (y0, y1) in ONE issue slot -- same FMA rate as two FFMAs
// (measured, tools/ffma2_probe.cu), half the instructions.  Operands are aligned 64-bit register pairs; the packs / unpacks
// below vanish when the register allocator keeps the two floats adjacent (accumulators that live in pairs across a loop do).
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi));
    return p;
}
// (a0, a1) += s * (y0, y1)  and  (a0, a1) += (x0, x1) * (y0, y1)  (the scalar form compiles to FFMA2 with a broadcast operand)
__device__ __forceinline__ void ffma2_s(float& a0, float& a1, float s, float y0, float y1) {
    unsigned long long a = pack2(a0, a1);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pack2(s, s)), "l"(pack2(y0, y1)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
__device__ __forceinline__ void ffma2_p(float& a0, float& a1, float x0, float x1, float y0, float y1) {
    unsigned long long a = pack2(a0, a1);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(pack2(x0, x1)), "l"(pack2(y0, y1)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
// (r0, r1) = (x0, x1) * (y0, y1) + (z0, z1)
__device__ __forceinline__ void fma2_p(float& r0, float& r1, float x0, float x1, float y0, float y1, float z0, float z1) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pack2(x0, x1)), "l"(pack2(y0, y1)), "l"(pack2(z0, z1)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
__device__ __forceinline__ void fmul2_p(float& r0, float& r1, float x0, float x1, float y0, float y1) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pack2(x0, x1)), "l"(pack2(y0, y1)));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
__device__ __forceinline__ void ffma2(float& a0, float& a1, unsigned long long x, unsigned long long y) {
    unsigned long long a = pack2(a0, a1);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(x), "l"(y));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
}
__device__ __forceinline__ float exp_acc(float x) { return __expf(x); }

}  // namespace nfdpf
