// Do an XU-bound warp and an FMA-bound warp on the SAME scheduler overlap?  256 threads per SM (two warps per scheduler): warps 0-3 run a
// MUFU stream (ex2 / rcp with the FADD2 / FFMA2 glue of the kernel's tanh), warps 4-7 a packed-FMA stream.  Times: each alone, both.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_overlap_probe tools/pipe_overlap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 P2(float lo, float hi) { u64 p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi)); return p; }
__device__ __forceinline__ void U2(u64 p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 tanh2(u64 a) {
    float x, y, ex, ey, sx, sy, rx, ry;
    U2(a, x, y);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(y));
    U2(add2(P2(ex, ey), P2(1.f, 1.f)), sx, sy);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(sx));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(sy));
    return fma2(P2(-2.f, -2.f), P2(rx, ry), P2(1.f, 1.f));
}
__global__ void __launch_bounds__(256) k(float* out, int iters, int mode) {   // mode 1: MUFU warps only, 2: FMA warps only, 3: both
    const int warp = threadIdx.x >> 5;
    float r = 0.f;
    if (warp < 4) {
        if (!(mode & 1)) return;
        u64 h[8];
        for (int i = 0; i < 8; ++i) h[i] = P2(0.01f * threadIdx.x + i, 0.3f * i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] = tanh2(h[i]);        // 32 MUFU per iteration
        }
        for (int i = 0; i < 8; ++i) { float a, b; U2(h[i], a, b); r += a + b; }
    } else {
        if (!(mode & 2)) return;
        u64 A[32];
        for (int i = 0; i < 32; ++i) A[i] = P2(threadIdx.x * 1e-3f + i, i * 0.5f);
        const u64 w = P2(1.0001f, 0.9999f), s = P2(1e-6f, 1e-6f);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) A[i] = fma2(w, s, A[i]);   // 32 FFMA2 per iteration
#pragma unroll
            for (int i = 0; i < 32; ++i) A[i] = fma2(A[i], w, s);   // 32 more (dependent on the first set, distance 32)
#pragma unroll
            for (int i = 0; i < 32; ++i) A[i] = fma2(w, A[i], s);
        }
        for (int i = 0; i < 32; ++i) { float a, b; U2(A[i], a, b); r += a + b; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 1; mode <= 3; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            k<<<148, 256>>>(d, iters, mode);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        const double cyc = best * 1e-3 * 1.965e9 / iters;
        printf("mode %d (%s): %.3f ms = %.0f cycles per iteration (MUFU warp: 32 MUFU = 256 XU cycles; FMA warp: 96 FFMA2 = 192 pipe cycles)\n", mode,
               mode == 1 ? "MUFU warps alone" : mode == 2 ? "FMA warps alone" : "both, one of each per scheduler", best, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
