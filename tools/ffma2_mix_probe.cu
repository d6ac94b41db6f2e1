// How close does a REALISTIC packed-FMA stream get to the FFMA2 pipe rate at low occupancy?  Mimics the weight-gradient / data-gradient
// block of the D = 2 coupling backward: per iteration 8 x (ld4 of a weight row from shared memory, then for two particles four
// d1 = fma2(w, bc(dj), d1) and four acc = fma2(bc(dj), h1, acc)) = 128 FFMA2 + 16 LDS.128 over 32 + 8 + 8 + ... live 64-bit registers.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_mix_probe tools/ffma2_mix_probe.cu ; run: one line per (warps per SM, variant)
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 P2(float lo, float hi) { u64 p; asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi)); return p; }
__device__ __forceinline__ void U2(u64 p, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ void acc_fma2(u64& acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
__device__ __forceinline__ u64 bc(float s) { return P2(s, s); }

// VAR 0: weights from shared memory (LDS.128 broadcast), broadcast-scalar operands (the kernel's form)
// VAR 1: same without the shared-memory loads (weights in registers)
// VAR 2: no broadcast scalars (all three operands are register pairs)
template <int VAR>
__global__ void k(float* out, int iters) {
    __shared__ __align__(16) float s_w[64];
    if (threadIdx.x < 64) s_w[threadIdx.x] = 1.0f + 1e-4f * threadIdx.x;
    __syncthreads();
    u64 A[32], d1[2][4], h1[2][4], d2[2][4];
    for (int i = 0; i < 32; ++i) A[i] = P2(threadIdx.x * 1e-3f + i, i * 0.5f);
    for (int u = 0; u < 2; ++u)
        for (int i = 0; i < 4; ++i) { d1[u][i] = P2(0.1f * i, 0.2f * u); h1[u][i] = P2(1e-3f * threadIdx.x, 0.5f + i); d2[u][i] = P2(1e-4f * (i + 1), 1e-4f * (u + 2)); }
    u64 wreg[4] = {P2(1.f, 1.0001f), P2(0.999f, 1.f), P2(1.0002f, 0.9998f), P2(1.f, 1.f)};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            u64 wr[4];
            if (VAR == 0) {
                const ulonglong2 a = reinterpret_cast<const ulonglong2*>(s_w + 8 * j)[0], b = reinterpret_cast<const ulonglong2*>(s_w + 8 * j)[1];
                wr[0] = a.x; wr[1] = a.y; wr[2] = b.x; wr[3] = b.y;
                asm volatile("" ::: "memory");
            } else {
                for (int i = 0; i < 4; ++i) wr[i] = wreg[i];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float lo, hi;
                U2(d2[u][j >> 1], lo, hi);
                const float dj = (j & 1) ? hi : lo;
                const u64 s = VAR == 2 ? d2[u][j >> 1] : bc(dj);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    d1[u][i] = fma2(wr[i], s, d1[u][i]);
                    acc_fma2(A[4 * (j & 7) + i], s, h1[u][i]);
                }
            }
        }
    }
    float r = 0;
    for (int i = 0; i < 32; ++i) { float a, b; U2(A[i], a, b); r += a + b; }
    for (int u = 0; u < 2; ++u) for (int i = 0; i < 4; ++i) { float a, b; U2(d1[u][i], a, b); r += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int VAR>
void run(float* d, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<VAR><<<148, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double ffma2 = 148.0 * threads * (double)iters * 128;           // packed instructions (thread level)
    const double cyc = best * 1e-3 * 1.965e9;                              // cycles at 1965 MHz
    const double per_sched = ffma2 / 32 / (148 * 4);                       // warp-level FFMA2 per scheduler
    printf("var %d  %2d warps/SM: %.3f ms, %.1f TFLOP/s, %.2f cycles per FFMA2 per scheduler (2.00 = pipe peak)\n", VAR, threads / 32, best,
           ffma2 * 4 / best / 1e9, cyc / per_sched);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 1024 * 4);
    for (int threads : {128, 256, 384, 512, 1024}) { run<0>(d, threads); run<1>(d, threads); run<2>(d, threads); }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
