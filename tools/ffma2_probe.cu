// FFMA vs packed FFMA2 (fma.rn.f32x2) throughput probe: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_probe tools/ffma2_probe.cu
// Measured on the B200: 72.5 TFLOP/s for both, i.e. FFMA2 has the same FMA rate at half the issue slots.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
    unsigned long long da = *reinterpret_cast<unsigned long long*>(&a), db = *reinterpret_cast<unsigned long long*>(&b), dd = *reinterpret_cast<unsigned long long*>(&d);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(da), "l"(db));
    d = *reinterpret_cast<float2*>(&dd);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float2 acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 a = make_float2(s, s * 0.99f), b = make_float2(1.0001f, 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) { acc[i].x = fmaf(a.x, b.x, acc[i].x); acc[i].y = fmaf(a.y, b.y, acc[i].y); }
            else ffma2(acc[i], a, b);
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, 20000, 1e-6f); else k<1><<<148 * 8, 256>>>(d, 20000, 1e-6f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * 20000.0 * 32;
            printf("mode %d: %.3f ms  %.1f TFLOP/s (fp32 FMA x2)\n", mode, ms, 2 * fma / ms / 1e9);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
