// Library runtime: error string, launch accounting, device queries.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace nfdpf {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
        return NFDPF_ERR_CUDA;
    }
    return NFDPF_OK;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace nfdpf

extern "C" int nfdpf_version(void) { return NFDPF_VERSION; }
extern "C" const char* nfdpf_last_error(void) { return nfdpf::g_err; }
extern "C" int64_t nfdpf_launch_count(void) { return nfdpf::g_launches.load(); }

// ---- pipe-peak probes: the roofline denominators for the SFU / FP32-bound kernels are MEASURED on the box ------------
// kind 0: dependent-free FFMA streams (2 flop each); kind 1: ex2.approx streams (1 op each).  out keeps the result live.
namespace nfdpf {
__global__ void __launch_bounds__(256) peak_probe_kernel(int kind, int iters, float* out) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0f + 1e-3f * (threadIdx.x + k);
    const float m = 1.0000001f, c = 1e-7f;
    if (kind == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[k]));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456f) out[0] = s;
}
}  // namespace nfdpf

extern "C" int64_t nfdpf_peak_probe(int kind, int iters, float* out, void* stream) {
    const int grid = nfdpf::sm_count() * 8;
    nfdpf::peak_probe_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kind, iters, out);
    if (nfdpf::check_launch("peak_probe")) return -1;
    return (int64_t)grid * 256 * 8 * (int64_t)iters;  // operations issued (FFMA = 2 flop each)
}
