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
