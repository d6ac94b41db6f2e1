// (K1) Fused conditional-RealNVP coupling stack: forward / inverse with log-det, and the backward of either.
// Replaces ~90 ATen launches per stack pass (nf/flows.py:215-239 via nf/models.py:45-61) and the (P,C) context
// materialisation of model/models.py:309-315, 338-346.  See coupling.cuh for the layout decisions.
#include "coupling.cuh"

namespace nfdpf {

constexpr int MAX_FCNN = 16;      // n_flows <= 4

// ------------------------------------------------------------------------------------------------- forward
constexpr int TPF = 256;          // forward: 8 warps per CTA (no tile, 50-80 registers): more resident warps to cover MUFU / LDS latency

template <int HALF, int CP, bool INVERSE, bool PAIRED_TANH>
__global__ void __launch_bounds__(TPF)
coupling_fwd_kernel(const float* __restrict__ packed, int n_flows, int C_row, const float* __restrict__ x,
                    const float* __restrict__ row_ctx, const float* __restrict__ part_ctx, int flags, int N, int chunk,
                    float* __restrict__ y, float* __restrict__ log_det) {
    using L = Lay<HALF, CP>;
    constexpr int D = 2 * HALF;
    extern __shared__ __align__(16) float smem[];
    const int n_fcnn = 4 * n_flows, tid = threadIdx.x;
    float* s_img = smem;                                  // [n_fcnn][L::SIZE]
    float* s_hb = s_img + n_fcnn * L::SIZE;               // [n_fcnn][8]
    float* s_w1r = s_hb + n_fcnn * H;                     // [n_fcnn][8][C_row]
    const int b = blockIdx.y;
    load_stack_images<HALF, CP>(packed, n_fcnn, C_row, s_img, s_w1r, tid, TPF);
    __syncthreads();
    hoist_row_context<HALF, CP>(s_img, s_w1r, row_ctx + (size_t)b * C_row, C_row, n_fcnn, s_hb, tid, TPF);
    __syncthreads();
    constexpr int inverse = INVERSE ? 1 : 0;  // flags bit 0 (compile-time here); bit 1: emit jac = -log_det instead of log_det
    const int n0 = blockIdx.x * chunk, n1 = min(N, n0 + chunk);
    if constexpr (HALF <= 2) {
        // narrow stacks: two particles per thread and iteration (n, n + TPF) share every weight load
        for (int n = n0 + tid; n < n1; n += 2 * TPF) {
            const bool two = n + TPF < n1;
            const size_t p[2] = {(size_t)b * N + n, (size_t)b * N + (two ? n + TPF : n)};
            float lo[2][HALF], up[2][HALF], pc[2][CP > 0 ? CP : 1], ld[2] = {0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int i = 0; i < HALF; ++i) { lo[q][i] = x[p[q] * D + i]; up[q][i] = x[p[q] * D + HALF + i]; }
#pragma unroll
                for (int i = 0; i < CP; ++i) pc[q][i] = part_ctx[p[q] * CP + i];
            }
            if (inverse) { swap_halves<HALF>(lo[0], up[0]); swap_halves<HALF>(lo[1], up[1]); }
#pragma unroll 1
            for (int st = 0; st < 2 * n_flows; ++st) {
                const int f = inverse ? n_flows - 1 - st / 2 : st / 2;
                const int pair = inverse ? 1 - (st & 1) : (st & 1);
                const float* im = s_img + (4 * f + 2 * pair) * L::SIZE;
                const float* hb = s_hb + (4 * f + 2 * pair) * H;
                stage_fwd_x2<HALF, CP, PAIRED_TANH>(im, im + L::SIZE, hb, hb + H, inverse != 0, lo, pc, up, ld);
                swap_halves<HALF>(lo[0], up[0]); swap_halves<HALF>(lo[1], up[1]);
            }
            if (inverse) { swap_halves<HALF>(lo[0], up[0]); swap_halves<HALF>(lo[1], up[1]); }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q == 1 && !two) break;
#pragma unroll
                for (int i = 0; i < HALF; ++i) { y[p[q] * D + i] = lo[q][i]; y[p[q] * D + HALF + i] = up[q][i]; }
                log_det[p[q]] = (flags & 2) ? -ld[q] : ld[q];
            }
        }
        return;
    }
    for (int n = n0 + tid; n < n1; n += TPF) {
        const size_t p = (size_t)b * N + n;
        float lo[HALF], up[HALF], pc[CP > 0 ? CP : 1], ld = 0.f;
#pragma unroll
        for (int i = 0; i < HALF; ++i) { lo[i] = x[p * D + i]; up[i] = x[p * D + HALF + i]; }
#pragma unroll
        for (int i = 0; i < CP; ++i) pc[i] = part_ctx[p * CP + i];
        // stage order: forward = flows 0..n-1, (t1,s1 | c = lower) then (t2,s2 | c = upper); inverse = the reverse walk.
        // A = conditioning half, Bv = transformed half; they swap after every stage (one inlined stage body).
        if (inverse) swap_halves<HALF>(lo, up);          // inverse starts with pair 1: c = upper
#pragma unroll 1
        for (int st = 0; st < 2 * n_flows; ++st) {
            const int f = inverse ? n_flows - 1 - st / 2 : st / 2;
            const int pair = inverse ? 1 - (st & 1) : (st & 1);
            const float* im = s_img + (4 * f + 2 * pair) * L::SIZE;
            const float* hb = s_hb + (4 * f + 2 * pair) * H;
            stage_fwd<HALF, CP, PAIRED_TANH>(im, im + L::SIZE, hb, hb + H, inverse != 0, lo, pc, up, ld);   // (c, v) = (lo, up) slots
            swap_halves<HALF>(lo, up);
        }
        // an even number of swaps leaves the slots in place for the forward walk; the inverse walk did one extra swap up front
        if (inverse) swap_halves<HALF>(lo, up);
#pragma unroll
        for (int i = 0; i < HALF; ++i) { y[p * D + i] = lo[i]; y[p * D + HALF + i] = up[i]; }
        log_det[p] = (flags & 2) ? -ld : ld;
    }
}

template <int HALF, int CP>
__global__ void __launch_bounds__(TP)
coupling_bwd_kernel(const float* __restrict__ packed, int n_flows, int C_row, const float* __restrict__ y,
                    const float* __restrict__ row_ctx, const float* __restrict__ part_ctx, int flags, int B, int N,
                    const float* __restrict__ g_y, const float* __restrict__ g_ld, float* __restrict__ d_x,
                    float* __restrict__ d_row_ctx, float* __restrict__ d_part_ctx, float* __restrict__ partials) {
    using L = Lay<HALF, CP>;
    using R = Rows<HALF, CP>;
    constexpr int D = 2 * HALF;
    extern __shared__ __align__(16) float smem[];
    const int n_fcnn = 4 * n_flows, tid = threadIdx.x, inverse = flags & 1;
    float* s_img = smem;
    float* s_hb = s_img + n_fcnn * L::SIZE;
    float* s_w1r = s_hb + n_fcnn * H;
    float* s_tile = s_w1r + (size_t)n_fcnn * H * C_row;
    constexpr int NW = TP / 32;
    const int warp = tid >> 5;
    float* s_acc = s_tile + R::TROWS * TSM;                      // [NW][n_fcnn][NOUT]: one accumulator copy per warp
    float* s_accR = s_acc + NW * n_fcnn * R::NOUT;
    float* s_d1row = s_accR + (size_t)n_fcnn * H * C_row;        // [NW][n_fcnn][8]
    float* s_ctx = s_d1row + NW * n_fcnn * H;
    const int pf = packed_fcnn_size(HALF, C_row + CP);
    for (int f = 0; f < n_fcnn; ++f)
        load_fcnn_image<HALF, CP>(packed + (size_t)f * pf, C_row, s_img + f * L::SIZE, s_w1r + (size_t)f * H * C_row, tid, TP);
    for (int e = tid; e < NW * n_fcnn * R::NOUT; e += TP) s_acc[e] = 0.f;
    for (int e = tid; e < n_fcnn * H * C_row; e += TP) s_accR[e] = 0.f;
    s_tile[R::ONE * TSM + tid] = 1.0f;
    for (int r = R::COUNT; r < R::TROWS; ++r) s_tile[r * TSM + tid] = 0.0f;   // padding rows of the dout tile + the ZERO row
    __syncthreads();

    SmemGradSink sink{s_acc + warp * n_fcnn * R::NOUT, s_d1row + warp * n_fcnn * H};
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        for (int e = tid; e < C_row; e += TP) s_ctx[e] = row_ctx[(size_t)b * C_row + e];
        for (int e = tid; e < NW * n_fcnn * H; e += TP) s_d1row[e] = 0.f;
        __syncthreads();
        hoist_row_context<HALF, CP>(s_img, s_w1r, s_ctx, C_row, n_fcnn, s_hb, tid, TP);
        __syncthreads();
        for (int n0 = 0; n0 < N; n0 += TP) {
            asm volatile("" ::: "memory");  // no LICM of shared-memory weight loads across particles
            const int n = n0 + tid;
            const bool live = n < N;
            const size_t p = (size_t)b * N + (live ? n : 0);
            float lo[HALF], up[HALF], glo[HALF], gup[HALF], pc[CP > 0 ? CP : 1], gpc[CP > 0 ? CP : 1];
#pragma unroll
            for (int i = 0; i < HALF; ++i) {
                lo[i] = y[p * D + i]; up[i] = y[p * D + HALF + i];
                glo[i] = live && g_y ? g_y[p * D + i] : 0.f;
                gup[i] = live && g_y ? g_y[p * D + HALF + i] : 0.f;
            }
            const float gld = live && g_ld ? ((flags & 2) ? -g_ld[p] : g_ld[p]) : 0.f;
#pragma unroll
            for (int i = 0; i < CP; ++i) { pc[i] = part_ctx[p * CP + i]; gpc[i] = 0.f; s_tile[(R::PC + i) * TSM + tid] = pc[i]; }
            // forward pass ran flows 0..n-1 (t1/s1 then t2/s2): walk back n-1..0 (t2/s2 then t1/s1);
            // inverse pass ran flows n-1..0 (t2/s2 then t1/s1): walk back 0..n-1 (t1/s1 then t2/s2).
            // (lo, glo) / (up, gup) are the (c, v) slots of the single stage body and swap after every stage.
            if (!inverse) { swap_halves<HALF>(lo, up); swap_halves<HALF>(glo, gup); }   // first stage has c = upper
#pragma unroll 1
            for (int st = 0; st < 2 * n_flows; ++st) {
                const int f = inverse ? st / 2 : n_flows - 1 - st / 2;
                const int pair = inverse ? (st & 1) : 1 - (st & 1);
                const float* im = s_img + (4 * f + 2 * pair) * L::SIZE;
                const float* hb = s_hb + (4 * f + 2 * pair) * H;
                stage_bwd<HALF, CP>(im, im + L::SIZE, hb, hb + H, 4 * f + 2 * pair, inverse != 0, live, lo, glo, pc, gpc, up, gup, gld,
                                    s_tile, sink);
                swap_halves<HALF>(lo, up);
                swap_halves<HALF>(glo, gup);
            }
            if (!inverse) { swap_halves<HALF>(lo, up); swap_halves<HALF>(glo, gup); }
            if (live) {
#pragma unroll
                for (int i = 0; i < HALF; ++i) { d_x[p * D + i] = glo[i]; d_x[p * D + HALF + i] = gup[i]; }
                if (CP > 0 && d_part_ctx) {
#pragma unroll
                    for (int i = 0; i < CP; ++i) d_part_ctx[p * CP + i] = gpc[i];
                }
            }
        }
        // row-context columns of W1 and the context gradient from this trajectory's layer-1 delta sums
        __syncthreads();   // every warp's delta sums of this trajectory are complete
        for (int e = tid; e < n_fcnn * H; e += TP) {   // fold the per-warp copies into copy 0 (fixed order)
            float a = s_d1row[e];
            for (int w = 1; w < NW; ++w) a += s_d1row[w * n_fcnn * H + e];
            s_d1row[e] = a;
        }
        __syncthreads();
        for (int e = tid; e < n_fcnn * H * C_row; e += TP) s_accR[e] = fmaf(s_d1row[e / C_row], s_ctx[e % C_row], s_accR[e]);
        if (d_row_ctx)
            for (int cidx = tid; cidx < C_row; cidx += TP) {
                float a = 0.f;
                for (int fk = 0; fk < n_fcnn * H; ++fk) a = fmaf(s_w1r[(size_t)fk * C_row + cidx], s_d1row[fk], a);
                d_row_ctx[(size_t)b * C_row + cidx] = a;
            }
        __syncthreads();
    }
    // per-CTA partial gradient in packed layout
    float* out = partials + (size_t)blockIdx.x * n_fcnn * pf;
    const int fin = HALF + C_row + CP;
    for (int e = tid; e < n_fcnn * R::NOUT; e += TP) {
        float a = s_acc[e];
        for (int w = 1; w < NW; ++w) a += s_acc[w * n_fcnn * R::NOUT + e];
        out[(size_t)(e / R::NOUT) * pf + packed_offset<HALF, CP>(e % R::NOUT, C_row)] = a * grad_out_scale<HALF, CP>(e % R::NOUT);
    }
    for (int e = tid; e < n_fcnn * H * C_row; e += TP) {
        const int fk = e / C_row, cidx = e % C_row;
        out[(size_t)(fk / H) * pf + (fk % H) * fin + HALF + cidx] = TANH_SCALE * s_accR[e];   // built from delta1 / scale
    }
}

// d_packed[i] = sum over CTAs (every parameter is written: the caller need not clear the buffer).  One warp per parameter: lane l sums partials l, l+32, ... in fp64, then a fixed
// butterfly -- the order depends only on (n_parts), so results are run-to-run deterministic.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_parts, int n_params, float* __restrict__ d_packed) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n_params) return;
    double a = 0.0;
    for (int c = lane; c < n_parts; c += 32) a += (double)partials[(size_t)c * n_params + i];
    a = warp_sum(a);
    if (lane == 0) d_packed[i] = (float)a;
}

int bwd_grid(int B) { return min(B, 2 * sm_count()); }

int launch_reduce_partials(const float* partials, int n_parts, int n_params, float* d_packed, cudaStream_t st) {
    reduce_partials_kernel<<<(n_params + 7) / 8, 256, 0, st>>>(partials, n_parts, n_params, d_packed);
    return check_launch("reduce_partials");
}

template <int HALF, int CP>
static int launch_fwd(const float* packed, int n_flows, int C_row, const float* x, const float* row_ctx, const float* part_ctx,
                      int inverse, int B, int N, float* y, float* log_det, cudaStream_t st) {
    using L = Lay<HALF, CP>;
    const int n_fcnn = 4 * n_flows;
    const size_t smem = ((size_t)n_fcnn * L::SIZE + n_fcnn * H + (size_t)n_fcnn * H * C_row) * sizeof(float);
    // paired tanh (3 MUFU per two activations) for the narrow two-particles-per-thread path: measured 87 -> 80 us at B = N = 1024
    constexpr bool PT = HALF <= 2;
    auto kern = (inverse & 1) ? coupling_fwd_kernel<HALF, CP, true, PT> : coupling_fwd_kernel<HALF, CP, false, PT>;
    if (smem > 48 * 1024) NFDPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // enough CTAs to fill the GPU even when B is small: split rows into chunks of >= TP particles
    int chunks = 1;
    const int target = 4 * sm_count();
    while (B * chunks < target && N / (chunks * 2) >= TPF) chunks *= 2;
    const int chunk = ((N + chunks - 1) / chunks + TPF - 1) / TPF * TPF;
    dim3 grid((N + chunk - 1) / chunk, B);
    kern<<<grid, TPF, smem, st>>>(packed, n_flows, C_row, x, row_ctx, part_ctx, inverse, N, chunk, y, log_det);
    return check_launch("coupling_fwd");
}

template <int HALF, int CP>
static int launch_bwd(const float* packed, int n_flows, int C_row, const float* y, const float* row_ctx, const float* part_ctx,
                      int inverse, int B, int N, const float* g_y, const float* g_ld, float* d_x, float* d_row_ctx,
                      float* d_part_ctx, float* d_packed, void* workspace, cudaStream_t st) {
    const int n_fcnn = 4 * n_flows;
    const size_t smem = BwdSmem<HALF, CP>::bytes(n_fcnn, C_row);
    if (smem > 220 * 1024) { set_error("coupling_bwd: stack too large for shared memory (%zu bytes)", smem); return NFDPF_ERR_UNSUPPORTED; }
    auto kern = coupling_bwd_kernel<HALF, CP>;
    if (smem > 48 * 1024) NFDPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = bwd_grid(B);
    const int n_params = n_fcnn * packed_fcnn_size(HALF, C_row + CP);
    kern<<<grid, TP, smem, st>>>(packed, n_flows, C_row, y, row_ctx, part_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_part_ctx,
                                 (float*)workspace);
    int rc = check_launch("coupling_bwd");
    if (rc) return rc;
    return launch_reduce_partials((const float*)workspace, grid, n_params, d_packed, st);
}

}  // namespace nfdpf

using namespace nfdpf;

#define NFDPF_COUPLING_SHAPES(X) X(1, 0) X(1, 4) X(1, 36) X(2, 0) X(2, 3) X(16, 0) X(16, 32)

extern "C" int nfdpf_coupling_fwd(const float* packed, int n_flows, int D, int C_row, int C_part, const float* x,
                                  const float* row_ctx, const float* part_ctx, int inverse, int B, int N, float* y, float* log_det,
                                  void* stream) {
    NFDPF_REQUIRE(packed && x && y && log_det, "coupling_fwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && D >= 2 && (D & 1) == 0, "coupling_fwd: need B,N > 0 and even D >= 2 (got %d,%d,%d)", B, N, D);
    NFDPF_REQUIRE(n_flows >= 1 && 4 * n_flows <= MAX_FCNN, "coupling_fwd: n_flows must be in 1..4 (got %d)", n_flows);
    NFDPF_REQUIRE(C_row >= 0 && C_part >= 0 && (C_row == 0 || row_ctx) && (C_part == 0 || part_ctx), "coupling_fwd: context pointers/sizes inconsistent");
    NFDPF_REQUIRE(C_row <= 64, "coupling_fwd: C_row <= 64 supported (got %d)", C_row);
#define X(HALF_, CP_) \
    if (D == 2 * HALF_ && C_part == CP_) return launch_fwd<HALF_, CP_>(packed, n_flows, C_row, x, row_ctx, part_ctx, inverse, B, N, y, log_det, (cudaStream_t)stream);
    NFDPF_COUPLING_SHAPES(X)
#undef X
    set_error("coupling_fwd: no kernel built for D=%d with %d per-particle context dims", D, C_part);
    return NFDPF_ERR_UNSUPPORTED;
}

extern "C" int64_t nfdpf_coupling_bwd_workspace(int n_flows, int D, int C_row, int C_part, int B, int N) {
    (void)N;
    if (n_flows < 1 || D < 2 || B < 1) return 0;
    if (D == 2 && C_part == 0)   // per-warp partial-gradient rows + per-CTA row-context partials (coupling_d2.cu)
        return (int64_t)coupling_bwd_d2_workspace_floats(n_flows, C_row, B) * (int64_t)sizeof(float);
    // per-CTA partial gradients
    return (int64_t)bwd_grid(B) * 4 * n_flows * packed_fcnn_size(D / 2, C_row + C_part) * (int64_t)sizeof(float);
}

// ---- deferred reduction for the D = 2 / row-context stacks (the headline shape): see include/nfdpf.h ----------------------------
extern "C" int64_t nfdpf_coupling_bwd_block_floats(int n_flows, int D, int C_row, int C_part, int B) {
    if (D != 2 || C_part != 0 || n_flows < 1 || 4 * n_flows > MAX_FCNN || B < 1 || C_row < 0 || C_row > 64) return 0;   // 0 = not offered
    return (int64_t)coupling_bwd_d2_block_floats(n_flows, C_row, B);
}
extern "C" int nfdpf_coupling_bwd_deferred(const float* packed, int n_flows, int D, int C_row, int C_part, const float* y,
                                           const float* row_ctx, int inverse, int B, int N, const float* g_y, const float* g_ld, float* d_x,
                                           float* d_row_ctx, float* block, void* workspace, void* stream) {
    NFDPF_REQUIRE(packed && y && d_x && block && workspace, "coupling_bwd_deferred: null pointer");
    NFDPF_REQUIRE(D == 2 && C_part == 0, "coupling_bwd_deferred: D = 2 stacks with row-constant context only");
    NFDPF_REQUIRE(B > 0 && N > 0 && n_flows >= 1 && 4 * n_flows <= MAX_FCNN && C_row >= 0 && C_row <= 64 && (C_row == 0 || row_ctx),
                  "coupling_bwd_deferred: bad sizes");
    return launch_coupling_bwd_d2_deferred(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, block, workspace,
                                           (cudaStream_t)stream);
}
extern "C" int nfdpf_coupling_bwd_reduce(int n_flows, int D, int C_row, int C_part, int B, const float* blocks, int n_calls,
                                         float* d_packed, void* stream) {
    NFDPF_REQUIRE(blocks && d_packed && n_calls >= 1, "coupling_bwd_reduce: bad arguments");
    NFDPF_REQUIRE(D == 2 && C_part == 0 && B > 0 && n_flows >= 1 && 4 * n_flows <= MAX_FCNN && C_row >= 0 && C_row <= 64,
                  "coupling_bwd_reduce: D = 2 stacks with row-constant context only");
    return launch_coupling_bwd_d2_reduce(n_flows, C_row, B, blocks, n_calls, d_packed, (cudaStream_t)stream);
}

extern "C" int nfdpf_coupling_bwd(const float* packed, int n_flows, int D, int C_row, int C_part, const float* y,
                                  const float* row_ctx, const float* part_ctx, int inverse, int B, int N, const float* g_y,
                                  const float* g_ld, float* d_x, float* d_row_ctx, float* d_part_ctx, float* d_packed,
                                  void* workspace, void* stream) {
    NFDPF_REQUIRE(packed && y && d_x && d_packed && workspace, "coupling_bwd: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0 && D >= 2 && (D & 1) == 0, "coupling_bwd: need B,N > 0 and even D >= 2");
    NFDPF_REQUIRE(n_flows >= 1 && 4 * n_flows <= MAX_FCNN, "coupling_bwd: n_flows must be in 1..4 (got %d)", n_flows);
    NFDPF_REQUIRE(C_row >= 0 && C_part >= 0 && (C_row == 0 || row_ctx) && (C_part == 0 || part_ctx), "coupling_bwd: context pointers/sizes inconsistent");
    NFDPF_REQUIRE(C_row <= 64, "coupling_bwd: C_row <= 64 supported (got %d)", C_row);
    if (D == 2 && C_part == 0)  // headline shape: register-accumulating kernel
        return launch_coupling_bwd_d2(packed, n_flows, C_row, y, row_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_packed, workspace,
                                      (cudaStream_t)stream);
#define X(HALF_, CP_) \
    if (D == 2 * HALF_ && C_part == CP_) return launch_bwd<HALF_, CP_>(packed, n_flows, C_row, y, row_ctx, part_ctx, inverse, B, N, g_y, g_ld, d_x, d_row_ctx, d_part_ctx, d_packed, workspace, (cudaStream_t)stream);
    NFDPF_COUPLING_SHAPES(X)
#undef X
    set_error("coupling_bwd: no kernel built for D=%d with %d per-particle context dims", D, C_part);
    return NFDPF_ERR_UNSUPPORTED;
}
