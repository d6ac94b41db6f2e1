// Per-step elementwise glue of the filter loop, fused so that the step is a short chain of libnfdpf launches:
//   motion_moments : x' = x + vel_b + noise (model/models.py:191-204) and the [mean | std] context of x' (309-310)
//   proposal_terms : prior / proposal log-densities of proposal_likelihood (model/models.py:369-376, utils.py:17-37)
#include "common.cuh"

namespace nfdpf {

__global__ void __launch_bounds__(256)
motion_moments_kernel(const float* __restrict__ x, const float* __restrict__ vel, const float* __restrict__ noise, int N,
                      float* __restrict__ out, float* __restrict__ ctx, int ctx_stride, int ctx_off) {
    __shared__ float s_red[33];
    const int b = blockIdx.x, tid = threadIdx.x;
    const float2 v = reinterpret_cast<const float2*>(vel)[b];
    const float2* xr = reinterpret_cast<const float2*>(x) + (size_t)b * N;
    const float2* nr = reinterpret_cast<const float2*>(noise) + (size_t)b * N;
    float2* orow = reinterpret_cast<float2*>(out) + (size_t)b * N;
    float sx = 0.f, sy = 0.f;
    for (int n = tid; n < N; n += 256) {
        const float2 p = xr[n], e = nr[n];
        const float2 o = make_float2((p.x + v.x) + e.x, (p.y + v.y) + e.y);   // (particles + vel) + noise, models.py:196-202
        orow[n] = o;
        sx += o.x; sy += o.y;
    }
    if (!ctx) return;
    const float mx = block_allreduce(sx, s_red, OpSum(), 0.f) / (float)N;
    const float my = block_allreduce(sy, s_red, OpSum(), 0.f) / (float)N;
    float vx = 0.f, vy = 0.f;
    for (int n = tid; n < N; n += 256) {   // the row was just written by this same thread: re-read it (L1/L2 hit)
        const float2 o = orow[n];
        vx = fmaf(o.x - mx, o.x - mx, vx); vy = fmaf(o.y - my, o.y - my, vy);
    }
    vx = block_allreduce(vx, s_red, OpSum(), 0.f);
    vy = block_allreduce(vy, s_red, OpSum(), 0.f);
    if (tid == 0) {
        float* c = ctx + (size_t)b * ctx_stride + ctx_off;
        c[0] = mx; c[1] = my; c[2] = sqrtf(vx / (float)(N - 1)); c[3] = sqrtf(vy / (float)(N - 1));
    }
}

// dens(v) = 2 log_c - 2 log sigma - |v|^2 / (2 sigma^2)  (d = 2: the velocity terms of utils.py:30-35 vanish)
// prior   = dens(back - (phys - noise)) - jac_back   (back may be NULL: then back = prop, jac_back = 0 -- NF off)
// propose = dens(noise) + jac_dyn + jac_prop
__global__ void proposal_terms_fwd_kernel(const float* __restrict__ back, const float* __restrict__ phys, const float* __restrict__ noise,
                                          const float* __restrict__ jac_back, const float* __restrict__ jac_dyn,
                                          const float* __restrict__ jac_prop, float c0, float inv2s2, size_t P, float* __restrict__ prior,
                                          float* __restrict__ propose) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float2 bk = reinterpret_cast<const float2*>(back)[i], ph = reinterpret_cast<const float2*>(phys)[i],
                 nz = reinterpret_cast<const float2*>(noise)[i];
    const float ux = bk.x - (ph.x - nz.x), uy = bk.y - (ph.y - nz.y);
    float pr = c0 - (ux * ux + uy * uy) * inv2s2;
    if (jac_back) pr -= jac_back[i];
    prior[i] = pr;
    float pp = c0 - (nz.x * nz.x + nz.y * nz.y) * inv2s2;
    if (jac_dyn) pp += jac_dyn[i];
    if (jac_prop) pp += jac_prop[i];
    propose[i] = pp;
}

// d_back = -g_prior * u / sigma^2, d_phys = -d_back, neg_g_prior = -g_prior (gradient of jac_back)
__global__ void proposal_terms_bwd_kernel(const float* __restrict__ g_prior, const float* __restrict__ back, const float* __restrict__ phys,
                                          const float* __restrict__ noise, float inv_s2, size_t P, float* __restrict__ d_back,
                                          float* __restrict__ d_phys, float* __restrict__ neg_g_prior) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float2 bk = reinterpret_cast<const float2*>(back)[i], ph = reinterpret_cast<const float2*>(phys)[i],
                 nz = reinterpret_cast<const float2*>(noise)[i];
    const float g = g_prior[i] * inv_s2;
    const float dx = -g * (bk.x - (ph.x - nz.x)), dy = -g * (bk.y - (ph.y - nz.y));
    reinterpret_cast<float2*>(d_back)[i] = make_float2(dx, dy);
    reinterpret_cast<float2*>(d_phys)[i] = make_float2(-dx, -dy);
    if (neg_g_prior) neg_g_prior[i] = -g_prior[i];
}

}  // namespace nfdpf

using namespace nfdpf;

extern "C" int nfdpf_motion_moments(const float* particles, const float* vel, const float* noise, int B, int N, int d, float* out,
                                    float* ctx, int ctx_stride, int ctx_off, void* stream) {
    NFDPF_REQUIRE(particles && vel && noise && out, "motion_moments: null pointer");
    NFDPF_REQUIRE(B > 0 && N > 0, "motion_moments: B and N must be positive");
    NFDPF_REQUIRE(!ctx || (ctx_off >= 0 && ctx_stride >= ctx_off + 4), "motion_moments: context row too short");
    if (d != 2) { set_error("motion_moments: built for state_dim 2, got %d", d); return NFDPF_ERR_UNSUPPORTED; }
    motion_moments_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(particles, vel, noise, N, out, ctx, ctx_stride, ctx_off);
    return check_launch("motion_moments");
}

extern "C" int nfdpf_proposal_terms_fwd(const float* back, const float* phys, const float* noise, const float* jac_back, const float* jac_dyn,
                                        const float* jac_prop, float sigma, int64_t P, float* prior, float* propose, void* stream) {
    NFDPF_REQUIRE(back && phys && noise && prior && propose && P > 0 && sigma > 0.f, "proposal_terms_fwd: bad arguments");
    const float c0 = -1.8378770664093453f - 2.0f * logf(sigma);
    proposal_terms_fwd_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(back, phys, noise, jac_back, jac_dyn, jac_prop, c0,
                                                                                             0.5f / (sigma * sigma), (size_t)P, prior, propose);
    return check_launch("proposal_terms_fwd");
}

extern "C" int nfdpf_proposal_terms_bwd(const float* g_prior, const float* back, const float* phys, const float* noise, float sigma, int64_t P,
                                        float* d_back, float* d_phys, float* neg_g_prior, void* stream) {
    NFDPF_REQUIRE(g_prior && back && phys && noise && d_back && d_phys && P > 0 && sigma > 0.f, "proposal_terms_bwd: bad arguments");
    proposal_terms_bwd_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g_prior, back, phys, noise, 1.0f / (sigma * sigma),
                                                                                             (size_t)P, d_back, d_phys, neg_g_prior);
    return check_launch("proposal_terms_bwd");
}
