// (f1) Block pseudo-likelihood of the semi-supervised objective -- reference losses.py:37-70
// (compute_block_density_nf; compute_block_density, losses.py:73-106, differs only in how the per-step prior term is formed,
// which the host mirror does elementwise before calling in here).
// At the end k of every block the reference walks each particle's ancestry back through the block with a chain of (B*N,)
// gathers (`lik.reshape(B*N)[index_a]`, `index_a = index_list[:, j].reshape(B*N)[index_a]`): ~6 ATen kernels and 5 (B,N)
// temporaries per step.  Here ONE launch walks all blocks: thread = particle, the ancestor pointer and the running sum stay in
// registers, the lists are read in place through their strides (they are transposed views of the filter's (T,B,N) buffers).
// Quirks mirrored: the running sum `logyita` is NOT reset between blocks (losses.py:47, 66); jac_list is gathered but unused.
//
// Backward: the coefficient c_k[n] = gQ/nb * sum_{block ends k' >= k} w[k'][n] enters at step k and is pushed down the ancestry,
// C_{j-1}[a'] = sum_{a : idx_j[a] = a'} C_j[a].  The resamplers' ancestor rows are sorted (soft: counts of a sorted prefix;
// OT / gate closed: identity), so the scatter is a sum over RUNS of equal keys: the head of a run adds its run in order -- fixed
// order, no atomics.  Rows that are not sorted (only hand-made index lists) take shared-memory atomics instead.
#include "common.cuh"

namespace nfdpf {

struct ListRef {          // element (b, t, n) of a (B,T,N) list lives at p[b * sb + t * st + n]
    long long sb, st;
};

__global__ void __launch_bounds__(1024)
block_density_fwd_kernel(const float* __restrict__ w, ListRef rw, const float* __restrict__ lik, ListRef rl,
                         const float* __restrict__ prior, ListRef rp, const int64_t* __restrict__ idx, ListRef ri, int B, int T,
                         int N, int block_len, float* __restrict__ Q, float* __restrict__ run_saved, int* __restrict__ bad) {
    __shared__ double s_red[33];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int nb = T / block_len;
    const long long P = (long long)B * N;
    double q = 0.0;
    int foreign = 0;
    for (int blk = 0; blk < nb; ++blk) {
        const int k = (blk + 1) * block_len - 1;
        for (int n = tid; n < N; n += nt) {
            // logyita carried over from the previous block (never reset, losses.py:47)
            float run = blk ? run_saved[((size_t)(blk - 1) * B + b) * N + n] : 0.f;
            long long a = (long long)b * N + n;                        // flat (B*N) position, as the reference indexes
            for (int j = k; j > k - block_len; --j) {
                const long long ab = a / N, an = a - ab * N;
                const float pr = prior[ab * rp.sb + j * rp.st + an], lk = lik[ab * rl.sb + j * rl.st + an];
                run = __fadd_rn(__fadd_rn(run, pr), lk);               // logyita + log_prior + lik_log, losses.py:64
                long long nx = idx[ab * ri.sb + j * ri.st + an];       // losses.py:52 / 59-60
                if (nx < 0 || nx >= P) { foreign = 2; nx = a; }
                else if (nx / N != ab) foreign |= 1;
                a = nx;
            }
            run_saved[((size_t)blk * B + b) * N + n] = run;
            q += (double)(w[(long long)b * rw.sb + k * rw.st + n] * run);   // torch.sum(w_k * logyita, -1), losses.py:65
        }
        __syncthreads();   // (nothing shared between blocks except run_saved[n], which its own thread wrote)
    }
    q = block_allreduce(q, s_red, OpSum(), 0.0);
    if (tid == 0) Q[b] = nb ? (float)(q / nb) : 0.f;                    // Q / b, losses.py:68
    if (foreign) atomicOr(bad, foreign);   // bit 0: an ancestor lives in another trajectory; bit 1: index out of range
}

// one CTA per trajectory; shared: C[N], Cn[N] (coefficients), ws[N] (suffix sums of the block-end weights), li[N] (local ancestors)
__global__ void __launch_bounds__(1024)
block_density_bwd_kernel(const float* __restrict__ gQ, const float* __restrict__ w, ListRef rw, const int64_t* __restrict__ idx,
                         ListRef ri, const float* __restrict__ run_saved, int B, int T, int N, int block_len,
                         float* __restrict__ d_w, float* __restrict__ d_lik, float* __restrict__ d_prior, ListRef ro) {
    extern __shared__ __align__(16) float sm[];
    float* C = sm;
    float* Cn = C + N;
    float* ws = Cn + N;
    int* li = reinterpret_cast<int*>(ws + N);
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int nb = T / block_len;
    const float coef = nb ? gQ[b] / (float)nb : 0.f;
    const long long ob = (long long)b * ro.sb;
    for (int n = tid; n < N; n += nt) ws[n] = 0.f;
    // steps outside every block carry no gradient; inside a block only its end step has a weight gradient
    for (int t = 0; t < T; ++t) {
        const bool covered = t < nb * block_len, end = covered && (t + 1) % block_len == 0;
        for (int n = tid; n < N; n += nt) {
            if (!end) d_w[ob + t * ro.st + n] = 0.f;
            if (!covered) { d_lik[ob + t * ro.st + n] = 0.f; d_prior[ob + t * ro.st + n] = 0.f; }
        }
    }
    for (int blk = nb - 1; blk >= 0; --blk) {
        const int k = (blk + 1) * block_len - 1;
        for (int n = tid; n < N; n += nt) {
            ws[n] += w[(long long)b * rw.sb + k * rw.st + n];
            C[n] = coef * ws[n];
            d_w[ob + k * ro.st + n] = coef * run_saved[((size_t)blk * B + b) * N + n];
        }
        __syncthreads();
        for (int j = k; j > k - block_len; --j) {
            for (int n = tid; n < N; n += nt) {
                const float c = C[n];
                d_lik[ob + j * ro.st + n] = c;
                d_prior[ob + j * ro.st + n] = c;
            }
            if (j == k - block_len + 1) break;
            int unsorted = 0;
            for (int n = tid; n < N; n += nt) {
                const long long a = idx[(long long)b * ri.sb + j * ri.st + n] - (long long)b * N;
                li[n] = a < 0 ? 0 : (a >= N ? N - 1 : (int)a);     // (foreign ancestors were reported by the forward)
                Cn[n] = 0.f;
            }
            __syncthreads();
            for (int n = tid + 1; n < N; n += nt) unsorted |= li[n] < li[n - 1];
            unsorted = __syncthreads_or(unsorted);
            if (!unsorted) {
                for (int n = tid; n < N; n += nt) {
                    const int key = li[n];
                    if (n && li[n - 1] == key) continue;            // not the head of its run
                    float s = 0.f;
                    for (int m = n; m < N && li[m] == key; ++m) s += C[m];
                    Cn[key] = s;
                }
            } else {
                for (int n = tid; n < N; n += nt) atomicAdd(&Cn[li[n]], C[n]);
            }
            __syncthreads();
            float* t_ = C; C = Cn; Cn = t_;
        }
        __syncthreads();
    }
}

}  // namespace nfdpf

using namespace nfdpf;

static int pick_nt(int N) {
    int t = 128;
    while (t < 1024 && t < N) t <<= 1;
    return t;
}

extern "C" int nfdpf_block_density_fwd(const float* w, int64_t w_sb, int64_t w_st, const float* lik, int64_t l_sb, int64_t l_st,
                                       const float* prior, int64_t p_sb, int64_t p_st, const int64_t* idx, int64_t i_sb, int64_t i_st,
                                       int B, int T, int N, int block_len, float* Q, float* run_saved, int32_t* bad, void* stream) {
    NFDPF_REQUIRE(w && lik && prior && idx && Q && run_saved && bad, "block_density_fwd: null pointer");
    NFDPF_REQUIRE(B > 0 && T > 0 && N > 0 && block_len > 0, "block_density_fwd: B, T, N, block_len must be positive");
    block_density_fwd_kernel<<<B, pick_nt(N), 0, (cudaStream_t)stream>>>(w, ListRef{w_sb, w_st}, lik, ListRef{l_sb, l_st}, prior,
                                                                          ListRef{p_sb, p_st}, idx, ListRef{i_sb, i_st}, B, T, N,
                                                                          block_len, Q, run_saved, bad);
    return check_launch("block_density_fwd");
}

extern "C" int nfdpf_block_density_bwd(const float* gQ, const float* w, int64_t w_sb, int64_t w_st, const int64_t* idx, int64_t i_sb,
                                       int64_t i_st, const float* run_saved, int B, int T, int N, int block_len, float* d_w,
                                       float* d_lik, float* d_prior, int64_t o_sb, int64_t o_st, void* stream) {
    NFDPF_REQUIRE(gQ && w && idx && run_saved && d_w && d_lik && d_prior, "block_density_bwd: null pointer");
    NFDPF_REQUIRE(B > 0 && T > 0 && N > 0 && block_len > 0, "block_density_bwd: B, T, N, block_len must be positive");
    const size_t smem = (size_t)N * 16;
    if (smem > 200 * 1024) { set_error("block_density_bwd: N=%d exceeds the shared-memory row limit (12800)", N); return NFDPF_ERR_UNSUPPORTED; }
    if (smem > 48 * 1024)
        NFDPF_CUDA(cudaFuncSetAttribute(block_density_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    block_density_bwd_kernel<<<B, pick_nt(N), smem, (cudaStream_t)stream>>>(gQ, w, ListRef{w_sb, w_st}, idx, ListRef{i_sb, i_st},
                                                                            run_saved, B, T, N, block_len, d_w, d_lik, d_prior,
                                                                            ListRef{o_sb, o_st});
    return check_launch("block_density_bwd");
}
