/* libnfdpf -- C-ABI of the B200-native NF-DPF particle update (hand-written sm_100a CUDA kernels).
 *
 * The reference (xiongjiechen/Normalizing-Flows-DPFs) has no FFI: its boundary is the Python module API
 * (SURVEY.md 8b).  The host-side mirror of that API lives in normalizing_flows_dpfs_b200/*.py and calls the
 * entry points below through ctypes; every entry point names the reference code it replaces (file:line in
 * the reference repository).
 *
 * Conventions
 *   - All pointers are DEVICE pointers to contiguous row-major fp32 (int64 where stated) owned by the
 *     caller; the library never allocates caller-visible memory and never copies to the host.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous and ordered on it.
 *   - Return value: 0 on success, <0 on error (NFDPF_ERR_*); nfdpf_last_error() gives the message of the
 *     last failure on the calling thread.  There is no CPU fallback and no silent dispatch: unsupported
 *     shapes return NFDPF_ERR_UNSUPPORTED.
 *   - B = trajectories (rows), N = particles per trajectory, P = B*N, d = state dimension.
 *   - "packed stack": the parameters of a coupling stack as one flat fp32 vector in state_dict order
 *     flows.k.{t1,s1,t2,s2}.network.{0,2,4}.{weight,bias}, k = 0..n_flows-1 (nf/flows.py:181-188,
 *     nf/models.py:37-43), Linear weights row-major (out,in), in = D/2 + C, hidden = 8.
 */
#ifndef NFDPF_H
#define NFDPF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NFDPF_VERSION 100
#define NFDPF_OK 0
#define NFDPF_ERR_INVALID (-1)     /* bad argument (null pointer, non-positive size, alpha out of range) */
#define NFDPF_ERR_UNSUPPORTED (-2) /* shape outside what the kernels are built for */
#define NFDPF_ERR_CUDA (-3)        /* a CUDA runtime call or launch failed */

#define NFDPF_HIDDEN 8 /* FCNN hidden width, nf/flows.py:123,183 */

int nfdpf_version(void);
const char* nfdpf_last_error(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
int64_t nfdpf_launch_count(void);

/* ---- (K3) soft resampling: resamplers/resamplers.py:20-60 soft_resampler ---------------------------
 * particles (B,N,d), probs (B,N), offsets (B,) the caller's U(0,1/N) draws (resamplers.py:43), markers (N,)
 * the caller's linspace(0,(N-1)/N,N) (resamplers.py:42).  alpha as a double so that (float)alpha and
 * (float)(1.0-alpha) are formed exactly as the reference's Python-scalar arithmetic does.
 * Outputs: particles_out (B,N,d), probs_out (B,N), idx_out (B,N) int64 FLAT indices (j + N*b),
 * saved (B,2) = {row sum of q before normalisation, row sum of the gathered importance weights}.
 * Indices are bit-exact with the reference CPU path (ATen cascade row-sum + fp64-accumulated cumsum). */
int nfdpf_soft_resample_fwd(const float* particles, const float* probs, const float* offsets, const float* markers,
                            double alpha, int B, int N, int d, float* particles_out, float* probs_out,
                            int64_t* idx_out, float* saved, float* logprobs_out, const int32_t* gate, void* stream);
/* gate: device int32 or NULL.  The ESS decision of DPFs.py:163-170 taken on the device (nfdpf_ess_gate): when *gate == 0 the
 * call passes particles / weights through with identity ancestor indices (the `else` branch, DPFs.py:168-170) -- the filter
 * loop needs no device-to-host synchronisation.  The backward must be given the same gate.
 * logprobs_out (B,N) or NULL: log of probs_out (DPFs.py:167), saves the filter loop one elementwise pass.
 * backward: g_particles (B,N,d), g_probs (B,N), g_logprobs (B,N) (each may be NULL = zero)
 * -> d_particles (B,N,d), d_probs (B,N) */
int nfdpf_soft_resample_bwd(const float* g_particles, const float* g_probs, const float* probs, const int64_t* idx,
                            const float* saved, double alpha, int B, int N, int d, float* d_particles,
                            float* d_probs, const float* g_logprobs, const int32_t* gate, void* stream);

/* ---- (K2, weight half) log-weight update + normalisation: DPFs.py:187-192, utils.py:39-44 ----------
 * logw = logw_prev + lki + prior - propose (NULL terms are skipped); probs = softmax_N(logw) + add_eps;
 * row_stats (B,2) = {sum_n logw (for obs_likelihood, DPFs.py:191), 1/sum_n probs^2 (ESS term, DPFs.py:163)}.
 * logw_out may be NULL.  normalize_log_probs alone = logw_prev only, add_eps = 0. */
int nfdpf_weight_update_fwd(const float* logw_prev, const float* lki, const float* prior, const float* propose,
                            float add_eps, int B, int N, float* logw_out, float* probs_out, float* row_stats,
                            void* stream);
/* backward: g_probs (B,N) or NULL, g_logw (B,N) or NULL, g_rowsum (B,) or NULL (grad of row_stats[:,0]);
 * probs = the forward output (with add_eps).  d_logw (B,N) is the gradient of every added term; d_neg (B,N) or NULL receives
 * its negation (the gradient of the subtracted proposal term) in the same pass.
 * particles (B,N,2) + g_pred (B,2), both or neither: the gradient of a prediction pred[b] = sum_n probs[b,n] particles[b,n]
 * (losses.py:22) that the measurement kernel formed in its epilogue reaches probs as <g_pred[b], particles[b,n]>. */
int nfdpf_weight_update_bwd(const float* g_probs, const float* g_logw, const float* g_rowsum, const float* probs,
                            float add_eps, int B, int N, float* d_logw, float* d_neg, const float* particles,
                            const float* g_pred, void* stream);

/* ---- (K1) fused coupling stack: nf/flows.py:155-179, 215-239; nf/models.py:11-30, 45-61 ---------------
 * x (P,D); context = [row_ctx (B,C_row) broadcast over the N particles of a row | part_ctx (P,C_part)],
 * C = C_row + C_part (either may be 0).  `inverse` is a bit field: bit 0 clear = flows 0..n-1 forward, set = flows
 * n-1..0 inverse; bit 1 set = return jac = -log_det in `log_det` (model/models.py:325, 350) -- the backward must be
 * called with the same value.
 * y (P,D), log_det (P,).  Supported: D = 2 with C_part = 0 (row-constant context hoisted out of layer 1);
 * D even <= 32 with any C <= 64 (general path). */
int nfdpf_coupling_fwd(const float* packed, int n_flows, int D, int C_row, int C_part, const float* x,
                       const float* row_ctx, const float* part_ctx, int inverse, int B, int N, float* y,
                       float* log_det, void* stream);
/* backward from the OUTPUT y (coupling layers are invertible: activations are recomputed walking the stack
 * backwards, nothing else is saved).  g_y (P,D), g_ld (P,) (NULL = 0).  d_x (P,D); d_row_ctx (B,C_row) /
 * d_part_ctx (P,C_part) may be NULL (context detached: model/models.py:309-313, 338-339, 360-361).
 * d_packed: fp32 vector of the packed-stack length; every entry is WRITTEN (sums formed in a fixed order: deterministic).
 * workspace: nfdpf_coupling_bwd_workspace() bytes of device scratch. */
int64_t nfdpf_coupling_bwd_workspace(int n_flows, int D, int C_row, int C_part, int B, int N);
int nfdpf_coupling_bwd(const float* packed, int n_flows, int D, int C_row, int C_part, const float* y,
                       const float* row_ctx, const float* part_ctx, int inverse, int B, int N, const float* g_y,
                       const float* g_ld, float* d_x, float* d_row_ctx, float* d_part_ctx, float* d_packed,
                       void* workspace, void* stream);

/* Deferred reduction for D = 2 stacks with row-constant context (C_part = 0; the shapes nf_dynamic_model / normalising_flow_propose
 * run, model/models.py:305-350).  nfdpf_coupling_bwd sums its per-CTA partial parameter gradients with one small launch per call;
 * a training step makes ~150 such calls on the same parameters.  Here a call leaves its partial rows in `block`
 * (nfdpf_coupling_bwd_block_floats() floats; 0 = shape not offered) and ONE nfdpf_coupling_bwd_reduce sums the blocks of n_calls
 * calls (consecutive in memory, same n_flows / C_row / B) into d_packed -- fixed order, fp64 accumulation, every entry written. */
int64_t nfdpf_coupling_bwd_block_floats(int n_flows, int D, int C_row, int C_part, int B);
int nfdpf_coupling_bwd_deferred(const float* packed, int n_flows, int D, int C_row, int C_part, const float* y,
                                const float* row_ctx, int inverse, int B, int N, const float* g_y, const float* g_ld, float* d_x,
                                float* d_row_ctx, float* block, void* workspace, void* stream);
int nfdpf_coupling_bwd_reduce(int n_flows, int D, int C_row, int C_part, int B, const float* blocks, int n_calls,
                              float* d_packed, void* stream);

/* ---- (K2) measurement log-likelihood fused with the log-weight update and normalisation ----------------
 * mode 0: measurement_model_Gaussian, model/models.py:237-254, with MultivariateNormal(loc = p0, cov = p1^2 I)
 *         (DPFs.py:84-86 uses p0 = 1, p1 = 10);  mode 1: measurement_model_cosine_distance, models.py:206-219;
 * mode 2: measurement_model_cnf, models.py:256-278: conditional RealNVP (D = C = hidden = 32, packed stack
 *         cnf_packed, n_flows) with prior N(p0, p1^2 I) (DPFs.py:75-76: p0 = 0, p1 = 2.5).
 * mode 3: measurement_model_NN, models.py:221-235: log of the Sigmoid MLP build_likelihood (models.py:119-128:
 *         Linear(64,64)-ReLU-Linear(64,64)-ReLU-Linear(64,1)) on [enc | particle encoding]; its 8385 parameters in
 *         state_dict order arrive in cnf_packed (and their gradient in d_cnf); no row-max shift, p0 / p1 / n_flows unused.
 * pe_packed: particle encoder Linear(2,16)-ReLU-Linear(16,32)-ReLU-Linear(32,32) (models.py:130-139) in
 * state_dict order (1648 floats).  enc (B,hidden) observation encodings, particles (B,N,2).
 * lki (B,N) = log-likelihood minus its row max (modes 0,2); argmax (B,) int32 = position of that max (needed
 * by the backward; may be NULL if no backward will follow).
 * If logw_prev != NULL the weight update of nfdpf_weight_update_fwd is fused in (prior / propose may be NULL):
 * logw_out (B,N) (may be NULL), probs_out (B,N), row_stats (B,2).
 * z_out (B,N,hidden) or NULL (mode 2 only): the flow output z, which nfdpf_measure_bwd can take as z_saved to walk the
 * stack backwards without re-running it forward (128 B / particle of HBM for ~13 % fewer backward instructions).
 * pred_out (B,2) or NULL (fused update only): the prediction of the supervised loss, sum_n probs[b,n] particles[b,n,:]
 * (losses.py:22), formed in the kernel's epilogue while the row is still L2-hot.
 * Implementation note: encoder layers 2-3 run on the tcgen05 tensor cores (3xTF32); every CTA of the forward / backward
 * allocates 128 / 256 tensor-memory columns for its lifetime (sm_100a only). */
int nfdpf_measure_fwd(int mode, const float* pe_packed, const float* cnf_packed, int n_flows, float p0, float p1,
                      const float* enc, const float* particles, int B, int N, int hidden, const float* logw_prev,
                      const float* prior, const float* propose, float add_eps, float* lki, int32_t* argmax,
                      float* logw_out, float* probs_out, float* row_stats, float* z_out, float* pred_out, void* stream);
/* backward of lki w.r.t. particles (B,N,2), enc (B,hidden; may be NULL: detached), particle-encoder and cnf
 * parameters (every entry of d_pe / d_cnf is WRITTEN, sums formed in a fixed order: deterministic).  g_lki (B,N) is the total gradient reaching lki
 * (the caller adds the weight-update gradient from nfdpf_weight_update_bwd when the update was fused).
 * g_pred (B,2) + probs (B,N), both or neither: the gradient of the fused prediction reaches the particles as g_pred[b] probs[b,n]
 * and is added to d_particles in the same pass. */
int64_t nfdpf_measure_bwd_workspace(int mode, int n_flows, int B, int N);
int nfdpf_measure_bwd(int mode, const float* pe_packed, const float* cnf_packed, int n_flows, float p0, float p1,
                      const float* enc, const float* particles, int B, int N, int hidden, const float* g_lki,
                      const int32_t* argmax, float* d_particles, float* d_enc, float* d_pe, float* d_cnf,
                      void* workspace, const float* z_saved, const float* g_pred, const float* probs, void* stream);

/* ---- per-trajectory particle moments: the detached flow context of model/models.py:309-310, 338-339 ----
 * out[b, out_off + k] = mean_n x[b,n,k], out[b, out_off + d + k] = unbiased std_n x[b,n,k]; out has row stride
 * out_stride floats (so the moments can be written straight into a wider (B,C_row) context buffer). */
int nfdpf_row_moments(const float* x, int B, int N, int d, float* out, int out_stride, int out_off, void* stream);
/* The same with the context row's leading columns filled in the same launch: out[b, :head_dim] = head[b, :head_dim] (the detached
 * observation encoding of the proposal's context, model/models.py:360-361), moments at out_off = head_dim. */
int nfdpf_row_moments_head(const float* x, int B, int N, int d, const float* head, int head_dim, float* out, int out_stride,
                           void* stream);

/* ---- (K4) entropy-regularised OT resampling: resamplers/resamplers.py:62-277 ---------------------------
 * particles (B,N,2), logw (B,N) = log of the (normalised) weights.  Log-domain Sinkhorn with epsilon-scaling
 * (eps_0 = squared extent of the scaled cloud, eps <- max(eps*scaling^2, eps)), simultaneous half-step updates,
 * the reference's batch-wide stop rule (all rows must want to continue; max_iter cap) evaluated on the device,
 * then particles_out = T particles with T column-normalised to N*w_j (resamplers.py:194-210, 254-264).
 * All arithmetic is fp32 (the reference runs this part in fp64, resamplers.py:76).
 * saved (B,N,4): what the backward needs (scaled positions and the row/column log-scalings of T, log2 units).
 * iters_out: device int32 or NULL, receives the reference's `total_iter + 2`.  workspace: nfdpf_ot_workspace() bytes. */
int64_t nfdpf_ot_workspace(int B, int N);
int nfdpf_ot_resample_fwd(const float* particles, const float* logw, float eps, float scaling, float threshold,
                          int max_iter, int B, int N, int d, float* particles_out, float* saved, int32_t* iters_out,
                          void* workspace, const int32_t* gate, void* stream);
/* backward: d_particles = T^T g_out; the plan itself carries no gradient (transport.backward returns None,
 * resamplers.py:240-245) and neither do the weights. */
int nfdpf_ot_resample_bwd(const float* g_out, const float* saved, float eps, int B, int N, int d, float* d_particles,
                          const int32_t* gate, void* stream);
/* gate (device int32 or NULL) as for soft resampling: when *gate == 0 every Sinkhorn launch exits at once and particles_out =
 * particles (d_particles = g_out).  The weights after a gated OT resample (1/N when it fired, the old ones otherwise,
 * DPFs.py:166-170) come from nfdpf_gate_weights_fwd: w_out, logw_out (B,N); backward d_probs = (g_w + g_logw / probs) if the
 * gate was closed, else 0. */
int nfdpf_gate_weights_fwd(const float* probs, const int32_t* gate, int B, int N, float* w_out, float* logw_out, void* stream);
int nfdpf_gate_weights_bwd(const float* g_w, const float* g_logw, const float* probs, const int32_t* gate, int B, int N,
                           float* d_probs, void* stream);

/* ---- pipe-peak probe (measurement aid for bench.py): launches streams of independent FFMA (kind 0) or
 * ex2.approx (kind 1) instructions; returns the number of instructions issued (time it with CUDA events). */
int64_t nfdpf_peak_probe(int kind, int iters, float* out, void* stream);

/* ---- per-step glue --------------------------------------------------------------------------------------
 * motion_moments: out = (particles + vel_b) + noise (model/models.py:191-204, noise injected by the caller) and, if
 * ctx != NULL, ctx[b, ctx_off .. ctx_off+3] = [mean_x, mean_y, std_x, std_y] of out (unbiased; models.py:309-310).
 * proposal_terms: prior = dens(back - (phys - noise)) - jac_back, propose = dens(noise) + jac_dyn + jac_prop with
 * dens the 2-d N(0, sigma^2 I) log-density (utils.py:17-37; model/models.py:369-376); jac_* may be NULL (= 0).
 * backward: d_back = -g_prior (back - phys + noise)/sigma^2, d_phys = -d_back, neg_g_prior = -g_prior (may be NULL). */
int nfdpf_motion_moments(const float* particles, const float* vel, const float* noise, int B, int N, int d, float* out,
                         float* ctx, int ctx_stride, int ctx_off, void* stream);
int nfdpf_proposal_terms_fwd(const float* back, const float* phys, const float* noise, const float* jac_back,
                             const float* jac_dyn, const float* jac_prop, float sigma, int64_t P, float* prior,
                             float* propose, void* stream);
int nfdpf_proposal_terms_bwd(const float* g_prior, const float* back, const float* phys, const float* noise, float sigma,
                             int64_t P, float* d_back, float* d_phys, float* neg_g_prior, void* stream);


/* ---- host-free step control and random draws (DPFs.py:163-165, model/models.py:199-200, resamplers.py:43,
 *      utils.py:46-62) ---------------------------------------------------------------------------------------
 * rng_state: device int64[2] = {seed, step counter}.  Draws are Philox4x32-10 keyed by the seed with counter = (element index,
 * step counter, consumer tag): reproducible for a given seed, independent of the launch geometry, CUDA-graph replayable (the
 * state lives on the device).  They are NOT the reference's CPU-generator stream: parity tests inject the reference's draws.
 * ess_gate: *gate_out = force if force is 0 / 1, else (mean_b ess_inv[b * ess_stride] < N/2) with ess_inv = column 1 of the
 * (B,2) row_stats of nfdpf_weight_update_fwd / nfdpf_measure_fwd (ess_stride = 2; DPFs.py:163-165); offsets_out (B,) or NULL
 * receives U(0, 1/N) resampling offsets (resamplers.py:43); if `advance`, the step counter is incremented afterwards (call
 * once per filter timestep). */
int nfdpf_ess_gate(const float* ess_inv, int ess_stride, int B, int N, int force, int64_t* rng_state, int advance,
                   int32_t* gate_out, float* offsets_out, void* stream);
/* motion_moments with the noise drawn in-kernel: noise_out (B,N,2) ~ N(0, sigma^2), out = (particles + vel_b) + noise_out. */
int nfdpf_motion_moments_rng(const float* particles, const float* vel, const int64_t* rng_state, float sigma, int B, int N,
                             int d, float* out, float* noise_out, float* ctx, int ctx_stride, int ctx_off, void* stream);
/* particle_initialization (utils.py:46-62): uniform on [-width/2, width/2)^2, or start[b, 0:2] + N(0, 1) if true_state. */
int nfdpf_init_particles_rng(const float* start, int start_stride, const int64_t* rng_state, float width, int true_state, int B,
                             int N, int d, float* out, void* stream);

/* ---- supervised-loss prediction (losses.py:22): pred[b,:] = sum_n probs[b,n] particles[b,n,:] for one timestep; backward
 * d_particles = g_pred[b] probs, d_probs = <g_pred[b], particles> (either output may be NULL). */
int nfdpf_weighted_mean_fwd(const float* particles, const float* probs, int B, int N, int d, float* pred, void* stream);
int nfdpf_weighted_mean_bwd(const float* g_pred, const float* particles, const float* probs, int B, int N, int d,
                            float* d_particles, float* d_probs, void* stream);

/* ---- block pseudo-likelihood of the semi-supervised objective (losses.py:37-70 compute_block_density_nf; the prior term of
 * losses.py:73-106 compute_block_density is formed by the caller).  Lists are (B,T,N) with the last dimension contiguous and
 * arbitrary element strides sb (between trajectories) / st (between steps): the filter's (T,B,N) buffers are read in place.
 * idx holds the FLAT ancestor indices (j + N*b) the resamplers return.  Q (B,) = (1/nb) sum_{blocks} sum_n w[b,k,n] logyita_k[b,n],
 * nb = T / block_len, logyita running over blocks WITHOUT a reset (as the reference).  run_saved (nb,B,N): logyita per block
 * (the backward's input).  bad: device int32, OR-ed with 1 if an ancestor lies in another trajectory (the backward supports only
 * own-row ancestry, which is what every resampler produces), 2 if an index is out of range.
 * backward: gQ (B,) -> d_w, d_lik, d_prior, each (B,T,N) written completely, with element strides o_sb / o_st. */
int nfdpf_block_density_fwd(const float* w, int64_t w_sb, int64_t w_st, const float* lik, int64_t l_sb, int64_t l_st,
                            const float* prior, int64_t p_sb, int64_t p_st, const int64_t* idx, int64_t i_sb, int64_t i_st,
                            int B, int T, int N, int block_len, float* Q, float* run_saved, int32_t* bad, void* stream);
int nfdpf_block_density_bwd(const float* gQ, const float* w, int64_t w_sb, int64_t w_st, const int64_t* idx, int64_t i_sb,
                            int64_t i_st, const float* run_saved, int B, int T, int N, int block_len, float* d_w, float* d_lik,
                            float* d_prior, int64_t o_sb, int64_t o_st, void* stream);

/* ---- fan-out of one tensor to several consumers: out[i] = a[i] + b[i] (+ c[i]) (+ d[i]), n floats; b, c, d may be NULL.
 * The backward of ops.fanout: the gradients of up to four consumers of one filter tensor are summed in ONE pass (autograd's own
 * accumulation is a chain of two-operand adds, each a launch and a full read-modify-write). */
int nfdpf_sum4(const float* a, const float* b, const float* c, const float* d, int64_t n, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NFDPF_H */
