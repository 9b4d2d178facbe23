/*
 * dpsttc.h — C ABI of libdpsttc.so, the B200 (sm_100a) kernels behind the dps-ttc hot path.
 *
 * The reference (vishnutez/dps-ttc) is pure Python/PyTorch: it has NO C/FFI boundary.  Its
 * plugin surface is a set of Python registries:
 *   guided_diffusion/measurements.py:20-32        register_operator / get_operator
 *   guided_diffusion/condition_methods.py:10-21   register_conditioning_method / get_conditioning_method
 *   guided_diffusion/gaussian_diffusion.py:19-56  register_sampler / create_sampler
 *   guided_diffusion/posterior_mean_variance.py:16-27,:137-148   mean / variance processors
 * The entry points below are what a ctypes binding inside those classes calls (the binding a
 * maintainer would add is shown in INTEGRATION.md; the shipped one is dps_ttc_b200/_lib.py).
 * Each entry point cites the reference code whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer named *_dev / every tensor argument is a DEVICE pointer owned by the caller
 *     (a torch tensor's data_ptr()), fp32, contiguous in its trailing (C,H,W) dims, 16-byte aligned;
 *   - "particle stride" arguments are in ELEMENTS (floats) between consecutive particles, so a
 *     channel-slice view such as model_output[:, :3] of an (N,6,H,W) tensor needs no copy;
 *   - all launches are asynchronous on the caller's stream (cudaStream_t passed as void*), never
 *     synchronise the host, never allocate: CUDA-graph capturable.  Only *_create/_destroy allocate;
 *   - return value: 0 (DPS_OK) or a negative DPS_ERR_* code; dps_last_error() gives the message
 *     (thread-local).  No CPU fallback exists anywhere in the library.
 */
#ifndef DPSTTC_H_
#define DPSTTC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPS_OK 0
#define DPS_ERR_INVALID (-1)     /* bad argument (null pointer, bad size, bad mode)            */
#define DPS_ERR_CUDA (-2)        /* a CUDA runtime call / kernel launch failed                 */
#define DPS_ERR_UNSUPPORTED (-3) /* shape or parameter outside what the kernels were built for */
#define DPS_ERR_ALIGN (-4)       /* pointer or stride not 16-byte aligned                      */

typedef void* dps_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------------- */
const char* dps_last_error(void);
int dps_version(void);            /* 100 * major + minor                                      */
int dps_compiled_sm(void);        /* 100 : the only architecture in the fat binary (sm_100a)  */
int dps_device_sm(int* sm_out);   /* compute capability (10*major+minor) of the current device */
/* number of kernel launches issued by this library in the calling process since load / reset  */
int64_t dps_launch_count(void);
void dps_launch_count_reset(void);

/* ---- where x̂₀ comes from --------------------------------------------------------------------
 * x̂₀ = clamp(c1·x − c2·ε, −1, 1)   posterior_mean_variance.py:120-123 (predict_xstart), :40-45 (clip)
 * Operator kernels evaluate this on the fly while loading their tile, so x̂₀ never has to be
 * materialised.  eps == NULL means "x already is the image" (c1,c2,clip ignored).            */
typedef struct dps_source {
  const float* x;     /* (N, C*H*W)                                                            */
  const float* eps;   /* (N, C*H*W) view, e.g. model_output[:, :3]; may be NULL                */
  int64_t x_stride;   /* elements between particles of x                                       */
  int64_t eps_stride; /* elements between particles of eps                                     */
  float c1;           /* f32(sqrt(1/ᾱ_t))       sqrt_recip_alphas_cumprod[t]                   */
  float c2;           /* f32(sqrt(1/ᾱ_t − 1))   sqrt_recipm1_alphas_cumprod[t]                 */
  int32_t clip;       /* clip_denoised (configs/diffusion_config.yaml:7)                       */
  int32_t pad_;
} dps_source;

/* ---- per-step scalar constants (fp64 tables indexed on the host, cast to fp32 after indexing,
 *      exactly like extract_and_expand, posterior_mean_variance.py:248-252) ------------------- */
typedef struct dps_step_consts {
  float p1;         /* posterior_mean_coef1[t]           posterior_mean_variance.py:106         */
  float p2;         /* posterior_mean_coef2[t]           :107                                   */
  float max_log;    /* log(beta_t)                       :232                                   */
  float min_log;    /* posterior_log_variance_clipped[t] :231                                   */
  float ddim_sa;    /* sqrt(f32(ᾱ_prev))                 gaussian_diffusion.py:496              */
  float ddim_sb;    /* sqrt(1 − f32(ᾱ_prev) − σ²)        :497                                   */
  float ddim_sigma; /* η·sqrt((1−ᾱ_prev)/(1−ᾱ))·sqrt(1−ᾱ/ᾱ_prev)   :488-492                     */
  int32_t noise_on; /* 0 at idx == 0 ("no noise when t == 0", :473, :501)                       */
  int32_t var_mode; /* 0 learned_range (:230-242)  1 fixed: log σ² = max_log  2 learned: log σ² = v */
  int32_t mean_mode; /* 0: μ = p1·x̂₀ + p2·x (epsilon / start_x processors, :110-118, :86-92)
                        1: μ = the model output itself (previous_x processor, :62-65)             */
} dps_step_consts;

/* x̂₀ alone (for callers that need the tensor: semantic embedder, pred_xstart output).
 * Replaces EpsilonXMeanProcessor.predict_xstart + process_xstart.  The other mean processors are the
 * same expression with other scalars in dps_source, bit for bit:
 *   start_x    (:86-92)  x̂₀ = out                         c1 = 0,              c2 = −1
 *   previous_x (:57-65)  x̂₀ = out/p1 − (p2/p1)·x          c1 = −f32(p2/p1),    c2 = −f32(1/p1)
 * (out = the model's first C channels, passed as `eps`); the guidance chain rule c1·g − c2·vjp holds as is. */
int dps_x0_from_eps(const dps_source* src, float* x0, int n_particles, int64_t chw,
                    dps_stream_t stream);

/* Fused posterior update, DDPM.  One pass over each particle tensor:
 *   x̂₀   = clamp(c1·x − c2·ε)                       posterior_mean_variance.py:120-129
 *   μ     = p1·x̂₀ + p2·x                             :110-118
 *   logσ² = ((v+1)/2)·max_log + (1−(v+1)/2)·min_log   :230-242
 *   s     = μ + exp(½logσ²)·z  (if noise_on)          gaussian_diffusion.py:468-476
 *   x'    = s − (c1·g − c2·vjp)                       condition_methods.py:103 + autograd chain
 * g  : cotangent w.r.t. the pre-clamp x̂₀ (already scaled by ζ, 1/‖r‖ …), may be NULL (no guidance:
 *      x' = s, i.e. this is DDPM.p_sample);  vjp : UNet VJP of g through ε, may be NULL.
 * x_next/sample_out/x0_out: outputs, the last two may be NULL.  v may be NULL iff var_mode == 1. */
int dps_posterior_update_ddpm(const dps_source* src, const float* v, int64_t v_stride,
                              const float* z, const float* g, int64_t g_stride, const float* vjp,
                              const dps_step_consts* k, float* x_next, float* sample_out,
                              float* x0_out, int n_particles, int64_t chw, dps_stream_t stream);

/* Same for DDIM (gaussian_diffusion.py:481-509): ε' = (c1·x − x̂₀)/c2,
 *   s = x̂₀·ddim_sa + ddim_sb·ε' (+ ddim_sigma·z if noise_on and sigma != 0).                    */
int dps_posterior_update_ddim(const dps_source* src, const float* z, const float* g,
                              int64_t g_stride, const float* vjp, const dps_step_consts* k,
                              float* x_next, float* sample_out, float* x0_out, int n_particles,
                              int64_t chw, dps_stream_t stream);

/* Extended form of the two updates (round 2).  Two independent options, selected by `ext`:
 *  (1) DEFERRED GUIDANCE COEFFICIENT (ext->partials != NULL).  The per-particle factor of the guidance gradient
 *      (∇‖r‖ = −Aᵀr/‖r‖, condition_methods.py:36-39; ∇‖r‖² = −2Aᵀr, :206-212) commutes with the adjoint operator, the
 *      clamp mask and the UNet VJP, which are all linear in the cotangent.  So g and vjp may be computed from the UNSCALED
 *      cotangent g = mask ⊙ Aᵀr (dps_operator_adjoint with coef = NULL, or dps_operator_guidance) and the factor applied
 *      here:  coef_n = −scale/‖r_n‖ (DPS_COEF_NORM; 0 where ‖r_n‖ = 0) or −2·scale (DPS_COEF_NORM_SQ), with ‖r_n‖² the
 *      fixed-order fp64 sum of the P partial sums the residual kernel wrote — the arithmetic of dps_guidance_coef, which
 *      this replaces (one launch and one N-vector round trip less per step);  x' = s − coef_n·(c1·g − c2·vjp).
 *      l2_out (nullable, N floats) receives ‖r_n‖.
 *  (2) DEVICE NOISE (ext->use_philox != 0, z == NULL): z ~ N(0,1) is generated in the kernel from the counter-based
 *      Philox4x32-10 generator — counter (element/4 lo, element/4 hi, particle_offset + n, step), key (seed lo, seed hi),
 *      Box–Muller on the four outputs — so the z tensor is neither written nor read (7T → 6T for DDPM) and a sharded
 *      run draws the same noise as an unsharded one.  Throughput mode: it does not reproduce torch's RNG stream; parity
 *      runs pass z (recorded draws) instead.                                                                          */
typedef struct dps_update_ext {
  const float* partials;   /* (N, P, 2) or NULL */
  int32_t P;
  int32_t coef_mode;       /* DPS_COEF_NORM | DPS_COEF_NORM_SQ */
  float scale;             /* ζ at this step (annealing folded in) */
  int32_t use_philox;
  float* l2_out;           /* nullable */
  uint64_t philox_seed;
  uint64_t philox_step;
  int64_t particle_offset; /* global index of particle 0 of this launch */
} dps_update_ext;
int dps_posterior_update_ddpm_ext(const dps_source* src, const float* v, int64_t v_stride, const float* z,
                                  const float* g, int64_t g_stride, const float* vjp, const dps_step_consts* k,
                                  const dps_update_ext* ext, float* x_next, int n_particles, int64_t chw,
                                  dps_stream_t stream);
int dps_posterior_update_ddim_ext(const dps_source* src, const float* z, const float* g, int64_t g_stride,
                                  const float* vjp, const dps_step_consts* k, const dps_update_ext* ext,
                                  float* x_next, int n_particles, int64_t chw, dps_stream_t stream);

/* DiffStateGrad hook (gaussian_diffusion.py:240-255, diffstategrad_utils.py:46-78): on a projection step the loop
 * needs the guidance gradient as a tensor, projects it (SVD of the sample: cuSOLVER/cuBLAS on the host side) and
 * applies the projected gradient — which has batch 1 upstream — to every particle.
 *   grad   = c1·g − c2·vjp                       the chain rule inside dps_posterior_update_*, materialised
 *   x_next = sample − grad[n·grad_stride]         grad_stride = 0 broadcasts one gradient over all particles (:255) */
int dps_guidance_grad(const float* g, int64_t g_stride, const float* vjp, float c1, float c2,
                      float* grad, int n_particles, int64_t chw, dps_stream_t stream);
int dps_apply_gradient(const float* sample, const float* grad, int64_t grad_stride, float* x_next,
                       int n_particles, int64_t chw, dps_stream_t stream);

/* Per-particle partial sums of (a − ref)² and |a − ref| (P per particle, layout (n, P, 2) like the
 * operators' partials; finish with dps_particle_norms).  ref_stride 0 broadcasts one reference.
 * Replaces compute_psnr_manual's mean((real − fake)²) (compute_metrics.py:93-98) and the drivers'
 * per-path distance bookkeeping (sample_condition_batched_ttc.py:183-196).                        */
int dps_particle_sqdiff(const float* a, int64_t a_stride, const float* ref, int64_t ref_stride,
                        int n_particles, int64_t chw, float* partials, int P, dps_stream_t stream);

/* q_sample (gaussian_diffusion.py:134-151): out = a·y + b·noise                                 */
int dps_q_sample(const float* y, const float* noise, float a, float b, float* out, int64_t n_elems,
                 dps_stream_t stream);

/* ---- measurement operators -------------------------------------------------------------------
 * An operator plan owns only immutable device tables (mask, taps, band tables, twiddles).        */
typedef struct dps_operator dps_operator;

#define DPS_OP_INPAINT 1
#define DPS_OP_BLUR_SEPARABLE 2
#define DPS_OP_BLUR_SPARSE 3
#define DPS_OP_RESIZE 4
#define DPS_OP_PHASE 5

/* InpaintingOperator (measurements.py:151-168): A x = mask ⊙ x, mask (H,W) broadcast over N,C. */
int dps_operator_create_inpainting(const float* mask_host, int C, int H, int W,
                                   dps_operator** out);
/* GaussialBlurOperator / MotionBlurOperator (measurements.py:93-149) + Blurkernel
 * (util/img_utils.py:268-308): ReflectionPad2d(k/2) + depthwise cross-correlation, same (k,k)
 * kernel for every channel.  The plan trims the kernel to its non-zero support and takes the
 * separable two-pass path when the kernel is rank-1 (Gaussian), else the sparse-tap 2-D path
 * (motion).  mode: 0 auto, 1 force separable (error if not rank-1), 2 force sparse.             */
int dps_operator_create_blur(const float* kernel_host, int ksize, int C, int H, int W, int mode,
                             dps_operator** out);
/* SuperResolutionOperator (measurements.py:76-91) + Resizer (util/resizer.py:55-74): per-dim
 * gather tables exactly as Resizer builds them: fov (taps, out_len) int32 and weights
 * (taps, out_len) fp32, row-major, for H (dim 2) and W (dim 3).                                 */
int dps_operator_create_resize(const int32_t* fov_h, const float* w_h, int taps_h, int out_h,
                               const int32_t* fov_w, const float* w_w, int taps_w, int out_w,
                               int C, int H, int W, dps_operator** out);
/* PhaseRetrievalOperator (measurements.py:179-189): zero-pad `pad` each side, centred ortho 2-D
 * FFT (util/fastmri_utils.py:67-89), magnitude.  Kernels exist for pad = 64 (the reference's
 * int(oversample/8·256) with oversample = 2) and H = W ∈ {256, 128, 64}: transform lengths 384, 256,
 * 192 = 8·8·{6, 4, 3}; other shapes return DPS_ERR_UNSUPPORTED.                                    */
int dps_operator_create_phase(int pad, int C, int H, int W, dps_operator** out);
void dps_operator_destroy(dps_operator* op);

typedef struct dps_operator_info {
  int32_t kind;                   /* DPS_OP_*                                                   */
  int32_t C, H, W;                /* input particle (C,H,W)                                     */
  int32_t out_C, out_H, out_W;    /* measurement shape per particle                             */
  int32_t partials_per_particle;  /* P: residual-norm partial sums written per particle         */
  int64_t aux_floats_per_particle;/* workspace the forward pass leaves for the adjoint (phase)  */
  int32_t taps;                   /* blur: non-zero taps (sparse) or 1-D support (separable)    */
  int32_t guidance_partials;      /* > 0: dps_operator_guidance runs ONE fused kernel for this operator and writes this many
                                     partial sums per particle; 0: it runs the forward and the adjoint kernel (P partials)   */
} dps_operator_info;
int dps_operator_get_info(const dps_operator* op, dps_operator_info* info);

/* Forward, fused with the residual (condition_methods.py:36-39):
 *   out = A(x̂₀)          if y == NULL
 *   out = y − A(x̂₀)      otherwise (y_stride = 0 broadcasts one measurement over all particles)
 * partials (nullable): (N, P, 2) fp32 — per-CTA partial Σout² and Σ|out| in a fixed tree order
 * that depends only on the particle, so sharded and unsharded runs are bit-identical.
 * aux (nullable unless the adjoint needs it): (N, aux_floats_per_particle).                     */
int dps_operator_forward(const dps_operator* op, const dps_source* src, const float* y,
                         int64_t y_stride, float* out, float* partials, float* aux,
                         int n_particles, dps_stream_t stream);

/* Residual AND unscaled cotangent in one call — the whole operator part of a guided step (condition_methods.py:33-39
 * differentiated through A and the clamp):
 *   r = y − A(x̂₀)   (kept on chip where a fused kernel exists; written to r_out if given),  partials: Σr², Σ|r| pieces,
 *   g = 1[−1 ≤ pre ≤ 1] ⊙ Aᵀ r        — WITHOUT the per-particle factor −ζ/‖r‖ (or −2ζ): the caller applies it in the
 *                                         posterior update (dps_update_ext), after the UNet VJP of g.
 * Operators with guidance_partials > 0 (super-resolution ×4/×8 at 256²) run one thread-block-cluster kernel: x, ε read once,
 * g written once, r never leaves shared memory (3T + M bytes instead of 5T + 2M and three launches).  The others run
 * dps_operator_forward + dps_operator_adjoint(coef = NULL) and need r_out (and aux where the operator has one).            */
int dps_operator_guidance(const dps_operator* op, const dps_source* src, const float* y, int64_t y_stride,
                          float* r_out, float* g, int64_t g_stride, float* partials, float* aux,
                          int n_particles, dps_stream_t stream);

/* Operator.project / ortho_project (measurements.py:48-54, :90-91) with the elementwise part in the
 * operator kernels' epilogues.  `y == NULL` selects ortho_project, else project:
 *   blur, inpainting (the reference's transpose is the identity):  ortho: out = d − A d   (one forward launch, the data as its
 *       own measurement);  project: out = (y − A y) − A d   (two forward launches; `scratch` holds n_y measurement planes);
 *   super-resolution (transpose = nearest ×F): out = (d − up(A d)) + up(y)  — ×4/×8 at 256²: ONE cluster-kernel launch
 *       (A d and y never leave shared memory); other shapes: forward + a combine kernel (`scratch`: n measurement planes);
 *   phase retrieval: DPS_ERR_UNSUPPORTED (image and measurement shapes differ; the reference cannot evaluate it either).
 * data: (n, C·H·W) with data_stride; y: n_y ∈ {1, n} measurement planes with y_stride; out: out_stride (dense for blur /
 * inpainting).  Bit-identical to the reference's composition evaluated with this library's forward kernel.               */
int dps_operator_project(const dps_operator* op, const float* data, int64_t data_stride, const float* y, int64_t y_stride,
                         int n_y, float* out, int64_t out_stride, float* scratch, int n_particles, dps_stream_t stream);

/* Adjoint / Jacobian-transpose, fused with the gradient scaling and the clamp backward
 * (SURVEY.md App. A.4):  g = 1[−1 ≤ c1·x−c2·ε ≤ 1] ⊙ (coef_n · Aᵀ r + extra)
 * coef (nullable → 1): per-particle fp32;  mask_src (nullable → no clamp mask);
 * extra (nullable): additive cotangent w.r.t. x̂₀ (e.g. semantic-guidance gradient).            */
int dps_operator_adjoint(const dps_operator* op, const float* r, const float* coef,
                         const dps_source* mask_src, const float* extra, int64_t extra_stride,
                         float* g, int64_t g_stride, const float* aux, int n_particles,
                         dps_stream_t stream);

/* Finish the per-particle reductions: l2[n] = sqrt(Σ partial sq), l1[n] = Σ partial abs.
 * (torch.linalg.norm(..., dim=-1), condition_methods.py:39; ord=1 gaussian_diffusion.py:563)    */
int dps_particle_norms(const float* partials, int P, int n_particles, float* l2, float* l1,
                       dps_stream_t stream);
/* Same plus the guidance coefficient folded into the adjoint:
 *   mode 1: coef = −scale/‖r‖      (∇‖r‖,  ps / ps_semantic / mcg)
 *   mode 2: coef = −2·scale        (∇‖r‖², ps_anneal, norm_exp == 2)                            */
#define DPS_COEF_NORM 1
#define DPS_COEF_NORM_SQ 2
/*   mode 3: coef = −scale/‖r‖_F over ALL particles (the Poisson-noise branch of grad_and_value,
 *           condition_methods.py:50-55; scale = ζ·mean(1/|y|) folded by the caller)              */
#define DPS_COEF_GLOBAL_NORM 3
int dps_guidance_coef(const float* partials, int P, int n_particles, int mode, float scale,
                      float* l2, float* coef, dps_stream_t stream);

/* ---- particle reweighting / resampling (gaussian_diffusion.py:537-552, :685-698) ------------ */
/* logw_i = −tau · (meas_scale · meas_i^meas_pow + sem_scale · sem_i^sem_pow); sem may be NULL.  */
int dps_particle_logweights(const float* meas, const float* sem, int n, float tau,
                            float meas_scale, int meas_pow, float sem_scale, int sem_pow,
                            float* logw, dps_stream_t stream);
/* Weights and CDF from log-weights.
 *   linear_mode 1: w_i = exp(logw_i)            (the reference's exp(−d/100), :690)
 *   linear_mode 0: w_i = exp(logw_i − max logw) (log-sum-exp form, immune to underflow)
 * weights_out (n) fp32 = w_i / Σw;  cdf (n) fp32 = cumsum_j≤i(w_j)/Σw with cdf[n−1] = 1, summed
 * sequentially in fp32 in index order and divided in fp32 (the arithmetic of torch.multinomial's
 * CPU kernel, which the reference calls: scalar_t accumulators, fp64 uniforms);
 * lse_out (1) fp32 = log Σ exp(logw);  degenerate_out (1) int32 = 1 iff max w == min w
 * (the "rs_potentials.max() != rs_potentials.min()" guard, :545, :693) or Σw is 0/non-finite.   */
int dps_weights_cdf(const float* logw, int n, int linear_mode, float* weights_out, float* cdf,
                    float* lse_out, int32_t* degenerate_out, dps_stream_t stream);
/* Ancestor indices.  multinomial: a_i = first j with cdf[j] >= u_i (n_draws fp64 uniforms);
 * systematic: one uniform u0, positions (i + u0)/n_draws.  If *degenerate (nullable) != 0 the
 * identity a_i = i is written (no resampling).  ancestors: int64 (torch.long).                  */
int dps_ancestors_multinomial(const float* cdf, int n, const double* uniforms, int n_draws,
                              const int32_t* degenerate, int64_t* ancestors, dps_stream_t stream);
int dps_ancestors_systematic(const float* cdf, int n, const double* u0, int n_draws,
                             const int32_t* degenerate, int64_t* ancestors, dps_stream_t stream);
/* dst[i] = src[ancestors[i]] for i in [0, n_dst); elems per particle; src and dst must not alias.
 * ancestors are indices into src (particle stride = elems).                                     */
int dps_gather_particles(const float* src, const int64_t* ancestors, float* dst, int n_dst,
                         int64_t elems, dps_stream_t stream);
/* Multi-GPU form (new: the reference is single-GPU): particle a = ancestors[i] lives on rank a / n_per_rank
 * at slot a % n_per_rank of that rank's particle buffer; peer_bases_dev[r] is a device pointer, valid in
 * THIS process, to rank r's buffer (CUDA IPC / symmetric memory).  The kernel reads the needed particles
 * straight from their owners over NVLink/NVSwitch — the exchange and the gather are one kernel, and only
 * n_dst·elems floats cross the fabric instead of the W·n_dst·elems of an all-gather.  The caller provides the
 * cross-rank ordering (a barrier before: buffers written; after: buffers free to be overwritten).           */
int dps_gather_particles_p2p(const float* const* peer_bases_dev, int n_per_rank, const int64_t* ancestors,
                             float* dst, int n_dst, int64_t elems, dps_stream_t stream);
/* The same exchange with the inter-GPU rendezvous inside the kernel (no barrier launches around it).
 * signal_pads_dev[r]: device pointer, valid in THIS process, to rank r's pad of `world` uint32 words (zeroed once);
 * `epoch` = 1, 2, 3, … counts the exchanges (identical on every rank).  Rank q announces that its particle buffer of
 * this exchange is complete by writing `epoch` into word q of every rank's pad; a CTA waits only for the owner of
 * the particle it copies.  The particle buffers are double-buffered by the caller: this exchange reads
 * peer_bases_dev[r] + slot_elems.  When the kernel has completed, every rank has finished exchange epoch − 1, so the
 * other half of the buffers may be overwritten (see resample.cu).  A missing peer traps (launch failure), never hangs. */
int dps_exchange_particles_p2p(const float* const* peer_bases_dev, uint32_t* const* signal_pads_dev, int rank,
                               int world, uint32_t epoch, int64_t slot_elems, int n_per_rank,
                               const int64_t* ancestors, float* dst, int n_dst, int64_t elems,
                               dps_stream_t stream);
/* Greedy search (SearchDDPM.p_sample_loop, :630-633): best = argmin costs (first minimum).      */
int dps_argmin(const float* costs, int n, int64_t* best, float* best_cost, dps_stream_t stream);
/* dst[i] = src[*index] for all i in [0, n_dst)  (img[best_path.repeat(n_paths)], :633)          */
int dps_broadcast_particle(const float* src, const int64_t* index, float* dst, int n_dst,
                           int64_t elems, dps_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DPSTTC_H_ */
