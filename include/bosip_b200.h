/* bosip_b200.h -- C ABI of libbosip_b200.so: the B200 (sm_100a) evaluator for BOSIP.jl's batched
 * GP-posterior / likelihood / acquisition / metric path.
 *
 * Every entry point is what the reference's host language (Julia, `ccall`) would bind for this path; the
 * reference interface each one replaces is cited as file:line relative to the BOSIP.jl v0.5.2 tree.
 * INTEGRATION.md shows the Julia-side `ccall` stubs.
 *
 * Conventions
 *   - All matrices are COLUMN-MAJOR Float64 exactly as Julia stores them: a d x M `Matrix{Float64}` is M
 *     contiguous d-tuples.  GP outputs are p x M (reference: test/unit/utils.jl:8-11).
 *   - Every function returns 0 (BOSIP_B200_OK) or a negative error code; the message is available from
 *     bosip_b200_last_error().  Nothing aborts the process.
 *   - Buffers are owned by the caller.  `mem` says where the *large* query/output buffers live:
 *     BOSIP_B200_HOST (plain or pinned host memory; the library stages H2D/D2H, overlapped with compute) or
 *     BOSIP_B200_DEVICE (device pointers on the context's GPU; used when inputs are already resident in HBM).
 *     Small parameter arrays (hyper-parameters, z_obs, priors ...) are always host pointers.
 *   - Calls are synchronous (they return when outputs are complete) and thread-safe per context.
 *   - There is NO CPU fallback: without a CUDA device bosip_b200_init() fails with BOSIP_B200_ECUDA.
 */
#ifndef BOSIP_B200_H
#define BOSIP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BOSIP_B200_VERSION 100

/* error codes */
#define BOSIP_B200_OK 0
#define BOSIP_B200_EINVAL (-1)   /* bad argument                                                         */
#define BOSIP_B200_ECUDA (-2)    /* CUDA runtime failure (message holds cudaGetErrorString)                  */
#define BOSIP_B200_ENOTPD (-3)   /* Cholesky met a non-positive pivot   (Julia: PosDefException)             */
#define BOSIP_B200_EDOMAIN (-4)  /* likelihood variance < -1e-8          (Julia: DomainError, src/posterior/likelihood.jl:61-65) */
#define BOSIP_B200_ENOMEM (-5)   /* device/host allocation failure                                           */

/* kernels: KernelFunctions.SqExponentialKernel / Matern32Kernel / Matern52Kernel under
 * with_lengthscale(k, lambda_j)  (examples/simple/main.jl:36, examples/interactive/toy_problem.jl:58) */
#define BOSIP_B200_KERNEL_SE 0
#define BOSIP_B200_KERNEL_MATERN32 1
#define BOSIP_B200_KERNEL_MATERN52 2

/* memory space of the large buffers */
#define BOSIP_B200_HOST 0
#define BOSIP_B200_DEVICE 1

/* `want` bitmask of bosip_b200_posterior_eval */
#define BOSIP_B200_WANT_APPROX 1u /* log_approx_posterior   src/posterior/log_posterior.jl:15-21 */
#define BOSIP_B200_WANT_MEAN 2u   /* log_posterior_mean     src/posterior/log_posterior.jl:35-41 */
#define BOSIP_B200_WANT_VAR 4u    /* log_posterior_variance src/posterior/log_posterior.jl:55-61 */

/* acquisitions: src/acquisitions/maxvar.jl:9-11, src/acquisitions/logmaxvar.jl:13-15 */
#define BOSIP_B200_ACQ_MAXVAR 0
#define BOSIP_B200_ACQ_LOGMAXVAR 1

/* univariate prior kinds of a product prior (src/types/problem.jl:33 `x_prior`; the examples use
 * Product of Uniform / Normal / truncated(Normal): examples/interactive/toy_problem.jl:43,
 * docs/src/example.md:52-55, examples/simple/toy_problem.jl:31) */
#define BOSIP_B200_PRIOR_UNIFORM 0     /* params: a, b          */
#define BOSIP_B200_PRIOR_NORMAL 1      /* params: m, s          */
#define BOSIP_B200_PRIOR_TRUNCNORMAL 2 /* params: m, s, a, b    */

/* candidate sources of bosip_b200_acq_argmax */
#define BOSIP_B200_CAND_POINTS 0 /* explicit d x M matrix (host or device)                                  */
#define BOSIP_B200_CAND_PHILOX 1 /* U(lb,ub)^d generated on device, Philox4x32-10 keyed by the global index */
#define BOSIP_B200_CAND_GRID 2   /* tensor grid: n_per_dim points per dimension, first coordinate fastest   */

typedef struct bosip_b200_ctx bosip_b200_ctx; /* one per (process, GPU) */
typedef struct bosip_b200_gp bosip_b200_gp;   /* a conditioned GP = BOSS.ModelPosterior handle */

/* Options whose BOSS-side semantics could not be verified: parameters, not constants. */
typedef struct {
  double jitter;     /* extra diagonal variance added to K (AbstractGPs FiniteGP default 1e-18); default 0      */
  int32_t pred_noise; /* 1: predictive variance includes sigma_j^2; default 0 (latent f)                          */
  int32_t clamp_var;  /* 1: negative predictive variances are clamped to 0 before sqrt; default 1                 */
} bosip_b200_gp_opts;

/* Closed-form likelihoods of the Gaussian family (one fused epilogue serves them all):
 *   NORMAL       NormalLikelihood(z_obs, std_obs)                 src/likelihoods/normal.jl:16-79
 *   LOGNORMAL    LogNormalLikelihood(log_z_obs, CV)               src/likelihoods/lognormal.jl:26-91   (z_obs = log_z_obs, std_obs = CV)
 *   NORMAL_DIFF  NormalDiffLikelihood(std_obs)                    src/likelihoods/normal_diff.jl:17-73 (z_obs ignored: 0)
 *   NORMAL_SUM   NormalSumLikelihood(sum_lengths, z_obs, std_obs) src/likelihoods/normal_sum.jl:18-115 (contiguous groups of GP outputs are summed)
 *   EXP          ExpLikelihood()                                  src/likelihoods/exp.jl:7-63          (the single GP output IS the log-likelihood)
 * `p` is the number of OBSERVATION dimensions (NORMAL_SUM: number of sums; the GP has sum(sum_lengths) outputs; EXP: 1). */
#define BOSIP_B200_LIKE_NORMAL 0
#define BOSIP_B200_LIKE_LOGNORMAL 1
#define BOSIP_B200_LIKE_NORMAL_DIFF 2
#define BOSIP_B200_LIKE_NORMAL_SUM 3
#define BOSIP_B200_LIKE_EXP 4
typedef struct {
  int32_t p;
  int32_t kind;               /* BOSIP_B200_LIKE_* (0 = NormalLikelihood) */
  const double* z_obs;        /* p   (LOGNORMAL: log_z_obs; NORMAL_DIFF, EXP: may be NULL) */
  const double* std_obs;      /* p, > 0 (LOGNORMAL: CV; EXP: may be NULL) */
  const int32_t* sum_lengths; /* NORMAL_SUM: p group sizes; otherwise NULL */
} bosip_b200_like_normal;
typedef bosip_b200_like_normal bosip_b200_like;

/* Product prior descriptor, or host-supplied log-prior values. */
typedef struct {
  int32_t d;
  const int32_t* kind;    /* d entries BOSIP_B200_PRIOR_*; NULL => no prior term (log prior = 0)             */
  const double* params;   /* d x 4 row-major: per dimension (p0,p1,p2,p3) as listed above                   */
  const double* logprior; /* optional: M precomputed log-prior values in the same memory space as Xq; when
                             non-NULL it replaces the descriptor (any MultivariateDistribution on the Julia side) */
} bosip_b200_prior;

typedef struct {
  int32_t kind;          /* BOSIP_B200_CAND_*                                                              */
  int32_t mem;           /* POINTS: memory space of `points`                                               */
  const double* points;  /* POINTS: d x M                                                                  */
  uint64_t seed;         /* PHILOX                                                                         */
  const double* lb;      /* PHILOX / GRID: d lower bounds                                                  */
  const double* ub;      /* PHILOX / GRID: d upper bounds                                                  */
  int64_t n_per_dim;     /* GRID                                                                           */
} bosip_b200_cand_src;

/* ---- lifecycle ------------------------------------------------------------------------------------- */
int bosip_b200_version(void);
/* Create a context on CUDA device `device`.  Fails (ECUDA) when no such device exists: there is no CPU path. */
int bosip_b200_init(int device, bosip_b200_ctx** out);
int bosip_b200_shutdown(bosip_b200_ctx* ctx);
/* Last error message of this context (ctx == NULL: of the last failed bosip_b200_init on this thread). */
const char* bosip_b200_last_error(const bosip_b200_ctx* ctx);
/* Launch the compute kernels on a caller-owned CUDA stream (a `cudaStream_t` passed as void*, e.g. the host framework's
 * current stream) so that the caller's events bracket them; NULL restores the context's own stream.  The H2D/D2H
 * staging streams stay internal.  Every entry point still returns only when its outputs are complete. */
int bosip_b200_set_stream(bosip_b200_ctx* ctx, void* cuda_stream);
/* Order the context's compute stream after everything already enqueued on `producer_stream` (a `cudaStream_t` passed as
 * void*; NULL = the legacy default stream).  Callers that hand over DEVICE buffers written (or freed and recycled) by work on
 * another stream call this first: the context's own streams are non-blocking, so nothing orders them implicitly.  HOST buffers
 * need no ordering. */
int bosip_b200_stream_wait(bosip_b200_ctx* ctx, void* producer_stream);
/* Number of kernels this context has launched since creation (bench.py: gpu_launches). */
int bosip_b200_launch_count(bosip_b200_ctx* ctx, int64_t* out);
/* Upper bound for the number of query points processed per internal chunk (0 = library default). */
int bosip_b200_set_chunk(bosip_b200_ctx* ctx, int64_t max_points_per_chunk);
/* Which kernel evaluates the variance contraction ||alpha^2 L^-1 k*||^2 of BOSS.var / mean_and_var
 * (src/likelihoods/normal.jl:37,43,60,68):
 *   BOSIP_B200_VAR_FP64  mma.sync FP64 tensor cores (DMMA), plain Float64 arithmetic;
 *   BOSIP_B200_VAR_INT8  tcgen05 INT8 tensor cores on error-free byte slices of both operands (Ozaki splitting):
 *                        `slices` bytes of fixed point per operand (6, 7 or 8: S (S + 1) / 2 exact INT8 GEMMs), or 0 = the
 *                        default: 7 bytes of alpha^2 L^-1 (2^-56 of the row scale) against 6 bytes of the cross-kernel
 *                        values (2^-48), 27 INT8 GEMMs -- the 28th slice pair of the 7 x 7 scheme only meets the most
 *                        significant byte of L^-1; integer products exact, rounding only in the epilogue.
 * Both satisfy the 1e-10 parity bound; results differ in the last bits.  The INT8 engine needs N < 8192 (exact integer
 * accumulation); BOSIP_B200_VAR_AUTO falls back to FP64 above that, an explicit INT8 request returns BOSIP_B200_EINVAL. */
#define BOSIP_B200_VAR_FP64 0
#define BOSIP_B200_VAR_INT8 1
#define BOSIP_B200_VAR_AUTO 2 /* default: INT8 for N >= 128 training points, FP64 below */
int bosip_b200_set_variance_engine(bosip_b200_ctx* ctx, int engine, int slices);
/* The current setting: engine (BOSIP_B200_VAR_*), byte slices of alpha^2 L^-1 and of the cross-kernel values (any pointer may be NULL). */
int bosip_b200_get_variance_engine(bosip_b200_ctx* ctx, int* engine, int* slices, int* kstar_slices);
/* How many threads issue tcgen05.mma in the INT8 engine of this context: 2 (default) only after a one-off self-check has
 * shown, on the first real chunk, that the two-issuer kernel reproduces the single-issuer kernel bit for bit (PTX orders
 * tcgen05.mma per issuing thread only); 1 after a mismatch or with BOSIP_B200_I8_ISSUERS=1; 0 = not decided yet (no INT8 launch
 * so far). */
int bosip_b200_i8_issuers(bosip_b200_ctx* ctx, int* out);
/* State of the paired (tcgen05 cta_group::2, two SMs per MMA stream) INT8 kernel of this context: 1 = in use (opt-in with
 * BOSIP_B200_I8_PAIR=1, slices = 7, after the same bitwise self-check against the single-issuer kernel), -1 = self-check failed,
 * 0 = not requested / not decided yet.  Same results bit for bit; measured slower than the unpaired kernel on B200. */
int bosip_b200_i8_pair(bosip_b200_ctx* ctx, int* out);
/* ---- AMIS importance sampler, inner loop (src/samplers/amis/amis_method.jl:26-83,126-134;
 *      src/samplers/amis/proposal_distributions/normal.jl:75-100; src/sampling.jl:55-57).  The samples, log_P, Delta and
 *      log_Omega can stay in HBM between iterations (mem = BOSIP_B200_DEVICE); log_pi(xs) is bosip_b200_posterior_eval. ---- */
/* `rand(q, M)` for q = MvNormal(mu, L L'): x_i = mu + L z_i, z from Philox4x32-10 keyed by (seed, offset + i) + Box-Muller.
 * mu (d) and L (d x d column-major, lower triangle read) are host pointers; out is d x M in `mem`. */
int bosip_b200_mvnormal_sample(bosip_b200_ctx* ctx, int d, const double* mu, const double* L, uint64_t seed, int64_t offset,
                               int64_t M, int mem, double* out);
/* delta[m] (+)= sum_q pdf(MvNormal(mu_q, Sigma_q), x_m)  (amis_method.jl:52,66-68,74).  Host parameter arrays: mus d x nq,
 * Linvs nq inverse Cholesky factors (d x d column-major, lower), lognorm[q] = -sum_i log L_q[i,i] - d/2 log(2 pi).
 * accumulate = 0 overwrites delta. */
int bosip_b200_mixture_pdf_sum(bosip_b200_ctx* ctx, int d, int nq, const double* mus, const double* Linvs, const double* lognorm,
                               const double* X, int64_t M, int mem, int accumulate, double* delta);
/* log_w[m] = log_p[m] - log(delta[m] / t)  (amis_method.jl:53,69,75) */
int bosip_b200_amis_log_weights(bosip_b200_ctx* ctx, const double* log_p, const double* delta, double t, int64_t M, int mem,
                                double* log_w);
/* exp_weights (amis_method.jl:126-134) + AnalyticalFitter / estimate_parameters! (normal.jl:75-100): w = exp(log_w - max),
 * mean = sum w x / V1, cov = sum w (x-mean)(x-mean)' / (V1 - V2/V1), V1 = sum w, V2 = sum w^2.  mean (d), cov (d x d) and
 * sums (2: V1, V2; optional) are host outputs; ws (optional, M, in `mem`) receives the normalised weights w / V1.
 * NaN or +Inf log-weights, or all weights zero, return BOSIP_B200_EDOMAIN ("AMIS: Numerical issues with sample weights."). */
int bosip_b200_weighted_moments(bosip_b200_ctx* ctx, int d, const double* X, const double* log_w, int64_t M, int mem, double* mean,
                                double* cov, double* sums, double* ws);
/* resample(xs, ws, count) (src/sampling.jl:55-57): `count` draws with replacement, P(i) = ws[i] / sum ws; draw j uses the
 * Philox uniform keyed by (seed, j).  out is d x count in `mem`; idx (optional, count, in `mem`) receives the chosen columns. */
int bosip_b200_resample(bosip_b200_ctx* ctx, int d, const double* X, const double* ws, int64_t M, uint64_t seed, int64_t count, int mem,
                        double* out, int64_t* idx);

/* Measured issue-rate ceiling of tcgen05.mma kind::i8 (M=128, N=256, operands resident in shared memory) on all SMs,
 * in TOPS counting 2 ops per multiply-accumulate: the roofline denominator of the INT8 engine.  burst = best of five 2.7 ms
 * launches; sustained (optional, may be NULL; takes ~2 s) = the rate the part holds under its power cap. */
int bosip_b200_int8_peak(bosip_b200_ctx* ctx, double* burst_tops, double* sustained_tops);
/* The same sustained loop (~2 s) with pseudo-random operand bytes instead of near-zero ones: the issue rate the part holds under
 * its power cap when the INT8 multipliers see operands like the engine's byte slices.  Context for the roofline fraction: the
 * switching energy of the multipliers depends on the operand bits, and in long runs it, not the schedule, sets the clock. */
int bosip_b200_int8_peak_random(bosip_b200_ctx* ctx, double* sustained_tops);

/* ---- GP conditioning: replaces BOSS.model_posterior(bosip.problem)
 *      (src/posterior/log_posterior.jl:81,115,148,182,214) --------------------------------------------- */
/* X d x N, Y p x N, lambda d x p, alpha p, sigma p (src/subset_problem.jl:38-50); mean_at_X p x N or NULL.
 * Builds K_j = alpha_j^2 k(X/lambda_j) + (sigma_j^2 + jitter) I, factorises it with the hand-written blocked
 * FP64 Cholesky, forms L_j^{-1} and a_j = K_j^{-1}(Y_j - m_j(X)) on the device. */
int bosip_b200_gp_fit(bosip_b200_ctx* ctx, int kernel_id, int d, int N, int p, const double* X, const double* Y,
                      const double* lambda, const double* alpha, const double* sigma, const double* mean_at_X,
                      const bosip_b200_gp_opts* opts, bosip_b200_gp** out);
/* Same, but the factorisation comes from the host (exact-parity runs): L is p matrices N x N column-major
 * (lower triangle read), a is N x p. */
int bosip_b200_gp_load_factors(bosip_b200_ctx* ctx, int kernel_id, int d, int N, int p, const double* X,
                               const double* lambda, const double* alpha, const double* sigma, const double* L,
                               const double* a, const bosip_b200_gp_opts* opts, bosip_b200_gp** out);
/* Download what the device holds (any pointer may be NULL): L (p x N x N col-major), Linv (same layout),
 * a (N x p).  For tests and for ncclBroadcast-free host replication. */
int bosip_b200_gp_get_factors(bosip_b200_gp* gp, double* L, double* Linv, double* a);
/* Shape of a conditioned GP (any pointer may be NULL): input dimension d, training points N, output dimension p. */
int bosip_b200_gp_dims(const bosip_b200_gp* gp, int* d, int* N, int* p);
int bosip_b200_gp_free(bosip_b200_gp* gp);

/* ---- BOSS.mean / var / mean_and_var (model_post, X::AbstractMatrix)  -> p x M
 *      (src/posterior/likelihood.jl:8,12,21,25; src/likelihoods/normal.jl:37,43,60,68) ------------------ */
/* mean_at_Xq: p x M prior-mean values m_j(Xq) or NULL (zero mean).  var may be NULL (mean only). */
int bosip_b200_gp_predict(bosip_b200_gp* gp, const double* Xq, int64_t M, int mem, const double* mean_at_Xq,
                          double* mu, double* var);

/* ---- fused posterior closures (src/posterior/log_posterior.jl:15-61 over src/posterior/likelihood.jl:19-59
 *      and src/likelihoods/normal.jl:28-79): ONE GP predict per point, all requested outputs at once. ------ */
/* Outputs (each M doubles or NULL, selected by `want`): log_approx_post, log_post_mean, log_post_var.
 * n_clamped (optional) counts likelihood variances in [-1e-8, 0) clamped to 0 (=> -Inf).  A variance
 * < -1e-8 returns BOSIP_B200_EDOMAIN like the reference's DomainError. */
int bosip_b200_posterior_eval(bosip_b200_gp* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                              const double* Xq, int64_t M, int mem, const double* mean_at_Xq, unsigned want,
                              double* log_approx_post, double* log_post_mean, double* log_post_var,
                              int64_t* n_clamped);

/* Same sweep, plus the acquisition argmax over the evaluated points in the same pass (acq = BOSIP_B200_ACQ_* or -1 for
 * none): what a SamplingAM-style maximiser needs when the per-point values are wanted too.  best_idx is
 * index_offset + (0-based position in Xq) with Julia's `argmax` semantics: first maximal index, -Inf is an ordinary value
 * (an all -Inf set returns its first index), a NaN value wins over everything (first NaN); -1 only when M == 0. */
int bosip_b200_posterior_eval_argmax(bosip_b200_gp* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                     const double* Xq, int64_t M, int mem, const double* mean_at_Xq, unsigned want,
                                     double* log_approx_post, double* log_post_mean, double* log_post_var,
                                     int64_t* n_clamped, int acq, int64_t index_offset, int64_t* best_idx,
                                     double* best_val);

/* Bayesian-inference variant (MultiFittedParams, src/posterior/log_posterior.jl:86-95,120-129,153-162): an ensemble of
 * n_samples conditioned GPs (hyper-parameter samples of the same data).  Each requested closure is
 * k*log_prior + log( (1/S) sum_s exp(loglike-term_s(x)) ), the average taken in linear space exactly as the reference's
 * `log.(mapreduce(f -> exp.(f(x)), .+, fs) ./ sample_count)` does.  Same outputs/argmax semantics as above. */
int bosip_b200_posterior_eval_bi(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like,
                                 const bosip_b200_prior* prior, const double* Xq, int64_t M, int mem, const double* mean_at_Xq,
                                 unsigned want, double* log_approx_post, double* log_post_mean, double* log_post_var,
                                 int64_t* n_clamped, int acq, int64_t index_offset, int64_t* best_idx, double* best_val);

/* ---- the Likelihood API on its own, for ANY ModelPosterior's outputs: NormalLikelihood given mu, var (p x M each;
 *      var == NULL means zero variance).  Replaces loglike_marginal / loglike (src/likelihoods/normal.jl:28-30,
 *      src/types/likelihood.jl:53-71), log_marginal_likelihood_mean (normal.jl:32-53), log_likelihood_mean
 *      (src/posterior/likelihood.jl:34-43), log_sq_likelihood_mean (normal.jl:55-79), log_likelihood_variance
 *      (src/posterior/likelihood.jl:47-59).  This is the seam the reference's unit tests use with MockModelPosterior
 *      (test/unit/utils.jl:1-11).  Any output may be NULL. --------------------------------------------------- */
int bosip_b200_normal_likelihood_eval(bosip_b200_ctx* ctx, const bosip_b200_like_normal* like, const double* mu,
                                      const double* var, int64_t M, int mem, double* loglike_marginal /* p x M */,
                                      double* loglike /* M */, double* log_marginal_mean /* p x M */,
                                      double* log_like_mean /* M */, double* log_sq_like_mean /* M */,
                                      double* log_like_var /* M */, int64_t* n_clamped);

/* ---- acquisition over a candidate set + argmax: MaxVar / LogMaxVar closure (src/acquisitions/maxvar.jl:9-11,
 *      logmaxvar.jl:13-15) evaluated by BOSS's SamplingAM / GridAM (docs/src/hyperparams.md:69). ----------- */
/* Candidates index_offset .. index_offset+M-1 of the source (so that ranks can shard one global set).
 * Julia `argmax` semantics: ties resolve to the lowest global index, -Inf is an ordinary value, NaN wins (first NaN).
 * best_idx is the GLOBAL 0-based index, best_x the d coordinates of the winner; best_idx = -1 only when M == 0. */
int bosip_b200_acq_argmax(bosip_b200_gp* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                          const bosip_b200_cand_src* src, int64_t M, int64_t index_offset, int acq,
                          int64_t* best_idx, double* best_val, double* best_x);
/* MultiFittedParams variants (src/posterior/log_posterior.jl:86-95,120-129,153-162 under src/acquisitions/maxvar.jl:9-11 and
 * src/posterior/evidence.jl:19-24): the same sweeps over an ensemble of n_samples conditioned GPs, averaged in linear space. */
int bosip_b200_acq_argmax_bi(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like,
                             const bosip_b200_prior* prior, const bosip_b200_cand_src* src, int64_t M, int64_t index_offset, int acq,
                             int64_t* best_idx, double* best_val, double* best_x);
int bosip_b200_evidence_sum_bi(bosip_b200_gp* const* gps, int n_samples, const bosip_b200_like_normal* like,
                               const bosip_b200_prior* prior, const double* xs, int64_t S, int mem, unsigned which, double* sum_out);
/* ---- marginal-integration sweep: `_compute_marginals_int` (ext/CairoExt.jl:146-246), the heaviest batched caller of the
 *      path.  lhc is the d x G Latin-hypercube sample matrix (host).  For every dimension k and grid value i (diagonal), and for
 *      every pair a < b and grid values (ia, ib), row k (rows a, b) of lhc is overwritten with range(lb, ub; length = grid_size)
 *      values, func = exp(closure selected by `which`) is evaluated on the G points and averaged: `mean(f(grid))`.
 *      The (d*gs + C(d,2)*gs^2) * G query points are generated on the device; only the means come back:
 *      diag[i + gs*k]  and  pairs[ia + gs*ib + gs*gs*pair]  with pairs ordered (1,2),(1,3),...,(d-1,d) as in the reference.
 *      Un-normalised (normalize_prob_vals!, ext/CairoExt.jl:410-430, stays on the host).  diag or pairs may be NULL. ------ */
int bosip_b200_marginals_int(bosip_b200_gp* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior, const double* lhc,
                             int64_t G, int grid_size, const double* lb, const double* ub, unsigned which, double* diag,
                             double* pairs);
/* Materialise candidates index_offset.. of a PHILOX/GRID source into `out` (d x M, memory space `mem`). */
int bosip_b200_candidates(bosip_b200_ctx* ctx, const bosip_b200_cand_src* src, int d, int64_t M,
                          int64_t index_offset, int mem, double* out);

/* ---- evidence (src/posterior/evidence.jl:19-24): sum_i post(x_i)/pdf(prior, x_i) over the given prior
 *      samples xs (d x S).  which = BOSIP_B200_WANT_APPROX or BOSIP_B200_WANT_MEAN selects the posterior.
 *      Returns the SUM and the count so that ranks can all-reduce; evidence = sum / S. -------------------- */
int bosip_b200_evidence_sum(bosip_b200_gp* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                            const double* xs, int64_t S, int mem, unsigned which, double* sum_out);

/* ---- MMD (src/metrics/mmd.jl:23-28): biased V-statistic.  A d x n, B d x m, lambda d (ARD length-scales of
 *      with_lengthscale).  sums_out[3] = {sum k(A,A), sum k(B,B), sum k(A,B)} over THIS rank's share of the
 *      tile grid (rank/world = 0/1 for the whole problem); mmd2 = sAA/n^2 + sBB/m^2 - 2 sAB/(n m). ------- */
int bosip_b200_mmd_sums(bosip_b200_ctx* ctx, int kernel_id, int d, const double* lambda, const double* A, int64_t n,
                        const double* B, int64_t m, int mem, int rank, int world, double* sums_out);
int bosip_b200_mmd(bosip_b200_ctx* ctx, int kernel_id, int d, const double* lambda, const double* A, int64_t n,
                   const double* B, int64_t m, int mem, double* mmd2_out);

/* ---- TV (src/metrics/tv.jl:59-73 with logsumexp/logmeanexp of src/utils/utils.jl:7-15) ------------------ */
/* stage 1: out[4] = {max(lw+a), sum exp(lw+a-max), max(lw+t), sum exp(lw+t-max)} over this rank's G points.
 * stage 2: given the global ev_approx/ev_true (logmeanexp), out[2] = {m, s} with
 *          exp(m) * s = sum_i exp(lw_i) |exp(a_i-ev_a) - exp(t_i-ev_t)|  (a scaled sum: pairs from several ranks merge like
 *          log-sum-exp pairs; TV = 0.5 * exp(m + log s - log G)).   bosip_b200_tv = both stages on one GPU: host vectors
 *          are uploaded once, the stages and their merges run back to back on the device, 8 bytes come back. */
int bosip_b200_tv_stage1(bosip_b200_ctx* ctx, const double* log_ws, const double* a, const double* t, int64_t G,
                         int mem, double* out4);
int bosip_b200_tv_stage2(bosip_b200_ctx* ctx, const double* log_ws, const double* a, const double* t, int64_t G,
                         int mem, double ev_approx, double ev_true, double* out2);
int bosip_b200_tv(bosip_b200_ctx* ctx, const double* log_ws, const double* a, const double* t, int64_t G, int mem,
                  double* tv_out);

/* ---- confidence sets (src/utils/confidence_set.jl:19-81) and the AEConfidence termination condition's core
 *      (src/term_conds/ae_confidence.jl:59-79).  All vectors have n entries in memory space `mem`; scalars come back on the host. */
/* quantile(vals, Distributions.weights(ws), p) as StatsBase defines it for non-frequency weights (zero weights dropped, values
 * sorted, h = p (W - w_1) + w_1, linear interpolation on the cumulative weights): what find_cutoff calls with p = 1 - q
 * (confidence_set.jl:24-26).  ws == NULL: unit weights.  Any NaN value -> NaN. */
int bosip_b200_weighted_quantile(bosip_b200_ctx* ctx, const double* vals, const double* ws, int64_t n, int mem, double p, double* out);
/* approx_cutoff_area (confidence_set.jl:52-56): sum of exp(log ws - log sum ws) over the entries with vals >= c. */
int bosip_b200_cutoff_area(bosip_b200_ctx* ctx, const double* vals, const double* ws, int64_t n, int mem, double c, double* out);
/* set_iou (confidence_set.jl:73-81): in_A, in_B are n bytes (Julia Vector{Bool}), logprior = logpdf.(x_prior, eachcol(xs));
 * ws = 1 ./ pdf, normalised; out = sum ws[A & B] / sum ws[A | B]. */
int bosip_b200_set_iou(bosip_b200_ctx* ctx, const uint8_t* in_A, const uint8_t* in_B, const double* logprior, int64_t n, int mem, double* out);
/* calculate(::AEConfidence, bosip) from the two closures' LOG values on the samples xs (log_f_approx = log_approx_posterior(xs),
 * log_f_expect = log_posterior_mean(xs), e.g. straight from bosip_b200_posterior_eval with mem = DEVICE): f = exp(log f),
 * ws = exp(log f - logprior), c = quantile(f, weights(ws), 1 - q) for each closure, IoU of {f_approx > c_approx} and
 * {f_expect > c_expect} under ws = 1 / pdf(prior).  c_approx / c_expect (optional) return the two cut-offs. */
int bosip_b200_confidence_iou(bosip_b200_ctx* ctx, const double* log_f_approx, const double* log_f_expect, const double* logprior,
                              int64_t n, int mem, double q, double* iou, double* c_approx, double* c_expect);

/* ---- multi-GPU from ONE host process: what a Julia host calling through BOSS.maximize_acquisition
 *      (src/main.jl:98-101, docs/src/hyperparams.md:69) or the posterior closures gets without torchrun.  bosip_b200_init_multi
 *      opens one context per device (devs == NULL: devices 0..ndev-1); the fit is replicated on every device; every `_multi`
 *      call shards its points into contiguous ranges (first M % ndev devices get one extra point, global indices preserved),
 *      runs the shards concurrently on one host thread per device and combines the few doubles that come back in fixed rank
 *      order.  Large buffers are HOST memory (a device pointer belongs to one GPU).  Per-point outputs and the argmax are
 *      bit-identical to the one-GPU entry points; sums differ only in the association of the per-device partials. ---------- */
typedef struct bosip_b200_multi bosip_b200_multi;       /* N contexts, one per device */
typedef struct bosip_b200_gp_multi bosip_b200_gp_multi; /* N replicas of one conditioned GP */
int bosip_b200_init_multi(int ndev, const int* devs, bosip_b200_multi** out);
int bosip_b200_shutdown_multi(bosip_b200_multi* mh);
int bosip_b200_multi_size(const bosip_b200_multi* mh);
/* Borrow the per-device context / GP replica (e.g. to select the variance engine or read timings); owned by the multi handle. */
bosip_b200_ctx* bosip_b200_multi_context(bosip_b200_multi* mh, int rank);
bosip_b200_gp* bosip_b200_gp_multi_replica(bosip_b200_gp_multi* gp, int rank);
/* Message of the last failed `_multi` call (mh == NULL: of the last failed bosip_b200_init_multi on this thread). */
const char* bosip_b200_multi_last_error(const bosip_b200_multi* mh);
/* bosip_b200_gp_fit on every device (src/posterior/log_posterior.jl:81). */
int bosip_b200_gp_fit_multi(bosip_b200_multi* mh, int kernel_id, int d, int N, int p, const double* X, const double* Y,
                            const double* lambda, const double* alpha, const double* sigma, const double* mean_at_X,
                            const bosip_b200_gp_opts* opts, bosip_b200_gp_multi** out);
int bosip_b200_gp_free_multi(bosip_b200_gp_multi* gp);
/* bosip_b200_posterior_eval_argmax over contiguous M ranges (src/posterior/log_posterior.jl:15-61); acq = -1: no argmax. */
int bosip_b200_posterior_eval_multi(bosip_b200_gp_multi* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                    const double* Xq, int64_t M, const double* mean_at_Xq, unsigned want, double* log_approx_post,
                                    double* log_post_mean, double* log_post_var, int64_t* n_clamped, int acq, int64_t index_offset,
                                    int64_t* best_idx, double* best_val);
/* bosip_b200_acq_argmax over contiguous candidate ranges (src/acquisitions/maxvar.jl:9-11, logmaxvar.jl:13-15). */
int bosip_b200_acq_argmax_multi(bosip_b200_gp_multi* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                const bosip_b200_cand_src* src, int64_t M, int64_t index_offset, int acq, int64_t* best_idx,
                                double* best_val, double* best_x);
/* bosip_b200_evidence_sum over contiguous sample ranges (src/posterior/evidence.jl:19-24). */
int bosip_b200_evidence_sum_multi(bosip_b200_gp_multi* gp, const bosip_b200_like_normal* like, const bosip_b200_prior* prior,
                                  const double* xs, int64_t S, unsigned which, double* sum_out);
/* calc_mmd with the tile grid dealt round-robin to the devices (src/metrics/mmd.jl:23-28); A, B host. */
int bosip_b200_mmd_multi(bosip_b200_multi* mh, int kernel_id, int d, const double* lambda, const double* A, int64_t n,
                         const double* B, int64_t m, double* mmd2_out);
/* calc_tv with the grid cut into contiguous slices, (max, sum) pairs merged between the stages (src/metrics/tv.jl:59-73). */
int bosip_b200_tv_multi(bosip_b200_multi* mh, const double* log_ws, const double* a, const double* t, int64_t G, double* tv_out);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------------------- */
/* FP64 pipe microbenchmark on this GPU: DMMA (mma.sync m8n8k4.f64) and DFMA TFLOP/s -- the roofline
 * denominators MEASURED_PEAKS.json lacks. */
int bosip_b200_fp64_peak(bosip_b200_ctx* ctx, double* dmma_tflops, double* dfma_tflops);
/* Self-test of the library's branch-free FP64 exp(-x)/sqrt(x) (csrc/fastmath.cuh) against the CUDA library versions on
 * n host values x >= 0: four host outputs of n doubles each.  variant 0: polynomial exp + two-step sqrt (MMD, TV, K build);
 * variant 1: table-driven exp + one-step sqrt (cross-kernel). */
int bosip_b200_math_selftest(bosip_b200_ctx* ctx, const double* x, int64_t n, int variant, double* exp_neg_fast, double* sqrt_fast,
                             double* exp_neg_ref, double* sqrt_ref);
/* Per-kernel device time (CUDA events on the launching stream) accumulated since the last reset:
 * ms_out[0]=cross-kernel+mean, [1]=variance GEMM, [2]=epilogue/argmax, launches_out[0..2] the launch counts. */
int bosip_b200_timing_reset(bosip_b200_ctx* ctx, int enable);
int bosip_b200_timing_get(bosip_b200_ctx* ctx, double* ms_out3, int64_t* launches_out3);

#ifdef __cplusplus
}
#endif
#endif /* BOSIP_B200_H */
