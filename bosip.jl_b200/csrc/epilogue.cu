// Fused epilogue: GP (mu, sigma^2) assembly + NormalLikelihood algebra + log-prior + posterior closures, and the
// block reductions that follow them (acquisition argmax, evidence sum), and device-side candidate generation.
//
// Reference statements restated here (paths relative to /root/reference):
//   normlogpdf                      Distributions.logpdf(Normal) = StatsFuns.normlogpdf           
//   plug-in log-likelihood          src/likelihoods/normal.jl:28-30 via src/posterior/likelihood.jl:19-29
//   log E[l]                        src/likelihoods/normal.jl:42-51 + src/posterior/likelihood.jl:34-43
//   log E[l^2]                      src/likelihoods/normal.jl:67-77
//   log V[l], clamp, DomainError    src/posterior/likelihood.jl:47-65
//   log-prior                       src/posterior/log_posterior.jl:228-229
//   log_approx_posterior / log_posterior_mean / log_posterior_variance   src/posterior/log_posterior.jl:15-61
#include <float.h>

#include "common.cuh"
#include "philox.cuh"

namespace bosip {

constexpr double LOG2PI = 1.8378770664093454835606594728112;
constexpr double MAX_NEG_VAR = 1e-8;  // src/posterior/likelihood.jl:3

__device__ __forceinline__ double normlogpdf(double mu, double s, double z) {
  const double zv = (z - mu) / s;
  return -(zv * zv + LOG2PI) / 2.0 - log(s);
}

__device__ __forceinline__ double log_prior_point(const EpiParams& ep, const double* __restrict__ x) {
  double lp = 0.0;
  for (int k = 0; k < ep.d; ++k) {
    const double xk = x[k];
    const int kind = ep.prior_kind[k];
    if (kind == BOSIP_B200_PRIOR_UNIFORM) {
      lp += (xk >= ep.prior_prm[k][0] && xk <= ep.prior_prm[k][1]) ? ep.prior_const[k] : -INFINITY;
    } else if (kind == BOSIP_B200_PRIOR_NORMAL) {
      lp += normlogpdf(ep.prior_prm[k][0], ep.prior_prm[k][1], xk);
    } else {
      lp += (xk >= ep.prior_prm[k][2] && xk <= ep.prior_prm[k][3])
                ? normlogpdf(ep.prior_prm[k][0], ep.prior_prm[k][1], xk) + ep.prior_const[k]
                : -INFINITY;
    }
  }
  return lp;
}

// counters[0] = #variances clamped to 0, counters[1] = min chunk-global index with variance < -1e-8 (init ~0ull)
__global__ void __launch_bounds__(256)
epilogue_kernel(const __grid_constant__ EpiParams ep, const double* __restrict__ xq, const double* __restrict__ logprior,
                const double* __restrict__ mean_at_xq, long long m_valid, long long mc, int nsplit, int T,
                const double* __restrict__ mu_part, int have_var, const double* __restrict__ q_part,
                double* __restrict__ mu_out, double* __restrict__ var_out, double* __restrict__ lp_out,
                double* __restrict__ log_approx, double* __restrict__ log_mean, double* __restrict__ log_var,
                unsigned long long* __restrict__ counters, unsigned long long index_base, int bi_mode) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= m_valid) return;
  const int p = ep.p;

  // bi_mode 0: one fit, outputs are the log-closures.  bi_mode 1/2: hyper-parameter sample s of a BI ensemble
  // (src/posterior/log_posterior.jl:86-95,120-129,153-162): the outputs ACCUMULATE exp(log-likelihood term) in linear space
  // (1: first sample, store; 2: add); the prior and the log are applied once by bi_finish_kernel.
  double lp = 0.0;
  if (bi_mode == 0) {
    if (ep.has_prior == 2) lp = logprior[m];
    else if (ep.has_prior == 1) lp = log_prior_point(ep, xq + m * ep.d);
  }

  // Observation i sums grp[i] consecutive model outputs (NormalSumLikelihood, src/likelihoods/normal_sum.jl:32-57; 1 otherwise).
  double plug = ep.c_approx, lm = ep.c_mean, lsq = ep.c_sq;
  double exp_mu = 0.0, exp_var = 0.0;  // ExpLikelihood: the single model output
  int j = 0;
  for (int i = 0; i < ep.q; ++i) {
    double mu_i = 0.0, var_i = 0.0;
    for (int gidx = 0; gidx < ep.grp[i]; ++gidx, ++j) {
      double mu = 0.0;
      for (int s = 0; s < nsplit; ++s) mu += mu_part[((size_t)s * p + j) * mc + m];
      if (mean_at_xq) mu += mean_at_xq[m * p + j];
      if (mu_out) mu_out[m * p + j] = mu;
      mu_i += mu;
      if (have_var) {
        double q = 0.0;
        for (int t = 0; t < T; ++t) q += q_part[((size_t)t * p + j) * mc + m];
        double var = ep.kxx[j] - q;
        if (ep.pred_noise) var += ep.sig2[j];
        if (ep.clamp_var && var < 0.0) var = 0.0;
        if (var_out) var_out[m * p + j] = var;
        var_i += var;
      }
    }
    if (ep.like_kind == BOSIP_B200_LIKE_EXP) { exp_mu = mu_i; exp_var = var_i; continue; }
    const double z = ep.z[i], so = ep.so[i];
    if (log_approx) plug += normlogpdf(mu_i, so, z);
    if (have_var) {
      const double so2 = so * so;
      lm += normlogpdf(mu_i, sqrt(so2 + var_i), z);
      lsq += normlogpdf(mu_i, sqrt((so2 + 2.0 * var_i) / 2.0), z);
    }
  }
  if (ep.like_kind == BOSIP_B200_LIKE_EXP) {
    // src/likelihoods/exp.jl:9-16,19-33,50-63: loglike = y; log E[l] = mu + sigma^2/2; log V[l] = 2 (mu + sigma^2) + log(1 - exp(-sigma^2))
    if (bi_mode == 0 && lp_out) lp_out[m] = lp;
    const double lv = 2.0 * (exp_mu + exp_var) + log(1.0 - exp(-exp_var));
    if (log_approx) log_approx[m] = bi_mode == 0 ? lp + exp_mu : (bi_mode == 1 ? exp(exp_mu) : log_approx[m] + exp(exp_mu));
    if (have_var) {
      const double lme = exp_mu + 0.5 * exp_var;
      if (log_mean) log_mean[m] = bi_mode == 0 ? lp + lme : (bi_mode == 1 ? exp(lme) : log_mean[m] + exp(lme));
      if (log_var) log_var[m] = bi_mode == 0 ? 2.0 * lp + lv : (bi_mode == 1 ? exp(lv) : log_var[m] + exp(lv));
    }
    return;
  }
  if (bi_mode == 0 && lp_out) lp_out[m] = lp;
  if (log_approx) log_approx[m] = bi_mode == 0 ? lp + plug : (bi_mode == 1 ? exp(plug) : log_approx[m] + exp(plug));
  if (have_var) {
    if (log_mean) log_mean[m] = bi_mode == 0 ? lp + lm : (bi_mode == 1 ? exp(lm) : log_mean[m] + exp(lm));
    if (log_var) {
      const double log_sq = ep.log_C + lsq;
      double v = exp(log_sq) - exp(2.0 * lm);
      if (v < 0.0) {
        if (v >= -MAX_NEG_VAR) {
          v = 0.0;
          atomicAdd(&counters[0], 1ull);
        } else {
          atomicMin(&counters[1], index_base + (unsigned long long)m);
        }
      }
      log_var[m] = bi_mode == 0 ? 2.0 * lp + log(v) : (bi_mode == 1 ? v : log_var[m] + v);
    }
  }
}

// BI ensemble: out = k * log_prior + log(acc / S)   -- log.(mapreduce(f -> exp.(f(x)), .+, fs) ./ sample_count)
__global__ void __launch_bounds__(256)
bi_finish_kernel(const __grid_constant__ EpiParams ep, const double* __restrict__ xq, const double* __restrict__ logprior,
                 long long m_valid, double inv_s, double* __restrict__ lp_out, double* __restrict__ la, double* __restrict__ lm,
                 double* __restrict__ lv) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= m_valid) return;
  double lp = 0.0;
  if (ep.has_prior == 2) lp = logprior[m];
  else if (ep.has_prior == 1) lp = log_prior_point(ep, xq + m * ep.d);
  if (lp_out) lp_out[m] = lp;
  if (la) la[m] = lp + log(la[m] * inv_s);
  if (lm) lm[m] = lp + log(lm[m] * inv_s);
  if (lv) lv[m] = 2.0 * lp + log(lv[m] * inv_s);
}

int launch_bi_finish(Ctx* ctx, const EpiParams& ep, const double* xq, const double* logprior, int64_t m_valid, int n_samples,
                     double* lp_out, double* la, double* lm, double* lv, cudaStream_t st) {
  if (m_valid <= 0) return 0;
  bi_finish_kernel<<<(unsigned)((m_valid + 255) / 256), 256, 0, st>>>(ep, xq, logprior, (long long)m_valid, 1.0 / n_samples, lp_out, la, lm, lv);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

int launch_epilogue(Ctx* ctx, const Gp* gp, const EpiParams& ep, const double* xq, const double* logprior,
                    const double* mean_at_xq, int64_t m_valid, int64_t mc, int nsplit, const double* mu_part,
                    bool have_var, const double* q_part, double* mu_out, double* var_out, double* lp_out,
                    double* log_approx, double* log_mean, double* log_var, unsigned long long* counters,
                    uint64_t index_base, int bi_mode, cudaStream_t st) {
  if (m_valid <= 0) return 0;
  const int block = 256;
  const unsigned grid = (unsigned)((m_valid + block - 1) / block);
  epilogue_kernel<<<grid, block, 0, st>>>(ep, xq, logprior, mean_at_xq, (long long)m_valid, (long long)mc, nsplit, 1 /* the GEMM writes one q per row */,
                                          mu_part, have_var ? 1 : 0, q_part, mu_out, var_out, lp_out, log_approx,
                                          log_mean, log_var, counters, (unsigned long long)index_base, bi_mode);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ likelihood from (mu, var)
// The Likelihood-API layer on its own, for ANY ModelPosterior's (mu, var) p x M outputs -- what the reference's unit tests
// exercise with MockModelPosterior (test/unit/utils.jl:1-11, test/unit/test/posterior/likelihood.jl,
// test/unit/test/likelihoods/normal.jl).  Outputs (any may be NULL): marg_approx p x M = loglike_marginal(mu),
// l_approx M = loglike(mu), marg_mean p x M = log_marginal_likelihood_mean, l_mean M, l_sq M = log_sq_likelihood_mean,
// l_var M = log_likelihood_variance.
__global__ void __launch_bounds__(256)
like_eval_kernel(const __grid_constant__ EpiParams ep, const double* __restrict__ mu_in, const double* __restrict__ var_in,
                 long long m_valid, double* __restrict__ marg_approx, double* __restrict__ l_approx,
                 double* __restrict__ marg_mean, double* __restrict__ l_mean, double* __restrict__ l_sq,
                 double* __restrict__ l_var, unsigned long long* __restrict__ counters) {
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= m_valid) return;
  const int p = ep.p, nq = ep.q;
  if (ep.like_kind == BOSIP_B200_LIKE_EXP) {  // src/likelihoods/exp.jl
    const double mu = mu_in[m], var = var_in ? var_in[m] : 0.0;
    if (marg_approx) marg_approx[m] = mu;
    if (marg_mean) marg_mean[m] = mu + 0.5 * var;
    if (l_approx) l_approx[m] = mu;
    if (l_mean) l_mean[m] = mu + 0.5 * var;
    if (l_sq) l_sq[m] = 2.0 * mu + 2.0 * var;
    if (l_var) l_var[m] = 2.0 * (mu + var) + log(1.0 - exp(-var));
    return;
  }
  double plug = ep.c_approx, lm = ep.c_mean, lsq = ep.c_sq;
  int j = 0;
  for (int i = 0; i < nq; ++i) {
    double mu = 0.0, var = 0.0;
    for (int gidx = 0; gidx < ep.grp[i]; ++gidx, ++j) { mu += mu_in[m * p + j]; if (var_in) var += var_in[m * p + j]; }
    const double z = ep.z[i], so = ep.so[i], so2 = so * so;
    const double t0 = normlogpdf(mu, so, z);
    const double t1 = normlogpdf(mu, sqrt(so2 + var), z);
    const double t2 = normlogpdf(mu, sqrt((so2 + 2.0 * var) / 2.0), z);
    // per-dimension constant: LogNormal's -log z_obs_i = -(z'_i - so^2/2) (z' is the shifted log-observation)
    const double ci = (ep.like_kind == BOSIP_B200_LIKE_LOGNORMAL) ? -(z - so2 / 2.0) : 0.0;
    if (marg_approx) marg_approx[m * nq + i] = t0 + ci;
    if (marg_mean) marg_mean[m * nq + i] = t1 + ci;
    plug += t0; lm += t1; lsq += t2;
  }
  if (l_approx) l_approx[m] = plug;
  if (l_mean) l_mean[m] = lm;
  const double log_sq = ep.log_C + lsq;
  if (l_sq) l_sq[m] = log_sq;
  if (l_var) {
    double v = exp(log_sq) - exp(2.0 * lm);
    if (v < 0.0) {
      if (v >= -MAX_NEG_VAR) { v = 0.0; atomicAdd(&counters[0], 1ull); }
      else atomicMin(&counters[1], (unsigned long long)m);
    }
    l_var[m] = log(v);
  }
}

int launch_like_eval(Ctx* ctx, const EpiParams& ep, const double* mu, const double* var, int64_t m, double* marg_approx,
                     double* l_approx, double* marg_mean, double* l_mean, double* l_sq, double* l_var,
                     unsigned long long* counters, cudaStream_t st) {
  if (m <= 0) return 0;
  like_eval_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(ep, mu, var, (long long)m, marg_approx, l_approx, marg_mean,
                                                               l_mean, l_sq, l_var, counters);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ argmax
// Julia `argmax` semantics of BOSS's candidate maximisers (findmax over `isless`): first maximal index; -Inf is an ordinary
// value (an all -Inf candidate set returns its first index); NaN orders above everything, so the first NaN wins.
struct Best { double v; long long i; };
__device__ __forceinline__ Best better(Best a, Best b) {
  if (b.i < 0) return a;
  if (a.i < 0) return b;
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an || bn) {
    if (an && bn) return b.i < a.i ? b : a;
    return bn ? b : a;
  }
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best y; y.v = __shfl_xor_sync(0xffffffffu, x.v, o); y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
    x = better(x, y);
  }
  return x;
}
__device__ __forceinline__ Best block_best(Best x) {
  __shared__ double sv[32];
  __shared__ long long si[32];
  x = warp_best(x);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (l == 0) { sv[w] = x.v; si[w] = x.i; }
  __syncthreads();
  if (w == 0) {
    Best y; y.v = l < nw ? sv[l] : 0.0; y.i = l < nw ? si[l] : -1;
    x = warp_best(y);
  }
  return x;
}

__global__ void __launch_bounds__(256)
argmax_stage1(const double* __restrict__ vals, long long m_valid, long long offset, int exp_vals,
              double* __restrict__ sv, long long* __restrict__ si) {
  Best b; b.v = 0.0; b.i = -1;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < m_valid; m += (long long)gridDim.x * blockDim.x) {
    double v = vals[m];
    if (exp_vals) v = exp(v);
    Best c; c.v = v; c.i = offset + m; b = better(b, c);
  }
  b = block_best(b);
  if (threadIdx.x == 0) { sv[blockIdx.x] = b.v; si[blockIdx.x] = b.i; }
}
__global__ void __launch_bounds__(256)
argmax_stage2(const double* __restrict__ sv, const long long* __restrict__ si, int nblocks, double* best_val, long long* best_idx) {
  Best b; b.v = 0.0; b.i = -1;
  for (int k = threadIdx.x; k < nblocks; k += blockDim.x) { Best c; c.v = sv[k]; c.i = si[k]; b = better(b, c); }
  b = block_best(b);
  if (threadIdx.x == 0) {
    Best cur; cur.v = *best_val; cur.i = *best_idx;
    cur = better(cur, b);
    *best_val = cur.v; *best_idx = cur.i;
  }
}

int launch_argmax_chunk(Ctx* ctx, const double* vals, int64_t m_valid, int64_t global_offset, int exp_vals,
                        double* scratch_val, long long* scratch_idx, double* best_val, long long* best_idx,
                        cudaStream_t st) {
  if (m_valid <= 0) return 0;
  const int nblocks = (int)((m_valid + 2047) / 2048 < 1024 ? (m_valid + 2047) / 2048 : 1024);
  argmax_stage1<<<nblocks, 256, 0, st>>>(vals, (long long)m_valid, (long long)global_offset, exp_vals, scratch_val, scratch_idx);
  argmax_stage2<<<1, 256, 0, st>>>(scratch_val, scratch_idx, nblocks, best_val, best_idx);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ evidence sum
// sum_i exp(log_post_i) / exp(log_prior_i)   (src/posterior/evidence.jl:20,22), fixed-order two-stage reduction
__global__ void __launch_bounds__(256)
ratio_sum_stage1(const double* __restrict__ log_post, const double* __restrict__ log_prior, long long m_valid,
                 double* __restrict__ partial) {
  __shared__ double sh[256];
  double s = 0.0;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < m_valid; m += (long long)gridDim.x * blockDim.x)
    s += exp(log_post[m]) / exp(log_prior[m]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void ratio_sum_stage2(const double* __restrict__ partial, int nblocks, double* accum) {
  double s = 0.0;
  for (int k = 0; k < nblocks; ++k) s += partial[k];
  *accum += s;
}

int launch_sum_chunk(Ctx* ctx, const double* log_post, const double* log_prior, int64_t m_valid, double* scratch,
                     double* accum, cudaStream_t st) {
  if (m_valid <= 0) return 0;
  const int nblocks = (int)((m_valid + 2047) / 2048 < 256 ? (m_valid + 2047) / 2048 : 256);
  ratio_sum_stage1<<<nblocks, 256, 0, st>>>(log_post, log_prior, (long long)m_valid, scratch);
  ratio_sum_stage2<<<1, 1, 0, st>>>(scratch, nblocks, accum);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ candidates
struct CandParams {
  int kind, d;
  unsigned long long seed;
  long long n_per_dim;
  double lb[MAX_D], ub[MAX_D];
  // kind 3 (marginal-integration sweep, ext/CairoExt.jl:146-246): LHC samples d x G on the device, grid_size values per axis
  const double* lhc;
  long long G;
  int n_diag_slots;  // d * grid_size when the diagonal marginals are requested, else 0
};

__device__ __forceinline__ double axis_value(const CandParams& cp, int k, long long i) {
  const long long n = cp.n_per_dim;
  if (n == 1) return cp.lb[k];
  if (i == n - 1) return cp.ub[k];
  return __dadd_rn(__dmul_rn((double)i, __ddiv_rn(__dsub_rn(cp.ub[k], cp.lb[k]), (double)(n - 1))), cp.lb[k]);
}

__global__ void __launch_bounds__(256)
candidates_kernel(const __grid_constant__ CandParams cp, long long m, long long offset, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long gi = (unsigned long long)(offset + i);
  if (cp.kind == BOSIP_B200_CAND_PHILOX) {
    for (int k = 0; k < cp.d; ++k) {
      const double u = philox_unit(cp.seed, gi, k);
      out[i * cp.d + k] = __dadd_rn(cp.lb[k], __dmul_rn(__dsub_rn(cp.ub[k], cp.lb[k]), u));
    }
  } else if (cp.kind == 3) {
    // slot = gi / G: diagonal slots first (dim * gs + i), then pair slots (pair (a<b) in the reference's order, ia fastest);
    // the point is LHC sample gi % G with row dim (or rows a, b) overwritten by the slot's grid value(s)
    const long long slot = (long long)(gi / (unsigned long long)cp.G), g = (long long)(gi % (unsigned long long)cp.G);
    const long long gs = cp.n_per_dim;
    int da = -1, db = -1; long long ia = 0, ib = 0;
    if (slot < cp.n_diag_slots) { da = (int)(slot / gs); ia = slot % gs; }
    else {
      long long r = slot - cp.n_diag_slots;
      long long pi = r / (gs * gs); r -= pi * gs * gs;
      ia = r % gs; ib = r / gs;
      int a = 0;
      while (pi >= cp.d - 1 - a) { pi -= cp.d - 1 - a; ++a; }
      da = a; db = a + 1 + (int)pi;
    }
    for (int k = 0; k < cp.d; ++k) {
      double x = cp.lhc[(size_t)g * cp.d + k];
      if (k == da) x = axis_value(cp, k, ia);
      if (k == db) x = axis_value(cp, k, ib);
      out[i * cp.d + k] = x;
    }
  } else {  // tensor grid, first coordinate fastest; matches numpy.linspace(lb, ub, n) per axis
    unsigned long long r = gi;
    const unsigned long long n = (unsigned long long)cp.n_per_dim;
    for (int k = 0; k < cp.d; ++k) {
      const unsigned long long ik = r % n;
      r /= n;
      double x;
      if (n == 1) x = cp.lb[k];
      else if (ik == n - 1) x = cp.ub[k];
      else x = __dadd_rn(__dmul_rn((double)ik, __ddiv_rn(__dsub_rn(cp.ub[k], cp.lb[k]), (double)(n - 1))), cp.lb[k]);
      out[i * cp.d + k] = x;
    }
  }
}

// mean over each segment of `seg_len` consecutive points of exp(log-values): out[seg] = (1/seg_len) sum_g exp(v[seg*seg_len + g])
// (the `mean(f(grid))` of ext/CairoExt.jl:184,221).  One block per segment, fixed summation order.
__global__ void __launch_bounds__(128)
segmean_exp_kernel(const double* __restrict__ v, long long seg_len, double* __restrict__ out) {
  __shared__ double sh[128];
  const double* seg = v + (size_t)blockIdx.x * seg_len;
  double s = 0.0;
  for (long long g = threadIdx.x; g < seg_len; g += blockDim.x) s += exp(seg[g]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0] / (double)seg_len;
}

int launch_segmean_exp(Ctx* ctx, const double* v, int64_t n_seg, int64_t seg_len, double* out, cudaStream_t st) {
  if (n_seg <= 0) return 0;
  segmean_exp_kernel<<<(unsigned)n_seg, 128, 0, st>>>(v, (long long)seg_len, out);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

int launch_candidates(Ctx* ctx, const bosip_b200_cand_src* src, int d, int64_t m, int64_t offset, double* out,
                      cudaStream_t st) {
  if (m <= 0) return 0;
  if (d > MAX_D) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "d exceeds MAX_D");
  if (!src->lb || !src->ub) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "candidate source needs lb and ub");
  CandParams cp;
  cp.kind = src->kind; cp.d = d; cp.seed = src->seed; cp.n_per_dim = src->n_per_dim;
  cp.lhc = src->points; cp.G = (long long)src->seed; cp.n_diag_slots = src->mem;  // kind 3 reuses these fields (internal source only)
  for (int k = 0; k < d; ++k) { cp.lb[k] = src->lb[k]; cp.ub[k] = src->ub[k]; }
  candidates_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(cp, (long long)m, (long long)offset, out);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace bosip
