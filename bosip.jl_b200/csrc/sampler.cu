// Device side of the AMIS importance sampler's inner loop (SURVEY.md section 8f rank 2;
// /root/reference/src/samplers/amis/amis_method.jl:26-83,126-134 and proposal_distributions/normal.jl:75-100):
//   * draws from the MvNormal proposal (Philox keyed by the global sample index),
//   * the proposal-mixture pdf sums  Delta[m] (+)= sum_i pdf(q_i, x_m)                      (amis_method.jl:52,66-68,74),
//   * the log-weights                log_Omega[m] = log_P[m] - log(Delta[m] / t)             (amis_method.jl:53,69,75),
//   * exp_weights + the AnalyticalFitter's weighted mean / covariance                         (amis_method.jl:126-134,
//     normal.jl:75-100) with fixed-order two-stage reductions,
//   * multinomial down-sampling `resample(xs, ws, count)`                                     (src/sampling.jl:55-57).
// The samples, log_P, Delta and log_Omega stay in HBM between iterations; log_pi(xs) is the fused posterior sweep (api.cu).
#include <math.h>

#include "common.cuh"
#include "philox.cuh"

namespace bosip {

constexpr int SMP_BLOCK = 256;
constexpr int SMP_MAXQ = 64;  // proposals per mixture-sum launch

// ---------------------------------------------------------------------------------------------- MvNormal draws
// x = mu + L z, z ~ N(0, I): Box-Muller on Philox uniforms; pair (2c, 2c+1) of coordinates shares one (u1, u2)
__global__ void __launch_bounds__(SMP_BLOCK) mvn_sample_kernel(int d, const double* __restrict__ mu, const double* __restrict__ L /* d x d col-major, lower */,
                                                               unsigned long long seed, long long offset, long long m, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const unsigned long long gi = (unsigned long long)(offset + i);
  double z[MAX_D];
  for (int c = 0; c < d; c += 2) {
    // (0, 1] so that log is finite
    const double u1 = 1.0 - philox_unit(seed, gi, c), u2 = philox_unit(seed, gi, c + 1);
    const double r = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    z[c] = r * cs;
    if (c + 1 < d) z[c + 1] = r * sn;
  }
  for (int a = 0; a < d; ++a) {
    double x = mu[a];
    for (int b = 0; b <= a; ++b) x = fma(L[a + (size_t)b * d], z[b], x);
    out[i * d + a] = x;
  }
}

// ---------------------------------------------------------------------------------------------- mixture pdf sums
struct MixParams {
  int d, nq;
  const double* mu;     // [nq][d]
  const double* Linv;   // [nq][d][d] col-major, lower: inverse Cholesky factor of Sigma_q
  const double* lognorm;  // [nq]  -sum log L_ii - d/2 log(2 pi)
};

__global__ void __launch_bounds__(SMP_BLOCK) mix_pdf_kernel(const MixParams mp, const double* __restrict__ X, long long m, int accumulate,
                                                            double* __restrict__ delta) {
  extern __shared__ double sh[];  // mu | Linv | lognorm of all nq proposals
  const int d = mp.d, per = d + d * d;
  for (int e = threadIdx.x; e < mp.nq * per; e += blockDim.x) {
    const int q = e / per, r = e % per;
    sh[e] = r < d ? mp.mu[(size_t)q * d + r] : mp.Linv[(size_t)q * d * d + (r - d)];
  }
  for (int e = threadIdx.x; e < mp.nq; e += blockDim.x) sh[mp.nq * per + e] = mp.lognorm[e];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double x[MAX_D];
  for (int a = 0; a < d; ++a) x[a] = X[i * d + a];
  double s = accumulate ? delta[i] : 0.0;
  for (int q = 0; q < mp.nq; ++q) {
    const double* mu = sh + (size_t)q * per;
    const double* Li = mu + d;
    double r2 = 0.0;
    for (int a = 0; a < d; ++a) {
      double z = 0.0;
      for (int b = 0; b <= a; ++b) z = fma(Li[a + b * d], x[b] - mu[b], z);
      r2 = fma(z, z, r2);
    }
    s += exp(sh[mp.nq * per + q] - 0.5 * r2);  // pdf, as the reference accumulates it (amis_method.jl:52,67,74)
  }
  delta[i] = s;
}

__global__ void __launch_bounds__(SMP_BLOCK) amis_logw_kernel(const double* __restrict__ logp, const double* __restrict__ delta, double t, long long m,
                                                              double* __restrict__ logw) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) logw[i] = logp[i] - log(delta[i] / t);
}

// ---------------------------------------------------------------------------------------------- weighted moments
// stage 1: per-block max of log w and flags (NaN / +Inf seen); stage 2 on one block
__global__ void __launch_bounds__(SMP_BLOCK) logw_max_kernel(const double* __restrict__ lw, long long m, double* __restrict__ part_max, int* __restrict__ flags) {
  __shared__ double sm[SMP_BLOCK];
  double mx = -INFINITY;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    const double v = lw[i];
    if (isnan(v) || v == INFINITY) bad = 1;
    mx = fmax(mx, v);
  }
  if (bad) atomicOr(flags, 1);
  sm[threadIdx.x] = mx;
  __syncthreads();
  for (int s = SMP_BLOCK / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) part_max[blockIdx.x] = sm[0];
}

// quantity q of the moment sums, w = exp(lw - max):
//   pass 0: q = 0: sum w, 1: sum w^2, 2 + a: sum w x_a
//   pass 1: q = pair (a >= b): sum w (x_a - mu_a)(x_b - mu_b)
// grid = (blocks, n_quantities); per-block partials, summed in a fixed order by moment_final_kernel
__global__ void __launch_bounds__(SMP_BLOCK) moment_part_kernel(int d, int pass, const double* __restrict__ X, const double* __restrict__ lw, long long m,
                                                                const double* __restrict__ gmax, const double* __restrict__ mean,
                                                                double* __restrict__ part) {
  __shared__ double sm[SMP_BLOCK];
  const int q = blockIdx.y;
  const double mx = gmax[0];
  int a = 0, b = 0;
  if (pass == 0) { a = q - 2; }
  else { int r = q; while (r > a) { r -= a + 1; ++a; } b = r; }  // q -> (a, b), b <= a, row-wise lower triangle
  const double ma = (pass == 1) ? mean[a] : 0.0, mb = (pass == 1) ? mean[b] : 0.0;
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    const double w = exp(lw[i] - mx);
    double f;
    if (pass == 0) f = (q == 0) ? 1.0 : (q == 1) ? w : X[i * d + a];
    else f = (X[i * d + a] - ma) * (X[i * d + b] - mb);
    s = fma(w, f, s);
  }
  sm[threadIdx.x] = s;
  __syncthreads();
  for (int t = SMP_BLOCK / 2; t > 0; t >>= 1) {
    if (threadIdx.x < t) sm[threadIdx.x] += sm[threadIdx.x + t];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[(size_t)q * gridDim.x + blockIdx.x] = sm[0];
}

__global__ void moment_final_kernel(int nq, int nblocks, const double* __restrict__ part, double* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += part[(size_t)q * nblocks + b];
  out[q] = s;
}

__global__ void max_final_kernel(int nblocks, const double* __restrict__ part, double* __restrict__ out) {
  double mx = -INFINITY;
  for (int b = 0; b < nblocks; ++b) mx = fmax(mx, part[b]);
  out[0] = mx;
}

// normalised weights ws = exp(lw - max) / sum (exp_weights, amis_method.jl:126-134)
__global__ void __launch_bounds__(SMP_BLOCK) normalise_w_kernel(const double* __restrict__ lw, long long m, const double* __restrict__ gmax,
                                                                const double* __restrict__ sums, double* __restrict__ ws) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) ws[i] = exp(lw[i] - gmax[0]) / sums[0];
}

// ---------------------------------------------------------------------------------------------- multinomial resampling
// inclusive prefix sums of ws in three steps (block scan, scan of the block totals by one block, add); then every output
// sample j draws u_j ~ U(0, total) (Philox keyed by j) and picks the first index with cdf > u_j by bisection.
__global__ void __launch_bounds__(SMP_BLOCK) scan_block_kernel(const double* __restrict__ ws, long long m, double* __restrict__ cdf, double* __restrict__ totals) {
  __shared__ double sm[SMP_BLOCK];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  sm[threadIdx.x] = i < m ? ws[i] : 0.0;
  __syncthreads();
  for (int off = 1; off < SMP_BLOCK; off <<= 1) {
    const double v = threadIdx.x >= off ? sm[threadIdx.x - off] : 0.0;
    __syncthreads();
    sm[threadIdx.x] += v;
    __syncthreads();
  }
  if (i < m) cdf[i] = sm[threadIdx.x];
  if (threadIdx.x == SMP_BLOCK - 1) totals[blockIdx.x] = sm[threadIdx.x];
}
__global__ void scan_totals_kernel(long long nblocks, double* __restrict__ totals) {  // one thread: exclusive scan in index order
  double run = 0.0;
  for (long long b = 0; b < nblocks; ++b) { const double v = totals[b]; totals[b] = run; run += v; }
  totals[nblocks] = run;
}
__global__ void __launch_bounds__(SMP_BLOCK) scan_add_kernel(long long m, const double* __restrict__ totals, double* __restrict__ cdf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) cdf[i] += totals[blockIdx.x];
}
__global__ void __launch_bounds__(SMP_BLOCK) resample_kernel(int d, const double* __restrict__ X, const double* __restrict__ cdf, long long m,
                                                             const double* __restrict__ total, unsigned long long seed, long long count,
                                                             double* __restrict__ out, long long* __restrict__ idx_out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= count) return;
  const double u = philox_unit(seed, (unsigned long long)j, 0) * total[0];
  long long lo = 0, hi = m - 1;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (cdf[mid] > u) hi = mid; else lo = mid + 1;
  }
  for (int a = 0; a < d; ++a) out[j * d + a] = X[lo * d + a];
  if (idx_out) idx_out[j] = lo;
}

// ---------------------------------------------------------------------------------------------- host entry points
static int stage_in(Ctx* ctx, const double* src, size_t n, int mem, double* dev_buf, const double** out, cudaStream_t st) {
  if (mem == BOSIP_B200_DEVICE) { *out = src; return 0; }
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dev_buf, src, n * 8, cudaMemcpyHostToDevice, st));
  *out = dev_buf;
  return 0;
}

int mvn_sample(Ctx* ctx, int d, const double* mu, const double* L, uint64_t seed, int64_t offset, int64_t m, int mem, double* out) {
  if (d < 1 || d > MAX_D || m < 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "mvnormal_sample: bad dimensions");
  if (m == 0) return 0;
  cudaStream_t st = ctx->s_comp;
  const size_t o_par = 0, o_out = round_up((size_t)(d + d * d) * 8, 256);
  BOSIP_TRY(ensure_ws(ctx, o_out + (mem == BOSIP_B200_HOST ? (size_t)m * d * 8 : 0)));
  double* par = reinterpret_cast<double*>(ctx->ws + o_par);
  BOSIP_CUDA(ctx, cudaMemcpyAsync(par, mu, (size_t)d * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(par + d, L, (size_t)d * d * 8, cudaMemcpyHostToDevice, st));
  double* dst = mem == BOSIP_B200_HOST ? reinterpret_cast<double*>(ctx->ws + o_out) : out;
  mvn_sample_kernel<<<(unsigned)((m + SMP_BLOCK - 1) / SMP_BLOCK), SMP_BLOCK, 0, st>>>(d, par, par + d, seed, (long long)offset, (long long)m, dst);
  BOSIP_CUDA(ctx, cudaGetLastError());
  if (mem == BOSIP_B200_HOST) BOSIP_CUDA(ctx, cudaMemcpyAsync(out, dst, (size_t)m * d * 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->launches += 1;
  return 0;
}

// mus: d x nq (host), Linvs: nq matrices d x d col-major (host), lognorm: nq (host); X, delta follow `mem`
int mix_pdf_sum(Ctx* ctx, int d, int nq, const double* mus, const double* Linvs, const double* lognorm, const double* X, int64_t m, int mem,
                int accumulate, double* delta) {
  if (d < 1 || d > MAX_D || nq < 1 || m < 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "mixture_pdf_sum: bad dimensions");
  if (m == 0) return 0;
  cudaStream_t st = ctx->s_comp;
  const size_t per = (size_t)d + (size_t)d * d;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_par = carve((size_t)SMP_MAXQ * (per + 1) * 8), o_x = carve(mem == BOSIP_B200_HOST ? (size_t)m * d * 8 : 0),
               o_d = carve(mem == BOSIP_B200_HOST ? (size_t)m * 8 : 0);
  BOSIP_TRY(ensure_ws(ctx, off));
  double* par = reinterpret_cast<double*>(ctx->ws + o_par);
  const double* Xd; double* dd = delta;
  BOSIP_TRY(stage_in(ctx, X, (size_t)m * d, mem, reinterpret_cast<double*>(ctx->ws + o_x), &Xd, st));
  if (mem == BOSIP_B200_HOST) {
    dd = reinterpret_cast<double*>(ctx->ws + o_d);
    if (accumulate) BOSIP_CUDA(ctx, cudaMemcpyAsync(dd, delta, (size_t)m * 8, cudaMemcpyHostToDevice, st));
  }
  for (int q0 = 0; q0 < nq; q0 += SMP_MAXQ) {
    const int nb = std::min(SMP_MAXQ, nq - q0);
    double* p_mu = par; double* p_li = par + (size_t)nb * d; double* p_ln = p_li + (size_t)nb * d * d;
    BOSIP_CUDA(ctx, cudaMemcpyAsync(p_mu, mus + (size_t)q0 * d, (size_t)nb * d * 8, cudaMemcpyHostToDevice, st));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(p_li, Linvs + (size_t)q0 * d * d, (size_t)nb * d * d * 8, cudaMemcpyHostToDevice, st));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(p_ln, lognorm + q0, (size_t)nb * 8, cudaMemcpyHostToDevice, st));
    MixParams mp{d, nb, p_mu, p_li, p_ln};
    const size_t smem = (size_t)nb * (per + 1) * 8;
    if (smem > 48 * 1024) BOSIP_CUDA(ctx, cudaFuncSetAttribute(mix_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mix_pdf_kernel<<<(unsigned)((m + SMP_BLOCK - 1) / SMP_BLOCK), SMP_BLOCK, smem, st>>>(mp, Xd, (long long)m, (accumulate || q0 > 0) ? 1 : 0, dd);
    BOSIP_CUDA(ctx, cudaGetLastError());
    BOSIP_CUDA(ctx, cudaStreamSynchronize(st));  // the parameter staging area is reused by the next batch
    ctx->launches += 1;
  }
  if (mem == BOSIP_B200_HOST) {
    BOSIP_CUDA(ctx, cudaMemcpyAsync(delta, dd, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  }
  return 0;
}

int amis_log_weights(Ctx* ctx, const double* logp, const double* delta, double t, int64_t m, int mem, double* logw) {
  if (m < 0 || !(t > 0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "amis_log_weights: bad arguments");
  if (m == 0) return 0;
  cudaStream_t st = ctx->s_comp;
  const bool host = mem == BOSIP_B200_HOST;
  BOSIP_TRY(ensure_ws(ctx, host ? 3 * round_up((size_t)m * 8, 256) : 0));
  const size_t stride = round_up((size_t)m * 8, 256);
  const double *pd, *dd; double* wd = logw;
  BOSIP_TRY(stage_in(ctx, logp, (size_t)m, mem, reinterpret_cast<double*>(ctx->ws), &pd, st));
  BOSIP_TRY(stage_in(ctx, delta, (size_t)m, mem, reinterpret_cast<double*>(ctx->ws + stride), &dd, st));
  if (host) wd = reinterpret_cast<double*>(ctx->ws + 2 * stride);
  amis_logw_kernel<<<(unsigned)((m + SMP_BLOCK - 1) / SMP_BLOCK), SMP_BLOCK, 0, st>>>(pd, dd, t, (long long)m, wd);
  BOSIP_CUDA(ctx, cudaGetLastError());
  if (host) BOSIP_CUDA(ctx, cudaMemcpyAsync(logw, wd, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->launches += 1;
  return 0;
}

// exp_weights + AnalyticalFitter moments.  out: mean[d], cov[d*d] (col-major, symmetric), sums[2] = {sum w, sum w^2} with
// w = exp(lw - max) (unnormalised), ws (optional, follows `mem`): normalised weights.
int weighted_moments(Ctx* ctx, int d, const double* X, const double* lw, int64_t m, int mem, double* mean, double* cov, double* sums,
                     double* ws) {
  if (d < 1 || d > MAX_D || m < 1) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "weighted_moments: bad dimensions");
  cudaStream_t st = ctx->s_comp;
  const bool host = mem == BOSIP_B200_HOST;
  const int nblocks = (int)std::min<int64_t>(4 * (int64_t)ctx->sms, (m + SMP_BLOCK - 1) / SMP_BLOCK);
  const int npair = d * (d + 1) / 2, nq0 = 2 + d;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_part = carve((size_t)std::max(npair, nq0) * nblocks * 8), o_res = carve((size_t)(4 + nq0 + npair + d) * 8), o_flag = carve(16),
               o_x = carve(host ? (size_t)m * d * 8 : 0), o_lw = carve(host ? (size_t)m * 8 : 0), o_ws = carve(host && ws ? (size_t)m * 8 : 0);
  BOSIP_TRY(ensure_ws(ctx, off));
  double* part = reinterpret_cast<double*>(ctx->ws + o_part);
  double* res = reinterpret_cast<double*>(ctx->ws + o_res);  // [0] max, [1..] pass-0 sums (nq0), then mean (d), then pair sums
  int* flags = reinterpret_cast<int*>(ctx->ws + o_flag);
  const double *Xd, *lwd;
  BOSIP_TRY(stage_in(ctx, X, (size_t)m * d, mem, reinterpret_cast<double*>(ctx->ws + o_x), &Xd, st));
  BOSIP_TRY(stage_in(ctx, lw, (size_t)m, mem, reinterpret_cast<double*>(ctx->ws + o_lw), &lwd, st));
  BOSIP_CUDA(ctx, cudaMemsetAsync(flags, 0, 16, st));
  logw_max_kernel<<<nblocks, SMP_BLOCK, 0, st>>>(lwd, (long long)m, part, flags);
  max_final_kernel<<<1, 1, 0, st>>>(nblocks, part, res);
  int hflag = 0; double hmax = 0.0;
  BOSIP_CUDA(ctx, cudaMemcpyAsync(&hflag, flags, 4, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(&hmax, res, 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  // amis_method.jl:112-114,128: NaN / +Inf weights or all weights zero are an error, not a fallback
  if (hflag || hmax == -INFINITY) BOSIP_FAIL(ctx, BOSIP_B200_EDOMAIN, "AMIS: Numerical issues with sample weights.");
  double* s0 = res + 1; double* mean_d = s0 + nq0; double* pairs = mean_d + d;
  moment_part_kernel<<<dim3(nblocks, nq0), SMP_BLOCK, 0, st>>>(d, 0, Xd, lwd, (long long)m, res, nullptr, part);
  moment_final_kernel<<<1, 64, 0, st>>>(nq0, nblocks, part, s0);
  std::vector<double> h0(nq0);
  BOSIP_CUDA(ctx, cudaMemcpyAsync(h0.data(), s0, (size_t)nq0 * 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  const double V1 = h0[0], V2 = h0[1];
  for (int a = 0; a < d; ++a) mean[a] = h0[2 + a] / V1;
  BOSIP_CUDA(ctx, cudaMemcpyAsync(mean_d, mean, (size_t)d * 8, cudaMemcpyHostToDevice, st));
  moment_part_kernel<<<dim3(nblocks, npair), SMP_BLOCK, 0, st>>>(d, 1, Xd, lwd, (long long)m, res, mean_d, part);
  moment_final_kernel<<<(npair + 63) / 64, 64, 0, st>>>(npair, nblocks, part, pairs);
  std::vector<double> hp(npair);
  BOSIP_CUDA(ctx, cudaMemcpyAsync(hp.data(), pairs, (size_t)npair * 8, cudaMemcpyDeviceToHost, st));
  if (ws) {
    double* wd = host ? reinterpret_cast<double*>(ctx->ws + o_ws) : ws;
    normalise_w_kernel<<<(unsigned)((m + SMP_BLOCK - 1) / SMP_BLOCK), SMP_BLOCK, 0, st>>>(lwd, (long long)m, res, s0, wd);
    if (host) BOSIP_CUDA(ctx, cudaMemcpyAsync(ws, wd, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    ctx->launches += 1;
  }
  BOSIP_CUDA(ctx, cudaGetLastError());
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  // unbiased "reliability weights" estimator (normal.jl:78-90): Sigma = sum w dd' / (V1 - V2 / V1)
  const double den = V1 - V2 / V1;
  int q = 0;
  for (int a = 0; a < d; ++a)
    for (int b = 0; b <= a; ++b, ++q) cov[a + (size_t)b * d] = cov[b + (size_t)a * d] = hp[q] / den;
  if (sums) { sums[0] = V1; sums[1] = V2; }
  ctx->launches += 6;
  return 0;
}

// resample(xs, ws, count) (src/sampling.jl:55-57): count draws with replacement, probability proportional to ws
int resample(Ctx* ctx, int d, const double* X, const double* ws, int64_t m, uint64_t seed, int64_t count, int mem, double* out, int64_t* idx) {
  if (d < 1 || d > MAX_D || m < 1 || count < 0) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "resample: bad dimensions");
  if (count == 0) return 0;
  cudaStream_t st = ctx->s_comp;
  const bool host = mem == BOSIP_B200_HOST;
  const long long nb = (m + SMP_BLOCK - 1) / SMP_BLOCK;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return o; };
  const size_t o_cdf = carve((size_t)m * 8), o_tot = carve((size_t)(nb + 1) * 8), o_x = carve(host ? (size_t)m * d * 8 : 0), o_w = carve(host ? (size_t)m * 8 : 0),
               o_out = carve(host ? (size_t)count * d * 8 : 0), o_idx = carve(host && idx ? (size_t)count * 8 : 0);
  BOSIP_TRY(ensure_ws(ctx, off));
  double* cdf = reinterpret_cast<double*>(ctx->ws + o_cdf); double* tot = reinterpret_cast<double*>(ctx->ws + o_tot);
  const double *Xd, *wd;
  BOSIP_TRY(stage_in(ctx, X, (size_t)m * d, mem, reinterpret_cast<double*>(ctx->ws + o_x), &Xd, st));
  BOSIP_TRY(stage_in(ctx, ws, (size_t)m, mem, reinterpret_cast<double*>(ctx->ws + o_w), &wd, st));
  scan_block_kernel<<<(unsigned)nb, SMP_BLOCK, 0, st>>>(wd, (long long)m, cdf, tot);
  scan_totals_kernel<<<1, 1, 0, st>>>(nb, tot);
  scan_add_kernel<<<(unsigned)nb, SMP_BLOCK, 0, st>>>((long long)m, tot, cdf);
  double* od = host ? reinterpret_cast<double*>(ctx->ws + o_out) : out;
  long long* id = idx ? (host ? reinterpret_cast<long long*>(ctx->ws + o_idx) : reinterpret_cast<long long*>(idx)) : nullptr;
  resample_kernel<<<(unsigned)((count + SMP_BLOCK - 1) / SMP_BLOCK), SMP_BLOCK, 0, st>>>(d, Xd, cdf, (long long)m, tot + nb, seed, (long long)count, od, id);
  BOSIP_CUDA(ctx, cudaGetLastError());
  if (host) {
    BOSIP_CUDA(ctx, cudaMemcpyAsync(out, od, (size_t)count * d * 8, cudaMemcpyDeviceToHost, st));
    if (idx) BOSIP_CUDA(ctx, cudaMemcpyAsync(idx, id, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
  }
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->launches += 4;
  return 0;
}

}  // namespace bosip
