// Confidence-set utilities on the device:
//   find_cutoff          /root/reference/src/utils/confidence_set.jl:19-27  -> weighted quantile (StatsBase.quantile(v, weights(w), p))
//   approx_cutoff_area   /root/reference/src/utils/confidence_set.jl:47-56
//   set_iou              /root/reference/src/utils/confidence_set.jl:73-81
//   AEConfidence.calculate  /root/reference/src/term_conds/ae_confidence.jl:59-79 (the three above fused on two closure outputs)
//
// The weighted quantile follows StatsBase's algorithm for non-frequency weights (an un-vendored dependency of the reference;
// restated from its published source): drop zero weights, sort the values (weights carried along), h = p (W - w_1) + w_1 on the
// cumulative weights S, k = first index with S_k > h, linear interpolation between v_{k-1} and v_k.  On the device:
//   1. values -> order-preserving 64-bit keys; a bitonic network sorts (key, original index) pairs, which makes the sort stable
//      and the result independent of the launch geometry (shared-memory stages for strides < 2048, one launch per wider stride);
//   2. the weights are gathered in sorted order and prefix-summed by a fixed three-kernel scan (block scan, scan of the block
//      totals, add-back), so S is reproducible bit for bit;
//   3. one thread binary-searches S for h and interpolates.
// Sizes are the termination condition's 1e4 .. 1e6 samples: everything is latency-bound, nothing here is on the hot path's roofline.
#include <math.h>
#include <vector>

#include "common.cuh"

namespace bosip {

constexpr int CS_CHUNK = 2048;       // elements sorted / merged per block in shared memory
constexpr int CS_THREADS = 1024;     // one compare-exchange per thread per stage
constexpr int CS_SCAN = 1024;        // elements per block of the prefix sum

struct KeyIdx { unsigned long long key; unsigned int idx; };

__device__ __forceinline__ unsigned long long f64_key(double v) {   // monotone map double -> uint64 (-0.0 < +0.0, NaN above +Inf)
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ bool ki_less(unsigned long long ka, unsigned int ia, unsigned long long kb, unsigned int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// keys[i] = key(vals[i]) for i < n (exp applied first when log_input), padding = all ones; nan_flag set when a NaN is met
__global__ void cs_keys_kernel(const double* __restrict__ vals, long long n, long long n_pad, int log_input, unsigned long long* __restrict__ keys,
                               unsigned int* __restrict__ idx, double* __restrict__ vals_out, int* __restrict__ nan_flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  if (i < n) {
    double v = vals[i];
    if (log_input) v = exp(v);
    if (vals_out) vals_out[i] = v;
    if (v != v) *nan_flag = 1;
    keys[i] = f64_key(v); idx[i] = (unsigned int)i;
  } else {
    keys[i] = ~0ull; idx[i] = 0xFFFFFFFFu;
  }
}

__device__ __forceinline__ void cs_cmpx(unsigned long long* k, unsigned int* x, int a, int b, bool up) {
  const bool sw = ki_less(k[b], x[b], k[a], x[a]) == up;
  if (sw) { const unsigned long long tk = k[a]; k[a] = k[b]; k[b] = tk; const unsigned int tx = x[a]; x[a] = x[b]; x[b] = tx; }
}

// full bitonic sort of every CS_CHUNK-element block in shared memory (direction alternates per block so that blocks pair up)
__global__ void __launch_bounds__(CS_THREADS) cs_sort_local(unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx) {
  __shared__ unsigned long long sk[CS_CHUNK];
  __shared__ unsigned int sx[CS_CHUNK];
  const long long base = (long long)blockIdx.x * CS_CHUNK;
  for (int e = threadIdx.x; e < CS_CHUNK; e += CS_THREADS) { sk[e] = keys[base + e]; sx[e] = idx[base + e]; }
  __syncthreads();
  for (int k = 2; k <= CS_CHUNK; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int t = threadIdx.x;
      const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;
      const bool up = (((base + a) & k) == 0);
      cs_cmpx(sk, sx, a, b, up);
      __syncthreads();
    }
  for (int e = threadIdx.x; e < CS_CHUNK; e += CS_THREADS) { keys[base + e] = sk[e]; idx[base + e] = sx[e]; }
}
// one global compare-exchange stage (stride j >= CS_CHUNK) of the merge with block size k
__global__ void cs_merge_global(unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx, long long n_pad, long long k, long long j) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pad / 2) return;
  const long long a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;
  const bool up = ((a & k) == 0);
  const unsigned long long ka = keys[a], kb = keys[b];
  const unsigned int ia = idx[a], ib = idx[b];
  if (ki_less(kb, ib, ka, ia) == up) { keys[a] = kb; keys[b] = ka; idx[a] = ib; idx[b] = ia; }
}
// strides CS_CHUNK/2 .. 1 of the merge with block size k, in shared memory
__global__ void __launch_bounds__(CS_THREADS) cs_merge_local(unsigned long long* __restrict__ keys, unsigned int* __restrict__ idx, long long k) {
  __shared__ unsigned long long sk[CS_CHUNK];
  __shared__ unsigned int sx[CS_CHUNK];
  const long long base = (long long)blockIdx.x * CS_CHUNK;
  for (int e = threadIdx.x; e < CS_CHUNK; e += CS_THREADS) { sk[e] = keys[base + e]; sx[e] = idx[base + e]; }
  __syncthreads();
  const bool up = ((base & k) == 0);   // k >= 2 * CS_CHUNK: the whole block lies in one direction
  for (int j = CS_CHUNK >> 1; j > 0; j >>= 1) {
    const int t = threadIdx.x;
    const int a = ((t & ~(j - 1)) << 1) | (t & (j - 1)), b = a | j;
    cs_cmpx(sk, sx, a, b, up);
    __syncthreads();
  }
  for (int e = threadIdx.x; e < CS_CHUNK; e += CS_THREADS) { keys[base + e] = sk[e]; idx[base + e] = sx[e]; }
}

// sorted weights: ws[idx[i]] (1 when ws == NULL; exp(log(v) - logprior) in the AEConfidence form), 0 for padding
__global__ void cs_gather_w_kernel(const unsigned int* __restrict__ idx, const double* __restrict__ ws, const double* __restrict__ vals,
                                   const double* __restrict__ logprior, long long n_pad, double* __restrict__ w_sorted) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const unsigned int s = idx[i];
  double w = 0.0;
  if (s != 0xFFFFFFFFu) w = logprior ? exp(log(vals[s]) - logprior[s]) : (ws ? ws[s] : 1.0);   // ae_confidence.jl:70-71
  w_sorted[i] = w;
}

// inclusive prefix sum, fixed order: (1) per-block Hillis-Steele scan + block total, (2) scan of the totals by one block,
// (3) add the preceding blocks' total
__global__ void __launch_bounds__(CS_SCAN) cs_scan_block(double* __restrict__ x, long long n, double* __restrict__ totals) {
  __shared__ double sh[2][CS_SCAN];
  const long long i = (long long)blockIdx.x * CS_SCAN + threadIdx.x;
  int cur = 0;
  sh[0][threadIdx.x] = i < n ? x[i] : 0.0;
  __syncthreads();
  for (int o = 1; o < CS_SCAN; o <<= 1) {
    sh[cur ^ 1][threadIdx.x] = threadIdx.x >= o ? sh[cur][threadIdx.x - o] + sh[cur][threadIdx.x] : sh[cur][threadIdx.x];
    cur ^= 1;
    __syncthreads();
  }
  if (i < n) x[i] = sh[cur][threadIdx.x];
  if (threadIdx.x == CS_SCAN - 1) totals[blockIdx.x] = sh[cur][threadIdx.x];
}
__global__ void cs_scan_totals(double* __restrict__ totals, int nblocks) {   // one thread: nblocks <= a few thousand
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) { const double t = totals[b]; totals[b] = s; s += t; }
}
__global__ void __launch_bounds__(CS_SCAN) cs_scan_add(double* __restrict__ x, long long n, const double* __restrict__ totals) {
  const long long i = (long long)blockIdx.x * CS_SCAN + threadIdx.x;
  if (i < n && blockIdx.x > 0) x[i] = totals[blockIdx.x] + x[i];
}

// StatsBase.quantile(v, w, p) on the sorted arrays: out[0] = quantile.  S = inclusive cumulative weights (non-decreasing).
__global__ void cs_quantile_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ idx,
                                   const double* __restrict__ vals, const double* __restrict__ w_sorted, const double* __restrict__ S,
                                   long long n, double p, const int* __restrict__ nan_flag, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  (void)keys;
  if (*nan_flag) { out[0] = NAN; return; }
  const double wsum = S[n - 1];
  long long first = 0;
  while (first < n && w_sorted[first] == 0.0) ++first;   // zero weights are dropped before sorting
  if (first == n) { out[0] = NAN; return; }                  // no weight at all (StatsBase throws; the caller never gets here with pdf values)
  const double w1 = w_sorted[first];
  const double h = p * (wsum - w1) + w1;
  // first k with S_k > h (the reference loops `while Sk <= h`)
  long long lo = first, hi = n;                              // invariant: S[lo-1] <= h (or lo == first), answer in [lo, hi]
  while (lo < hi) { const long long mid = (lo + hi) >> 1; if (S[mid] > h) hi = mid; else lo = mid + 1; }
  long long k = lo;
  if (k >= n) {                                              // h >= total: the last kept value
    long long last = n - 1;
    while (last > first && w_sorted[last] == 0.0) --last;
    out[0] = vals[idx[last]];
    return;
  }
  while (k < n && w_sorted[k] == 0.0) ++k;                  // S jumps at a kept element; be safe against -0 weights
  long long kp = k - 1;
  while (kp >= first && w_sorted[kp] == 0.0) --kp;          // previous kept element
  const double Sk = S[k], Skold = kp >= first ? S[kp] : 0.0;
  const double vk = vals[idx[k]], vkold = kp >= first ? vals[idx[kp]] : 0.0;
  out[0] = vkold + (h - Skold) / (Sk - Skold) * (vk - vkold);
}

// fixed-order block sums of up to 4 masked weight series: part[block][c]
template <int NC>
__device__ __forceinline__ void cs_block_sum(double (&s)[NC], double* __restrict__ part) {
  __shared__ double sh[NC][256];
#pragma unroll
  for (int c = 0; c < NC; ++c) sh[c][threadIdx.x] = s[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int c = 0; c < NC; ++c) sh[c][threadIdx.x] += sh[c][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) part[(size_t)blockIdx.x * NC + c] = sh[c][0];
  }
}

// approx_cutoff_area: part = {sum ws, sum ws log-normalised later}: pass 1 sums ws; pass 2 sums exp(log w - log W)[vals >= c]
__global__ void __launch_bounds__(256) cs_area_kernel(const double* __restrict__ vals, const double* __restrict__ ws, long long n, double c,
                                                      const double* __restrict__ total, double* __restrict__ part) {
  double s[1] = {0.0};
  const double logW = total ? log(total[0]) : 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double w = ws ? ws[i] : 1.0;
    if (!total) s[0] += w;
    else if (vals[i] >= c) s[0] += exp(log(w) - logW);      // confidence_set.jl:53-54
  }
  cs_block_sum<1>(s, part);
}

// set_iou: ws = 1 / pdf(prior), normalised; part = {sum ws, sum ws[A & B], sum ws[A | B]} (normalisation applied per element in pass 2)
__global__ void __launch_bounds__(256) cs_iou_kernel(const double* __restrict__ va, const double* __restrict__ ca, const double* __restrict__ vb,
                                                     const double* __restrict__ cb, const unsigned char* __restrict__ in_a,
                                                     const unsigned char* __restrict__ in_b, const double* __restrict__ logprior, long long n,
                                                     const double* __restrict__ total, double* __restrict__ part) {
  double s[2] = {0.0, 0.0};
  const double ta = ca ? ca[0] : 0.0, tb = cb ? cb[0] : 0.0;
  const double W = total ? total[0] : 1.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double w = 1.0 / exp(logprior[i]);                 // 1 ./ pdf.(Ref(x_prior), eachcol(xs))
    if (!total) { s[0] += w; continue; }
    const bool a = in_a ? in_a[i] != 0 : va[i] > ta;         // ae_confidence.jl:76-77: strict >
    const bool b = in_b ? in_b[i] != 0 : vb[i] > tb;
    const double wn = w / W;
    if (a && b) s[0] += wn;
    if (a || b) s[1] += wn;
  }
  cs_block_sum<2>(s, part);
}
__global__ void cs_part_sum(const double* __restrict__ part, int nblocks, int nc, double* __restrict__ out) {
  if (threadIdx.x >= nc || blockIdx.x != 0) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += part[(size_t)b * nc + threadIdx.x];
  out[threadIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------ host drivers
struct CsBufs {
  long long n, n_pad;
  unsigned long long* keys; unsigned int* idx; double *vals, *w, *totals, *scal; int* nan_flag;
  const double *d_in[4];
};

static long long cs_pad(long long n) { long long p = CS_CHUNK; while (p < n) p <<= 1; return p; }

// sort + weights + scan + quantile for values `vals_dev` (device, n); result (device scalar) at out_dev
static int cs_quantile_device(Ctx* ctx, const double* vals_dev, int log_input, const double* ws_dev, const double* logprior_dev, long long n,
                              double p, unsigned long long* keys, unsigned int* idx, double* vals_lin, double* w_sorted, double* totals,
                              int* nan_flag, double* out_dev, cudaStream_t st) {
  const long long n_pad = cs_pad(n);
  BOSIP_CUDA(ctx, cudaMemsetAsync(nan_flag, 0, sizeof(int), st));
  cs_keys_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(vals_dev, n, n_pad, log_input, keys, idx, vals_lin, nan_flag);
  cs_sort_local<<<(unsigned)(n_pad / CS_CHUNK), CS_THREADS, 0, st>>>(keys, idx);
  for (long long k = 2 * CS_CHUNK; k <= n_pad; k <<= 1) {
    for (long long j = k >> 1; j >= CS_CHUNK; j >>= 1)
      cs_merge_global<<<(unsigned)((n_pad / 2 + 255) / 256), 256, 0, st>>>(keys, idx, n_pad, k, j);
    cs_merge_local<<<(unsigned)(n_pad / CS_CHUNK), CS_THREADS, 0, st>>>(keys, idx, k);
  }
  cs_gather_w_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(idx, ws_dev, vals_lin, logprior_dev, n_pad, w_sorted);
  // scan over the n real entries (padding sorts last and carries zero weight); keep an unscanned copy of the weights
  double* S = w_sorted + n_pad;
  BOSIP_CUDA(ctx, cudaMemcpyAsync(S, w_sorted, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
  const int nb = (int)((n + CS_SCAN - 1) / CS_SCAN);
  cs_scan_block<<<nb, CS_SCAN, 0, st>>>(S, n, totals);
  cs_scan_totals<<<1, 1, 0, st>>>(totals, nb);
  cs_scan_add<<<nb, CS_SCAN, 0, st>>>(S, n, totals);
  cs_quantile_kernel<<<1, 32, 0, st>>>(keys, idx, vals_lin, w_sorted, S, n, p, nan_flag, out_dev);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

struct CsWs {   // workspace carve-up for one quantile problem of n values
  size_t keys, idx, vals, w, totals, flag, bytes;
};
static CsWs cs_layout(long long n, size_t base) {
  const long long n_pad = cs_pad(n);
  CsWs L; size_t off = base;
  auto carve = [&](size_t b) { size_t o = off; off += (b + 255) / 256 * 256; return o; };
  L.keys = carve((size_t)n_pad * 8); L.idx = carve((size_t)n_pad * 4); L.vals = carve((size_t)n * 8);
  L.w = carve((size_t)2 * n_pad * 8); L.totals = carve(((size_t)(n + CS_SCAN - 1) / CS_SCAN + 1) * 8); L.flag = carve(8);
  L.bytes = off;
  return L;
}

// upload helper: returns the device view of a length-n double array given in `mem`
static int cs_stage(Ctx* ctx, const double* src, long long n, int mem, size_t* off, const double** dev) {
  if (!src) { *dev = nullptr; return 0; }
  if (mem == BOSIP_B200_DEVICE) { *dev = src; return 0; }
  double* d = reinterpret_cast<double*>(ctx->ws + *off);
  *off += ((size_t)n * 8 + 255) / 256 * 256;
  BOSIP_CUDA(ctx, cudaMemcpyAsync(d, src, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->s_comp));
  *dev = d;
  return 0;
}

int cs_weighted_quantile(Ctx* ctx, const double* vals, const double* ws, int64_t n, int mem, double p, double* out) {
  if (!vals || !out || n < 1 || n >= (1ll << 31)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "weighted_quantile: need 1 <= n < 2^31 values");
  if (!(p >= 0.0 && p <= 1.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "weighted_quantile: p must be in [0, 1]");
  const size_t stage = mem == BOSIP_B200_HOST ? 2 * (((size_t)n * 8 + 255) / 256 * 256) : 0;
  const CsWs L = cs_layout(n, stage + 256);
  BOSIP_TRY(ensure_ws(ctx, L.bytes));
  size_t off = 256;
  const double *dv, *dw;
  BOSIP_TRY(cs_stage(ctx, vals, n, mem, &off, &dv));
  BOSIP_TRY(cs_stage(ctx, ws, n, mem, &off, &dw));
  char* w = ctx->ws;
  double* res = reinterpret_cast<double*>(w);
  BOSIP_TRY(cs_quantile_device(ctx, dv, 0, dw, nullptr, n, p, reinterpret_cast<unsigned long long*>(w + L.keys), reinterpret_cast<unsigned int*>(w + L.idx),
                               reinterpret_cast<double*>(w + L.vals), reinterpret_cast<double*>(w + L.w), reinterpret_cast<double*>(w + L.totals),
                               reinterpret_cast<int*>(w + L.flag), res, ctx->s_comp));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(out, res, 8, cudaMemcpyDeviceToHost, ctx->s_comp));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(ctx->s_comp));
  ctx->launches += 8;
  return 0;
}

int cs_cutoff_area(Ctx* ctx, const double* vals, const double* ws, int64_t n, int mem, double c, double* out) {
  if (!vals || !out || n < 1) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "cutoff_area: bad arguments");
  const int nblocks = (int)std::min<int64_t>((n + 2047) / 2048, 1024);
  const size_t stage = mem == BOSIP_B200_HOST ? 2 * (((size_t)n * 8 + 255) / 256 * 256) : 0;
  BOSIP_TRY(ensure_ws(ctx, 256 + stage + (size_t)nblocks * 8 + 256));
  size_t off = 256;
  const double *dv, *dw;
  BOSIP_TRY(cs_stage(ctx, vals, n, mem, &off, &dv));
  BOSIP_TRY(cs_stage(ctx, ws, n, mem, &off, &dw));
  double* scal = reinterpret_cast<double*>(ctx->ws);           // [0] total weight, [1] area
  double* part = reinterpret_cast<double*>(ctx->ws + off);
  cudaStream_t st = ctx->s_comp;
  cs_area_kernel<<<nblocks, 256, 0, st>>>(dv, dw, n, c, nullptr, part);
  cs_part_sum<<<1, 32, 0, st>>>(part, nblocks, 1, scal);
  cs_area_kernel<<<nblocks, 256, 0, st>>>(dv, dw, n, c, scal, part);
  cs_part_sum<<<1, 32, 0, st>>>(part, nblocks, 1, scal + 1);
  BOSIP_CUDA(ctx, cudaGetLastError());
  BOSIP_CUDA(ctx, cudaMemcpyAsync(out, scal + 1, 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->launches += 4;
  return 0;
}

// IoU on device-resident inputs; thresholds ca/cb are device scalars (or NULL with explicit masks)
static int cs_iou_device(Ctx* ctx, const double* va, const double* ca, const double* vb, const double* cb, const unsigned char* ma,
                         const unsigned char* mb, const double* lp, long long n, double* part, int nblocks, double* scal3, cudaStream_t st) {
  cs_iou_kernel<<<nblocks, 256, 0, st>>>(va, ca, vb, cb, ma, mb, lp, n, nullptr, part);
  cs_part_sum<<<1, 32, 0, st>>>(part, nblocks, 2, scal3);        // scal3[0] = sum ws
  cs_iou_kernel<<<nblocks, 256, 0, st>>>(va, ca, vb, cb, ma, mb, lp, n, scal3, part);
  cs_part_sum<<<1, 32, 0, st>>>(part, nblocks, 2, scal3 + 1);    // scal3[1] = intersection, scal3[2] = union
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

int cs_set_iou(Ctx* ctx, const unsigned char* in_a, const unsigned char* in_b, const double* logprior, int64_t n, int mem, double* out) {
  if (!in_a || !in_b || !logprior || !out || n < 1) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "set_iou: bad arguments");
  const int nblocks = (int)std::min<int64_t>((n + 2047) / 2048, 1024);
  const size_t nb8 = ((size_t)n * 8 + 255) / 256 * 256, nb1 = ((size_t)n + 255) / 256 * 256;
  BOSIP_TRY(ensure_ws(ctx, 256 + (mem == BOSIP_B200_HOST ? nb8 + 2 * nb1 : 0) + (size_t)nblocks * 16 + 256));
  cudaStream_t st = ctx->s_comp;
  size_t off = 256;
  const double* dlp;
  BOSIP_TRY(cs_stage(ctx, logprior, n, mem, &off, &dlp));
  const unsigned char *da = in_a, *db = in_b;
  if (mem == BOSIP_B200_HOST) {
    unsigned char* a = reinterpret_cast<unsigned char*>(ctx->ws + off); off += nb1;
    unsigned char* b = reinterpret_cast<unsigned char*>(ctx->ws + off); off += nb1;
    BOSIP_CUDA(ctx, cudaMemcpyAsync(a, in_a, (size_t)n, cudaMemcpyHostToDevice, st));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(b, in_b, (size_t)n, cudaMemcpyHostToDevice, st));
    da = a; db = b;
  }
  double* scal = reinterpret_cast<double*>(ctx->ws);
  BOSIP_TRY(cs_iou_device(ctx, nullptr, nullptr, nullptr, nullptr, da, db, dlp, n, reinterpret_cast<double*>(ctx->ws + off), nblocks, scal, st));
  double h[3];
  BOSIP_CUDA(ctx, cudaMemcpyAsync(h, scal, 24, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->launches += 4;
  *out = h[1] / h[2];
  return 0;
}

// AEConfidence.calculate (ae_confidence.jl:59-79) on the log-values of the two closures: everything stays on the device,
// three doubles come back.
int cs_confidence_iou(Ctx* ctx, const double* log_fa, const double* log_fb, const double* logprior, int64_t n, int mem, double q, double* iou,
                      double* c_a, double* c_b) {
  if (!log_fa || !log_fb || !logprior || !iou || n < 1 || n >= (1ll << 31)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "confidence_iou: bad arguments");
  if (!(q >= 0.0 && q <= 1.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "confidence_iou: q must be in [0, 1]");
  const size_t nb8 = ((size_t)n * 8 + 255) / 256 * 256;
  const size_t stage = mem == BOSIP_B200_HOST ? 3 * nb8 : 0;
  const CsWs L = cs_layout(n, 256 + stage + nb8);              // + a second linear-value array (closure b)
  const int nblocks = (int)std::min<int64_t>((n + 2047) / 2048, 1024);
  BOSIP_TRY(ensure_ws(ctx, L.bytes + (size_t)nblocks * 16 + 256));
  cudaStream_t st = ctx->s_comp;
  size_t off = 256;
  const double *dfa, *dfb, *dlp;
  BOSIP_TRY(cs_stage(ctx, log_fa, n, mem, &off, &dfa));
  BOSIP_TRY(cs_stage(ctx, log_fb, n, mem, &off, &dfb));
  BOSIP_TRY(cs_stage(ctx, logprior, n, mem, &off, &dlp));
  char* w = ctx->ws;
  double* vals_b = reinterpret_cast<double*>(w + off);          // linear values of closure b (closure a's live in L.vals until b's sort)
  double* scal = reinterpret_cast<double*>(w);                  // [0] c_a, [1] c_b, [2..4] iou sums
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(w + L.keys);
  unsigned int* idx = reinterpret_cast<unsigned int*>(w + L.idx);
  double* vals_a = reinterpret_cast<double*>(w + L.vals);
  double* wbuf = reinterpret_cast<double*>(w + L.w);
  double* totals = reinterpret_cast<double*>(w + L.totals);
  int* flag = reinterpret_cast<int*>(w + L.flag);
  // find_cutoff(f, xs, ws, q) = quantile(vals, weights(ws), 1 - q) with ws = exp(log f - logpdf(prior))  (ae_confidence.jl:70-74)
  BOSIP_TRY(cs_quantile_device(ctx, dfb, 1, nullptr, dlp, n, 1.0 - q, keys, idx, vals_b, wbuf, totals, flag, scal + 1, st));
  BOSIP_TRY(cs_quantile_device(ctx, dfa, 1, nullptr, dlp, n, 1.0 - q, keys, idx, vals_a, wbuf, totals, flag, scal + 0, st));
  BOSIP_TRY(cs_iou_device(ctx, vals_a, scal + 0, vals_b, scal + 1, nullptr, nullptr, dlp, n, reinterpret_cast<double*>(w + L.bytes), nblocks, scal + 2, st));
  double h[5];
  BOSIP_CUDA(ctx, cudaMemcpyAsync(h, scal, 40, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->launches += 20;
  if (c_a) *c_a = h[0];
  if (c_b) *c_b = h[1];
  *iou = h[3] / h[4];
  return 0;
}

}  // namespace bosip
