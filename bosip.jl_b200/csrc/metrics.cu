// Metrics of /root/reference/src/metrics: the MMD pairwise-kernel sums (mmd.jl:23-28) and the TV distance on a weighted
// grid (tv.jl:59-73 with logsumexp/logmeanexp of src/utils/utils.jl:7-15).
#include <math.h>
#include <vector>

#include "common.cuh"
#include "fastmath.cuh"

namespace bosip {

// ================================================================================================ MMD
// sum_{i,j} k(P_i, Q_j) for a 128-row tile x a strip of column tiles.  Thread = row point (coordinates in registers),
// column tile staged in shared memory (SoA) and read as broadcasts.  In symmetric mode (P == Q) only column tiles
// >= row tile are visited, off-diagonal tiles weighted 2: the result is still the full V-statistic incl. diagonal.
//
// The kernel is FP64-pipe bound, so the instruction count per pair is what matters.  Points are stored centred and scaled
// with their squared norm as an extra SoA row; when every norm is small (MMD_EXPAND_LIMIT) the squared distance is formed
// as |p|^2 + |q|^2 - 2 p.q (d + 1 instructions, absolute error <= ~(d + 2) 2^-53 max|p|^2, i.e. <= 1e-13 relative in k)
// instead of sum (p_k - q_k)^2 (2 d instructions).  Data that is spread over many length-scales takes the difference form;
// the choice is made on the device from the maximum norm the scaling pass recorded, identically on every rank.
constexpr int MMD_TILE = 128;
constexpr int MMD_STRIP = 16;            // column tiles per CTA
constexpr int MMD_EXP_REP = 1;           // copies of the exp table in shared memory (fastmath.cuh: replication buys nothing)
constexpr double MMD_EXPAND_LIMIT = 128.0;  // largest scaled squared norm for which the expanded distance is used

template <int KERNEL, int DP, bool EXPAND>
__device__ __forceinline__ double mmd_tile_sum(const double (&pi)[DP], double pn, const double (*sq)[MMD_TILE], int jmax,
                                               const double* exp_tab) {
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  int jj = 0;
  for (; jj + 4 <= jmax; jj += 4) {
    double r2[4];
    if (EXPAND) {
      const double2 n01 = *reinterpret_cast<const double2*>(&sq[DP][jj]);
      const double2 n23 = *reinterpret_cast<const double2*>(&sq[DP][jj + 2]);
      r2[0] = pn + n01.x; r2[1] = pn + n01.y; r2[2] = pn + n23.x; r2[3] = pn + n23.y;
    } else {
      r2[0] = r2[1] = r2[2] = r2[3] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      const double2 q01 = *reinterpret_cast<const double2*>(&sq[k][jj]);
      const double2 q23 = *reinterpret_cast<const double2*>(&sq[k][jj + 2]);
      if (EXPAND) {   // pi holds -2 p
        r2[0] = fma(pi[k], q01.x, r2[0]); r2[1] = fma(pi[k], q01.y, r2[1]); r2[2] = fma(pi[k], q23.x, r2[2]); r2[3] = fma(pi[k], q23.y, r2[3]);
      } else {
        const double t0 = pi[k] - q01.x, t1 = pi[k] - q01.y, t2 = pi[k] - q23.x, t3 = pi[k] - q23.y;
        r2[0] = fma(t0, t0, r2[0]); r2[1] = fma(t1, t1, r2[1]); r2[2] = fma(t2, t2, r2[2]); r2[3] = fma(t3, t3, r2[3]);
      }
    }
    // expanded form: r2 <= 4 * MMD_EXPAND_LIMIT, so exp cannot underflow (no cut-off); rounding can leave r2 a few ulps of
    // the norms below 0, which the square root of the Matern kernels must not see
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[e] += kernel_from_r2_t<KERNEL, MMD_EXP_REP, !EXPAND>((EXPAND && KERNEL != 0) ? fabs(r2[e]) : r2[e], exp_tab);
  }
  for (; jj < jmax; ++jj) {
    double r2 = EXPAND ? pn + sq[DP][jj] : 0.0;
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      if (EXPAND) r2 = fma(pi[k], sq[k][jj], r2);
      else { const double t = pi[k] - sq[k][jj]; r2 = fma(t, t, r2); }
    }
    acc[0] += kernel_from_r2_t<KERNEL, MMD_EXP_REP, !EXPAND>((EXPAND && KERNEL != 0) ? fabs(r2) : r2, exp_tab);
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}

// Ps/Qs: centred, scaled points, SoA [dp + 1][np] / [dp + 1][nq] (row dp = squared norm; zero padded to multiples of 128
// in count; padded points are masked).  norm_max_bits: bit pattern of the largest squared norm of either set.
template <int KERNEL, int DP>
__global__ void __launch_bounds__(MMD_TILE)
mmd_kernel(const double* __restrict__ Ps, long long np, long long np_pad, const double* __restrict__ Qs, long long nq,
           long long nq_pad, int symmetric, int rank, int world, const unsigned long long* __restrict__ norm_max_bits,
           double* __restrict__ partial) {
  __shared__ __align__(16) double sq[DP + 1][MMD_TILE];
  __shared__ double red[MMD_TILE];
  __shared__ double exp_tab_all[EXP_TAB_N * MMD_EXP_REP];
  load_exp_table<MMD_EXP_REP>(exp_tab_all);
  const double* exp_tab = exp_table_lane<MMD_EXP_REP>(exp_tab_all);
  const long long cta = (long long)blockIdx.y * gridDim.x + blockIdx.x;
  const int rt = blockIdx.x;                    // row tile
  const int ct0 = blockIdx.y * MMD_STRIP;       // first column tile of the strip
  const int n_ct = (int)(nq_pad / MMD_TILE);
  double total = 0.0;
  const bool mine = (cta % world) == rank;
  const unsigned long long nmax_bits = __ldg(norm_max_bits);
  if (mine && !(symmetric && ct0 + MMD_STRIP - 1 < rt)) {
    const bool expand = nmax_bits <= (unsigned long long)__double_as_longlong(MMD_EXPAND_LIMIT);
    const long long i = (long long)rt * MMD_TILE + threadIdx.x;
    double pi[DP];
#pragma unroll
    for (int k = 0; k < DP; ++k) pi[k] = Ps[(size_t)k * np_pad + i] * (expand ? -2.0 : 1.0);
    const double pn = Ps[(size_t)DP * np_pad + i];
    const bool row_ok = i < np;
    for (int ct = ct0; ct < min(ct0 + MMD_STRIP, n_ct); ++ct) {
      if (symmetric && ct < rt) continue;
      __syncthreads();
#pragma unroll
      for (int k = 0; k <= DP; ++k) sq[k][threadIdx.x] = Qs[(size_t)k * nq_pad + (size_t)ct * MMD_TILE + threadIdx.x];
      __syncthreads();
      const int jmax = (int)min((long long)MMD_TILE, nq - (long long)ct * MMD_TILE);
      const double s = expand ? mmd_tile_sum<KERNEL, DP, true>(pi, pn, sq, jmax, exp_tab)
                              : mmd_tile_sum<KERNEL, DP, false>(pi, pn, sq, jmax, exp_tab);
      const double w = (symmetric && ct != rt) ? 2.0 : 1.0;
      if (row_ok) total += w * s;
    }
  }
  red[threadIdx.x] = total;
  __syncthreads();
  for (int o = MMD_TILE / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[cta] = nmax_bits > 0x7ff0000000000000ull ? NAN : red[0];
}

// fixed-order tree sum of `n` partials into out[0] (one block)
__global__ void __launch_bounds__(1024) tree_sum_kernel(const double* __restrict__ partial, long long n, double* out) {
  __shared__ double red[1024];
  double s = 0.0;
  for (long long k = threadIdx.x; k < n; k += 1024) s += partial[k];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0];
}

// centre[k] = mean of coordinate k over the first min(n, 4096) points of P (one block; any fixed point near the data would do:
// the kernels are stationary, the centre only keeps the norms small)
__global__ void __launch_bounds__(256) mmd_centre_kernel(const double* __restrict__ P, long long n, int d, double* __restrict__ centre) {
  __shared__ double red[256];
  const int cnt = (int)min(n, 4096LL);
  for (int k = 0; k < d; ++k) {
    double s = 0.0;
    for (int i = threadIdx.x; i < cnt; i += 256) s += P[(size_t)k + (size_t)d * i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) { const double c = red[0] / cnt; centre[k] = isfinite(c) ? c : 0.0; }
    __syncthreads();
  }
}

// AoS d x n (column-major) -> centred, scaled SoA [dp + 1][n_pad] with the squared norm in row dp; records the largest norm
__global__ void __launch_bounds__(256)
mmd_scale_kernel(const double* __restrict__ P, long long n, long long n_pad, int d, int dp, const double* __restrict__ scale,
                 const double* __restrict__ centre, double* __restrict__ Ps, unsigned long long* __restrict__ norm_max_bits) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double nrm = 0.0;
  if (i < n_pad) {
    for (int k = 0; k < dp; ++k) {
      const double v = (i < n && k < d) ? (P[(size_t)k + (size_t)d * i] - centre[k]) * scale[k] : 0.0;
      Ps[(size_t)k * n_pad + i] = v;
      nrm = fma(v, v, nrm);
    }
    Ps[(size_t)dp * n_pad + i] = nrm;
  }
  // positive doubles order like their bit patterns; +Inf marks "too wide for the expanded form", a NaN pattern (above +Inf)
  // marks NaN coordinates, which poison the result as they do in the reference
  unsigned long long b = (unsigned long long)__double_as_longlong(nrm);
  if (!(nrm <= MMD_EXPAND_LIMIT) && nrm == nrm) b = 0x7ff0000000000000ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const unsigned long long y = __shfl_xor_sync(0xffffffffu, b, o); b = y > b ? y : b; }
  if ((threadIdx.x & 31) == 0 && b) atomicMax(norm_max_bits, b);
}

template <int KERNEL, int DP>
static void mmd_launch(const double* Ps, long long np, long long npp, const double* Qs, long long nq, long long nqp, int sym,
                       int rank, int world, const unsigned long long* nmax, double* partial, dim3 grid, cudaStream_t st) {
  mmd_kernel<KERNEL, DP><<<grid, MMD_TILE, 0, st>>>(Ps, np, npp, Qs, nq, nqp, sym, rank, world, nmax, partial);
}

template <int KERNEL>
static int mmd_launch_k(Ctx* ctx, int dp, const double* Ps, long long np, long long npp, const double* Qs, long long nq,
                        long long nqp, int sym, int rank, int world, const unsigned long long* nmax, double* partial, dim3 grid,
                        cudaStream_t st) {
#define BOSIP_DP_CASE(DPV) \
  case DPV: mmd_launch<KERNEL, DPV>(Ps, np, npp, Qs, nq, nqp, sym, rank, world, nmax, partial, grid, st); break;
  switch (dp) {
    BOSIP_DP_CASE(2) BOSIP_DP_CASE(3) BOSIP_DP_CASE(4) BOSIP_DP_CASE(5) BOSIP_DP_CASE(6) BOSIP_DP_CASE(8)
    BOSIP_DP_CASE(10) BOSIP_DP_CASE(12) BOSIP_DP_CASE(16) BOSIP_DP_CASE(24) BOSIP_DP_CASE(32)
    default: BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unsupported dimension");
  }
#undef BOSIP_DP_CASE
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

static int mmd_pair(Ctx* ctx, int kernel_id, int dp, const double* Ps, long long np, long long npp, const double* Qs,
                    long long nq, long long nqp, int sym, int rank, int world, const unsigned long long* nmax, double* partial,
                    double* out, cudaStream_t st) {
  dim3 grid((unsigned)(npp / MMD_TILE), (unsigned)((nqp / MMD_TILE + MMD_STRIP - 1) / MMD_STRIP));
  const long long nparts = (long long)grid.x * grid.y;
  switch (kernel_id) {
    case 0: BOSIP_TRY((mmd_launch_k<0>(ctx, dp, Ps, np, npp, Qs, nq, nqp, sym, rank, world, nmax, partial, grid, st))); break;
    case 1: BOSIP_TRY((mmd_launch_k<1>(ctx, dp, Ps, np, npp, Qs, nq, nqp, sym, rank, world, nmax, partial, grid, st))); break;
    case 2: BOSIP_TRY((mmd_launch_k<2>(ctx, dp, Ps, np, npp, Qs, nq, nqp, sym, rank, world, nmax, partial, grid, st))); break;
    default: BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown kernel id");
  }
  tree_sum_kernel<<<1, 1024, 0, st>>>(partial, nparts, out);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

int mmd_sums(Ctx* ctx, int kernel_id, int d, const double* lambda, const double* A, int64_t n, const double* B,
             int64_t m, int mem, int rank, int world, double* sums) {
  if (d < 1 || d > MAX_D || n < 1 || m < 1 || !A || !B || !lambda || !sums) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad MMD arguments");
  if (world < 1 || rank < 0 || rank >= world) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad rank/world");
  const int dp = pad_dim(d);
  const long long np = round_up(n, MMD_TILE), mp = round_up(m, MMD_TILE);
  cudaStream_t st = ctx->s_comp;
  const double ck = kernel_id == BOSIP_B200_KERNEL_SE ? sqrt(0.5) : kernel_id == BOSIP_B200_KERNEL_MATERN32 ? sqrt(3.0) : sqrt(5.0);
  std::vector<double> h_scale(dp, 0.0);
  for (int k = 0; k < d; ++k) {
    if (!(lambda[k] > 0.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "length-scales must be > 0");
    h_scale[k] = ck / lambda[k];
  }
  const long long max_tiles = (np > mp ? np : mp) / MMD_TILE;
  const long long max_parts = max_tiles * ((max_tiles + MMD_STRIP - 1) / MMD_STRIP);
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_As = carve((size_t)(dp + 1) * np * 8), o_Bs = carve((size_t)(dp + 1) * mp * 8), o_scale = carve((size_t)dp * 8),
               o_centre = carve((size_t)dp * 8), o_nmax = carve(8), o_part = carve((size_t)max_parts * 8), o_out = carve(3 * 8);
  const size_t o_dA = mem == BOSIP_B200_HOST ? carve((size_t)d * n * 8) : 0, o_dB = mem == BOSIP_B200_HOST ? carve((size_t)d * m * 8) : 0;
  BOSIP_TRY(ensure_ws(ctx, off));
  auto D = [&](size_t o) { return reinterpret_cast<double*>(ctx->ws + o); };
  double *As = D(o_As), *Bs = D(o_Bs), *dscale = D(o_scale), *partial = D(o_part), *dout = D(o_out);
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dscale, h_scale.data(), dp * 8, cudaMemcpyHostToDevice, st));
  const double *pA = A, *pB = B;
  if (mem == BOSIP_B200_HOST) {
    BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_dA), A, (size_t)d * n * 8, cudaMemcpyHostToDevice, st));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(D(o_dB), B, (size_t)d * m * 8, cudaMemcpyHostToDevice, st));
    pA = D(o_dA); pB = D(o_dB);
  }
  double* dcentre = D(o_centre);
  unsigned long long* nmax = reinterpret_cast<unsigned long long*>(ctx->ws + o_nmax);
  BOSIP_CUDA(ctx, cudaMemsetAsync(nmax, 0, 8, st));
  mmd_centre_kernel<<<1, 256, 0, st>>>(pA, n, d, dcentre);
  mmd_scale_kernel<<<(unsigned)((np + 255) / 256), 256, 0, st>>>(pA, n, np, d, dp, dscale, dcentre, As, nmax);
  mmd_scale_kernel<<<(unsigned)((mp + 255) / 256), 256, 0, st>>>(pB, m, mp, d, dp, dscale, dcentre, Bs, nmax);
  BOSIP_CUDA(ctx, cudaGetLastError());
  BOSIP_TRY(mmd_pair(ctx, kernel_id, dp, As, n, np, As, n, np, 1, rank, world, nmax, partial, dout + 0, st));
  BOSIP_TRY(mmd_pair(ctx, kernel_id, dp, Bs, m, mp, Bs, m, mp, 1, rank, world, nmax, partial, dout + 1, st));
  BOSIP_TRY(mmd_pair(ctx, kernel_id, dp, As, n, np, Bs, m, mp, 0, rank, world, nmax, partial, dout + 2, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(sums, dout, 3 * 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}

// ================================================================================================ TV
// Online log-sum-exp pair (max, sum exp(. - max)); -Inf-safe merge in a fixed order.
struct Lse { double m, s; };
__host__ __device__ inline Lse lse_merge(Lse a, Lse b) {
  if (b.m == -INFINITY) return a;
  if (a.m == -INFINITY) return b;
  Lse r;
  if (a.m >= b.m) { r.m = a.m; r.s = (a.m == INFINITY) ? a.s : a.s + b.s * exp(b.m - a.m); }
  else { r.m = b.m; r.s = (b.m == INFINITY) ? b.s : b.s + a.s * exp(a.m - b.m); }
  return r;
}
// Scaled-sum accumulator: value = exp(m) * s.  Batches of TV_B elements are folded with ONE rescale and TV_B independent
// branch-free exps (fastmath.cuh), instead of a serial, branchy online update per element.
constexpr int TV_B = 4;
__device__ __forceinline__ void lse_push_batch(Lse& acc, const double (&v)[TV_B], const double* exp_tab) {
  double vm = v[0];
#pragma unroll
  for (int e = 1; e < TV_B; ++e) vm = fmax(vm, v[e]);
  if (vm == -INFINITY) return;                        // nothing to add
  if (vm == INFINITY) { acc.m = INFINITY; acc.s = 1.0; return; }
  const double nm = fmax(acc.m, vm);
  double s = (acc.m == -INFINITY) ? 0.0 : acc.s * fast_exp_neg_t<1>(nm - acc.m, exp_tab);
#pragma unroll
  for (int e = 0; e < TV_B; ++e) s += fast_exp_neg_t<1>(nm - v[e], exp_tab);   // v = -Inf -> exp(-Inf) = 0
  acc.m = nm; acc.s = s;
}
__device__ __forceinline__ Lse lse_block(Lse x, Lse* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Lse y; y.m = __shfl_down_sync(0xffffffffu, x.m, o); y.s = __shfl_down_sync(0xffffffffu, x.s, o);
    x = lse_merge(x, y);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = x;
  __syncthreads();
  if (w == 0) {
    Lse y; y.m = -INFINITY; y.s = 0.0;
    if (l < (int)(blockDim.x >> 5)) y = sh[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Lse z; z.m = __shfl_down_sync(0xffffffffu, y.m, o); z.s = __shfl_down_sync(0xffffffffu, y.s, o);
      y = lse_merge(y, z);
    }
    x = y;
  }
  return x;
}

// Both TV kernels are two passes over three vectors that do not fit the 126 MB L2 at the benchmark size (2.4 GB at G = 1e8):
// HBM-bound.  Every thread loads 4 consecutive doubles per vector with ONE 32-byte `ld.global.nc.v4.f64` (L1 bypassed: the data is
// read once per pass) and keeps two such groups per vector in flight; a block owns a contiguous slice whose start is a multiple
// of 1024 elements, so the vector loads are aligned whenever the base pointers are (VEC = false is the scalar path for slices
// that start at an odd element, e.g. a rank's shard of a host vector).  Stage 2 walks its slice backwards: the tail of every
// slice is what stage 1 left in L2.
__device__ __forceinline__ void ld4(const double* __restrict__ p, double (&v)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];\n" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
template <bool VEC>
__device__ __forceinline__ void tv_load4(const double* __restrict__ p, long long i0, long long hi, double fill, double (&v)[4]) {
  if (VEC && i0 + 4 <= hi) { ld4(p + i0, v); return; }
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = (i0 + e < hi) ? __ldg(p + i0 + e) : fill;
}

// out[block] = {max_a, sum_a, max_t, sum_t}: log-sum-exp pairs of lw+a and lw+t over the block's contiguous slice
template <bool VEC>
__global__ void __launch_bounds__(256)
tv_stage1_kernel(const double* __restrict__ lw, const double* __restrict__ a, const double* __restrict__ t, long long G, long long per,
                 double* __restrict__ out) {
  __shared__ Lse sh[8];
  __shared__ double exp_tab[EXP_TAB_N];
  load_exp_table<1>(exp_tab);
  __syncthreads();
  Lse la{-INFINITY, 0.0}, lt{-INFINITY, 0.0};
  const long long lo = (long long)blockIdx.x * per, hi = min(G, lo + per);
  for (long long i0 = lo + (long long)threadIdx.x * TV_B; i0 < hi; i0 += 2ll * blockDim.x * TV_B) {
    const long long i1 = i0 + (long long)blockDim.x * TV_B;
    double w0[4], a0[4], t0[4], w1[4], a1[4], t1[4];
    tv_load4<VEC>(lw, i0, hi, 0.0, w0); tv_load4<VEC>(a, i0, hi, -INFINITY, a0); tv_load4<VEC>(t, i0, hi, -INFINITY, t0);
    tv_load4<VEC>(lw, i1, hi, 0.0, w1); tv_load4<VEC>(a, i1, hi, -INFINITY, a1); tv_load4<VEC>(t, i1, hi, -INFINITY, t1);
    double va[TV_B], vt[TV_B];
#pragma unroll
    for (int e = 0; e < TV_B; ++e) { va[e] = w0[e] + a0[e]; vt[e] = w0[e] + t0[e]; }
    lse_push_batch(la, va, exp_tab);
    lse_push_batch(lt, vt, exp_tab);
#pragma unroll
    for (int e = 0; e < TV_B; ++e) { va[e] = w1[e] + a1[e]; vt[e] = w1[e] + t1[e]; }
    lse_push_batch(la, va, exp_tab);
    lse_push_batch(lt, vt, exp_tab);
  }
  la = lse_block(la, sh);
  lt = lse_block(lt, sh);
  if (threadIdx.x == 0) { out[4 * blockIdx.x + 0] = la.m; out[4 * blockIdx.x + 1] = la.s; out[4 * blockIdx.x + 2] = lt.m; out[4 * blockIdx.x + 3] = lt.s; }
}

// one batch of 4 grid points of stage 2: |e^(ua - nm) - e^(ut - nm)| added at the running scale nm
__device__ __forceinline__ void tv_push_diff(Lse& ld, const double (&ua)[TV_B], const double (&ut)[TV_B], const double* exp_tab) {
  double vm = -INFINITY;
#pragma unroll
  for (int e = 0; e < TV_B; ++e) vm = fmax(vm, fmax(ua[e], ut[e]));
  if (vm == -INFINITY) return;
  if (vm == INFINITY) { ld.m = INFINITY; ld.s = 1.0; return; }
  const double nm = fmax(ld.m, vm);
  double s = (ld.m == -INFINITY) ? 0.0 : ld.s * fast_exp_neg_t<1>(nm - ld.m, exp_tab);
#pragma unroll
  for (int e = 0; e < TV_B; ++e) s += fabs(fast_exp_neg_t<1>(nm - ua[e], exp_tab) - fast_exp_neg_t<1>(nm - ut[e], exp_tab));
  ld.m = nm; ld.s = s;
}

// out[block] = {m, s} with exp(m) * s = sum_i exp(lw_i) |exp(a_i - eva) - exp(t_i - evt)|  (tv.jl:64-70).  The log of the
// reference's `log.(diffs)` is never taken: each term is formed directly at the running scale m >= lw + max(a~, t~).
template <bool VEC>
__global__ void __launch_bounds__(256)
tv_stage2_kernel(const double* __restrict__ lw, const double* __restrict__ a, const double* __restrict__ t, long long G, long long per,
                 double eva_v, double evt_v, const double* __restrict__ ev_dev, double* __restrict__ out) {
  __shared__ Lse sh[8];
  __shared__ double exp_tab[EXP_TAB_N];
  load_exp_table<1>(exp_tab);
  __syncthreads();
  // the normalisers come by value (staged calls, combined across ranks on the host) or from the merge kernel of stage 1
  const double eva = ev_dev ? ev_dev[0] : eva_v, evt = ev_dev ? ev_dev[1] : evt_v;
  const bool a_inf = isinf(eva), t_inf = isinf(evt);   // all-(-Inf) vector: its normalised density is all -Inf (tv.jl:64-66)
  Lse ld{-INFINITY, 0.0};
  const long long lo = (long long)blockIdx.x * per, hi = min(G, lo + per);
  const long long stride = 2ll * blockDim.x * TV_B;
  const long long n_it = (hi - lo + stride - 1) / stride;
  for (long long it = n_it - 1; it >= 0; --it) {       // backwards: the end of the slice is what stage 1 touched last
    const long long i0 = lo + it * stride + (long long)threadIdx.x * TV_B, i1 = i0 + (long long)blockDim.x * TV_B;
    double w0[4], a0[4], t0[4], w1[4], a1[4], t1[4];
    tv_load4<VEC>(lw, i0, hi, -INFINITY, w0); tv_load4<VEC>(a, i0, hi, 0.0, a0); tv_load4<VEC>(t, i0, hi, 0.0, t0);
    tv_load4<VEC>(lw, i1, hi, -INFINITY, w1); tv_load4<VEC>(a, i1, hi, 0.0, a1); tv_load4<VEC>(t, i1, hi, 0.0, t1);
    double ua[TV_B], ut[TV_B];
#pragma unroll
    for (int e = 0; e < TV_B; ++e) {
      ua[e] = a_inf ? -INFINITY : w1[e] + (a1[e] - eva);   // log of the weighted normalised approx density
      ut[e] = t_inf ? -INFINITY : w1[e] + (t1[e] - evt);
    }
    tv_push_diff(ld, ua, ut, exp_tab);
#pragma unroll
    for (int e = 0; e < TV_B; ++e) {
      ua[e] = a_inf ? -INFINITY : w0[e] + (a0[e] - eva);
      ut[e] = t_inf ? -INFINITY : w0[e] + (t0[e] - evt);
    }
    tv_push_diff(ld, ua, ut, exp_tab);
  }
  ld = lse_block(ld, sh);
  if (threadIdx.x == 0) { out[2 * blockIdx.x + 0] = ld.m; out[2 * blockIdx.x + 1] = ld.s; }
}

static int tv_blocks(Ctx* ctx, int64_t G) {   // 80 registers x 256 threads: 3 resident CTAs per SM -> two full waves at most
  return (int)((G + 2047) / 2048 < (int64_t)ctx->sms * 6 ? (G + 2047) / 2048 : (int64_t)ctx->sms * 6);
}
// elements per block: a multiple of 1024 so that every thread's groups of 4 stay 32-byte aligned relative to the base pointers
static long long tv_per(int64_t G, int nblocks) { return ((G + nblocks - 1) / nblocks + 1023) / 1024 * 1024; }
static bool tv_aligned(const double* lw, const double* a, const double* t) {
  return ((reinterpret_cast<uintptr_t>(lw) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(t)) & 31u) == 0;
}
static void tv_launch1(const double* lw, const double* a, const double* t, int64_t G, int nblocks, double* dpart, cudaStream_t st) {
  const long long per = tv_per(G, nblocks);
  if (tv_aligned(lw, a, t)) tv_stage1_kernel<true><<<nblocks, 256, 0, st>>>(lw, a, t, (long long)G, per, dpart);
  else tv_stage1_kernel<false><<<nblocks, 256, 0, st>>>(lw, a, t, (long long)G, per, dpart);
}
static void tv_launch2(const double* lw, const double* a, const double* t, int64_t G, int nblocks, double eva, double evt, const double* ev_dev,
                       double* dpart, cudaStream_t st) {
  const long long per = tv_per(G, nblocks);
  if (tv_aligned(lw, a, t)) tv_stage2_kernel<true><<<nblocks, 256, 0, st>>>(lw, a, t, (long long)G, per, eva, evt, ev_dev, dpart);
  else tv_stage2_kernel<false><<<nblocks, 256, 0, st>>>(lw, a, t, (long long)G, per, eva, evt, ev_dev, dpart);
}
static size_t tv_part_bytes(int nblocks) { return ((size_t)nblocks * 4 * 8 + 3 * 8 + 255) / 256 * 256; }

// one stage on device-resident vectors; dpart: nblocks * 4 doubles of scratch; out: the merged (max, sum) pairs
static int tv_run(Ctx* ctx, const double* pl, const double* pa, const double* pt, int64_t G, int stage, double eva, double evt,
                  double* dpart, int nblocks, double* out) {
  cudaStream_t st = ctx->s_comp;
  const int width = stage == 1 ? 4 : 2;
  if (stage == 1) tv_launch1(pl, pa, pt, G, nblocks, dpart, st);
  else tv_launch2(pl, pa, pt, G, nblocks, eva, evt, nullptr, dpart, st);
  BOSIP_CUDA(ctx, cudaGetLastError());
  std::vector<double> h((size_t)nblocks * width);
  BOSIP_CUDA(ctx, cudaMemcpyAsync(h.data(), dpart, h.size() * 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  for (int c = 0; c < width / 2; ++c) {
    Lse acc{-INFINITY, 0.0};
    for (int b = 0; b < nblocks; ++b) acc = lse_merge(acc, Lse{h[(size_t)b * width + 2 * c], h[(size_t)b * width + 2 * c + 1]});
    out[2 * c] = acc.m; out[2 * c + 1] = acc.s;
  }
  return 0;
}

// workspace = [partials][lw][a][t]; host vectors are uploaded ONCE (both stages of tv_full read the same device copies)
static int tv_stage_inputs(Ctx* ctx, const double*& lw, const double*& a, const double*& t, int64_t G, int mem, int nblocks, double** dpart) {
  if (G < 1 || !lw || !a || !t) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad TV arguments");
  const size_t part_bytes = tv_part_bytes(nblocks);
  const size_t vec_bytes = ((size_t)G * 8 + 255) / 256 * 256;
  BOSIP_TRY(ensure_ws(ctx, part_bytes + (mem == BOSIP_B200_HOST ? 3 * vec_bytes : 0)));
  *dpart = reinterpret_cast<double*>(ctx->ws);
  if (mem == BOSIP_B200_HOST) {
    cudaStream_t st = ctx->s_comp;
    double* dl = reinterpret_cast<double*>(ctx->ws + part_bytes);
    double* da = reinterpret_cast<double*>(ctx->ws + part_bytes + vec_bytes);
    double* dt = reinterpret_cast<double*>(ctx->ws + part_bytes + 2 * vec_bytes);
    BOSIP_CUDA(ctx, cudaMemcpyAsync(dl, lw, (size_t)G * 8, cudaMemcpyHostToDevice, st));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(da, a, (size_t)G * 8, cudaMemcpyHostToDevice, st));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(dt, t, (size_t)G * 8, cudaMemcpyHostToDevice, st));
    lw = dl; a = da; t = dt;
  }
  return 0;
}

int tv_stage1(Ctx* ctx, const double* lw, const double* a, const double* t, int64_t G, int mem, double* out4) {
  if (!out4) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad TV arguments");
  const int nblocks = tv_blocks(ctx, G);
  double* dpart = nullptr;
  BOSIP_TRY(tv_stage_inputs(ctx, lw, a, t, G, mem, nblocks, &dpart));
  return tv_run(ctx, lw, a, t, G, 1, 0.0, 0.0, dpart, nblocks, out4);
}
int tv_stage2(Ctx* ctx, const double* lw, const double* a, const double* t, int64_t G, int mem, double eva, double evt, double* out2) {
  if (!out2) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "bad TV arguments");
  const int nblocks = tv_blocks(ctx, G);
  double* dpart = nullptr;
  BOSIP_TRY(tv_stage_inputs(ctx, lw, a, t, G, mem, nblocks, &dpart));
  return tv_run(ctx, lw, a, t, G, 2, eva, evt, dpart, nblocks, out2);
}

// Device-side merge of the per-block (max, sum) pairs in a fixed order (strided partial merges, then the block tree).
// stage 1: ev[0..1] = logmeanexp of lw + a and of lw + t;  stage 2: ev[2] = TV = exp(logmeanexp(...)) / 2.
// logmeanexp = logsumexp - log(G); logsumexp returns the max itself when it is infinite (src/utils/utils.jl:9)
__global__ void __launch_bounds__(256) tv_merge_kernel(const double* __restrict__ part, int nblocks, int stage, double log_g, double* __restrict__ ev) {
  __shared__ Lse sh[8];
  const int width = stage == 1 ? 4 : 2;
  for (int c = 0; c < width / 2; ++c) {
    Lse acc{-INFINITY, 0.0};
    for (int b = threadIdx.x; b < nblocks; b += 256) acc = lse_merge(acc, Lse{part[(size_t)b * width + 2 * c], part[(size_t)b * width + 2 * c + 1]});
    acc = lse_block(acc, sh);
    if (threadIdx.x == 0) {
      const double lme = isinf(acc.m) ? acc.m : acc.m + log(acc.s) - log_g;
      if (stage == 1) ev[c] = lme; else ev[2] = 0.5 * exp(lme);
    }
    __syncthreads();
  }
}

// calc_tv on one GPU (tv.jl:59-73): inputs uploaded once, both stages and their merges queued back to back, one 8-byte read-back
int tv_full(Ctx* ctx, const double* lw, const double* a, const double* t, int64_t G, int mem, double* tv_out) {
  const int nblocks = tv_blocks(ctx, G);
  double* dpart = nullptr;
  BOSIP_TRY(tv_stage_inputs(ctx, lw, a, t, G, mem, nblocks, &dpart));
  cudaStream_t st = ctx->s_comp;
  double* ev = dpart + (size_t)nblocks * 4;   // 3 doubles behind the partials (tv_part_bytes reserves them)
  const double lg = log((double)G);
  tv_launch1(lw, a, t, G, nblocks, dpart, st);
  tv_merge_kernel<<<1, 256, 0, st>>>(dpart, nblocks, 1, lg, ev);
  tv_launch2(lw, a, t, G, nblocks, 0.0, 0.0, ev, dpart, st);
  tv_merge_kernel<<<1, 256, 0, st>>>(dpart, nblocks, 2, lg, ev);
  BOSIP_CUDA(ctx, cudaGetLastError());
  BOSIP_CUDA(ctx, cudaMemcpyAsync(tv_out, ev + 2, 8, cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  return 0;
}

}  // namespace bosip
