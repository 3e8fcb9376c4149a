// GP conditioning on the device: replaces BOSS.model_posterior(bosip.problem)
// (/root/reference/src/posterior/log_posterior.jl:81,115,148): K_j = alpha_j^2 k(X/lambda_j) + sigma_j^2 I,
// L_j = chol(K_j), L_j^{-1}, a_j = K_j^{-1}(Y_j - m_j(X))   (SURVEY.md Appendix A.2).
//
// Hand-written blocked FP64 factorisation, all p outputs batched in every launch:
//   * right-looking Cholesky, NB = 64: diagonal block factorised + inverted by one CTA in shared memory,
//     panel  L21 = A21 inv(L11)^T  and trailing update  A22 -= L21 L21^T  by a strided batched DMMA GEMM;
//   * triangular inverse by recursive doubling: the 64x64 diagonal inverses come out of the Cholesky for free,
//     then level s = 64,128,...: X21 = -X22 (L21 X11), two batched GEMMs per level (ceil(log2(P/64)) levels).  Both products have
//     a triangular factor (X11 resp. X22), so every CTA walks only the non-zero half of its k-range; a ragged last pair
//     (P is NOT rounded to a power of two) is one more launch per level with a shorter X22;
//   * a = K^{-1} y by two blocked triangular SOLVES with L (forward, then backward substitution; the 64x64 diagonal blocks are
//     substituted exactly, the off-diagonal part is a GEMV) -- not through the explicit inverse;
//   * L^{-1} scaled by alpha^2 is repacked tile-major + XOR-swizzled for the variance GEMM's bulk copies.
// Matrices are padded to P = 64 * ceil(N/64) with an identity tail.
#include <math.h>
#include <stdlib.h>
#include <time.h>
#include <utility>
#include <vector>

#include "common.cuh"
#include "fastmath.cuh"

namespace bosip {

constexpr int NB = 64;
constexpr size_t POTRF_SMEM = NB * (NB + 1) * sizeof(double);

// ------------------------------------------------------------------------------------------------ generic GEMM
// C(m,n) = beta*C(m,n) + alpha * sum_k A(m,k) B(k,n), two-level batch (z = z1 + nb1*z2).  A and C are column-major (unit row
// stride, leading dimensions sak / scn); B is addressed by (sbk, sbn) with ONE of them equal to 1 (template BK: k is the unit
// direction).  M, N multiples of 64, K multiple of 16, every base pointer and leading dimension a multiple of 2 doubles.
//   lower_only : skip CTA tiles that lie entirely above the diagonal (SYRK-style trailing update);
//   b_lower    : B(k,n) = 0 for k < n  (lower-triangular right factor)  -> the k-loop starts at the tile's first column;
//   a_lower    : A(m,k) = 0 for k > m  (lower-triangular left factor)   -> the k-loop ends at the tile's last row.
// CTA tile 128 x 64, 8 warps x (32 x 32) of DMMA m8n8k4, k-slabs of 16 through a 3-stage cp.async ring (16-byte copies along the
// unit-stride direction).  Shared-memory rows are padded by 4 doubles: the 8-row x 4-column fragment loads of a half-warp then
// fall into 16 distinct bank pairs.
struct GemmArgs {
  const double* A; const double* B; double* C;
  long long sam, sak, sbk, sbn, scm, scn;
  long long bA1, bA2, bB1, bB2, bC1, bC2;
  int nb1, K, lower_only, a_lower, b_lower, M, N;
  double alpha, beta;
};

constexpr int GM = 128, GN = 64, GK = 16, GSTAGES = 3;
constexpr int SA_LD = GM + 4;          // sA[k][m]
constexpr int SB_LD_N = GN + 4;        // sB[k][n]  (B with unit n-stride)
constexpr int SB_LD_K = GK + 4;        // sB[n][k]  (B with unit k-stride)
constexpr int G_STAGE_DOUBLES = GK * SA_LD + GN * SB_LD_K;   // 2112 + 1280 (>= GK * SB_LD_N)
constexpr size_t GEMM_SMEM = (size_t)GSTAGES * G_STAGE_DOUBLES * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <bool BK>
__global__ void __launch_bounds__(256) gemm_dmma_kernel(const __grid_constant__ GemmArgs g) {
  const int m0 = blockIdx.x * GM, n0 = blockIdx.y * GN;
  if (g.lower_only && n0 >= m0 + GM) return;
  extern __shared__ __align__(16) unsigned char gemm_smem_raw[];
  double* smem = reinterpret_cast<double*>(gemm_smem_raw);
  const int z1 = blockIdx.z % g.nb1, z2 = blockIdx.z / g.nb1;
  const double* A = g.A + z1 * g.bA1 + z2 * g.bA2 + m0;
  const double* B = g.B + z1 * g.bB1 + z2 * g.bB2 + (long long)n0 * g.sbn;
  double* C = g.C + z1 * g.bC1 + z2 * g.bC2 + m0 + (long long)n0 * g.scn;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  const bool m_hi_valid = m0 + 64 < g.M;            // M is a multiple of 64: the upper half of the tile is all-or-nothing
  int k_lo = 0, k_hi = g.K;
  if (g.b_lower) k_lo = n0 & ~(GK - 1);
  if (g.a_lower) k_hi = min(g.K, m0 + (m_hi_valid ? GM : 64));
  const int nk = (k_hi - k_lo + GK - 1) / GK;

  auto load_slab = [&](int kt, int stage) {
    double* sA = smem + (size_t)stage * G_STAGE_DOUBLES;
    double* sB = sA + GK * SA_LD;
    const int k0 = k_lo + kt * GK;
#pragma unroll
    for (int it = 0; it < 4; ++it) {          // A: 128 m x 16 k = 1024 chunks of 2 doubles
      const int c = tid + it * 256, k = c >> 6, m2 = (c & 63) << 1;
      if (m2 < 64 || m_hi_valid) cp_async16(sA + k * SA_LD + m2, A + m2 + (long long)(k0 + k) * g.sak);
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) {          // B: 64 n x 16 k = 512 chunks
      const int c = tid + it * 256;
      if (BK) { const int n = c >> 3, k2 = (c & 7) << 1; cp_async16(sB + n * SB_LD_K + k2, B + (k0 + k2) + (long long)n * g.sbn); }
      else { const int k = c >> 5, n2 = (c & 31) << 1; cp_async16(sB + k * SB_LD_N + n2, B + (long long)(k0 + k) * g.sbk + n2); }
    }
  };

  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) { c[i][jn][0] = 0.0; c[i][jn][1] = 0.0; }

#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) { if (s < nk) load_slab(s, s); cp_async_commit(); }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();                             // slab kt has landed for every thread; slab kt-1's readers are done
    if (kt + GSTAGES - 1 < nk) load_slab(kt + GSTAGES - 1, (kt + GSTAGES - 1) % GSTAGES);
    cp_async_commit();
    const double* sA = smem + (size_t)(kt % GSTAGES) * G_STAGE_DOUBLES;
    const double* sB = sA + GK * SA_LD;
#pragma unroll
    for (int ks = 0; ks < GK; ks += 4) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[(ks + tig) * SA_LD + wm + 8 * i + gq];
#pragma unroll
      for (int jn = 0; jn < 4; ++jn) b[jn] = BK ? sB[(wn + 8 * jn + gq) * SB_LD_K + ks + tig] : sB[(ks + tig) * SB_LD_N + wn + 8 * jn + gq];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) dmma884(c[i][jn][0], c[i][jn][1], a[i], b[jn]);
    }
  }
  cp_async_wait<0>();
  if (wm >= 64 && !m_hi_valid) return;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jn = 0; jn < 4; ++jn)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long m = wm + 8 * i + gq, n = wn + 8 * jn + 2 * tig + h;
        double* dst = C + m + n * g.scn;
        const double v = g.alpha * c[i][jn][h];
        *dst = (g.beta == 0.0) ? v : fma(g.beta, *dst, v);
      }
}

static int gemm64(Ctx* ctx, GemmArgs g, int M, int N, int batch, cudaStream_t st) {
  if (M <= 0 || N <= 0 || batch <= 0) return 0;
  g.M = M; g.N = N;
  if (g.sam != 1 || g.scm != 1 || (g.sbk != 1 && g.sbn != 1)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "gemm: unsupported operand layout");
  dim3 grid((unsigned)((M + GM - 1) / GM), (unsigned)(N / GN), (unsigned)batch);
  if (g.sbk == 1) {   // the attribute is per device: set it on every launch (a host-side table write)
    BOSIP_CUDA(ctx, cudaFuncSetAttribute(gemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    gemm_dmma_kernel<true><<<grid, 256, GEMM_SMEM, st>>>(g);
  } else {
    BOSIP_CUDA(ctx, cudaFuncSetAttribute(gemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    gemm_dmma_kernel<false><<<grid, 256, GEMM_SMEM, st>>>(g);
  }
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ small kernels

// Us[j][k][n] = X[k + d n] * scale[j][k]
__global__ void scale_inputs_kernel(const double* __restrict__ X, const double* __restrict__ scale, int d, int dp, int N,
                                    int Ns, double* __restrict__ Us) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y, j = blockIdx.z;
  if (n >= Ns) return;
  Us[((size_t)j * dp + k) * Ns + n] = (n < N && k < d) ? X[(size_t)k + (size_t)d * n] * scale[j * dp + k] : 0.0;
}

// Kp[j][r + c P2] = alpha_j^2 k(U_r, U_c) + delta_rc (sigma_j^2 + jitter) for r,c < N; identity tail.
__global__ void build_K_kernel(int kernel_id, const double* __restrict__ Us, int dp, int N, int Ns, int P2,
                               const double* __restrict__ alpha2, const double* __restrict__ diag_add,
                               double* __restrict__ Kp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, j = blockIdx.z;
  if (r >= P2) return;
  double v;
  if (r < N && c < N) {
    const double* U = Us + (size_t)j * dp * Ns;
    double r2 = 0.0;
    for (int k = 0; k < dp; ++k) { const double t = U[(size_t)k * Ns + r] - U[(size_t)k * Ns + c]; r2 = fma(t, t, r2); }
    double kv = kernel_id == BOSIP_B200_KERNEL_SE ? kernel_from_r2<0>(r2) : kernel_id == BOSIP_B200_KERNEL_MATERN32 ? kernel_from_r2<1>(r2) : kernel_from_r2<2>(r2);
    v = alpha2[j] * kv + (r == c ? diag_add[j] : 0.0);
  } else {
    v = (r == c) ? 1.0 : 0.0;
  }
  Kp[(size_t)j * P2 * P2 + (size_t)r + (size_t)c * P2] = v;
}

// Upload path: Lp[j] (P2 x P2, identity tail) from a host-layout N x N col-major L (lower triangle read).
__global__ void pad_L_kernel(const double* __restrict__ Lin, int N, int P2, double* __restrict__ Lp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, j = blockIdx.z;
  if (r >= P2) return;
  double v = (r == c) ? 1.0 : 0.0;
  if (r < N && c < N) v = (c <= r) ? Lin[(size_t)j * N * N + (size_t)r + (size_t)c * N] : 0.0;
  Lp[(size_t)j * P2 * P2 + (size_t)r + (size_t)c * P2] = v;
}

// Factorise the 64x64 diagonal block b in place (lower), zero its strict upper part, and write its inverse into the
// same block of Xp.  One CTA of 256 threads per output.  Thread (tr, tq) = (tid & 63, tid >> 6) keeps the 16 elements
// (tr, tq + 4 i) of its row in REGISTERS for the whole sweep; what a column step needs from other threads -- the pivot column,
// resp. the finished row of the inverse -- goes through a double-buffered 64-entry line in shared memory, so a step is ONE barrier,
// 17 broadcast loads and up to 16 independent FMAs (the first version kept the block in shared memory with a read-modify-write loop
// of variable length and two barriers per column: ~1000 clk per column, 65 us per block -- 4.2 of the 8.9 ms of an N = 4096
// Cholesky; same operations in the same order, so L and L^-1 are bit-identical to it).  flag[0] = 1 + block index on a bad pivot.
__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ Ap, double* __restrict__ Xp, int P2, int b,
                                                         int do_factor, int* __restrict__ flag, const double* __restrict__ piv_floor) {
  extern __shared__ __align__(16) unsigned char potrf_smem[];
  double (*s)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potrf_smem);   // the finished factor (for the inverse sweep)
  __shared__ double line[2][NB];
  __shared__ double rdiag[NB];
  const int j = blockIdx.x, tid = threadIdx.x, tr = tid & 63, tq = tid >> 6;
  double* A = Ap + (size_t)j * P2 * P2 + (size_t)b * NB + (size_t)b * NB * P2;
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = A[tr + (size_t)(tq + 4 * i) * P2];
  bool ok = true;
  if (do_factor) {
    // "not positive definite" = a pivot at or below rounding level of the original diagonal K_cc = alpha^2 + sigma^2 + jitter
    // (LAPACK / Julia's cholesky test `<= 0`, which for an exactly singular K is decided by rounding noise; the floor makes the
    // PosDefException deterministic and never fires for a noise variance above 1e-14 of the prior variance)
    const double floor_j = piv_floor ? piv_floor[j] : 0.0;
    if (tq == 0) line[0][tr] = a[0];
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      const double* col = line[c & 1];            // column c as the previous steps left it: rows >= c are final
      const double piv = col[c];
      if (!(piv > floor_j)) { if (tid == 0) atomicCAS(flag, 0, 1 + b); ok = false; break; }  // uniform: every thread reads the same pivot
      // sqrt and reciprocal from ONE Goldschmidt iteration (hardware rsqrt seed, two coupled steps, residual correction: <= 2 ulp,
      // fastmath.cuh) instead of the library's sqrt + divide, which put ~800 dependent cycles into each of the 64 serial steps
      double yv;
      asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(yv) : "d"(piv));
      double gg = piv * yv, hh = 0.5 * yv;
      double rr_ = fma(-hh, gg, 0.5);
      gg = fma(gg, rr_, gg); hh = fma(hh, rr_, hh);
      rr_ = fma(-hh, gg, 0.5);
      gg = fma(gg, rr_, gg); hh = fma(hh, rr_, hh);
      const double lcc = fma(fma(-gg, gg, piv), hh, gg);          // sqrt(piv)
      const double rinv = 2.0 * fma(hh, fma(-2.0 * hh, lcc, 1.0), hh);   // 1 / lcc: one Newton step on 2 h
      const double l_rc = col[tr] * rinv;                           // L[tr][c] for tr > c
      if (tq == (c & 3)) a[c >> 2] = tr > c ? l_rc : (tr == c ? lcc : 0.0);
      // rank-1 update of the trailing lower triangle: this thread's columns cc = tq + 4 i with c < cc <= tr
#pragma unroll
      for (int i = (c >> 2); i < 16; ++i) {
        const int cc = tq + 4 * i;
        if (cc > c && cc <= tr) a[i] -= l_rc * (col[cc] * rinv);
      }
      if (c + 1 < NB) {
        if (tq == ((c + 1) & 3) && tr >= c + 1) line[(c + 1) & 1][tr] = a[(c + 1) >> 2];
        __syncthreads();
      }
    }
    if (ok) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { const int cc = tq + 4 * i; A[tr + (size_t)cc * P2] = cc <= tr ? a[i] : 0.0; }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 16; ++i) { const int cc = tq + 4 * i; s[tr][cc] = (cc == tr && !ok) ? 1.0 : a[i]; }
  __syncthreads();
  // inverse: solve L X = I as 64 rank-1 sweeps: X[k][:] = B[k][:] / L[k][k], then B[i][:] -= L[i][k] X[k][:] for the rows below
  // (only columns <= k of row k are non-zero); x[i] = X[tr][tq + 4 i] in registers, row k travels through `line`
  if (tid < NB) rdiag[tid] = 1.0 / s[tid][tid];     // the 64 reciprocals at once (exact divisions, off the serial path)
  double x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (tq + 4 * i == tr) ? 1.0 : 0.0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (tr == k) {
      const double rk = rdiag[k];
#pragma unroll
      for (int i = 0; i <= (k >> 2); ++i) { const int cc = tq + 4 * i; if (cc <= k) x[i] = x[i] * rk; line[k & 1][cc] = cc <= k ? x[i] : 0.0; }
    }
    __syncthreads();
    if (tr > k) {
      const double l_ik = s[tr][k];
#pragma unroll
      for (int i = 0; i <= (k >> 2); ++i) { const int cc = tq + 4 * i; if (cc <= k) x[i] -= l_ik * line[k & 1][cc]; }
    }
  }
  double* X = Xp + (size_t)j * P2 * P2 + (size_t)b * NB + (size_t)b * NB * P2;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const int cc = tq + 4 * i; X[tr + (size_t)cc * P2] = cc <= tr ? x[i] : 0.0; }
}

// a = K^{-1} y = L^{-T} (L^{-1} y) by two blocked triangular SOLVES with the Cholesky factor itself (not through the explicit
// inverse).  One launch per 64-row block and direction, all outputs batched: every CTA substitutes the 64 x 64 diagonal system
// exactly (warp 0: two rows per lane, the pivot row's value broadcast by shuffle, multiplied by the reciprocal diagonal, which is
// formed once per block by 64 parallel divisions -- redundant across CTAs, which
// costs nothing but keeps the step to ONE kernel), CTA 0 stores the block's solution, CTA c >= 1 applies the block's solution
// to its share of the remaining right-hand side (a GEMV slab with coalesced reads).  rhs and sol are [p][P], zero padded; the
// block's solution goes to a SEPARATE array: the other CTAs of the launch still read the block's right-hand side.
__device__ __forceinline__ void trsv_load_block(const double* __restrict__ L, int P, int r0, double (*D)[65]) {
  for (int e = threadIdx.x; e < 64 * 64; e += 256) { const int rr = e & 63, cc = e >> 6; D[rr][cc] = L[(size_t)(r0 + rr) + (size_t)(r0 + cc) * P]; }
}
__global__ void __launch_bounds__(256) trsv_fwd_step(const double* __restrict__ Lp, int P, int b, double* __restrict__ rhs,
                                                     double* __restrict__ sol) {
  __shared__ double D[64][65];
  __shared__ double w[64];
  const int j = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r0 = b * 64;
  const double* L = Lp + (size_t)j * P * P;
  double* y = rhs + (size_t)j * P;
  __shared__ double rd[64];
  trsv_load_block(L, P, r0, D);
  __syncthreads();
  if (tid < 64) rd[tid] = 1.0 / D[tid][tid];   // the 64 reciprocals at once: no division on the 64-step serial path
  __syncthreads();
  if (warp == 0) {
    double x0 = y[r0 + lane], x1 = y[r0 + lane + 32];
    for (int k = 0; k < 64; ++k) {
      const double rk = __shfl_sync(0xffffffffu, k < 32 ? x0 : x1, k & 31);
      const double wk = rk * rd[k];
      if (lane == (k & 31)) { if (k < 32) x0 = wk; else x1 = wk; }
      if (lane > k) x0 = fma(-D[lane][k], wk, x0);
      if (lane + 32 > k) x1 = fma(-D[lane + 32][k], wk, x1);
    }
    w[lane] = x0; w[lane + 32] = x1;
  }
  __syncthreads();
  if (blockIdx.x == 0) { if (tid < 64) sol[(size_t)j * P + r0 + tid] = w[tid]; return; }
  const int row = r0 + 64 + (blockIdx.x - 1) * 256 + tid;
  if (row >= P) return;
  const double* Lr = L + (size_t)row + (size_t)r0 * P;
  double acc = 0.0;
#pragma unroll 8
  for (int k = 0; k < 64; ++k) acc = fma(Lr[(size_t)k * P], w[k], acc);
  y[row] -= acc;
}
__global__ void __launch_bounds__(256) trsv_bwd_step(const double* __restrict__ Lp, int P, int b, double* __restrict__ rhs,
                                                     double* __restrict__ sol) {
  __shared__ double D[64][65];
  __shared__ double a[64];
  const int j = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, r0 = b * 64;
  const double* L = Lp + (size_t)j * P * P;
  double* w = rhs + (size_t)j * P;
  __shared__ double rd[64];
  trsv_load_block(L, P, r0, D);
  __syncthreads();
  if (tid < 64) rd[tid] = 1.0 / D[tid][tid];
  __syncthreads();
  if (warp == 0) {
    double x0 = w[r0 + lane], x1 = w[r0 + lane + 32];
    for (int k = 63; k >= 0; --k) {
      const double rk = __shfl_sync(0xffffffffu, k < 32 ? x0 : x1, k & 31);
      const double ak = rk * rd[k];
      if (lane == (k & 31)) { if (k < 32) x0 = ak; else x1 = ak; }
      if (lane < k) x0 = fma(-D[k][lane], ak, x0);          // column `lane` of row k of L^T = L[k][lane]
      if (lane + 32 < k) x1 = fma(-D[k][lane + 32], ak, x1);
    }
    a[lane] = x0; a[lane + 32] = x1;
  }
  __syncthreads();
  if (blockIdx.x == 0) { if (tid < 64) sol[(size_t)j * P + r0 + tid] = a[tid]; return; }
  // rows jj < r0 of the transposed system: w[jj] -= sum_k L[r0 + k][jj] a[k]; one warp per row, lanes along k (contiguous in memory)
  const double a0 = a[lane], a1 = a[lane + 32];
  for (int q = 0; q < 8; ++q) {
    const int jj = (blockIdx.x - 1) * 64 + warp * 8 + q;
    if (jj >= r0) break;
    const double* Lc = L + (size_t)r0 + (size_t)jj * P;
    double acc = fma(Lc[lane], a0, Lc[lane + 32] * a1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) w[jj] -= acc;
  }
}
// rhs[j][n] = y[j][n] for n < N, 0 for the padding rows
__global__ void pad_rhs_kernel(const double* __restrict__ y, int N, int P, double* __restrict__ rhs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (n < P) rhs[(size_t)j * P + n] = n < N ? y[(size_t)j * N + n] : 0.0;
}
__global__ void unpad_rhs_kernel(const double* __restrict__ rhs, int N, int P, double* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (n < N) out[(size_t)j * N + n] = rhs[(size_t)j * P + n];
}

// LinvT[j][t][kc][r][swz] = alpha_j^2 Linv[128 t + r][16 kc + kk]
__global__ void pack_linv_kernel(const double* __restrict__ Xp, int P2, int N, int T, int KC,
                                 const double* __restrict__ alpha2, double* __restrict__ LinvT) {
  const size_t per_out = (size_t)T * KC * TILE_N * TILE_K;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (e >= per_out) return;
  const int kk = (int)(e & 15), r = (int)((e >> 4) & 127);
  const size_t chunk = e >> 11;
  const int kc = (int)(chunk % KC), t = (int)(chunk / KC);
  const int n = t * TILE_N + r, k = kc * TILE_K + kk;
  double v = 0.0;
  if (n < N && k <= n) v = alpha2[j] * Xp[(size_t)j * P2 * P2 + (size_t)n + (size_t)k * P2];
  LinvT[(size_t)j * per_out + (chunk * TILE_N + r) * TILE_K + swz(r, kk)] = v;
}

__global__ void pack_a_kernel(const double* __restrict__ a, int N, int Ns, const double* __restrict__ alpha2,
                              double* __restrict__ As) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (n >= Ns) return;
  As[(size_t)j * Ns + n] = (n < N) ? alpha2[j] * a[(size_t)j * N + n] : 0.0;
}

// extract N x N col-major (lower, zero upper) from a padded P2 x P2 buffer
__global__ void unpad_kernel(const double* __restrict__ Pp, int P2, int N, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y, j = blockIdx.z;
  if (r >= N) return;
  out[(size_t)j * N * N + (size_t)r + (size_t)c * N] = (c <= r) ? Pp[(size_t)j * P2 * P2 + (size_t)r + (size_t)c * P2] : 0.0;
}

// ------------------------------------------------------------------------------------------------ driver
void gp_destroy(Gp* gp) {
  if (!gp) return;
  cudaFree(gp->block);   // Us, scale, As, LinvT, L, Linv, a live in one allocation
  cudaFree(gp->LinvI8); cudaFree(gp->LinvI8P); cudaFree(gp->cscale);
  delete gp;
}

static int padded_size(int N) { return (N + NB - 1) / NB * NB; }   // multiple of 64, NOT a power of two

int gp_build(Ctx* ctx, int kernel_id, int d, int N, int p, const double* X, const double* Y, const double* lambda,
             const double* alpha, const double* sigma, const double* mean_at_X, const double* L_host,
             const double* a_host, const bosip_b200_gp_opts* opts_in, Gp** out) {
  if (kernel_id < 0 || kernel_id > 2) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown kernel id");
  if (d < 1 || d > MAX_D) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "d must be in 1..32");
  if (p < 1 || p > MAX_P) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "p must be in 1..64");
  if (N < 1) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "N must be >= 1");
  if (!X || !lambda || !alpha || !sigma) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "null hyper-parameter/data pointer");
  if (!L_host && !Y) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "Y is required");
  if (L_host && !a_host) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "a is required with L");
  for (int j = 0; j < p; ++j) {
    if (!(alpha[j] > 0.0) || !(sigma[j] >= 0.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "alpha must be > 0 and sigma >= 0");
    for (int k = 0; k < d; ++k)
      if (!(lambda[k + d * j] > 0.0)) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "length-scales must be > 0");
  }

  // BOSIP_B200_FIT_TRACE=1: stage wall times (stream synchronised at every mark) on stderr -- development aid
  static const bool trace = getenv("BOSIP_B200_FIT_TRACE") != nullptr;
  std::vector<std::pair<const char*, double>> marks;
  auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
  const double t_begin = trace ? now() : 0.0;
  auto mark = [&](const char* name) { if (trace) { cudaStreamSynchronize(ctx->s_comp); marks.emplace_back(name, now()); } };

  Gp* gp = new Gp();
  gp->ctx = ctx; gp->kernel_id = kernel_id; gp->d = d; gp->dp = pad_dim(d); gp->N = N; gp->p = p;
  gp->Ns = (int)round_up(N, TILE_K); gp->T = (N + TILE_N - 1) / TILE_N; gp->KC = gp->Ns / TILE_K;
  bosip_b200_gp_opts dflt = {0.0, 0, 1};
  gp->opts = opts_in ? *opts_in : dflt;
  const int dp = gp->dp, Ns = gp->Ns, T = gp->T, KC = gp->KC, P2 = padded_size(N);
  cudaStream_t st = ctx->s_comp, sb = ctx->s_in, ss = ctx->s_out;   // main | trailing-update lookahead | triangular solves
  struct Guard { Gp* g; bool ok = false; ~Guard() { if (!ok) gp_destroy(g); } } guard{gp};

  // ---- small host-side parameter prep
  const double ck = kernel_id == BOSIP_B200_KERNEL_SE ? sqrt(0.5) : kernel_id == BOSIP_B200_KERNEL_MATERN32 ? sqrt(3.0) : sqrt(5.0);
  std::vector<double> h_scale((size_t)p * dp, 0.0), h_alpha2(p), h_diag(p), h_floor(p), h_resid;
  for (int j = 0; j < p; ++j) {
    for (int k = 0; k < d; ++k) h_scale[(size_t)j * dp + k] = ck / lambda[k + d * j];
    h_alpha2[j] = alpha[j] * alpha[j];
    gp->alpha2[j] = h_alpha2[j];
    gp->sig2[j] = sigma[j] * sigma[j];
    h_diag[j] = sigma[j] * sigma[j] + gp->opts.jitter;
    h_floor[j] = 1e-14 * (h_alpha2[j] + h_diag[j]);
  }

  const size_t pp = (size_t)p * P2 * P2;
  double *dX = nullptr, *dAlpha2 = nullptr, *dDiag = nullptr, *dFloor = nullptr, *dY = nullptr, *dW = nullptr, *Lp = nullptr, *Xp = nullptr, *Tmp = nullptr,
         *dLin = nullptr;
  int* dFlag = nullptr;
  {
    // what the conditioned GP keeps: ONE allocation (cudaMalloc / cudaFree of these ~1.6 GB at N = 4096, p = 4 cost milliseconds each)
    size_t off = 0;
    auto carve = [&](size_t n_doubles) { size_t o = off; off += (n_doubles * 8 + 255) / 256 * 256; return o; };
    const size_t o_Us = carve((size_t)p * dp * Ns), o_scale = carve((size_t)p * dp), o_As = carve((size_t)p * Ns),
                 o_LinvT = carve((size_t)p * T * KC * TILE_N * TILE_K), o_L = carve((size_t)p * N * N), o_Linv = carve((size_t)p * N * N),
                 o_a = carve((size_t)p * N);
    BOSIP_CUDA(ctx, cudaMalloc(&gp->block, off));
    char* base = reinterpret_cast<char*>(gp->block);
    auto G = [&](size_t o) { return reinterpret_cast<double*>(base + o); };
    gp->Us = G(o_Us); gp->scale = G(o_scale); gp->As = G(o_As); gp->LinvT = G(o_LinvT); gp->L = G(o_L); gp->Linv = G(o_Linv); gp->a = G(o_a);
    // temporaries of the factorisation: the context's grow-only workspace (no allocation after the first fit of a size)
    off = 0;
    const size_t w_X = carve((size_t)d * N), w_alpha = carve(p), w_diag = carve(p), w_floor = carve(p), w_Y = carve((size_t)p * N), w_W = carve((size_t)3 * p * P2),
                 w_Lp = carve(pp), w_Xp = carve(pp), w_Tmp = carve((size_t)p * ((size_t)P2 * P2 / 2 + (size_t)64 * P2)),
                 w_Lin = L_host ? carve((size_t)p * N * N) : 0, w_flag = carve(1);
    BOSIP_TRY(ensure_ws(ctx, off));
    auto Wp = [&](size_t o) { return reinterpret_cast<double*>(ctx->ws + o); };
    dX = Wp(w_X); dAlpha2 = Wp(w_alpha); dDiag = Wp(w_diag); dFloor = Wp(w_floor); dY = Wp(w_Y); dW = Wp(w_W); Lp = Wp(w_Lp); Xp = Wp(w_Xp); Tmp = Wp(w_Tmp);
    if (L_host) dLin = Wp(w_Lin);
    dFlag = reinterpret_cast<int*>(ctx->ws + w_flag);
  }
  BOSIP_CUDA(ctx, cudaMemsetAsync(Xp, 0, pp * 8, st));
  BOSIP_CUDA(ctx, cudaMemsetAsync(dFlag, 0, sizeof(int), st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dX, X, (size_t)d * N * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(gp->scale, h_scale.data(), h_scale.size() * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dAlpha2, h_alpha2.data(), p * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dDiag, h_diag.data(), p * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dFloor, h_floor.data(), p * 8, cudaMemcpyHostToDevice, st));

  mark("alloc+upload");
  scale_inputs_kernel<<<dim3((Ns + 127) / 128, dp, p), 128, 0, st>>>(dX, gp->scale, d, dp, N, Ns, gp->Us);
  BOSIP_CUDA(ctx, cudaGetLastError());

  const int nb = P2 / NB;
  BOSIP_CUDA(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)POTRF_SMEM));
  if (!L_host) {
    // residuals y_j - m_j(X), laid out [p][N]  (Y is p x N column-major: Y[j + p n])
    h_resid.resize((size_t)p * N);
    for (int j = 0; j < p; ++j)
      for (int n = 0; n < N; ++n)
        h_resid[(size_t)j * N + n] = Y[j + (size_t)p * n] - (mean_at_X ? mean_at_X[j + (size_t)p * n] : 0.0);
    BOSIP_CUDA(ctx, cudaMemcpyAsync(dY, h_resid.data(), h_resid.size() * 8, cudaMemcpyHostToDevice, st));
    build_K_kernel<<<dim3((P2 + 127) / 128, P2, p), 128, 0, st>>>(kernel_id, gp->Us, dp, N, Ns, P2, dAlpha2, dDiag, Lp);
    BOSIP_CUDA(ctx, cudaGetLastError());
    mark("build_K");
    // ---- blocked right-looking Cholesky, two levels: inside a 256-column panel the 64-column sub-panels update only the rest of
    //      that panel (K = 64); the big trailing update then runs ONCE per panel with K = 256, which quarters the read-modify-write
    //      traffic on the trailing matrix (the part that does not fit the L2) and gives the GEMM 16 k-slabs per tile instead of 4
    const int PANEL = 4 * NB;
    bool big_pending = false;
    for (int J = 0; J < P2; J += PANEL) {
      const int Jw = std::min(PANEL, P2 - J);
      for (int c0i = J; c0i < J + Jw; c0i += NB) {
        potrf_diag_kernel<<<p, 256, POTRF_SMEM, st>>>(Lp, Xp, P2, c0i / NB, 1, dFlag, dFloor);
        BOSIP_CUDA(ctx, cudaGetLastError());
        const int rem = P2 - (c0i + NB);
        if (rem == 0) break;
        const long long r0 = c0i + NB, c0 = c0i;
        GemmArgs g{};
        // panel: L21 = A21 * inv(L11)^T       (in place; a CTA has all of its A rows in shared memory before it writes them)
        g.A = Lp + r0 + c0 * P2; g.sam = 1; g.sak = P2;
        g.B = Xp + c0 + c0 * P2; g.sbk = P2; g.sbn = 1;   // B(k,n) = Dinv(n,k)
        g.C = Lp + r0 + c0 * P2; g.scm = 1; g.scn = P2;
        g.bA1 = g.bB1 = g.bC1 = (long long)P2 * P2; g.nb1 = p; g.K = NB; g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0;
        BOSIP_TRY(gemm64(ctx, g, rem, NB, p, st));
        // inside the panel: A[r0:, r0:J+Jw] -= L21 L21^T  (lower tiles only)
        const int n_in = (int)(J + Jw - r0);
        if (n_in > 0) {
          g.A = Lp + r0 + c0 * P2; g.sam = 1; g.sak = P2;
          g.B = Lp + r0 + c0 * P2; g.sbk = P2; g.sbn = 1;   // B(k,n) = L21(n,k)
          g.C = Lp + r0 + r0 * P2; g.scm = 1; g.scn = P2;
          g.alpha = -1.0; g.beta = 1.0; g.lower_only = 1;
          BOSIP_TRY(gemm64(ctx, g, rem, n_in, p, st));
        }
      }
      const long long rt = J + Jw;
      const int remt = (int)(P2 - rt);
      if (remt > 0) {   // trailing: A[rt:, rt:] -= L[rt:, J:rt] L[rt:, J:rt]^T, with LOOKAHEAD:
        // (i) the next panel's columns [rt, rt + 256) on the main stream -- the sub-panel chain of the next panel (64 latency-bound
        // launches of ~100 us with a handful of CTAs each) depends on nothing else; (ii) the rest of the trailing matrix on a side
        // stream, where its big GEMM overlaps that chain.  (i) and the next (ii) read-modify-write what the previous (ii) wrote.
        GemmArgs g{};
        g.sam = 1; g.sak = P2; g.sbk = P2; g.sbn = 1; g.scm = 1; g.scn = P2;
        g.bA1 = g.bB1 = g.bC1 = (long long)P2 * P2; g.nb1 = p; g.K = Jw; g.alpha = -1.0; g.beta = 1.0; g.lower_only = 1;
        const int n_i = std::min(PANEL, remt);
        if (big_pending) BOSIP_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_comp[1], 0));   // the previous panel's (ii) has finished
        g.A = Lp + rt + (long long)J * P2; g.B = g.A; g.C = Lp + rt + rt * P2;
        BOSIP_TRY(gemm64(ctx, g, remt, n_i, p, st));
        const int rest = remt - n_i;
        if (rest > 0) {
          BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_comp[0], st));
          BOSIP_CUDA(ctx, cudaStreamWaitEvent(sb, ctx->ev_comp[0], 0));
          const long long r2 = rt + n_i;
          g.A = Lp + r2 + (long long)J * P2; g.B = g.A; g.C = Lp + r2 + r2 * P2;
          BOSIP_TRY(gemm64(ctx, g, rest, rest, p, sb));
          BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_comp[1], sb));
          big_pending = true;
        }
      }
    }
    if (big_pending) BOSIP_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_comp[1], 0));
  } else {
    BOSIP_CUDA(ctx, cudaMemcpyAsync(dLin, L_host, (size_t)p * N * N * 8, cudaMemcpyHostToDevice, st));
    pad_L_kernel<<<dim3((P2 + 127) / 128, P2, p), 128, 0, st>>>(dLin, N, P2, Lp);
    BOSIP_CUDA(ctx, cudaGetLastError());
    for (int b = 0; b < nb; ++b) {
      potrf_diag_kernel<<<p, 256, POTRF_SMEM, st>>>(Lp, Xp, P2, b, 0, dFlag, nullptr);
      BOSIP_CUDA(ctx, cudaGetLastError());
    }
  }
  mark("cholesky");
  int h_flag = 0;
  BOSIP_CUDA(ctx, cudaMemcpyAsync(&h_flag, dFlag, sizeof(int), cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  if (h_flag != 0)
    BOSIP_FAIL(ctx, BOSIP_B200_ENOTPD, "Cholesky: non-positive pivot in diagonal block " + std::to_string(h_flag - 1) +
                                         " (matrix is not positive definite)");

  // the triangular solves for `a` need only L: they run on a side stream, underneath the inverse's GEMMs
  if (!L_host) {
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_in[0], st));
    BOSIP_CUDA(ctx, cudaStreamWaitEvent(ss, ctx->ev_in[0], 0));
    pad_rhs_kernel<<<dim3((P2 + 255) / 256, p), 256, 0, ss>>>(dY, N, P2, dW);
    for (int b = 0; b < nb; ++b)
      trsv_fwd_step<<<dim3(1 + (P2 - (b + 1) * NB + 255) / 256, p), 256, 0, ss>>>(Lp, P2, b, dW, dW + (size_t)p * P2);
    for (int b = nb - 1; b >= 0; --b)
      trsv_bwd_step<<<dim3(1 + (b * NB + 63) / 64, p), 256, 0, ss>>>(Lp, P2, b, dW + (size_t)p * P2, dW + (size_t)2 * p * P2);
    unpad_rhs_kernel<<<dim3((N + 255) / 256, p), 256, 0, ss>>>(dW + (size_t)2 * p * P2, N, P2, gp->a);
    BOSIP_CUDA(ctx, cudaGetLastError());
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_in[1], ss));
  }

  // ---- triangular inverse by recursive doubling: X21 = -X22 (L21 X11).  Level s pairs block [2is, 2is+s) with
  //      [2is+s, 2is+2s); when P2 is not a multiple of 2s the last pair's second block is shorter (m rows) -- one more launch.
  if (nb > 1) {
    for (int s = NB; s < P2; s <<= 1) {
      const int npairs = P2 / (2 * s);
      const long long pair_stride = (long long)2 * s * (P2 + 1);
      const long long r_off = (long long)npairs * 2 * s;                 // first row of the ragged pair
      const int m_rag = (P2 - r_off > s) ? (int)(P2 - r_off - s) : 0;    // rows of its X22 (multiple of 64)
      double* Tmp_rag = Tmp + (size_t)p * npairs * s * s;
      GemmArgs g{};
      g.K = s; g.lower_only = 0;
      if (npairs > 0) {
        g.nb1 = npairs;
        // T = L21 X11   (X11 lower triangular: k >= n)
        g.A = Lp + s; g.sam = 1; g.sak = P2; g.bA1 = pair_stride; g.bA2 = (long long)P2 * P2;
        g.B = Xp; g.sbk = 1; g.sbn = P2; g.bB1 = pair_stride; g.bB2 = (long long)P2 * P2;
        g.C = Tmp; g.scm = 1; g.scn = s; g.bC1 = (long long)s * s; g.bC2 = (long long)npairs * s * s;
        g.alpha = 1.0; g.beta = 0.0; g.a_lower = 0; g.b_lower = 1;
        BOSIP_TRY(gemm64(ctx, g, s, s, npairs * p, st));
        // X21 = -X22 T  (X22 lower triangular: k <= m)
        g.A = Xp + (long long)s * (P2 + 1); g.sam = 1; g.sak = P2; g.bA1 = pair_stride; g.bA2 = (long long)P2 * P2;
        g.B = Tmp; g.sbk = 1; g.sbn = s; g.bB1 = (long long)s * s; g.bB2 = (long long)npairs * s * s;
        g.C = Xp + s; g.scm = 1; g.scn = P2; g.bC1 = pair_stride; g.bC2 = (long long)P2 * P2;
        g.alpha = -1.0; g.beta = 0.0; g.a_lower = 1; g.b_lower = 0;
        BOSIP_TRY(gemm64(ctx, g, s, s, npairs * p, st));
      }
      if (m_rag > 0) {
        g.nb1 = 1;
        g.A = Lp + (r_off + s) + r_off * P2; g.sam = 1; g.sak = P2; g.bA1 = 0; g.bA2 = (long long)P2 * P2;
        g.B = Xp + r_off * (P2 + 1); g.sbk = 1; g.sbn = P2; g.bB1 = 0; g.bB2 = (long long)P2 * P2;
        g.C = Tmp_rag; g.scm = 1; g.scn = m_rag; g.bC1 = 0; g.bC2 = (long long)m_rag * s;
        g.K = s; g.alpha = 1.0; g.beta = 0.0; g.a_lower = 0; g.b_lower = 1;
        BOSIP_TRY(gemm64(ctx, g, m_rag, s, p, st));
        g.A = Xp + (r_off + s) * (P2 + 1); g.sam = 1; g.sak = P2; g.bA1 = 0; g.bA2 = (long long)P2 * P2;
        g.B = Tmp_rag; g.sbk = 1; g.sbn = m_rag; g.bB1 = 0; g.bB2 = (long long)m_rag * s;
        g.C = Xp + (r_off + s) + r_off * P2; g.scm = 1; g.scn = P2; g.bC1 = 0; g.bC2 = (long long)P2 * P2;
        g.K = m_rag; g.alpha = -1.0; g.beta = 0.0; g.a_lower = 1; g.b_lower = 0;
        BOSIP_TRY(gemm64(ctx, g, m_rag, s, p, st));
      }
    }
  }

  mark("inverse");
  // ---- weights a = K^{-1} y: forward + backward substitution with L (issued above on the side stream); join here
  if (!L_host) {
    BOSIP_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_in[1], 0));
  } else {
    // a_host is N x p column-major == [p][N]
    BOSIP_CUDA(ctx, cudaMemcpyAsync(gp->a, a_host, (size_t)p * N * 8, cudaMemcpyHostToDevice, st));
  }

  mark("solves");
  // ---- pack for the evaluator, keep plain copies for gp_get_factors
  {
    const size_t per_out = (size_t)T * KC * TILE_N * TILE_K;
    pack_linv_kernel<<<dim3((unsigned)((per_out + 255) / 256), p), 256, 0, st>>>(Xp, P2, N, T, KC, dAlpha2, gp->LinvT);
    pack_a_kernel<<<dim3((Ns + 127) / 128, p), 128, 0, st>>>(gp->a, N, Ns, dAlpha2, gp->As);
    unpad_kernel<<<dim3((N + 127) / 128, N, p), 128, 0, st>>>(Lp, P2, N, gp->L);
    unpad_kernel<<<dim3((N + 127) / 128, N, p), 128, 0, st>>>(Xp, P2, N, gp->Linv);
    BOSIP_CUDA(ctx, cudaGetLastError());
  }
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  mark("pack");
  if (trace) {
    double prev = t_begin;
    fprintf(stderr, "[fit trace N=%d p=%d P=%d]", N, p, P2);
    for (auto& m : marks) { fprintf(stderr, " %s %.2f ms |", m.first, m.second - prev); prev = m.second; }
    fprintf(stderr, " (frees follow)\n");
  }
  guard.ok = true;
  *out = gp;
  return 0;
}

}  // namespace bosip
