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
#include <vector>

#include "common.cuh"
#include "fastmath.cuh"

namespace bosip {

constexpr int NB = 64;
constexpr size_t POTRF_SMEM = 2 * NB * (NB + 1) * sizeof(double);

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
// same block of Xp.  One CTA of 64 threads per output.  flag[0] = 1 + block index on a non-positive pivot.
__global__ void __launch_bounds__(64) potrf_diag_kernel(double* __restrict__ Ap, double* __restrict__ Xp, int P2, int b,
                                                        int do_factor, int* __restrict__ flag) {
  extern __shared__ __align__(16) unsigned char potrf_smem[];
  double (*s)[NB + 1] = reinterpret_cast<double (*)[NB + 1]>(potrf_smem);
  double (*x)[NB + 1] = s + NB;
  const int j = blockIdx.x, tid = threadIdx.x;
  double* A = Ap + (size_t)j * P2 * P2 + (size_t)b * NB + (size_t)b * NB * P2;
  for (int c = 0; c < NB; ++c) s[tid][c] = A[tid + (size_t)c * P2];
  __syncthreads();
  if (do_factor) {
    for (int c = 0; c < NB; ++c) {
      const double piv = s[c][c];
      if (!(piv > 0.0)) { if (tid == 0) atomicCAS(flag, 0, 1 + b); break; }  // uniform: every thread reads the same pivot
      const double lcc = sqrt(piv);
      __syncthreads();
      if (tid == c) s[c][c] = lcc;
      else if (tid > c) s[tid][c] = s[tid][c] / lcc;
      __syncthreads();
      // rank-1 update of the trailing lower triangle: row tid, columns c+1..tid
      if (tid > c) {
        const double l_rc = s[tid][c];
        for (int cc = c + 1; cc <= tid; ++cc) s[tid][cc] -= l_rc * s[cc][c];
      }
      __syncthreads();
    }
    for (int c = 0; c < NB; ++c) A[tid + (size_t)c * P2] = (c <= tid) ? s[tid][c] : 0.0;
  }
  __syncthreads();
  // inverse by forward substitution, thread = column of the identity
  for (int r = 0; r < NB; ++r) {
    double acc = (r == tid) ? 1.0 : 0.0;
    for (int k = tid; k < r; ++k) acc -= s[r][k] * x[k][tid];
    x[r][tid] = (r >= tid) ? acc / s[r][r] : 0.0;
  }
  __syncthreads();
  double* X = Xp + (size_t)j * P2 * P2 + (size_t)b * NB + (size_t)b * NB * P2;
  for (int c = 0; c < NB; ++c) X[tid + (size_t)c * P2] = x[tid][c];
}

// a = K^{-1} y = L^{-T} (L^{-1} y) by two blocked triangular solves with the Cholesky factor itself (not through the explicit
// inverse).  One CTA of 256 threads per output; 64-row blocks: the off-diagonal part is a GEMV with coalesced column reads, the
// 64 x 64 diagonal block is substituted exactly by warp 0 (two rows per lane, the pivot row's value broadcast by shuffle).
__global__ void __launch_bounds__(256) trsv_chol_kernel(const double* __restrict__ Lp, int P, int N, const double* __restrict__ y,
                                                        double* __restrict__ w_out, double* __restrict__ a_out) {
  extern __shared__ __align__(16) unsigned char trsv_smem[];
  double* w = reinterpret_cast<double*>(trsv_smem);     // [P]   forward solution, then overwritten by the backward one
  double (*part)[64] = reinterpret_cast<double (*)[64]>(w + P);   // [4][64] GEMV partials
  double* r = &part[0][0] + 4 * 64;                     // [64]  right-hand side of the diagonal block
  double (*D)[65] = reinterpret_cast<double (*)[65]>(r + 64);     // [64][65] diagonal block
  const int j = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* L = Lp + (size_t)j * P * P;
  const int nb = P / 64;
  // ---------------- forward: L w = y
  for (int b = 0; b < nb; ++b) {
    const int r0 = b * 64, row = tid & 63, q = tid >> 6;
    double acc = 0.0;
    for (int c = q; c < r0; c += 4) acc = fma(L[(size_t)(r0 + row) + (size_t)c * P], w[c], acc);
    part[q][row] = acc;
    for (int e = tid; e < 64 * 64; e += 256) { const int rr = e & 63, cc = e >> 6; D[rr][cc] = L[(size_t)(r0 + rr) + (size_t)(r0 + cc) * P]; }
    __syncthreads();
    if (tid < 64) {
      const double yv = (r0 + tid < N) ? y[(size_t)j * N + r0 + tid] : 0.0;
      r[tid] = yv - (((part[0][tid] + part[1][tid]) + part[2][tid]) + part[3][tid]);
    }
    __syncthreads();
    if (warp == 0) {
      double x0 = r[lane], x1 = r[lane + 32];
      for (int k = 0; k < 64; ++k) {
        const double rk = __shfl_sync(0xffffffffu, k < 32 ? x0 : x1, k & 31);
        const double wk = rk / D[k][k];
        if (lane == (k & 31)) { if (k < 32) x0 = wk; else x1 = wk; }
        if (lane > k) x0 = fma(-D[lane][k], wk, x0);
        if (lane + 32 > k) x1 = fma(-D[lane + 32][k], wk, x1);
      }
      w[r0 + lane] = x0; w[r0 + lane + 32] = x1;
    }
    __syncthreads();
  }
  if (w_out) for (int e = tid; e < N; e += 256) w_out[(size_t)j * N + e] = w[e];
  // ---------------- backward: L^T a = w  (w is overwritten block by block from the end)
  for (int b = nb - 1; b >= 0; --b) {
    const int r0 = b * 64;
    // r[col] = w[r0 + col] - sum_{c >= r0 + 64} L[c][r0 + col] a[c]: warp `warp` owns columns warp, warp + 8, ..., lanes run along c
    for (int col = warp; col < 64; col += 8) {
      double acc = 0.0;
      const double* Lc = L + (size_t)(r0 + col) * P;
      for (int c = r0 + 64 + lane; c < P; c += 32) acc = fma(Lc[c], w[c], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) r[col] = w[r0 + col] - acc;
    }
    for (int e = tid; e < 64 * 64; e += 256) { const int rr = e & 63, cc = e >> 6; D[rr][cc] = L[(size_t)(r0 + rr) + (size_t)(r0 + cc) * P]; }
    __syncthreads();
    if (warp == 0) {
      double x0 = r[lane], x1 = r[lane + 32];
      for (int k = 63; k >= 0; --k) {
        const double rk = __shfl_sync(0xffffffffu, k < 32 ? x0 : x1, k & 31);
        const double ak = rk / D[k][k];
        if (lane == (k & 31)) { if (k < 32) x0 = ak; else x1 = ak; }
        if (lane < k) x0 = fma(-D[k][lane], ak, x0);          // column `lane` of L^T row k = L[k][lane]
        if (lane + 32 < k) x1 = fma(-D[k][lane + 32], ak, x1);
      }
      w[r0 + lane] = x0; w[r0 + lane + 32] = x1;
    }
    __syncthreads();
  }
  for (int e = tid; e < N; e += 256) a_out[(size_t)j * N + e] = w[e];
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
  cudaFree(gp->Us); cudaFree(gp->scale); cudaFree(gp->As); cudaFree(gp->LinvT);
  cudaFree(gp->L); cudaFree(gp->Linv); cudaFree(gp->a);
  cudaFree(gp->LinvI8); cudaFree(gp->cscale);
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

  Gp* gp = new Gp();
  gp->ctx = ctx; gp->kernel_id = kernel_id; gp->d = d; gp->dp = pad_dim(d); gp->N = N; gp->p = p;
  gp->Ns = (int)round_up(N, TILE_K); gp->T = (N + TILE_N - 1) / TILE_N; gp->KC = gp->Ns / TILE_K;
  bosip_b200_gp_opts dflt = {0.0, 0, 1};
  gp->opts = opts_in ? *opts_in : dflt;
  const int dp = gp->dp, Ns = gp->Ns, T = gp->T, KC = gp->KC, P2 = padded_size(N);
  cudaStream_t st = ctx->s_comp;
  struct Guard { Gp* g; bool ok = false; ~Guard() { if (!ok) gp_destroy(g); } } guard{gp};

  // ---- small host-side parameter prep
  const double ck = kernel_id == BOSIP_B200_KERNEL_SE ? sqrt(0.5) : kernel_id == BOSIP_B200_KERNEL_MATERN32 ? sqrt(3.0) : sqrt(5.0);
  std::vector<double> h_scale((size_t)p * dp, 0.0), h_alpha2(p), h_diag(p), h_resid;
  for (int j = 0; j < p; ++j) {
    for (int k = 0; k < d; ++k) h_scale[(size_t)j * dp + k] = ck / lambda[k + d * j];
    h_alpha2[j] = alpha[j] * alpha[j];
    gp->alpha2[j] = h_alpha2[j];
    gp->sig2[j] = sigma[j] * sigma[j];
    h_diag[j] = sigma[j] * sigma[j] + gp->opts.jitter;
  }

  const size_t pp = (size_t)p * P2 * P2;
  double *dX = nullptr, *dAlpha2 = nullptr, *dDiag = nullptr, *dY = nullptr, *dW = nullptr, *Lp = nullptr, *Xp = nullptr, *Tmp = nullptr,
         *dLin = nullptr;
  int* dFlag = nullptr;
  struct Tmps { double** v[9]; int** f; ~Tmps() { for (auto q : v) if (q && *q) cudaFree(*q); if (f && *f) cudaFree(*f); } }
      tmps{{&dX, &dAlpha2, &dDiag, &dY, &dW, &Lp, &Xp, &Tmp, &dLin}, &dFlag};

  BOSIP_CUDA(ctx, cudaMalloc(&gp->Us, (size_t)p * dp * Ns * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&gp->scale, (size_t)p * dp * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&gp->As, (size_t)p * Ns * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&gp->LinvT, (size_t)p * T * KC * TILE_N * TILE_K * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&gp->L, (size_t)p * N * N * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&gp->Linv, (size_t)p * N * N * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&gp->a, (size_t)p * N * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&dX, (size_t)d * N * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&dAlpha2, p * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&dDiag, p * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&dY, (size_t)p * N * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&dW, (size_t)p * N * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&Lp, pp * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&Xp, pp * 8));
  BOSIP_CUDA(ctx, cudaMalloc(&dFlag, sizeof(int)));
  BOSIP_CUDA(ctx, cudaMemsetAsync(Xp, 0, pp * 8, st));
  BOSIP_CUDA(ctx, cudaMemsetAsync(dFlag, 0, sizeof(int), st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dX, X, (size_t)d * N * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(gp->scale, h_scale.data(), h_scale.size() * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dAlpha2, h_alpha2.data(), p * 8, cudaMemcpyHostToDevice, st));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dDiag, h_diag.data(), p * 8, cudaMemcpyHostToDevice, st));

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
    // ---- blocked right-looking Cholesky
    for (int b = 0; b < nb; ++b) {
      potrf_diag_kernel<<<p, NB, POTRF_SMEM, st>>>(Lp, Xp, P2, b, 1, dFlag);
      BOSIP_CUDA(ctx, cudaGetLastError());
      const int rem = (nb - b - 1) * NB;
      if (rem == 0) break;
      const long long r0 = (long long)(b + 1) * NB, c0 = (long long)b * NB;
      GemmArgs g{};
      // panel: L21 = A21 * inv(L11)^T       (in place; a CTA reads its whole 64x64 A block before writing it)
      g.A = Lp + r0 + c0 * P2; g.sam = 1; g.sak = P2;
      g.B = Xp + c0 + c0 * P2; g.sbk = P2; g.sbn = 1;   // B(k,n) = Dinv(n,k)
      g.C = Lp + r0 + c0 * P2; g.scm = 1; g.scn = P2;
      g.bA1 = g.bB1 = g.bC1 = (long long)P2 * P2; g.nb1 = p; g.K = NB; g.alpha = 1.0; g.beta = 0.0; g.lower_only = 0;
      BOSIP_TRY(gemm64(ctx, g, rem, NB, p, st));
      // trailing: A22 -= L21 L21^T  (lower tiles only)
      g.A = Lp + r0 + c0 * P2; g.sam = 1; g.sak = P2;
      g.B = Lp + r0 + c0 * P2; g.sbk = P2; g.sbn = 1;   // B(k,n) = L21(n,k)
      g.C = Lp + r0 + r0 * P2; g.scm = 1; g.scn = P2;
      g.alpha = -1.0; g.beta = 1.0; g.lower_only = 1;
      BOSIP_TRY(gemm64(ctx, g, rem, rem, p, st));
    }
  } else {
    BOSIP_CUDA(ctx, cudaMalloc(&dLin, (size_t)p * N * N * 8));
    BOSIP_CUDA(ctx, cudaMemcpyAsync(dLin, L_host, (size_t)p * N * N * 8, cudaMemcpyHostToDevice, st));
    pad_L_kernel<<<dim3((P2 + 127) / 128, P2, p), 128, 0, st>>>(dLin, N, P2, Lp);
    BOSIP_CUDA(ctx, cudaGetLastError());
    for (int b = 0; b < nb; ++b) {
      potrf_diag_kernel<<<p, NB, POTRF_SMEM, st>>>(Lp, Xp, P2, b, 0, dFlag);
      BOSIP_CUDA(ctx, cudaGetLastError());
    }
  }
  int h_flag = 0;
  BOSIP_CUDA(ctx, cudaMemcpyAsync(&h_flag, dFlag, sizeof(int), cudaMemcpyDeviceToHost, st));
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  if (h_flag != 0)
    BOSIP_FAIL(ctx, BOSIP_B200_ENOTPD, "Cholesky: non-positive pivot in diagonal block " + std::to_string(h_flag - 1) +
                                         " (matrix is not positive definite)");

  // ---- triangular inverse by recursive doubling: X21 = -X22 (L21 X11).  Level s pairs block [2is, 2is+s) with
  //      [2is+s, 2is+2s); when P2 is not a multiple of 2s the last pair's second block is shorter (m rows) -- one more launch.
  if (nb > 1) {
    BOSIP_CUDA(ctx, cudaMalloc(&Tmp, (size_t)p * ((size_t)P2 * P2 / 2 + (size_t)64 * P2) * 8));
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

  // ---- weights a = K^{-1} y: forward + backward substitution with L (trsv_chol_kernel)
  if (!L_host) {
    const size_t trsv_smem = ((size_t)P2 + 4 * 64 + 64 + 64 * 65) * sizeof(double);
    if (trsv_smem > 220 * 1024) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "N too large for the in-shared-memory triangular solve (N <= 23000)");
    BOSIP_CUDA(ctx, cudaFuncSetAttribute(trsv_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsv_smem));
    trsv_chol_kernel<<<p, 256, trsv_smem, st>>>(Lp, P2, N, dY, dW, gp->a);
    BOSIP_CUDA(ctx, cudaGetLastError());
  } else {
    // a_host is N x p column-major == [p][N]
    BOSIP_CUDA(ctx, cudaMemcpyAsync(gp->a, a_host, (size_t)p * N * 8, cudaMemcpyHostToDevice, st));
  }

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
  guard.ok = true;
  *out = gp;
  return 0;
}

}  // namespace bosip
