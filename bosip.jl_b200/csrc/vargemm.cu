// Predictive-variance contraction  q[m] = || L_j^{-1} k*_m ||^2   on the FP64 tensor-core pipe (DMMA).
// Replaces the TRSM `L \ K*'` + column sum of squares inside BOSS.var / mean_and_std
// (call sites /root/reference/src/likelihoods/normal.jl:37,43,60,68; SURVEY.md section 3.2 "HOT LOOP #2").
//
// GEMM view: W[m, n] = sum_k Ks[m, k] * Linv[n, k]  (Linv lower triangular => only k <= n contributes), then
// q[m] = sum_n W[m, n]^2.  W is never written: each CTA owns a 128(m) x 128(n) tile, keeps it in registers
// (8 warps x [16 x 128], 64 FP64 accumulators per thread) and reduces its squares to one partial per row.
//
// Pipeline: a dedicated producer warp streams k-chunks of 16 with two 16 KB 1-D TMA bulk copies per stage
// (A = K* chunk, B = L^{-1} chunk; both stored chunk-major and XOR-swizzled so a stage is contiguous in global
// memory AND conflict-free for the m8n8k4 fragment loads), full/empty mbarriers, STAGES deep.
// Triangular structure: tile t only walks k < min(128 (t+1), K); inside the diagonal block the 8-wide n-subtiles
// that lie entirely left of the current k-step are skipped (warp-uniform predicate), so executed DMMA work is
// ~ N^2/2 + 6 N per row instead of N^2.
#include "common.cuh"

namespace bosip {

constexpr int VG_STAGES = 5;
constexpr int VG_CONSUMER_WARPS = 8;
constexpr int VG_THREADS = (VG_CONSUMER_WARPS + 1) * 32;
constexpr int VG_STAGE_DOUBLES = (TILE_M + TILE_N) * TILE_K;  // A then B
constexpr size_t VG_SMEM = (size_t)VG_STAGES * VG_STAGE_DOUBLES * sizeof(double) + 2 * VG_STAGES * sizeof(uint64_t);

__global__ void __launch_bounds__(VG_THREADS, 1)
vargemm_kernel(const double* __restrict__ Ks, const double* __restrict__ LinvT, int KC, int T, long long mc,
               double* __restrict__ q_part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)VG_STAGES * VG_STAGE_DOUBLES * sizeof(double));
  uint64_t* empty = full + VG_STAGES;

  const int t = T - 1 - (int)blockIdx.x;  // heavy (large-k) tiles are scheduled first
  const int mt = blockIdx.y, j = blockIdx.z, p = gridDim.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nk = min((t + 1) * (TILE_N / TILE_K), KC);  // k-chunks this tile needs
  const int diag_kc = t * (TILE_N / TILE_K);            // first chunk of the diagonal block

  if (threadIdx.x == 0) {
    for (int s = 0; s < VG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], VG_CONSUMER_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  const double* gA = Ks + ((size_t)j * KC * mc + (size_t)mt * TILE_M) * TILE_K;           // + kc * mc * 16
  const double* gB = LinvT + (((size_t)j * T + t) * KC) * (size_t)(TILE_N * TILE_K);       // + kc * 128 * 16

  if (warp == VG_CONSUMER_WARPS) {
    // ------------------------------------------------------------- producer warp (one elected lane)
    if (lane == 0) {
      for (int it = 0; it < nk; ++it) {
        const int s = it % VG_STAGES;
        if (it >= VG_STAGES) mbar_wait(&empty[s], ((it / VG_STAGES) - 1) & 1);
        double* sA = stage_base + (size_t)s * VG_STAGE_DOUBLES;
        double* sB = sA + TILE_M * TILE_K;
        mbar_expect_tx(&full[s], (uint32_t)(VG_STAGE_DOUBLES * sizeof(double)));
        bulk_g2s(sA, gA + (size_t)it * mc * TILE_K, TILE_M * TILE_K * sizeof(double), &full[s]);
        bulk_g2s(sB, gB + (size_t)it * (TILE_N * TILE_K), TILE_N * TILE_K * sizeof(double), &full[s]);
      }
    }
    return;
  }

  // --------------------------------------------------------------- consumer warps: 16 rows x 128 cols each
  const int g = lane >> 2, tig = lane & 3;
  double c[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; c[i][2] = 0.0; c[i][3] = 0.0; }

  const int sw = g & 3;  // (row & 3) for rows g, g+8 of A and rows 8i+g of B alike
  const int a_off0 = (warp * 16 + g) * TILE_K + tig;
  const int a_off1 = a_off0 + 8 * TILE_K;
  const int b_off = TILE_M * TILE_K + g * TILE_K + tig;

#pragma unroll 1
  for (int it = 0; it < nk; ++it) {
    const int s = it % VG_STAGES;
    mbar_wait(&full[s], (it / VG_STAGES) & 1);
    const double* sS = stage_base + (size_t)s * VG_STAGE_DOUBLES;
    if (it < diag_kc) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int col = ((ks ^ sw) << 2);
        const double a0 = sS[a_off0 + col], a1 = sS[a_off1 + col];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const double b = sS[b_off + i * 8 * TILE_K + col];
          dmma884(c[i][0], c[i][1], a0, b);
          dmma884(c[i][2], c[i][3], a1, b);
        }
      }
    } else {
      const int kloc = (it - diag_kc) * TILE_K;  // k offset inside the diagonal block
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int col = ((ks ^ sw) << 2);
        const double a0 = sS[a_off0 + col], a1 = sS[a_off1 + col];
        const int imin = (kloc + 4 * ks) >> 3;  // n-subtiles i < imin hold only zeros of L^{-1} at this k-step
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (i >= imin) {
            const double b = sS[b_off + i * 8 * TILE_K + col];
            dmma884(c[i][0], c[i][1], a0, b);
            dmma884(c[i][2], c[i][3], a1, b);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // --------------------------------------------------------------- epilogue: row sums of squares
  double s0 = 0.0, s1 = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    s0 = fma(c[i][0], c[i][0], s0); s0 = fma(c[i][1], c[i][1], s0);
    s1 = fma(c[i][2], c[i][2], s1); s1 = fma(c[i][3], c[i][3], s1);
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  if (tig == 0) {
    double* dst = q_part + ((size_t)t * p + j) * mc + (size_t)mt * TILE_M + warp * 16 + g;
    dst[0] = s0;
    dst[8] = s1;
  }
}

int launch_vargemm(Ctx* ctx, const Gp* gp, const double* Ks, int64_t mc, int64_t m_tiles, double* q_part, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    BOSIP_CUDA(ctx, cudaFuncSetAttribute(vargemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VG_SMEM));
    attr_set = true;
  }
  dim3 grid((unsigned)gp->T, (unsigned)m_tiles, (unsigned)gp->p);
  vargemm_kernel<<<grid, VG_THREADS, VG_SMEM, st>>>(Ks, gp->LinvT, gp->KC, gp->T, (long long)mc, q_part);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace bosip
