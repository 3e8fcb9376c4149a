// Predictive-variance contraction  q[m] = || L_j^{-1} k*_m ||^2   on the FP64 tensor-core pipe (DMMA).
// Replaces the TRSM `L \ K*'` + column sum of squares inside BOSS.var / mean_and_std
// (call sites /root/reference/src/likelihoods/normal.jl:37,43,60,68; SURVEY.md section 3.2 "HOT LOOP #2").
//
// GEMM view: W[m, n] = sum_k Ks[m, k] * Linv[n, k]  (Linv lower triangular => only k <= n contributes), then
// q[m] = sum_n W[m, n]^2.  W is never written.
//
// Persistent kernel, one CTA per SM.  A CTA owns whole GROUPS = (128-row m-tile, output j) and walks the T n-tiles of a
// group back to back: per n-tile a 128(m) x 128(n) block of W lives in registers (8 consumer warps x [16 x 128], 64 FP64
// accumulators per thread), is squared and folded into two per-thread running row sums, and the accumulators are reused
// for the next n-tile.  One q per row is written at the end of the group (fixed summation order, no atomics).
//
// A dedicated producer warp streams k-chunks of 16 through a STAGES-deep ring with two 16 KB 1-D TMA bulk copies per
// stage (A = K* chunk, B = L^{-1} chunk; both stored chunk-major and XOR-swizzled so a stage is contiguous in global
// memory AND conflict-free for the m8n8k4 fragment loads) and full/empty mbarriers.  Because the producer follows the
// same static schedule as the consumers it runs ahead across n-tile and group boundaries: the pipeline fills once per
// kernel, not once per tile.
//
// Triangular structure: n-tile t only walks k < min(128 (t+1), K); inside the diagonal block the 8-wide n-subtiles that
// lie entirely left of the current k-step are skipped, and in the last n-tile the subtiles that only hold the zero rows
// padding N up to a multiple of 128 are skipped too (warp-uniform predicates), so executed DMMA work is
// ~ N^2/2 + 6 N per row instead of N^2.
#include <stdlib.h>

#include "common.cuh"

namespace bosip {

#ifndef BOSIP_VG_STAGES
#define BOSIP_VG_STAGES 5
#endif
#ifndef BOSIP_VG_RING
#define BOSIP_VG_RING 12
#endif
#ifndef BOSIP_VG_LAG
#define BOSIP_VG_LAG 2
#endif
constexpr int VG_STAGES = BOSIP_VG_STAGES;
constexpr int VG_CONSUMER_WARPS = 8;
constexpr int VG_THREADS = VG_CONSUMER_WARPS * 32;  // exactly 2 warps per SM sub-partition: up to 255 registers/thread
constexpr int VG_LAG = BOSIP_VG_LAG;                           // the refill of a stage trails its release by this many chunks
constexpr int VG_STAGE_DOUBLES = (TILE_M + TILE_N) * TILE_K;  // A then B
constexpr int VG_SUB = TILE_N / 8;                            // 16 n-subtiles of 8 columns
constexpr size_t VG_SMEM = (size_t)VG_STAGES * VG_STAGE_DOUBLES * sizeof(double) + 2 * VG_STAGES * sizeof(uint64_t);

__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
  return v;
}

// One k-chunk (4 k-steps) of one warp's 16 x 128 block, hand-scheduled: every operand load is an `asm volatile` LDS issued a
// full k-step (up to 32 DMMAs = 512 pipe cycles) ahead of its use, into the register its previous value has just left
// (the compiler's own schedule keeps only two B fragments in flight, 2 DMMAs ahead, and stalls on the LDS latency).
//   DIAG_C < 0 : chunk left of the diagonal block -- every subtile 0..IMAX-1 at every k-step
//   DIAG_C >= 0: chunk DIAG_C of the diagonal block -- subtiles i < 2*DIAG_C + (ks>>1) hold only zeros of L^{-1}: skipped
//   IMAX       : number of 8-row subtiles of this n-tile that hold real rows of L^{-1} (16 except in a ragged last tile)
template <int DIAG_C, int IMAX>
struct VgSched {  // compile-time schedule of the (k-step, subtile) pairs of one chunk
  static __host__ __device__ constexpr int imin(int ks) { return DIAG_C < 0 ? 0 : 2 * DIAG_C + (ks >> 1); }
  static __host__ __device__ constexpr int count(int ks) { return IMAX > imin(ks) ? IMAX - imin(ks) : 0; }
  static __host__ __device__ constexpr int total() { return count(0) + count(1) + count(2) + count(3); }
  static __host__ __device__ constexpr int ks_of(int q) {
    return q < count(0) ? 0 : q < count(0) + count(1) ? 1 : q < count(0) + count(1) + count(2) ? 2 : 3;
  }
  static __host__ __device__ constexpr int first(int ks) { return ks == 0 ? 0 : ks == 1 ? count(0) : ks == 2 ? count(0) + count(1) : count(0) + count(1) + count(2); }
  static __host__ __device__ constexpr int i_of(int q) { return imin(ks_of(q)) + q - first(ks_of(q)); }
};

constexpr int VG_RING = BOSIP_VG_RING;  // B fragments in flight per thread (look-ahead of 11 pairs = 22 DMMAs = 352 pipe cycles)

template <int DIAG_C, int IMAX>
__device__ __forceinline__ void vg_chunk_fast(double (&c)[VG_SUB][4], uint32_t sbase, uint32_t a_off0b, uint32_t a_off1b,
                                              uint32_t b_offb, const uint32_t (&colb)[4]) {
  using S = VgSched<DIAG_C, IMAX>;
  constexpr int NP = S::total();
  double b[VG_RING];
  double a0[2], a1[2];  // double-buffered A fragments (current / next k-step)
  a0[0] = lds_f64(sbase + a_off0b + colb[0]); a1[0] = lds_f64(sbase + a_off1b + colb[0]);
#pragma unroll
  for (int q = 0; q < VG_RING - 1; ++q)
    if (q < NP) b[q % VG_RING] = lds_f64(sbase + b_offb + colb[S::ks_of(q)] + S::i_of(q) * (8 * TILE_K * 8));
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int ks = S::ks_of(q), i = S::i_of(q);
    if (q == S::first(ks) && ks < 3 && S::count(ks + 1) > 0) {  // entering k-step ks: fetch the A fragments of the next one
      a0[(ks + 1) & 1] = lds_f64(sbase + a_off0b + colb[(ks + 1) & 3]);
      a1[(ks + 1) & 1] = lds_f64(sbase + a_off1b + colb[(ks + 1) & 3]);
    }
    dmma884(c[i][0], c[i][1], a0[ks & 1], b[q % VG_RING]);
    dmma884(c[i][2], c[i][3], a1[ks & 1], b[q % VG_RING]);
    const int qn = q + VG_RING - 1;  // refill the slot the PREVIOUS pair has just left
    if (qn < NP) b[qn % VG_RING] = lds_f64(sbase + b_offb + colb[S::ks_of(qn)] + S::i_of(qn) * (8 * TILE_K * 8));
  }
}

template <int IMAX>
__device__ __forceinline__ void vg_chunk_left(double (&c)[VG_SUB][4], uint32_t sbase, uint32_t a0b, uint32_t a1b, uint32_t bb,
                                              const uint32_t (&colb)[4]) {
  vg_chunk_fast<-1, IMAX>(c, sbase, a0b, a1b, bb, colb);
}

// generic fallback: runtime bounds [imin(ks), imax) (diagonal chunks of a ragged last tile)
__device__ __forceinline__ void vg_chunk_guard(double (&c)[VG_SUB][4], const double* __restrict__ sS, int a_off0, int a_off1,
                                               int b_off, int sw, int kloc, int imax) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const int col = ((ks ^ sw) << 2);
    const double a0 = sS[a_off0 + col], a1 = sS[a_off1 + col];
    const int imin = (kloc + 4 * ks) >> 3;  // arithmetic shift: kloc < 0 => no left skip
#pragma unroll
    for (int i = 0; i < VG_SUB; ++i) {
      if (i >= imin && i < imax) {
        const double b = sS[b_off + i * 8 * TILE_K + col];
        dmma884(c[i][0], c[i][1], a0, b);
        dmma884(c[i][2], c[i][3], a1, b);
      }
    }
  }
}

// Walks the CTA's static chunk schedule: groups g = blockIdx.x + i*gridDim.x, n-tiles t = T-1..0, chunks it = 0..nk(t)-1.
struct VgCursor {
  int g, t, it, nk;
  const double* gA;  // K* chunk 0 of the group's m-tile
  const double* gB;  // L^{-1} chunk 0 of n-tile t
};

__global__ void __launch_bounds__(VG_THREADS, 1)
vargemm_kernel(const double* __restrict__ Ks, const double* __restrict__ LinvT, int KC, int T, int N, long long mc,
               int m_tiles, int groups, double* __restrict__ q_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* stage_base = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)VG_STAGES * VG_STAGE_DOUBLES * sizeof(double));
  uint64_t* empty = full + VG_STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CPT = TILE_N / TILE_K;  // k-chunks per n-tile width (8)
  const bool is_producer = (threadIdx.x == 0);

  // chunks per group and in this CTA's whole schedule
  int chunks_per_group = 0;
  for (int t = 0; t < T; ++t) chunks_per_group += min((t + 1) * CPT, KC);
  const int my_groups = (groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t total_chunks = (uint32_t)my_groups * (uint32_t)chunks_per_group;

  // ---- producer state (thread 0): a cursor that runs VG_STAGES - VG_LAG chunks ahead of the consumers
  VgCursor pc;
  uint32_t pn = 0;  // next chunk to load
  auto cursor_set_group = [&](VgCursor& cu, int g) {
    cu.g = g; cu.t = T - 1; cu.it = 0; cu.nk = min(T * CPT, KC);
    const int j = g / m_tiles, mt = g - j * m_tiles;
    cu.gA = Ks + ((size_t)j * KC * mc + (size_t)mt * TILE_M) * TILE_K;
    cu.gB = LinvT + (((size_t)j * T + cu.t) * KC) * (size_t)(TILE_N * TILE_K);
  };
  auto cursor_advance = [&](VgCursor& cu) {
    if (++cu.it < cu.nk) return;
    cu.it = 0;
    if (--cu.t >= 0) {
      cu.nk = min((cu.t + 1) * CPT, KC);
      cu.gB -= (size_t)KC * (TILE_N * TILE_K);
    } else {
      cursor_set_group(cu, cu.g + (int)gridDim.x);
    }
  };
  auto issue_load = [&](const VgCursor& cu, uint32_t n) {
    const uint32_t s = n % VG_STAGES;
    double* sA = stage_base + (size_t)s * VG_STAGE_DOUBLES;
    mbar_expect_tx(&full[s], (uint32_t)(VG_STAGE_DOUBLES * sizeof(double)));
    bulk_g2s(sA, cu.gA + (size_t)cu.it * mc * TILE_K, TILE_M * TILE_K * sizeof(double), &full[s]);
    bulk_g2s(sA + TILE_M * TILE_K, cu.gB + (size_t)cu.it * (TILE_N * TILE_K), TILE_N * TILE_K * sizeof(double), &full[s]);
  };

  if (is_producer) {
    for (int s = 0; s < VG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], VG_CONSUMER_WARPS); }
    mbar_fence_init();
    cursor_set_group(pc, blockIdx.x);
    for (; pn < (uint32_t)VG_STAGES && pn < total_chunks; ++pn) { issue_load(pc, pn); cursor_advance(pc); }
  }
  __syncthreads();

  // ---- all 8 warps consume: 16 rows x 128 cols each
  const int gq = lane >> 2, tig = lane & 3;
  const int sw = gq & 3;  // (row & 3) for rows g, g+8 of A and rows 8i+g of B alike
  const int a_off0 = (warp * 16 + gq) * TILE_K + tig;
  const int a_off1 = a_off0 + 8 * TILE_K;
  const int b_off = TILE_M * TILE_K + gq * TILE_K + tig;
  const uint32_t stage_u32 = smem_u32(stage_base);
  const uint32_t a0b = a_off0 * 8u, a1b = a_off1 * 8u, bb = b_off * 8u;
  const uint32_t colb[4] = {(uint32_t)((0 ^ sw) << 5), (uint32_t)((1 ^ sw) << 5), (uint32_t)((2 ^ sw) << 5), (uint32_t)((3 ^ sw) << 5)};

  double c[VG_SUB][4];
#pragma unroll
  for (int i = 0; i < VG_SUB; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; c[i][2] = 0.0; c[i][3] = 0.0; }

  uint32_t n = 0;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int j = g / m_tiles, mt = g - j * m_tiles;
    double s0 = 0.0, s1 = 0.0;
    for (int t = T - 1; t >= 0; --t) {
      const int nk = min((t + 1) * CPT, KC);
      const int diag_kc = t * CPT;                                   // first chunk of the diagonal block
      const int imax = min(VG_SUB, (N - t * TILE_N + 7) >> 3);       // subtiles that hold real rows of L^{-1}
      const bool partial = imax < VG_SUB;
#pragma unroll 1
      for (int it = 0; it < nk; ++it, ++n) {
        // refill the stage released VG_LAG chunks ago with the chunk VG_STAGES - VG_LAG ahead of this one
        if (is_producer && n >= (uint32_t)VG_LAG && pn < total_chunks) {
          const uint32_t r = n - VG_LAG;  // chunk whose stage is recycled
          mbar_wait(&empty[r % VG_STAGES], (r / VG_STAGES) & 1);
          issue_load(pc, pn);
          cursor_advance(pc);
          ++pn;
        }
        const uint32_t s = n % VG_STAGES;
        mbar_wait(&full[s], (n / VG_STAGES) & 1);
        const uint32_t sb = stage_u32 + s * (uint32_t)(VG_STAGE_DOUBLES * sizeof(double));
        if (it < diag_kc) {
          if (!partial) vg_chunk_left<16>(c, sb, a0b, a1b, bb, colb);
          else {
            switch ((imax + 1) >> 1) {  // ragged last tile: subtile count rounded up to even
              case 1: vg_chunk_left<2>(c, sb, a0b, a1b, bb, colb); break;
              case 2: vg_chunk_left<4>(c, sb, a0b, a1b, bb, colb); break;
              case 3: vg_chunk_left<6>(c, sb, a0b, a1b, bb, colb); break;
              case 4: vg_chunk_left<8>(c, sb, a0b, a1b, bb, colb); break;
              case 5: vg_chunk_left<10>(c, sb, a0b, a1b, bb, colb); break;
              case 6: vg_chunk_left<12>(c, sb, a0b, a1b, bb, colb); break;
              default: vg_chunk_left<14>(c, sb, a0b, a1b, bb, colb); break;
            }
          }
        } else if (!partial) {
          switch (it - diag_kc) {
            case 0: vg_chunk_fast<0, 16>(c, sb, a0b, a1b, bb, colb); break;
            case 1: vg_chunk_fast<1, 16>(c, sb, a0b, a1b, bb, colb); break;
            case 2: vg_chunk_fast<2, 16>(c, sb, a0b, a1b, bb, colb); break;
            case 3: vg_chunk_fast<3, 16>(c, sb, a0b, a1b, bb, colb); break;
            case 4: vg_chunk_fast<4, 16>(c, sb, a0b, a1b, bb, colb); break;
            case 5: vg_chunk_fast<5, 16>(c, sb, a0b, a1b, bb, colb); break;
            case 6: vg_chunk_fast<6, 16>(c, sb, a0b, a1b, bb, colb); break;
            default: vg_chunk_fast<7, 16>(c, sb, a0b, a1b, bb, colb); break;
          }
        } else {
          vg_chunk_guard(c, stage_base + (size_t)s * VG_STAGE_DOUBLES, a_off0, a_off1, b_off, sw, (it - diag_kc) * TILE_K, imax);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      // fold this n-tile's squares into the running row sums and recycle the accumulators
#pragma unroll
      for (int i = 0; i < VG_SUB; ++i) {
        s0 = fma(c[i][0], c[i][0], s0); s0 = fma(c[i][1], c[i][1], s0);
        s1 = fma(c[i][2], c[i][2], s1); s1 = fma(c[i][3], c[i][3], s1);
        c[i][0] = 0.0; c[i][1] = 0.0; c[i][2] = 0.0; c[i][3] = 0.0;
      }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    if (tig == 0) {
      double* dst = q_out + (size_t)j * mc + (size_t)mt * TILE_M + warp * 16 + gq;
      dst[0] = s0;
      dst[8] = s1;
    }
  }
}

int launch_vargemm(Ctx* ctx, const Gp* gp, const double* Ks, int64_t mc, int64_t m_tiles, double* q_part, cudaStream_t st) {
  // per launch: the attribute is per device, and a process may hold contexts on several GPUs
  BOSIP_CUDA(ctx, cudaFuncSetAttribute(vargemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VG_SMEM));
  const int groups = (int)(m_tiles * gp->p);
  const int grid = groups < ctx->sms ? groups : ctx->sms;
  static const bool noimax = getenv("BOSIP_VG_NOIMAX") != nullptr;  // dev A/B switch
  vargemm_kernel<<<grid, VG_THREADS, VG_SMEM, st>>>(Ks, gp->LinvT, gp->KC, gp->T, noimax ? gp->T * TILE_N : gp->N, (long long)mc, (int)m_tiles, groups, q_part);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace bosip
