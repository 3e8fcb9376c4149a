// Cross-kernel K(X*, X) + predictive mean  (replaces BOSS.mean(model_post, X): K* formation and the GEMV
// mu = K* a, call sites /root/reference/src/posterior/likelihood.jl:8,12,21,25 and src/likelihoods/normal.jl:37-68).
//
// One CTA = 128 query points (one per thread) x one n-span of training points x one output j.
// The scaled training inputs U_j (SoA, d rows) and the weights alpha_j^2 a_j of the span are staged into shared
// memory by 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) signalled on an mbarrier; every thread then walks the span
// four training points at a time (broadcast LDS.128), evaluates the ARD SE / Matern kernel in FP64, accumulates the
// mean privately (no shuffles needed: a thread owns its query row) and, when the variance is wanted, writes the four
// K* values as one 32-byte group into the k-chunk-major, XOR-swizzled workspace the variance GEMM bulk-loads.
#include "common.cuh"
#include "fastmath.cuh"

namespace bosip {


// 4 x 4 byte transpose: out[b] = { byte b of x0, byte b of x1, byte b of x2, byte b of x3 }  (8 PRMT)
__device__ __forceinline__ void byte_transpose4(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t (&o)[4]) {
  const uint32_t a = __byte_perm(x0, x1, 0x5140), b = __byte_perm(x0, x1, 0x7362);
  const uint32_t c = __byte_perm(x2, x3, 0x5140), e = __byte_perm(x2, x3, 0x7362);
  o[0] = __byte_perm(a, c, 0x5410); o[1] = __byte_perm(a, c, 0x7632);
  o[2] = __byte_perm(b, e, 0x5410); o[3] = __byte_perm(b, e, 0x7632);
}

#ifndef KSTAR_EXP_REP
#define KSTAR_EXP_REP 1   // copies of the exp table (fastmath.cuh: replication measured, no gain)
#endif

// MODE 0: mean only; 1: FP64 K* chunks (vargemm.cu); 2: S unsigned byte slices of round(k * 2^(8S-1)) written as
// SWIZZLE_128B tile images [j][m-tile][k-block][slice][128 rows][128 B] for the INT8 engine (vargemm_i8.cu)
template <int KERNEL, int DP, int MODE, int S>
__global__ void __launch_bounds__(TILE_M) kstar_kernel(const double* __restrict__ Us, const double* __restrict__ scale,
                                                       const double* __restrict__ As, int Ns, int d,
                                                       const double* __restrict__ xq, long long m_valid, long long mc,
                                                       int nspan, double* __restrict__ Ks, double* __restrict__ mu_part) {
  constexpr bool WRITE_K = (MODE == 1);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* su = reinterpret_cast<double*>(smem_raw);          // [DP][nspan]
  double* sa = su + (size_t)DP * nspan;                       // [nspan]
  __shared__ __align__(8) uint64_t bar;
  __shared__ double exp_tab_all[EXP_TAB_N * KSTAR_EXP_REP];
  load_exp_table<KSTAR_EXP_REP>(exp_tab_all);
  const double* exp_tab = exp_table_lane<KSTAR_EXP_REP>(exp_tab_all);

  const int j = blockIdx.z, split = blockIdx.y, p = gridDim.z;
  const int n0 = split * nspan;
  const int len = min(nspan, Ns - n0);  // multiple of 16
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    const uint32_t row_bytes = (uint32_t)len * 8u;
    mbar_expect_tx(&bar, row_bytes * (DP + 1));
#pragma unroll 1
    for (int k = 0; k < DP; ++k) bulk_g2s(su + (size_t)k * nspan, Us + ((size_t)j * DP + k) * Ns + n0, row_bytes, &bar);
    bulk_g2s(sa, As + (size_t)j * Ns + n0, row_bytes, &bar);
  }

  const long long m = (long long)blockIdx.x * TILE_M + tid;
  const long long m_src = m < m_valid ? m : m_valid - 1;
  double uq[DP];
#pragma unroll
  for (int k = 0; k < DP; ++k) uq[k] = (k < d) ? xq[m_src * d + k] * scale[j * DP + k] : 0.0;

  if constexpr (MODE == 2) {  // the 32-wide loop may read up to 16 entries past `len`: give them finite values and zero weights
    if (len & 31) {
      for (int e = tid; e < 16 * (DP + 1); e += TILE_M) {
        const int k = e >> 4, i = len + (e & 15);
        if (k < DP) su[(size_t)k * nspan + i] = 0.0; else sa[i] = 0.0;
      }
    }
  }
  __syncthreads();  // barrier init visible to all waiters
  mbar_wait(&bar, 0);
  if constexpr (MODE == 2) {
    // The kernel values are formed already scaled by 2^(8S-1) (the fixed-point scale of the byte slices, folded into the exponent
    // patch of exp: one FP64 multiply per value less); the weights take the inverse power of two, so every product k * a of the
    // mean is the same Float64 number as in the other modes (exact scalings), and the mean stays bit-identical across engines.
    __syncthreads();
    for (int i = tid; i < len; i += TILE_M) sa[i] *= ((S == 8) ? 0x1p-63 : (S == 7 ? 0x1p-55 : 0x1p-47));
    __syncthreads();
  }

  double acc = 0.0;
  const int row_sw = (int)(m & 3);
  if constexpr (MODE == 2) {
    const int KB = (Ns + 127) >> 7;
    unsigned char* tile0 = reinterpret_cast<unsigned char*>(Ks) +
                           (((size_t)j * (size_t)(mc >> 7) + blockIdx.x) * KB) * (size_t)(S * 16384) + (tid >> 3) * 1024 + (tid & 7) * 128;
    // 32 training points per iteration: the two 16-byte chunks of a slice form one aligned 32-byte sector of the swizzled row
    const int len32 = (len + 31) & ~31;
    const int odd = tid & 1;  // chunk c sits at c ^ (row & 7): for odd rows the two halves of a sector swap places
#pragma unroll 1
    for (int n = 0; n < len32; n += 32) {
      uint32_t w[S][8];
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        double r2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          const double2 u01 = *reinterpret_cast<const double2*>(su + (size_t)k * nspan + n + 4 * q4);
          const double2 u23 = *reinterpret_cast<const double2*>(su + (size_t)k * nspan + n + 4 * q4 + 2);
          double t0 = uq[k] - u01.x, t1 = uq[k] - u01.y, t2 = uq[k] - u23.x, t3 = uq[k] - u23.y;
          r2[0] = fma(t0, t0, r2[0]); r2[1] = fma(t1, t1, r2[1]); r2[2] = fma(t2, t2, r2[2]); r2[3] = fma(t3, t3, r2[3]);
        }
        double kv[4];
        unsigned long long iv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { kv[e] = kernel_from_r2_t<KERNEL, KSTAR_EXP_REP, true, 8 * S - 1>(r2[e], exp_tab); iv[e] = __double2ull_rn(kv[e]); }
        const double2 a01 = *reinterpret_cast<const double2*>(sa + n + 4 * q4);
        const double2 a23 = *reinterpret_cast<const double2*>(sa + n + 4 * q4 + 2);
        acc = fma(kv[0], a01.x, acc); acc = fma(kv[1], a01.y, acc); acc = fma(kv[2], a23.x, acc); acc = fma(kv[3], a23.y, acc);
        uint32_t lo[4], hi[4];
        byte_transpose4((uint32_t)iv[0], (uint32_t)iv[1], (uint32_t)iv[2], (uint32_t)iv[3], lo);
        byte_transpose4((uint32_t)(iv[0] >> 32), (uint32_t)(iv[1] >> 32), (uint32_t)(iv[2] >> 32), (uint32_t)(iv[3] >> 32), hi);
#pragma unroll
        for (int a = 0; a < S; ++a) {  // slice a (0 = most significant) is byte S-1-a
          const int byte = S - 1 - a;
          w[a][q4] = byte < 4 ? lo[byte] : hi[byte - 4];
        }
      }
      const int ng = n0 + n;  // global training index of the 32-group (n0 is a multiple of 32 in this mode)
      unsigned char* dst = tile0 + (size_t)(ng >> 7) * (size_t)(S * 16384) + ((((((ng & 127) >> 4) ^ tid) & 7) & ~1) << 4);
#pragma unroll
      for (int a = 0; a < S; ++a) {
        const uint32_t x0 = odd ? w[a][4] : w[a][0], x1 = odd ? w[a][5] : w[a][1], x2 = odd ? w[a][6] : w[a][2], x3 = odd ? w[a][7] : w[a][3];
        const uint32_t x4 = odd ? w[a][0] : w[a][4], x5 = odd ? w[a][1] : w[a][5], x6 = odd ? w[a][2] : w[a][6], x7 = odd ? w[a][3] : w[a][7];
        asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n" ::"l"(dst + (size_t)a * 16384), "r"(x0), "r"(x1), "r"(x2), "r"(x3),
                     "r"(x4), "r"(x5), "r"(x6), "r"(x7) : "memory");
      }
    }
    mu_part[((size_t)split * p + j) * mc + m] = acc;
    return;
  }
#pragma unroll 1
  for (int n = 0; n < len; n += 4) {
    double r2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < DP; ++k) {
      const double2 u01 = *reinterpret_cast<const double2*>(su + (size_t)k * nspan + n);
      const double2 u23 = *reinterpret_cast<const double2*>(su + (size_t)k * nspan + n + 2);
      double t0 = uq[k] - u01.x, t1 = uq[k] - u01.y, t2 = uq[k] - u23.x, t3 = uq[k] - u23.y;
      r2[0] = fma(t0, t0, r2[0]); r2[1] = fma(t1, t1, r2[1]); r2[2] = fma(t2, t2, r2[2]); r2[3] = fma(t3, t3, r2[3]);
    }
    double kv[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) kv[e] = kernel_from_r2_t<KERNEL, KSTAR_EXP_REP>(r2[e], exp_tab);
    const double2 a01 = *reinterpret_cast<const double2*>(sa + n);
    const double2 a23 = *reinterpret_cast<const double2*>(sa + n + 2);
    acc = fma(kv[0], a01.x, acc); acc = fma(kv[1], a01.y, acc); acc = fma(kv[2], a23.x, acc); acc = fma(kv[3], a23.y, acc);
    if (WRITE_K) {
      const int ng = n0 + n;            // global training index of the group (multiple of 4)
      const int kc = ng >> 4;           // k-chunk
      const int grp = ((ng >> 2) ^ row_sw) & 3;
      double* dst = Ks + (((size_t)j * (Ns >> 4) + kc) * mc + m) * TILE_K + grp * 4;
      *reinterpret_cast<double2*>(dst) = make_double2(kv[0], kv[1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(kv[2], kv[3]);
    }
  }
  mu_part[((size_t)split * p + j) * mc + m] = acc;
}

template <int KERNEL, int DP, int MODE, int S>
static int launch_kds(Ctx* ctx, const Gp* gp, const double* xq, int64_t m_valid, int64_t mc, int64_t m_tiles, int nsplit, int nspan,
                      void* Ks, double* mu_part, cudaStream_t st) {
  dim3 grid((unsigned)m_tiles, (unsigned)nsplit, (unsigned)gp->p);
  size_t smem = (size_t)(DP + 1) * nspan * sizeof(double);
  auto kern = kstar_kernel<KERNEL, DP, MODE, S>;
  BOSIP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, TILE_M, smem, st>>>(gp->Us, gp->scale, gp->As, gp->Ns, gp->d, xq, m_valid, mc, nspan, reinterpret_cast<double*>(Ks), mu_part);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

template <int KERNEL, int DP>
static int launch_kd(Ctx* ctx, const Gp* gp, const double* xq, int64_t m_valid, int64_t mc, int64_t m_tiles, int nsplit, int nspan,
                     void* Ks, double* mu_part, int write_k, int slices, cudaStream_t st) {
  if (write_k == 0) return launch_kds<KERNEL, DP, 0, 0>(ctx, gp, xq, m_valid, mc, m_tiles, nsplit, nspan, Ks, mu_part, st);
  if (write_k == 1) return launch_kds<KERNEL, DP, 1, 0>(ctx, gp, xq, m_valid, mc, m_tiles, nsplit, nspan, Ks, mu_part, st);
  switch (slices) {
    case 6: return launch_kds<KERNEL, DP, 2, 6>(ctx, gp, xq, m_valid, mc, m_tiles, nsplit, nspan, Ks, mu_part, st);
    case 7: return launch_kds<KERNEL, DP, 2, 7>(ctx, gp, xq, m_valid, mc, m_tiles, nsplit, nspan, Ks, mu_part, st);
    case 8: return launch_kds<KERNEL, DP, 2, 8>(ctx, gp, xq, m_valid, mc, m_tiles, nsplit, nspan, Ks, mu_part, st);
  }
  BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unsupported slice count");
}

template <int KERNEL>
static int launch_k(Ctx* ctx, const Gp* gp, const double* xq, int64_t m_valid, int64_t mc, int64_t m_tiles, int nsplit, int nspan,
                    void* Ks, double* mu_part, int write_k, int slices, cudaStream_t st) {
#define BOSIP_DP_CASE(DPV) \
  case DPV: return launch_kd<KERNEL, DPV>(ctx, gp, xq, m_valid, mc, m_tiles, nsplit, nspan, Ks, mu_part, write_k, slices, st);
  switch (gp->dp) {
    BOSIP_DP_CASE(2) BOSIP_DP_CASE(3) BOSIP_DP_CASE(4) BOSIP_DP_CASE(5) BOSIP_DP_CASE(6) BOSIP_DP_CASE(8)
    BOSIP_DP_CASE(10) BOSIP_DP_CASE(12) BOSIP_DP_CASE(16) BOSIP_DP_CASE(24) BOSIP_DP_CASE(32)
  }
#undef BOSIP_DP_CASE
  BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unsupported padded dimension");
}

// The 165 instantiations (3 kernels x 11 padded dimensions x 5 output modes) dominate the build time, so build.py compiles this file
// once per kernel family (-DKSTAR_KERNEL_ID=0/1/2, in parallel); without the macro one translation unit holds all of them.
#define KSTAR_ARGS_DECL Ctx* ctx, const Gp* gp, const double* xq, int64_t m_valid, int64_t mc, int64_t m_tiles, int nsplit, int nspan, \
                        void* Ks, double* mu_part, int write_k, int slices, cudaStream_t st
#define KSTAR_ARGS ctx, gp, xq, m_valid, mc, m_tiles, nsplit, nspan, Ks, mu_part, write_k, slices, st
#ifdef KSTAR_KERNEL_ID
int launch_kstar_k0(KSTAR_ARGS_DECL);
int launch_kstar_k1(KSTAR_ARGS_DECL);
int launch_kstar_k2(KSTAR_ARGS_DECL);
#if KSTAR_KERNEL_ID == 0
int launch_kstar_k0(KSTAR_ARGS_DECL) { return launch_k<0>(KSTAR_ARGS); }
#elif KSTAR_KERNEL_ID == 1
int launch_kstar_k1(KSTAR_ARGS_DECL) { return launch_k<1>(KSTAR_ARGS); }
#else
int launch_kstar_k2(KSTAR_ARGS_DECL) { return launch_k<2>(KSTAR_ARGS); }
#endif
#else
static int launch_kstar_k0(KSTAR_ARGS_DECL) { return launch_k<0>(KSTAR_ARGS); }
static int launch_kstar_k1(KSTAR_ARGS_DECL) { return launch_k<1>(KSTAR_ARGS); }
static int launch_kstar_k2(KSTAR_ARGS_DECL) { return launch_k<2>(KSTAR_ARGS); }
#endif

#if !defined(KSTAR_KERNEL_ID) || KSTAR_KERNEL_ID == 0
int launch_kstar(KSTAR_ARGS_DECL) {
  static_assert(BOSIP_B200_KERNEL_SE == 0 && BOSIP_B200_KERNEL_MATERN32 == 1 && BOSIP_B200_KERNEL_MATERN52 == 2, "kernel ids");
  switch (gp->kernel_id) {
    case BOSIP_B200_KERNEL_SE: return launch_kstar_k0(KSTAR_ARGS);
    case BOSIP_B200_KERNEL_MATERN32: return launch_kstar_k1(KSTAR_ARGS);
    case BOSIP_B200_KERNEL_MATERN52: return launch_kstar_k2(KSTAR_ARGS);
  }
  BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "unknown kernel id");
}
#endif

}  // namespace bosip
