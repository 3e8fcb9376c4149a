// Shared declarations for libbosip_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/bosip_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libbosip_b200 is written for sm_100a (B200) only"
#endif

namespace bosip {

// ------------------------------------------------------------------ tiling constants
constexpr int TILE_M = 128;   // query rows per variance-GEMM CTA tile and per cross-kernel CTA
constexpr int TILE_N = 128;   // L^{-1} rows (output columns) per variance-GEMM CTA tile
constexpr int TILE_K = 16;    // k-chunk per pipeline stage; one chunk of one row = 128 B = 4 groups of 4 doubles
constexpr int MAX_P = 64;     // max output dimensions held in the epilogue parameter block
constexpr int MAX_D = 32;     // max parameter dimensions

// Element (row, kk) of a [rows][16] chunk is stored at row*16 + swz(row, kk): the four 32-byte groups of a row
// are XOR-permuted by (row & 3), which makes the m8n8k4 fragment loads (8 rows x 4 consecutive doubles)
// bank-conflict free without padding, so a whole [128][16] stage is ONE contiguous 16 KB bulk copy.
__host__ __device__ __forceinline__ int swz(int row, int kk) { return ((((kk >> 2) ^ row) & 3) << 2) | (kk & 3); }

struct Ctx;

// Parameters of the fused epilogue (likelihood + prior), passed by value to kernels.
struct EpiParams {
  int p, d;
  int has_prior;            // 0: none, 1: descriptor, 2: logprior array
  int pred_noise, clamp_var;
  double z[MAX_P], so[MAX_P], kxx[MAX_P], sig2[MAX_P];
  double log_C;             // -sum log(2 sqrt(pi) so_j)      (src/likelihoods/normal.jl:75)
  int like_kind;            // BOSIP_B200_LIKE_*: 0..3 share the Gaussian-family formulas, 4 = ExpLikelihood
  int q;                    // observation dims; observation i sums grp[i] consecutive GP outputs (1 except NormalSum)
  int grp[MAX_P];
  double c_approx, c_mean, c_sq;  // additive constants of the plug-in / E[l] / E[l^2] closures (LogNormal: -sum log z_obs terms)
  int prior_kind[MAX_D];
  double prior_prm[MAX_D][4];
  double prior_const[MAX_D];  // per-dimension additive constant (uniform: -log(b-a); truncnormal: -logtp)
};

// A conditioned GP on the device.
struct Gp {
  Ctx* ctx;
  int kernel_id, d, dp, N, Ns, p, T, KC;
  bosip_b200_gp_opts opts;
  void* block = nullptr;    // the one device allocation that holds Us .. a
  double* Us = nullptr;     // [p][dp][Ns]  training inputs scaled by c_kernel / lambda_j (SoA, zero padded)
  double* scale = nullptr;  // [p][dp]      c_kernel / lambda_jk (0 for padded dims)
  double* As = nullptr;     // [p][Ns]      alpha_j^2 * a_j (zero padded)
  double* LinvT = nullptr;  // [p][T][KC][128][16] alpha_j^2 * L_j^{-1}, tile-major + swizzled (see swz)
  double* L = nullptr;      // [p][N][N] col-major lower Cholesky factor (kept for gp_get_factors)
  double* Linv = nullptr;   // [p][N][N] col-major L^{-1}
  double* a = nullptr;      // [p][N]
  double alpha2[MAX_P], sig2[MAX_P];
  // INT8 variance engine (vargemm_i8.cu), built on first use from Linv
  int i8_S = 0, i8_T64 = 0;
  size_t i8_b_per_out = 0;
  uint8_t* LinvI8 = nullptr;  // [p][sets][S][64 x 128 B]  byte slices of alpha_j^2 L_j^{-1} (SWIZZLE_128B tile images)
  double* cscale = nullptr;   // [p][T64*64]               2^(f_n - 13): row scale of the fixed-point representation
  double* i8_qscale = nullptr;   // [p]         value of one unit of the integer row sums (inside the cscale allocation)
  uint32_t* i8_nib = nullptr;    // [p][T64*8]  f_n - f_base, four bits per row (inside the cscale allocation)
  uint8_t* LinvI8P = nullptr; // [p][sets][2][S x 8 KB]    the same slices as the per-rank operand images of the paired (cta_group::2) kernel
};

struct Timing {
  bool enabled = false;
  double ms[3] = {0, 0, 0};
  int64_t launches[3] = {0, 0, 0};
};

struct Ctx {
  int device = 0;
  int sms = 0;
  std::mutex mu;
  std::string err;
  cudaStream_t s_comp = nullptr, s_own = nullptr, s_in = nullptr, s_out = nullptr;  // s_comp == s_own unless set_stream
  int64_t launches = 0;
  cudaEvent_t ev_in[2] = {}, ev_comp[2] = {}, ev_out[2] = {}, ev_t0 = nullptr, ev_t1 = nullptr, ev_dep = nullptr;
  int64_t user_chunk = 0;
  int var_engine = 0, var_slices = 7;  // BOSIP_B200_VAR_*: which kernel computes ||L^-1 k*||^2
  // INT8 engine: byte slices of K* (var_slices = slices of alpha^2 L^-1).  Default 6 against 7: K* in 2^-47 fixed point.  Its lowest byte
  // of a 7-byte K* meets only the MOST significant byte of L^-1 (pairs with a + b >= 7 are dropped anyway), which is small for all but a
  // row's largest entries, so leaving that one slice pair out (27 instead of 28 INT8 GEMMs, one operand tile in seven never written, fetched
  // or multiplied) costs 2-8 x the slicing error of the symmetric 7 x 7 scheme and stays 25 x below the 1e-10 gate
  // (tests/test_int8_model.py, tests/test_gpu_int8_engine.py).  An explicit `slices` request (6, 7, 8) selects the symmetric scheme.
  int var_a_slices = 6;
  int i8_issuers = 0;                  // MMA-issuing threads of the INT8 engine: 0 = self-check pending, else 1 or 2 (vargemm_i8.cu)
  int i8_pair = 0;                     // paired (cta_group::2) INT8 kernel: 0 = self-check pending, 1 = in use, -1 = not used
  // chunk workspace (grown on demand)
  size_t ws_bytes = 0;
  char* ws = nullptr;
  // pinned host staging for callers that hand over PAGEABLE host buffers (Julia's Array{Float64}); grown on demand
  size_t pin_bytes = 0;
  char* pin = nullptr;
  Timing timing;
};

#define BOSIP_CUDA(ctx, call)                                                                          \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess) {                                                                          \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +       \
                   std::to_string(__LINE__) + ")";                                                     \
      return (e__ == cudaErrorMemoryAllocation) ? BOSIP_B200_ENOMEM : BOSIP_B200_ECUDA;                \
    }                                                                                                  \
  } while (0)

#define BOSIP_FAIL(ctx, code, msg) \
  do {                             \
    (ctx)->err = (msg);            \
    return (code);                 \
  } while (0)

#define BOSIP_TRY(expr)          \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int pad_dim(int d) {
  const int opts[] = {2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32};
  for (int o : opts) if (d <= o) return o;
  return -1;
}

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif

// ------------------------------------------------------------------ cross-TU host entry points
// api.cu: grow-only device workspace of the context (valid until the next call that needs a bigger one)
int ensure_ws(Ctx* ctx, size_t bytes);
// fill the likelihood part of the epilogue parameters (gp_p = GP output dimension the likelihood must match)
int make_like_params(Ctx* ctx, const bosip_b200_like_normal* like, int gp_p, EpiParams* ep);
// kstar.cu
// write_k: 0 = mean only, 1 = FP64 K* chunks for vargemm.cu, 2 = byte slices for vargemm_i8.cu (Ks is then the slice buffer)
int launch_kstar(Ctx* ctx, const Gp* gp, const double* xq, int64_t m_valid, int64_t mc, int64_t m_tiles, int nsplit,
                 int nspan, void* Ks, double* mu_part, int write_k, int slices, cudaStream_t st);
// vargemm.cu
int launch_vargemm(Ctx* ctx, const Gp* gp, const double* Ks, int64_t mc, int64_t m_tiles, double* q_part, cudaStream_t st);
// vargemm_i8.cu
int i8_prepare(Ctx* ctx, Gp* gp, int S, cudaStream_t st);
int i8_nsub(const Gp* gp);
int launch_vargemm_i8(Ctx* ctx, const Gp* gp, const uint8_t* Ks8, int a_slices, int64_t mc, int64_t m_tiles, double* q_part, double* q,
                      cudaStream_t st);
int i8_peak(Ctx* ctx, double* burst, double* sustained, double* sustained_random);
// epilogue.cu
int launch_epilogue(Ctx* ctx, const Gp* gp, const EpiParams& ep, const double* xq, const double* logprior,
                    const double* mean_at_xq, int64_t m_valid, int64_t mc, int nsplit, const double* mu_part,
                    bool have_var, const double* q_part, double* mu_out, double* var_out, double* lp_out,
                    double* log_approx, double* log_mean, double* log_var, unsigned long long* counters,
                    uint64_t index_base, int bi_mode, cudaStream_t st);
int launch_bi_finish(Ctx* ctx, const EpiParams& ep, const double* xq, const double* logprior, int64_t m_valid, int n_samples,
                     double* lp_out, double* la, double* lm, double* lv, cudaStream_t st);
int launch_argmax_chunk(Ctx* ctx, const double* vals, int64_t m_valid, int64_t global_offset, int exp_vals,
                        double* scratch_val, long long* scratch_idx, double* best_val, long long* best_idx,
                        cudaStream_t st);
int launch_sum_chunk(Ctx* ctx, const double* log_post, const double* log_prior, int64_t m_valid, double* scratch,
                     double* accum, cudaStream_t st);
int launch_like_eval(Ctx* ctx, const EpiParams& ep, const double* mu, const double* var, int64_t m, double* marg_approx,
                     double* l_approx, double* marg_mean, double* l_mean, double* l_sq, double* l_var,
                     unsigned long long* counters, cudaStream_t st);
int launch_segmean_exp(Ctx* ctx, const double* v, int64_t n_seg, int64_t seg_len, double* out, cudaStream_t st);
int launch_candidates(Ctx* ctx, const bosip_b200_cand_src* src, int d, int64_t m, int64_t offset, double* out,
                      cudaStream_t st);
// fit.cu
int gp_build(Ctx* ctx, int kernel_id, int d, int N, int p, const double* X, const double* Y, const double* lambda,
             const double* alpha, const double* sigma, const double* mean_at_X, const double* L_host,
             const double* a_host, const bosip_b200_gp_opts* opts, Gp** out);
void gp_destroy(Gp* gp);
// metrics.cu
int mmd_sums(Ctx* ctx, int kernel_id, int d, const double* lambda, const double* A, int64_t n, const double* B,
             int64_t m, int mem, int rank, int world, double* sums);
int tv_stage1(Ctx* ctx, const double* lw, const double* a, const double* t, int64_t G, int mem, double* out4);
int tv_stage2(Ctx* ctx, const double* lw, const double* a, const double* t, int64_t G, int mem, double eva,
              double evt, double* out2);
int tv_full(Ctx* ctx, const double* lw, const double* a, const double* t, int64_t G, int mem, double* tv_out);
// sampler.cu
int mvn_sample(Ctx* ctx, int d, const double* mu, const double* L, uint64_t seed, int64_t offset, int64_t m, int mem, double* out);
int mix_pdf_sum(Ctx* ctx, int d, int nq, const double* mus, const double* Linvs, const double* lognorm, const double* X, int64_t m, int mem,
                int accumulate, double* delta);
int amis_log_weights(Ctx* ctx, const double* logp, const double* delta, double t, int64_t m, int mem, double* logw);
int weighted_moments(Ctx* ctx, int d, const double* X, const double* lw, int64_t m, int mem, double* mean, double* cov, double* sums,
                     double* ws);
int resample(Ctx* ctx, int d, const double* X, const double* ws, int64_t m, uint64_t seed, int64_t count, int mem, double* out, int64_t* idx);
// confset.cu
int cs_weighted_quantile(Ctx* ctx, const double* vals, const double* ws, int64_t n, int mem, double p, double* out);
int cs_cutoff_area(Ctx* ctx, const double* vals, const double* ws, int64_t n, int mem, double c, double* out);
int cs_set_iou(Ctx* ctx, const unsigned char* in_a, const unsigned char* in_b, const double* logprior, int64_t n, int mem, double* out);
int cs_confidence_iou(Ctx* ctx, const double* log_fa, const double* log_fb, const double* logprior, int64_t n, int mem, double q, double* iou,
                      double* c_a, double* c_b);
// peaks.cu
int fp64_peak(Ctx* ctx, double* dmma, double* dfma);
int math_selftest(Ctx* ctx, const double* x, int64_t n, int variant, double* e_fast, double* q_fast, double* e_ref, double* q_ref);

}  // namespace bosip
