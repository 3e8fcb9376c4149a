// FP64 pipe microbenchmark: the roofline denominators for this path.  MEASURED_PEAKS.json holds HBM GB/s and bf16
// TFLOP/s only; SURVEY.md section 8(d) asks for a DMMA and a DFMA figure measured on the box.
// (tools/fp64_microbench.cu is the long form: all mma.sync f64 shapes lower to DMMA.8x8x4 on sm_100a.)
#include "common.cuh"
#include "fastmath.cuh"

namespace bosip {

__global__ void __launch_bounds__(256) peak_dmma_kernel(double* out, int iters, double av, double bv) {
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) { c[t][0] = threadIdx.x; c[t][1] = t; }
  const double a = av + (threadIdx.x & 31) * 1e-4, b = bv + (threadIdx.x & 31) * 1e-5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) dmma884(c[t][0], c[t][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) peak_dfma_kernel(double* out, int iters, double x, double y) {
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

int fp64_peak(Ctx* ctx, double* dmma, double* dfma) {
  cudaStream_t st = ctx->s_comp;
  double* d_out = nullptr;
  BOSIP_CUDA(ctx, cudaMalloc(&d_out, 64));
  const int iters = 20000, grid = ctx->sms * 2, block = 256;
  float best_mma = 1e30f, best_fma = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    float ms = 0;
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t0, st));
    peak_dmma_kernel<<<grid, block, 0, st>>>(d_out, iters, 1.0000001, 0.9999999);
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t1, st));
    BOSIP_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    BOSIP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    if (rep > 0 && ms < best_mma) best_mma = ms;
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t0, st));
    peak_dfma_kernel<<<grid, block, 0, st>>>(d_out, iters, 1.0000001, 1e-9);
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t1, st));
    BOSIP_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    BOSIP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    if (rep > 0 && ms < best_fma) best_fma = ms;
  }
  cudaFree(d_out);
  const double warps = (double)grid * block / 32;
  if (dmma) *dmma = 2.0 * 256.0 * 8 * iters * warps / best_mma * 1e-9;
  if (dfma) *dfma = 2.0 * 8 * iters * (double)grid * block / best_fma * 1e-9;
  return 0;
}

// self-test of fastmath.cuh against the CUDA library (host buffers)
__global__ void math_selftest_kernel(const double* __restrict__ x, long long n, int variant, double* __restrict__ e_fast,
                                     double* __restrict__ q_fast, double* __restrict__ e_ref, double* __restrict__ q_ref) {
  __shared__ double tab_all[EXP_TAB_N * 8];
  load_exp_table<8>(tab_all);
  const double* tab = exp_table_lane<8>(tab_all);
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = x[i];
  if (variant == 0) { e_fast[i] = fast_exp_neg(v); q_fast[i] = fast_sqrt(v); }
  else { e_fast[i] = fast_exp_neg_t<8>(v, tab); q_fast[i] = fast_sqrt1(v); }
  e_ref[i] = exp(-v); q_ref[i] = sqrt(v);
}

int math_selftest(Ctx* ctx, const double* x, int64_t n, int variant, double* e_fast, double* q_fast, double* e_ref, double* q_ref) {
  if (n <= 0) return 0;
  cudaStream_t st = ctx->s_comp;
  double* d = nullptr;
  BOSIP_CUDA(ctx, cudaMalloc(&d, (size_t)n * 5 * 8));
  cudaError_t e = cudaMemcpyAsync(d, x, (size_t)n * 8, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    math_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, (long long)n, variant, d + n, d + 2 * n, d + 3 * n, d + 4 * n);
    double* outs[4] = {e_fast, q_fast, e_ref, q_ref};
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) e = cudaMemcpyAsync(outs[k], d + (size_t)(k + 1) * n, (size_t)n * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  }
  cudaFree(d);
  BOSIP_CUDA(ctx, e);
  return 0;
}

}  // namespace bosip
