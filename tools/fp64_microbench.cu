// FP64 pipe microbenchmarks for B200 (sm_100a): DFMA peak, DMMA peak per mma.sync shape,
// DMMA+DFMA co-issue, fragment-layout check, exp/sqrt element rates.
// Standalone (no torch): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
// The measured DMMA figure is the roofline denominator for the variance GEMM
// (MEASURED_PEAKS.json holds no FP64 number; SURVEY.md §8d asks for this microbenchmark).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma_m8n8k4(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma_m16n8k4(double (&c)[4], const double (&a)[2], double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma_m16n8k8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma_m16n8k16(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
               "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---------------------------------------------------------------- DFMA peak
template <int CHAINS>
__global__ void __launch_bounds__(1024) k_dfma(double* out, int iters, double x, double y) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], x, y);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- DMMA peak
// SHAPE: 0=m8n8k4 1=m16n8k4 2=m16n8k8 3=m16n8k16 ; TILES independent accumulator tiles per warp
template <int SHAPE, int TILES>
__global__ void __launch_bounds__(1024) k_dmma(double* out, int iters, double av, double bv) {
  double c[TILES][4];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c[t][0] = threadIdx.x; c[t][1] = t; c[t][2] = 1; c[t][3] = 2; }
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = av + i * 1e-3 + (threadIdx.x & 31) * 1e-4;
#pragma unroll
  for (int i = 0; i < 4; ++i) b[i] = bv + i * 1e-3 + (threadIdx.x & 31) * 1e-5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      if (SHAPE == 0) { double (&cc)[2] = *reinterpret_cast<double (*)[2]>(&c[t][0]); mma_m8n8k4(cc, a[0], b[0]); }
      if (SHAPE == 1) { double (&aa)[2] = *reinterpret_cast<double (*)[2]>(&a[0]); mma_m16n8k4(c[t], aa, b[0]); }
      if (SHAPE == 2) { double (&aa)[4] = *reinterpret_cast<double (*)[4]>(&a[0]);
                        double (&bb)[2] = *reinterpret_cast<double (*)[2]>(&b[0]); mma_m16n8k8(c[t], aa, bb); }
      if (SHAPE == 3) { mma_m16n8k16(c[t], a, b); }
    }
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
  if (s == 12345.678) out[0] = s;
}

// DMMA (m16n8k4) + independent DFMA chains in the same warp: do the two share one pipe?
template <int TILES, int CHAINS>
__global__ void __launch_bounds__(1024) k_mixed(double* out, int iters, double av, double bv) {
  double c[TILES][4];
#pragma unroll
  for (int t = 0; t < TILES; ++t) { c[t][0] = threadIdx.x; c[t][1] = t; c[t][2] = 1; c[t][3] = 2; }
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-9 + i;
  double a[2] = {av, av + 1e-3}, b = bv;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < TILES; ++t) mma_m16n8k4(c[t], a, b);
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int t = 0; t < TILES; ++t) s += c[t][0] + c[t][1] + c[t][2] + c[t][3];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- transcendental element rates
// MODE 0: exp(-s) ; 1: sqrt(r2) ; 2: matern52 from r2 (sqrt + poly + exp)
template <int MODE, int ILP>
__global__ void __launch_bounds__(256) k_transc(double* out, int iters, double x0) {
  double v[ILP], acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { v[i] = x0 + 1e-3 * i + 1e-6 * threadIdx.x; acc[i] = 0; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      double r;
      if (MODE == 0) r = exp(-v[i]);
      if (MODE == 1) r = sqrt(v[i]);
      if (MODE == 2) { double s = sqrt(v[i]); r = (1.0 + s + v[i] * (1.0 / 3.0)) * exp(-s); }
      acc[i] += r;
      v[i] += 1e-7;
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  if (s == 12345.678) out[0] = s;
}

// ---------------------------------------------------------------- fragment layout check
// C[16x8] = A[16xK] (row-major) * B[KxN] where B is given as Bt[8][K] (n-major, k contiguous)
template <int SHAPE>
__global__ void k_layout(const double* A, const double* Bt, double* C, int K) {
  int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  double c[4] = {0, 0, 0, 0};
  if (SHAPE == 1) {
    for (int k0 = 0; k0 < K; k0 += 4) {
      double a[2] = {A[g * K + k0 + tig], A[(g + 8) * K + k0 + tig]};
      double b = Bt[g * K + k0 + tig];
      mma_m16n8k4(c, a, b);
    }
  } else if (SHAPE == 2) {
    for (int k0 = 0; k0 < K; k0 += 8) {
      double a[4], b[2];
      for (int v1 = 0; v1 < 2; ++v1) {
        a[2 * v1] = A[g * K + k0 + tig + 4 * v1];
        a[2 * v1 + 1] = A[(g + 8) * K + k0 + tig + 4 * v1];
        b[v1] = Bt[g * K + k0 + tig + 4 * v1];
      }
      mma_m16n8k8(c, a, b);
    }
  } else if (SHAPE == 3) {
    for (int k0 = 0; k0 < K; k0 += 16) {
      double a[8], b[4];
      for (int v1 = 0; v1 < 4; ++v1) {
        a[2 * v1] = A[g * K + k0 + tig + 4 * v1];
        a[2 * v1 + 1] = A[(g + 8) * K + k0 + tig + 4 * v1];
        b[v1] = Bt[g * K + k0 + tig + 4 * v1];
      }
      mma_m16n8k16(c, a, b);
    }
  } else {  // m8n8k4: only rows 0..7
    double cc[2] = {0, 0};
    for (int k0 = 0; k0 < K; k0 += 4) mma_m8n8k4(cc, A[g * K + k0 + tig], Bt[g * K + k0 + tig]);
    c[0] = cc[0]; c[1] = cc[1];
    // rows 8..15 by a second pass
    double c2[2] = {0, 0};
    for (int k0 = 0; k0 < K; k0 += 4) mma_m8n8k4(c2, A[(g + 8) * K + k0 + tig], Bt[g * K + k0 + tig]);
    c[2] = c2[0]; c[3] = c2[1];
  }
  C[g * 8 + 2 * tig] = c[0];
  C[g * 8 + 2 * tig + 1] = c[1];
  C[(g + 8) * 8 + 2 * tig] = c[2];
  C[(g + 8) * 8 + 2 * tig + 1] = c[3];
}

template <typename F>
static float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

template <int SHAPE, int TILES>
static void run_dmma(double* d_out, int sms, int warps, int ctas_per_sm, const char* name) {
  const int iters = 20000;
  const double fma_per_mma = (SHAPE == 0 ? 8. * 8 * 4 : SHAPE == 1 ? 16. * 8 * 4 : SHAPE == 2 ? 16. * 8 * 8 : 16. * 8 * 16);
  int grid = sms * ctas_per_sm, block = warps * 32;
  float ms = time_ms([&] { k_dmma<SHAPE, TILES><<<grid, block>>>(d_out, iters, 1.0000001, 0.9999999); });
  double flops = 2.0 * fma_per_mma * TILES * (double)iters * warps * grid;
  printf("{\"bench\":\"dmma\",\"shape\":\"%s\",\"tiles\":%d,\"warps_per_cta\":%d,\"ctas_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n",
         name, TILES, warps, ctas_per_sm, ms, flops / ms * 1e-9);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d,\"cc\":\"%d.%d\",\"clock_khz\":%d}\n", prop.name, sms, prop.major, prop.minor, prop.clockRate);
  double* d_out; CK(cudaMalloc(&d_out, 1024));

  // --- layout checks
  {
    const int K = 32;
    std::vector<double> A(16 * K), Bt(8 * K), C(128), Cr(128, 0.0);
    for (int i = 0; i < 16 * K; ++i) A[i] = sin(0.37 * i) + 0.01 * i;
    for (int i = 0; i < 8 * K; ++i) Bt[i] = cos(0.73 * i) - 0.02 * i;
    for (int m = 0; m < 16; ++m) for (int n = 0; n < 8; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * Bt[n * K + k]; Cr[m * 8 + n] = s; }
    double *dA, *dB, *dC; CK(cudaMalloc(&dA, A.size() * 8)); CK(cudaMalloc(&dB, Bt.size() * 8)); CK(cudaMalloc(&dC, 128 * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bt.data(), Bt.size() * 8, cudaMemcpyHostToDevice));
    for (int shape = 0; shape < 4; ++shape) {
      CK(cudaMemset(dC, 0, 128 * 8));
      if (shape == 0) k_layout<0><<<1, 32>>>(dA, dB, dC, K);
      if (shape == 1) k_layout<1><<<1, 32>>>(dA, dB, dC, K);
      if (shape == 2) k_layout<2><<<1, 32>>>(dA, dB, dC, K);
      if (shape == 3) k_layout<3><<<1, 32>>>(dA, dB, dC, K);
      CK(cudaMemcpy(C.data(), dC, 128 * 8, cudaMemcpyDeviceToHost));
      double err = 0; for (int i = 0; i < 128; ++i) err = fmax(err, fabs(C[i] - Cr[i]));
      printf("{\"bench\":\"layout\",\"shape\":%d,\"max_abs_err\":%.3e}\n", shape, err);
    }
  }

  // --- DFMA peak
  for (int block : {256, 512, 1024}) {
    for (int cps : {1, 2}) {
      if (block * cps > 2048) continue;
      const int iters = 20000;
      float ms = time_ms([&] { k_dfma<8><<<sms * cps, block>>>(d_out, iters, 1.0000001, 1e-9); });
      double flops = 2.0 * 8 * (double)iters * block * sms * cps;
      printf("{\"bench\":\"dfma\",\"chains\":8,\"block\":%d,\"ctas_per_sm\":%d,\"ms\":%.4f,\"tflops\":%.3f}\n", block, cps, ms, flops / ms * 1e-9);
    }
  }
  // --- DMMA peak per shape
  run_dmma<0, 8>(d_out, sms, 8, 1, "m8n8k4");
  run_dmma<0, 8>(d_out, sms, 16, 1, "m8n8k4");
  run_dmma<1, 8>(d_out, sms, 8, 1, "m16n8k4");
  run_dmma<1, 8>(d_out, sms, 16, 1, "m16n8k4");
  run_dmma<1, 16>(d_out, sms, 8, 1, "m16n8k4");
  run_dmma<1, 4>(d_out, sms, 4, 1, "m16n8k4");
  run_dmma<2, 8>(d_out, sms, 8, 1, "m16n8k8");
  run_dmma<2, 8>(d_out, sms, 16, 1, "m16n8k8");
  run_dmma<3, 8>(d_out, sms, 8, 1, "m16n8k16");
  run_dmma<3, 8>(d_out, sms, 16, 1, "m16n8k16");
  run_dmma<3, 4>(d_out, sms, 4, 1, "m16n8k16");
  run_dmma<3, 2>(d_out, sms, 8, 1, "m16n8k16");
  run_dmma<3, 1>(d_out, sms, 8, 1, "m16n8k16");
  run_dmma<3, 1>(d_out, sms, 4, 1, "m16n8k16");
  // --- mixed
  {
    const int iters = 20000;
    float ms = time_ms([&] { k_mixed<8, 8><<<sms, 256>>>(d_out, iters, 1.0000001, 1e-9); });
    double f_mma = 2.0 * 512 * 8 * (double)iters * 8 * sms, f_fma = 2.0 * 8 * (double)iters * 256 * sms;
    printf("{\"bench\":\"mixed\",\"tiles\":8,\"chains\":8,\"warps\":8,\"ms\":%.4f,\"dmma_tflops\":%.3f,\"dfma_tflops\":%.3f,\"sum_tflops\":%.3f}\n",
           ms, f_mma / ms * 1e-9, f_fma / ms * 1e-9, (f_mma + f_fma) / ms * 1e-9);
    ms = time_ms([&] { k_mixed<8, 32><<<sms, 256>>>(d_out, iters, 1.0000001, 1e-9); });
    f_fma = 2.0 * 32 * (double)iters * 256 * sms;
    printf("{\"bench\":\"mixed\",\"tiles\":8,\"chains\":32,\"warps\":8,\"ms\":%.4f,\"dmma_tflops\":%.3f,\"dfma_tflops\":%.3f,\"sum_tflops\":%.3f}\n",
           ms, f_mma / ms * 1e-9, f_fma / ms * 1e-9, (f_mma + f_fma) / ms * 1e-9);
    ms = time_ms([&] { k_mixed<8, 32><<<sms, 512>>>(d_out, iters, 1.0000001, 1e-9); });
    f_mma *= 2; f_fma *= 2;
    printf("{\"bench\":\"mixed\",\"tiles\":8,\"chains\":32,\"warps\":16,\"ms\":%.4f,\"dmma_tflops\":%.3f,\"dfma_tflops\":%.3f,\"sum_tflops\":%.3f}\n",
           ms, f_mma / ms * 1e-9, f_fma / ms * 1e-9, (f_mma + f_fma) / ms * 1e-9);
  }
  // --- transcendental rates (elements/s, whole chip)
  {
    const int iters = 2000;
    const int grid = sms * 4, block = 256;
    double elems = 4.0 * iters * block * (double)grid;
    float ms = time_ms([&] { k_transc<0, 4><<<grid, block>>>(d_out, iters, 0.5); });
    printf("{\"bench\":\"exp\",\"ms\":%.4f,\"gelem_per_s\":%.2f}\n", ms, elems / ms * 1e-6);
    ms = time_ms([&] { k_transc<1, 4><<<grid, block>>>(d_out, iters, 0.5); });
    printf("{\"bench\":\"sqrt\",\"ms\":%.4f,\"gelem_per_s\":%.2f}\n", ms, elems / ms * 1e-6);
    ms = time_ms([&] { k_transc<2, 4><<<grid, block>>>(d_out, iters, 0.5); });
    printf("{\"bench\":\"matern52\",\"ms\":%.4f,\"gelem_per_s\":%.2f}\n", ms, elems / ms * 1e-6);
  }
  CK(cudaFree(d_out));
  return 0;
}
