// Does the DMMA pipe sustain 16 cycles/instruction with a GEMM-like register pattern (64 accumulators, 2 A, 16 B fragments)?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int NB>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters, double av, double bv) {
  double c[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  double b[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) b[i] = bv + i * 1e-3;
  double a0 = av, a1 = av + 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      dmma884(c[i][0], c[i][1], a0, b[i % NB]);
      dmma884(c[i][2], c[i][3], a1, b[i % NB]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}
template <int NB> void run(double* d, int sms, int warps) {
  const int iters = 4000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NB><<<sms, warps * 32>>>(d, iters, 1.0000001, 0.999999);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    cudaEventRecord(e0); k<NB><<<sms, warps * 32>>>(d, iters, 1.0000001, 0.999999); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double flops = 2.0 * 256 * 32 * (double)iters * warps * sms;
  printf("{\"pattern\":\"64acc\",\"distinct_b\":%d,\"warps\":%d,\"tflops\":%.3f}\n", NB, warps, flops / best * 1e-9);
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  double* d; cudaMalloc(&d, 64);
  run<1>(d, p.multiProcessorCount, 8); run<2>(d, p.multiProcessorCount, 8); run<16>(d, p.multiProcessorCount, 8);
  run<16>(d, p.multiProcessorCount, 4);
  return 0;
}
