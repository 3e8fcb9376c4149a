// Dev probe for the INT8 tcgen05 path (sm_100a): (1) correctness of the shared-memory/instruction descriptors for
// K-major SWIZZLE_128B int8 operands against a host reference, (2) issue-rate ceiling of tcgen05.mma.kind::i8 for
// N = 128 / 256 with cta_group::1.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8mma_probe tools/i8mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 2000000000ll) { printf("TIMEOUT tag %d block %d thread %d\n", tag, blockIdx.x, threadIdx.x); asm volatile("trap;"); }
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// image of a [rows][128 B] K-major tile in the SWIZZLE_128B layout: 16-byte chunk c of row r sits at chunk c ^ (r & 7)
__host__ __device__ inline size_t sw128(int r, int kbyte) { return (size_t)(r >> 3) * 1024 + (r & 7) * 128 + ((((kbyte >> 4) ^ (r & 7)) & 7) << 4) + (kbyte & 15); }

// ---------------------------------------------------------------- check kernel: D[128][N] = A[128][K] * B[N][K]^T, K = KB * 128
template <int N>
__global__ void __launch_bounds__(128) check_kernel(const int8_t* __restrict__ Aimg, const int8_t* __restrict__ Bimg, int KB, int32_t* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full, done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sA = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + (size_t)KB * 16384;
  if (tid == 0) { mbar_init(&full, 1); mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    mbar_expect_tx(&full, (uint32_t)KB * (16384 + N * 128));
    for (int kb = 0; kb < KB; ++kb) {
      bulk_g2s(sA + (size_t)kb * 16384, Aimg + (size_t)kb * 16384, 16384, &full);
      bulk_g2s(sB + (size_t)kb * N * 128, Bimg + (size_t)kb * N * 128, N * 128, &full);
    }
    mbar_wait(&full, 0, 1);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t idesc = make_idesc(128, N);
    for (int kb = 0; kb < KB; ++kb)
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_desc(smem_u32(sA + (size_t)kb * 16384) + k * 32);
        const uint64_t db = make_desc(smem_u32(sB + (size_t)kb * N * 128) + k * 32);
        mma_i8(tb, da, db, idesc, (kb | k) ? 1u : 0u);
      }
    mma_commit(&done);
  }
  __syncthreads();
  mbar_wait(&done, 0, 2);
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) D[(size_t)tid * N + c0 + i] = (int32_t)v[i];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

// ---------------------------------------------------------------- peak kernel: issue `iters` x 4 MMAs on resident operands
template <int N>
__global__ void __launch_bounds__(128) peak_kernel(int iters, int nacc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sA = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + 4 * 16384;
  for (int i = tid; i < (4 * 16384 + 4 * N * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0x01010101u * (i & 3);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, N);
    for (int it = 0; it < iters; ++it) {
      const int st = it & 3;
      const uint32_t d = tb + (uint32_t)((it % nacc) * N);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_i8(d, make_desc(smem_u32(sA + st * 16384) + k * 32), make_desc(smem_u32(sB + st * N * 128) + k * 32), idesc, 1u);
    }
    mma_commit(&done);
    mbar_wait(&done, 0, 3);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

template <int N>
static int run_check(int KB) {
  const int K = KB * 128;
  std::vector<int8_t> A((size_t)128 * K), B((size_t)N * K), Aimg((size_t)KB * 16384), Bimg((size_t)KB * N * 128);
  srand(1234 + N + KB);
  for (auto& x : A) x = (int8_t)(rand() % 129 - 64);
  for (auto& x : B) x = (int8_t)(rand() % 129 - 64);
  for (int r = 0; r < 128; ++r) for (int k = 0; k < K; ++k) Aimg[(size_t)(k >> 7) * 16384 + sw128(r, k & 127)] = A[(size_t)r * K + k];
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) Bimg[(size_t)(k >> 7) * N * 128 + sw128(r, k & 127)] = B[(size_t)r * K + k];
  int8_t *dA, *dB; int32_t* dD;
  CK(cudaMalloc(&dA, Aimg.size())); CK(cudaMalloc(&dB, Bimg.size())); CK(cudaMalloc(&dD, (size_t)128 * N * 4));
  CK(cudaMemcpy(dA, Aimg.data(), Aimg.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bimg.data(), Bimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, (size_t)128 * N * 4));
  const size_t smem = (size_t)KB * (16384 + N * 128) + 1024;
  CK(cudaFuncSetAttribute(check_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  check_kernel<N><<<1, 128, smem>>>(dA, dB, KB, dD);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> D((size_t)128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0; int first = -1;
  for (int r = 0; r < 128; ++r) for (int c = 0; c < N; ++c) {
    int32_t ref = 0;
    for (int k = 0; k < K; ++k) ref += (int32_t)A[(size_t)r * K + k] * (int32_t)B[(size_t)c * K + k];
    if (ref != D[(size_t)r * N + c]) { if (first < 0) first = r * N + c; ++bad; }
  }
  printf("check N=%d K=%d: %lld mismatches of %d%s\n", N, K, bad, 128 * N, bad ? " (FAIL)" : " (exact)");
  if (bad) printf("  first mismatch at row %d col %d: got %d\n", first / N, first % N, D[first]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return bad ? 1 : 0;
}

template <int N>
static void run_peak(int blocks, int iters, int nacc) {
  const size_t smem = 4 * 16384 + 4 * (size_t)N * 128 + 1024;
  CK(cudaFuncSetAttribute(peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  peak_kernel<N><<<blocks, 128, smem>>>(100, nacc);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a));
    peak_kernel<N><<<blocks, 128, smem>>>(iters, nacc);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  const double macs = (double)blocks * iters * 4.0 * 128.0 * N * 32.0;
  printf("peak N=%d blocks=%d iters=%d nacc=%d: %.3f ms  %.1f TOPS (2*MAC)  %.0f MAC/clk/SM @1.965GHz\n", N, blocks, iters, nacc, best,
         2.0 * macs / (best * 1e-3) / 1e12, macs / blocks / (best * 1e-3) / 1.965e9);
}


// ---------------------------------------------------------------- pattern kernel: the per-k-step MMA sequence of the S-slice engine
// mode 0: chunks of <= 4 slices from the top (4+3, 4+2, 4+1, 4, 3, 2, 1); mode 1: balanced halves (4+3, 3+3, 3+2, 4, 3, 2, 1)
__global__ void __launch_bounds__(128) pattern_kernel(int iters, int S, int mode, int nmax) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sA = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + 4 * 16384;
  for (int i = tid; i < (4 * 16384 + 8 * 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0x01010101u * (i & 3);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    for (int it = 0; it < iters; ++it) {
      const int ks = it & 3;
      for (int a = 0; a < S; ++a) {
        const uint64_t da = make_desc(smem_u32(sA + (a & 3) * 16384) + ks * 32);
        const int cnt = S - a;
        if (mode == 2) {  // single-N sweep: one MMA of N = nmax per "row"
          mma_i8(tb, da, make_desc(smem_u32(sB) + ks * 32), make_idesc(128, nmax), 1u);
          continue;
        }
        int n0 = cnt <= 4 ? cnt : (mode == 0 ? 4 : (cnt + 1) / 2), n1 = cnt - n0;
        mma_i8(tb, da, make_desc(smem_u32(sB + a * 8192) + ks * 32), make_idesc(128, 64 * n0), 1u);
        if (n1 > 0) mma_i8(tb + 64 * n0, da, make_desc(smem_u32(sB + (a + n0) * 8192) + ks * 32), make_idesc(128, 64 * n1), 1u);
      }
    }
    mma_commit(&done);
    mbar_wait(&done, 0, 4);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

static void run_pattern(int blocks, int iters, int S, int mode, int nmax) {
  const size_t smem = 4 * 16384 + 8 * 8192 + 1024;
  CK(cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  pattern_kernel<<<blocks, 128, smem>>>(50, S, mode, nmax);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    pattern_kernel<<<blocks, 128, smem>>>(iters, S, mode, nmax);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  const double clk = best * 1e-3 * 1.965e9 / iters;
  if (mode == 2) printf("single N=%d: %.1f clk per MMA (M=128,K=32)\n", nmax, clk / S);
  else printf("pattern S=%d mode=%d: %.1f clk per k-step (%d pairs; ideal %.0f)\n", S, mode, clk, S * (S + 1) / 2, S * (S + 1) / 2 * 32.0);
}

// ---------------------------------------------------------------- compile-time MMA sequences (tight issue loop): cost model of tcgen05.mma kind::i8
// PAT 0: the engine's S=7 k-step (256+192, 256+128, 256+64, 256, 192, 128, 64); 1: 7 x N=256; 2: 14 x N=128; 3: 28 x N=64;
// 4: 9 x N=192 + 1 x 64; 5: S=7 k-step with all MMAs on D column 0; 6: pattern 0 with a tcgen05.commit after every slice row
// of every 4th k-step (what the engine's A ring needs); 7: pattern 0 with a commit after every k-step only
template <int PAT>
__global__ void __launch_bounds__(128) seq_kernel(int iters) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done;
  __shared__ __align__(8) uint64_t sink[8];
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  unsigned char* sA = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + 7 * 16384;
  if (tid == 0) for (int i = 0; i < 8; ++i) mbar_init(&sink[i], 1);
  for (int i = tid; i < (7 * 16384 + 7 * 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0x01010101u * (i & 3);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int it = 0; it < iters; ++it) {
      const uint32_t ko = (it & 3) * 32;
      if (PAT == 0 || PAT == 5 || PAT == 6 || PAT == 7) {
#pragma unroll
        for (int a = 0; a < 7; ++a) {
          const uint64_t da = make_desc(a0 + a * 16384 + ko);
#pragma unroll
          for (int c0 = a; c0 < 7; c0 += 4) {
            const int nb = (7 - c0) < 4 ? (7 - c0) : 4;
            mma_i8(tb + (PAT != 5 ? 64 * (c0 - a) : 0), da, make_desc(b0 + c0 * 8192 + ko), make_idesc(128, 64 * nb), 1u);
          }
          if (PAT == 6 && (it & 3) == 3) mma_commit(&sink[a]);
        }
        if (PAT == 7) mma_commit(&sink[7]);
      } else if (PAT >= 8 && PAT <= 11) {
        // the engine's real order: one iteration = one k-block = rows a = 6..0, each row its 4 k-steps, then a commit;
        // PAT 8: commit per row + one per k-block (what vargemm_i8 does); 9: commit per 2 rows; 10: per 3 rows; 11: one per k-block
        const int every = PAT == 8 ? 1 : PAT == 9 ? 2 : PAT == 10 ? 3 : 100;
#pragma unroll
        for (int a = 6; a >= 0; --a) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = make_desc(a0 + a * 16384 + ks * 32);
#pragma unroll
            for (int c0 = a; c0 < 7; c0 += 4) {
              const int nb = (7 - c0) < 4 ? (7 - c0) : 4;
              mma_i8(tb + 64 * (c0 - a), da, make_desc(b0 + c0 * 8192 + ks * 32), make_idesc(128, 64 * nb), 1u);
            }
          }
          if ((6 - a) % every == every - 1) mma_commit(&sink[a]);
        }
        mma_commit(&sink[7]);
      } else if (PAT == 1) {
#pragma unroll
        for (int a = 0; a < 7; ++a) mma_i8(tb + (a & 1) * 256, make_desc(a0 + a * 16384 + ko), make_desc(b0 + (a % 4) * 8192 + ko), make_idesc(128, 256), 1u);
      } else if (PAT == 2) {
#pragma unroll
        for (int a = 0; a < 14; ++a) mma_i8(tb + (a & 3) * 128, make_desc(a0 + (a % 7) * 16384 + ko), make_desc(b0 + (a % 6) * 8192 + ko), make_idesc(128, 128), 1u);
      } else if (PAT == 3) {
#pragma unroll
        for (int a = 0; a < 28; ++a) mma_i8(tb + (a & 7) * 64, make_desc(a0 + (a % 7) * 16384 + ko), make_desc(b0 + (a % 7) * 8192 + ko), make_idesc(128, 64), 1u);
      } else if (PAT == 4) {
#pragma unroll
        for (int a = 0; a < 9; ++a) mma_i8(tb + (a & 1) * 192, make_desc(a0 + (a % 7) * 16384 + ko), make_desc(b0 + (a % 5) * 8192 + ko), make_idesc(128, 192), 1u);
        mma_i8(tb + 448, make_desc(a0 + ko), make_desc(b0 + ko), make_idesc(128, 64), 1u);
      }
    }
    mma_commit(&done);
    mbar_wait(&done, 0, 5);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

template <int PAT>
static void run_seq(int blocks, int iters, const char* what) {
  const size_t smem = 7 * 16384 + 7 * 8192 + 1024;
  CK(cudaFuncSetAttribute(seq_kernel<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  seq_kernel<PAT><<<blocks, 128, smem>>>(50);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(a));
    seq_kernel<PAT><<<blocks, 128, smem>>>(iters);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
  }
  printf("seq %d (%s): %.1f clk per k-step (1792 output columns; math floor 896)\n", PAT, what, best * 1e-3 * 1.965e9 / iters / (PAT >= 8 ? 4 : 1));
}

// ---------------------------------------------------------------- what other warps can do while the tensor pipe is busy
// warp 0 lane 0 issues N=256 MMAs back to back (or nothing when mma_on == 0); warps 1..4 (one per SM sub-partition) each run `reps`
// iterations of a 32-instruction dependent-free loop of one instruction type and report their cycles.
template <int KIND>  // 0 DFMA, 1 FFMA, 2 IMAD (32-bit), 3 IMAD.WIDE, 4 SHFL, 5 LOP3/IADD3 mix, 6 LDS
__global__ void __launch_bounds__(160) contend_kernel(int iters, int mma_on, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base;
  __shared__ double sd[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned char* sA = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + 4 * 16384;
  for (int i = tid; i < (4 * 16384 + 4 * 256 * 128) / 4; i += 160) reinterpret_cast<uint32_t*>(sA)[i] = 0x01010101u * (i & 3);
  if (tid < 64) sd[tid] = tid;
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) { mbar_init(&done, 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (warp == 0) {
    if (lane == 0 && mma_on) {
      for (int it = 0; it < iters; ++it) {
        const int st = it & 3;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          mma_i8(tb + (uint32_t)((it & 1) * 256), make_desc(smem_u32(sA + st * 16384) + k * 32), make_desc(smem_u32(sB + st * 256 * 128) + k * 32), make_idesc(128, 256), 1u);
      }
      mma_commit(&done);
      mbar_wait(&done, 0, 6);
    }
  } else {
    double d0 = lane, d1 = 1.5, d2 = 0.25, d3 = 3.0; float f0 = lane, f1 = 1.5f, f2 = 0.25f, f3 = 3.f;
    int i0 = lane, i1 = 3, i2 = 5, i3 = 7; long long l0 = lane, l1 = 1, l2 = 2, l3 = 3;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (KIND == 0) { d0 = fma(d0, 1.0000001, 0.5); d1 = fma(d1, 1.0000001, 0.5); d2 = fma(d2, 1.0000001, 0.5); d3 = fma(d3, 1.0000001, 0.5); }
        if (KIND == 1) { f0 = fmaf(f0, 1.0001f, 0.5f); f1 = fmaf(f1, 1.0001f, 0.5f); f2 = fmaf(f2, 1.0001f, 0.5f); f3 = fmaf(f3, 1.0001f, 0.5f); }
        if (KIND == 2) { i0 = i0 * 3 + 1; i1 = i1 * 3 + 1; i2 = i2 * 3 + 1; i3 = i3 * 3 + 1; }
        if (KIND == 3) { l0 = (long long)i0 * 257 + l0; l1 = (long long)i1 * 257 + l1; l2 = (long long)i2 * 257 + l2; l3 = (long long)i3 * 257 + l3; i0 ^= (int)l3; }
        if (KIND == 4) { i0 = __shfl_sync(0xffffffffu, i0, (lane + 1) & 31); i1 = __shfl_sync(0xffffffffu, i1, (lane + 2) & 31); i2 = __shfl_sync(0xffffffffu, i2, (lane + 3) & 31); i3 = __shfl_sync(0xffffffffu, i3, (lane + 4) & 31); }
        if (KIND == 5) { i0 = (i0 ^ i1) + i2; i1 = (i1 & i3) + i0; i2 = (i2 | i0) - i1; i3 = (i3 ^ i2) + 7; }
        if (KIND == 6) { d0 += sd[(i0++) & 63]; d1 += sd[(i1++) & 63]; d2 += sd[(i2++) & 63]; d3 += sd[(i3++) & 63]; }
        if (KIND == 7) { l0 += __double_as_longlong(__ll2double_rn(l0)) >> 40; l1 += __double_as_longlong(__ll2double_rn(l1)) >> 40; l2 += __double_as_longlong(__ll2double_rn(l2)) >> 40; l3 += __double_as_longlong(__ll2double_rn(l3)) >> 40; }
        if (KIND == 8) { i0 += __double2hiint(__int2double_rn(i0)) >> 20; i1 += __double2hiint(__int2double_rn(i1)) >> 20; i2 += __double2hiint(__int2double_rn(i2)) >> 20; i3 += __double2hiint(__int2double_rn(i3)) >> 20; }
      }
    }
    const long long t1 = clock64();
    if (lane == 0) out[warp - 1] = t1 - t0;
    if (d0 + d1 + d2 + d3 + f0 + f1 + f2 + f3 + i0 + i1 + i2 + i3 + l0 + l1 + l2 + l3 == 12345.678) out[7] = 1;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

template <int KIND>
static void run_contend(const char* what) {
  const size_t smem = 4 * 16384 + 4 * 256 * 128 + 1024;
  CK(cudaFuncSetAttribute(contend_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* d; CK(cudaMalloc(&d, 64)); long long h[2][8];
  for (int on = 0; on < 2; ++on) {
    CK(cudaMemset(d, 0, 64));
    contend_kernel<KIND><<<1, 160, smem>>>(4000, on, 2000, d);   // 16000 MMAs of 134 clk cover the 2000 x 32 instructions of the other warps
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h[on], d, 64, cudaMemcpyDeviceToHost));
  }
  printf("contend %-9s: clk per instruction per warp, tensor idle %.2f %.2f %.2f %.2f | tensor busy %.2f %.2f %.2f %.2f  (warps on sub-partitions 1,2,3,0)\n", what,
         h[0][0] / 64000.0, h[0][1] / 64000.0, h[0][2] / 64000.0, h[0][3] / 64000.0, h[1][0] / 64000.0, h[1][1] / 64000.0, h[1][2] / 64000.0, h[1][3] / 64000.0);
  cudaFree(d);
}

int main(int argc, char** argv) {
  if (argc > 1 && !strcmp(argv[1], "contend")) {
    run_contend<0>("DFMA"); run_contend<1>("FFMA"); run_contend<2>("IMAD"); run_contend<3>("IMAD.WIDE"); run_contend<4>("SHFL"); run_contend<5>("LOP/IADD"); run_contend<6>("LDS+DADD"); run_contend<7>("I2F.64.s64"); run_contend<8>("I2F.64.s32");
    return 0;
  }
  int rc = 0;
  rc |= run_check<128>(1);
  rc |= run_check<128>(3);
  rc |= run_check<256>(2);
  rc |= run_check<64>(2);
  if (rc) { printf("descriptor check FAILED; skipping peak\n"); return 1; }
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  run_peak<128>(1, 20000, 1);
  run_peak<128>(1, 20000, 4);
  run_peak<256>(1, 10000, 2);
  run_peak<128>(sms, 20000, 4);
  run_peak<256>(sms, 10000, 2);
  run_peak<64>(sms, 20000, 4);
  run_seq<0>(sms, 3000, "S=7 engine pattern"); run_seq<5>(sms, 3000, "same, all on D column 0"); run_seq<1>(sms, 3000, "7 x N=256");
  run_seq<6>(sms, 3000, "S=7 pattern + commit per slice row every 4th k-step"); run_seq<7>(sms, 3000, "S=7 pattern + commit per k-step");
  run_seq<8>(sms, 1000, "engine order, commit per row + per k-block"); run_seq<9>(sms, 1000, "engine order, commit per 2 rows"); run_seq<10>(sms, 1000, "engine order, commit per 3 rows"); run_seq<11>(sms, 1000, "engine order, one commit per k-block");
  run_seq<2>(sms, 3000, "14 x N=128"); run_seq<3>(sms, 3000, "28 x N=64"); run_seq<4>(sms, 3000, "9 x N=192 + N=64");
  for (int n = 32; n <= 256; n += 32) run_pattern(sms, 4000, 4, 2, n);
  for (int S = 6; S <= 8; ++S) for (int mode = 0; mode < 2; ++mode) run_pattern(sms, 2000, S, mode, 0);
  return 0;
}
