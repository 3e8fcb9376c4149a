// Dev probe for tcgen05.mma.cta_group::2.kind::i8 (CTA pairs, M = 256) on sm_100a -- the semantics the paired INT8 variance GEMM
// relies on, pinned against a host reference before the real kernel uses them:
//   (1) TMEM allocation with cta_group::2 from the same warp of both CTAs,
//   (2) which rows of B each CTA contributes (rank 0: D columns [0, N/2), rank 1: [N/2, N)) and where D lands (each CTA's own TMEM),
//   (3) a cp.async.bulk issued by the PEER CTA into its own shared memory that completes its bytes on the LEADER's mbarrier,
//   (4) tcgen05.commit ... multicast::cluster arriving on a barrier in both CTAs, remote mbarrier.arrive from the peer,
//   (5) per-instruction cost of the MMA shapes of the slice engine in pair mode (operands resident, one issuing thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8mma2_probe tools/i8mma2_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) { asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope: the arrive came from the peer CTA
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, int tag) {
  long long t0 = clock64();
  while (!mbar_try_cluster(bar, parity)) {
    if (clock64() - t0 > 400000000ll) { printf("TIMEOUT tag %d block %d thread %d\n", tag, blockIdx.x, threadIdx.x); return false; }
  }
  return true;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 400000000ll) { printf("TIMEOUT tag %d block %d thread %d\n", tag, blockIdx.x, threadIdx.x); return false; }
  }
  return true;
}
// dst and bar are shared::cluster addresses (a shared::cta address is the executing CTA's own window)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {   // D = s32, A = s8 (bit 7), B = s8 (bit 10)
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma2_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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
__host__ __device__ inline size_t sw128(int r, int kbyte) { return (size_t)(r >> 3) * 1024 + (r & 7) * 128 + ((((kbyte >> 4) ^ (r & 7)) & 7) << 4) + (kbyte & 15); }

// ---------------------------------------------------------------- check: D[256][N] = A[256][128] * B[N][128]^T on a CTA pair
// rank r holds A rows [128 r, 128 r + 128) and B rows [r N/2, (r+1) N/2); remote_tx != 0: the peer's copies complete on the LEADER's barrier
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) check2_kernel(const int8_t* __restrict__ Aimg, const int8_t* __restrict__ Bimg,
                                                                                 int32_t* __restrict__ D, int remote_tx, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full, done, peer_ready;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cta_rank();
  unsigned char* sA = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + 16384;
  if (tid == 0) {
    // leader's `full`: one arrive (its own expect_tx) ; bytes of both CTAs when remote_tx, else only its own and the peer says "ready" by a remote arrive
    mbar_init(&full, 1); mbar_init(&done, 1); mbar_init(&peer_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  const uint32_t my_bytes = 16384 + (N / 2) * 128;
  bool ok = true;
  if (tid == 0) {
    const int8_t* Asrc = Aimg + (size_t)rank * 16384;
    const int8_t* Bsrc = Bimg + (size_t)rank * (N / 2) * 128;
    if (rank == 0) {
      mbar_expect_tx(&full, remote_tx ? 2 * my_bytes : my_bytes);
      bulk_g2s(smem_u32(sA), Asrc, 16384, smem_u32(&full));
      bulk_g2s(smem_u32(sB), Bsrc, (N / 2) * 128, smem_u32(&full));
    } else if (remote_tx) {
      const uint32_t lead_full = mapa(smem_u32(&full), 0);
      bulk_g2s(mapa(smem_u32(sA), 1), Asrc, 16384, lead_full);
      bulk_g2s(mapa(smem_u32(sB), 1), Bsrc, (N / 2) * 128, lead_full);
    } else {
      mbar_expect_tx(&full, my_bytes);
      bulk_g2s(smem_u32(sA), Asrc, 16384, smem_u32(&full));
      bulk_g2s(smem_u32(sB), Bsrc, (N / 2) * 128, smem_u32(&full));
      ok = mbar_wait(&full, 0, 10);
      mbar_arrive_remote(mapa(smem_u32(&peer_ready), 0));   // release.cluster: the landed tile is visible to the leader's async proxy? (tested)
    }
    if (rank == 0) {
      ok = mbar_wait(&full, 0, 1);
      if (!remote_tx) ok = ok && mbar_wait_cluster(&peer_ready, 0, 2);
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const uint32_t idesc = make_idesc(256, N);
      if (ok)
        for (int k = 0; k < 4; ++k) mma2_i8(tb, make_desc(smem_u32(sA) + k * 32), make_desc(smem_u32(sB) + k * 32), idesc, k ? 1u : 0u);
      mma2_commit_mc(&done, 0b11);
    }
  }
  __syncthreads();
  ok = mbar_wait(&done, 0, 3) && ok;
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 32; ++i) D[((size_t)rank * 128 + tid) * N + c0 + i] = (int32_t)v[i];
  }
  if (!ok) atomicAdd(status, 1);
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

template <int N>
static int run_check2(int remote_tx) {
  const int K = 128;
  std::vector<int8_t> A((size_t)256 * K), B((size_t)N * K), Aimg((size_t)2 * 16384), Bimg((size_t)N * 128);
  srand(99 + N + remote_tx);
  for (auto& x : A) x = (int8_t)(rand() % 255 - 127);
  for (auto& x : B) x = (int8_t)(rand() % 255 - 127);
  for (int r = 0; r < 256; ++r) for (int k = 0; k < K; ++k) Aimg[(size_t)(r >> 7) * 16384 + sw128(r & 127, k)] = A[(size_t)r * K + k];
  for (int r = 0; r < N; ++r) for (int k = 0; k < K; ++k) Bimg[sw128(r, k)] = B[(size_t)r * K + k];
  int8_t *dA, *dB; int32_t* dD; int* dS;
  CK(cudaMalloc(&dA, Aimg.size())); CK(cudaMalloc(&dB, Bimg.size())); CK(cudaMalloc(&dD, (size_t)256 * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, Aimg.data(), Aimg.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bimg.data(), Bimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, (size_t)256 * N * 4)); CK(cudaMemset(dS, 0, 4));
  const size_t smem = 16384 + (size_t)(N / 2) * 128 + 1024;
  CK(cudaFuncSetAttribute(check2_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  check2_kernel<N><<<2, 128, smem>>>(dA, dB, dD, remote_tx, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("check2 N=%d remote_tx=%d: launch failed: %s\n", N, remote_tx, cudaGetErrorString(e)); return 2; }
  std::vector<int32_t> D((size_t)256 * N); int st = 0;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  long long bad = 0; int first = -1;
  for (int r = 0; r < 256; ++r) for (int c = 0; c < N; ++c) {
    int32_t ref = 0;
    for (int k = 0; k < K; ++k) ref += (int32_t)A[(size_t)r * K + k] * (int32_t)B[(size_t)c * K + k];
    if (ref != D[(size_t)r * N + c]) { if (first < 0) first = r * N + c; ++bad; }
  }
  printf("check2 N=%d remote_tx=%d: %lld mismatches of %d%s, timeouts %d\n", N, remote_tx, bad, 256 * N, bad ? " (FAIL)" : " (exact)", st);
  if (bad) printf("  first mismatch at row %d col %d: got %d\n", first / N, first % N, D[first]);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return bad ? 1 : 0;
}

// ---------------------------------------------------------------- cost of the MMA shapes in pair mode (operands resident)
// pattern 0: `iters` x 4 k-steps of one shape N; pattern 1: the 12-MMA k-step of the paired S = 7 engine (4 x 256, 4 x 128, 4 x 64);
// pattern 2: the 10-MMA k-step of the unpaired engine issued in pair mode (256+192, 256+128, 256+64, 256, 192, 128, 64)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) cost2_kernel(int iters, int pattern, int N, long long* __restrict__ clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cta_rank();
  unsigned char* sA = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + 7 * 16384;
  for (int i = tid; i < (7 * 16384 + 7 * 8192) / 4; i += 128) reinterpret_cast<uint32_t*>(sA)[i] = 0x01010101u * (i & 3);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) { mbar_init(&done, pattern == 4 ? 2 : 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base;
  if (pattern >= 3) {   // the real kernel's order: row-outer, k-step inner; pattern 3: one issuing thread, 4: two (rows {6,5,3,0} / {4,2,1})
    __shared__ long long t_start;
    if (rank == 0 && (tid == 0 || (pattern == 4 && tid == 32))) {
      const int q = tid ? 1 : 0;
      const long long t0 = clock64();
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      for (int it = 0; it < iters; ++it) {
        for (int a = 6; a >= 0; --a) {
          const int own = (a == 0 || a == 3 || a == 5 || a == 6) ? 0 : 1;
          if (pattern == 4 && own != q) continue;
          const uint32_t A = a0 + a * 16384;
          for (int k = 0; k < 4; ++k) {
            if (a <= 3) mma2_i8(tb, make_desc(A + k * 32), make_desc(b0 + a * 8192 + k * 32), make_idesc(256, 256), 1u);
            if (a == 0 || a == 1 || a == 4 || a == 5) mma2_i8(tb + 64 * (5 - a), make_desc(A + k * 32), make_desc(b0 + 5 * 8192 + k * 32), make_idesc(256, 128), 1u);
            if (a == 0 || a == 4) mma2_i8(tb + 64 * (4 - a), make_desc(A + k * 32), make_desc(b0 + 6 * 8192 + 4096 + k * 32), make_idesc(256, 64), 1u);
            if (a == 2 || a == 6) mma2_i8(tb + 64 * (6 - a), make_desc(A + k * 32), make_desc(b0 + 6 * 8192 + k * 32), make_idesc(256, 64), 1u);
          }
        }
      }
      mma2_commit_mc(&done, 0b11);
      if (tid == 0) t_start = t0;
    }
    if (tid == 0) { mbar_wait(&done, 0, 3); if (rank == 0) clk[0] = (clock64() - t_start) * 4; }   // x4: the host divides by iters*4 (k-steps), a k-block is 4 k-steps
    __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
    return;
  }
  if (tid == 0 && rank == 0) {
    const long long t0 = clock64();
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int it = 0; it < iters; ++it) {
      for (int k = 0; k < 4; ++k) {
        if (pattern == 0) {
          mma2_i8(tb, make_desc(a0 + (it & 3) * 16384 + k * 32), make_desc(b0 + k * 32), make_idesc(256, N), 1u);
        } else if (pattern == 1) {
          // rows 0..3: N = 256 on slots a..; rows 0,1,4,5: N = 128 on slot 5; rows 0,4: N = 64 on slot 6 + 4 KB; rows 2,6: N = 64 on slot 6
          for (int a = 6; a >= 0; --a) {
            const uint32_t A = a0 + a * 16384 + k * 32;
            if (a <= 3) mma2_i8(tb, make_desc(A), make_desc(b0 + a * 8192 + k * 32), make_idesc(256, 256), 1u);
            if (a == 0 || a == 1 || a == 4 || a == 5) mma2_i8(tb + (a == 0 ? 320 : a == 1 ? 256 : a == 4 ? 64 : 0), make_desc(A), make_desc(b0 + 5 * 8192 + k * 32), make_idesc(256, 128), 1u);
            if (a == 0 || a == 4) mma2_i8(tb + (a == 0 ? 256 : 0), make_desc(A), make_desc(b0 + 6 * 8192 + 4096 + k * 32), make_idesc(256, 64), 1u);
            if (a == 2 || a == 6) mma2_i8(tb + (a == 2 ? 256 : 0), make_desc(A), make_desc(b0 + 6 * 8192 + k * 32), make_idesc(256, 64), 1u);
          }
        } else {
          for (int a = 6; a >= 0; --a) {
            const uint32_t A = a0 + a * 16384 + k * 32;
            for (int c0 = a; c0 < 7; c0 += 4) {
              const int nb = (7 - c0) < 4 ? (7 - c0) : 4;
              mma2_i8(tb + 64 * (c0 - a), make_desc(A), make_desc(b0 + c0 * 4096 + k * 32), make_idesc(256, 64 * nb), 1u);
            }
          }
        }
      }
    }
    mma2_commit_mc(&done, 0b11);
    mbar_wait(&done, 0, 3);
    clk[0] = clock64() - t0;
  }
  if (!(tid == 0 && rank == 0)) { if (tid == 0) mbar_wait(&done, 0, 4); }
  __syncthreads();
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

static void run_cost2(int pattern, int N, int iters) {
  const size_t smem = 7 * 16384 + 7 * 8192 + 1024;
  CK(cudaFuncSetAttribute(cost2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* dC; CK(cudaMalloc(&dC, 8)); CK(cudaMemset(dC, 0, 8));
  cost2_kernel<<<2, 128, smem>>>(200, pattern, N, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cost2 pattern %d N %d: launch failed: %s\n", pattern, N, cudaGetErrorString(e)); exit(3); }
  cost2_kernel<<<2, 128, smem>>>(iters, pattern, N, dC);
  CK(cudaDeviceSynchronize());
  long long c = 0; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
  if (pattern == 0) printf("cost2 shape M=256 N=%d: %.1f clk per MMA (math floor %.1f)\n", N, (double)c / (iters * 4.0), 128.0 * N * 32.0 / 8192.0);
  else printf("cost2 pattern %d (%s): %.1f clk per k-step (math floor 896)\n", pattern, pattern == 1 ? "paired 12-MMA k-step, k-step outer" : pattern == 2 ? "unpaired 10-MMA k-step in pair mode" : pattern == 3 ? "paired, row outer (kernel order), 1 issuer" : "paired, row outer, 2 issuers",
              (double)c / (iters * 4.0));
  cudaFree(dC);
}

int main(int argc, char** argv) {
  int fails = 0;
  fails += run_check2<256>(0);
  fails += run_check2<128>(0);
  fails += run_check2<64>(0);
  fails += run_check2<192>(0);
  if (argc > 1 && !strcmp(argv[1], "remote")) {   // may be rejected by the hardware: run separately
    fails += run_check2<256>(1);
    fails += run_check2<64>(1);
    return fails ? 1 : 0;
  }
  for (int N : {256, 192, 128, 64, 32}) run_cost2(0, N, 2000);
  run_cost2(1, 0, 2000);
  run_cost2(2, 0, 2000);
  run_cost2(3, 0, 2000);
  run_cost2(4, 0, 2000);
  return fails ? 1 : 0;
}
