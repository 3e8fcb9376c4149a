// Variance contraction q = ||alpha^2 L^{-1} k*||^2 on the INT8 tcgen05 tensor cores (UTCIMMA, accumulators in TMEM):
// the second engine behind BOSS.var / mean_and_var (call sites /root/reference/src/likelihoods/normal.jl:37,43,60,68).
//
// Error-free splitting (Ozaki scheme, base 256).  Both operands are written as fixed-point integers
//     k(x_i, X_k)             = IA[i,k] * 2^-(8S-1),        IA in [0, 2^(8S-1)]   (kernel values lie in [0, 1])
//     alpha^2 Linv[n,k]       = IB[n,k] * 2^(f_n-(8S-2)),   |IB| <= 2^(8S-2)      (f_n: per-row power-of-two scale)
// and cut into S byte slices (A unsigned, B balanced signed).  A slice product A_a . B_b^T is an exact int32 GEMM; all pairs
// with a + b = g share the weight 2^-8g and accumulate in ONE TMEM accumulator.  Pairs with a + b >= S are below the kept
// precision (2^-8S relative to the row scale) and are skipped, which leaves S(S+1)/2 INT8 GEMMs.  The FP64 value
//     v[i,n] = 2^(f_n-13) * sum_g 2^-8g D_g[i,n]
// is assembled in the epilogue (one int64 per column for S <= 7, an FP64 Horner chain for S = 8), squared and added to the row sum.
// Template parameter SA < S: K* carries only SA byte slices (IA = rn(k 2^(8SA-1)); slice a keeps its weight 2^(-8a-7)), rows
// a = SA .. S-1 of the slice-pair triangle do not exist.  The library's default is S = 7, SA = 6 (27 pairs): the 28th pair would
// multiply the lowest byte of K* with the most significant byte of alpha^2 L^{-1} only (Ctx::var_a_slices in common.cuh).
//
// Shape of one CTA step: query tile 128 rows (MMA M) x 64 rows of L^{-1} (n-tile) x one 128-byte k-block.  All S group
// accumulators of the n-tile are live at once (S x 64 TMEM columns <= 512), so every operand tile is loaded exactly once per
// (n-tile, k-block).  For slice a of the query tile the B operands it meets are the slices b = 0..S-1-a; they are stored
// contiguously in descending b, so up to four of them form ONE tcgen05.mma of N = 256 whose output columns are four
// adjacent group accumulators.  Roles: warp 0 = TMA producer (cp.async.bulk, full/empty mbarriers), warp 1 = TMEM allocation,
// warps 1 and 10 = one MMA-issuing thread each on disjoint slice rows, warps 2..9 = epilogue (tcgen05.ld, int64 assembly of the
// groups, FP64 row sums).
#include <stdlib.h>

#include "common.cuh"

namespace bosip {

constexpr int I8_BM = 128, I8_BN = 64, I8_BK = 128;
constexpr int I8_A_TILE = I8_BM * I8_BK;  // 16 KB: [128 rows][128 B], SWIZZLE_128B image
constexpr int I8_B_TILE = I8_BN * I8_BK;  //  8 KB: [ 64 rows][128 B], SWIZZLE_128B image
constexpr int I8_THREADS = 352;   // warp 0 TMA producer, warps 1 and 10 MMA issuers, warps 2..9 epilogue
constexpr int I8_ISSUERS = 2;      // MMA-issuing threads of the default kernel; NI = 1 is kept as a cross-check (BOSIP_B200_I8_ISSUERS=1)
constexpr int I8_EPI_WARPS = 8;
constexpr int I8_DEFAULT_MODE = 0;  // MMA schedule of the default kernel (see i8_issue): 0 plain, 1 phased, 2 split last k-block

template <int S> struct I8Cfg {
  // which of the two MMA threads issues slice a (a row of the slice-pair triangle); balanced in MMA count per k-step
  static constexpr int owner(int a, int ni) {
    return ni == 1 ? 0 : (S == 6) ? ((a == 0 || a == 3 || a == 4) ? 0 : 1)
         : (S == 7) ? ((a == 0 || a == 3 || a == 5 || a == 6) ? 0 : 1)
                    : ((a == 0 || a == 3 || a == 4 || a == 7) ? 0 : 1);
  }
  // A tiles live in one ring PER issuing thread (ring(ni) slots each): every full/empty barrier then has exactly one producer and
  // one consumer, which is what the one-bit phase parity of an mbarrier can express.  (With a shared ring a thread can poll the
  // NEXT fill of a slot whose current fill -- owned by the other thread -- has not landed yet; the parity test then passes early.
  // That race dead-locked builds with a shallow shared ring.)
  // Sized so that A rings + two B sets fill the 227 KB of shared memory: S = 6: 4 + 4, S = 7: 4 + 3, S = 8: 3 + 3 tiles of 16 KB.
  static constexpr int A_SLOTS = (S == 6) ? 8 : (S == 7) ? 7 : 6;
#ifdef BOSIP_I8_RING_SWAP   // dev A/B: the odd slot goes to issuer 1 instead of issuer 0
  static constexpr int ring(int q, int ni) { return ni == 1 ? A_SLOTS : (q == 0 ? A_SLOTS / 2 : (A_SLOTS + 1) / 2); }
  static constexpr int ring_base(int q, int ni) { return (ni == 1 || q == 0) ? 0 : A_SLOTS / 2; }
#else
  static constexpr int ring(int q, int ni) { return ni == 1 ? A_SLOTS : (q == 0 ? (A_SLOTS + 1) / 2 : A_SLOTS / 2); }
  static constexpr int ring_base(int q, int ni) { return (ni == 1 || q == 0) ? 0 : (A_SLOTS + 1) / 2; }
#endif
  static constexpr int B_SET = S * I8_B_TILE;
  static constexpr size_t SMEM = 1024 + (size_t)A_SLOTS * I8_A_TILE + 2 * (size_t)B_SET + 256;
};

// byte offset of (row r, k-byte c) inside a SWIZZLE_128B K-major tile image: 16-byte chunk (c >> 4) is stored at chunk ^ (r & 7)
__host__ __device__ __forceinline__ uint32_t sw128(int r, int c) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 4) ^ r) & 7) << 4) + (c & 15));
}
// number of (n-tile, k-block) operand sets stored before 64-row tile t: tile u owns (u >> 1) + 1 k-blocks (lower triangle)
__host__ __device__ __forceinline__ int i8_tile_off(int t) { const int v = t >> 1; return (t & 1) ? (v + 1) * (v + 1) : v * (v + 1); }
// n-tiles are dealt to the `nsub` work items of a (output, query tile) in zig-zag order from the last (most expensive) tile
__host__ __device__ __forceinline__ int i8_subset(int t, int T64, int nsub) {
  const int r = (T64 - 1 - t) % (2 * nsub);
  return r < nsub ? r : 2 * nsub - 1 - r;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread; a suspended thread is woken several hundred clk after the phase completes)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_spin_b(uint64_t* bar, uint32_t parity) {   // bounded spin on test_wait: latency-critical hand-overs
  if (mbar_test(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_test(bar, parity))
    if (clock64() - t0 > 8000000000ll) asm volatile("trap;");
}
// bounded wait: a protocol error traps (-> cudaErrorLaunchFailure) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, int tag) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
#ifdef BOSIP_I8_DEBUG_TIMEOUT
    if (clock64() - t0 > 400000000ll) {
      printf("i8 timeout: block %d thread %d tag %d parity %u\n", blockIdx.x, threadIdx.x, tag, parity);
      asm volatile("trap;");
    }
#else
    (void)tag;
    if (clock64() - t0 > 8000000000ll) asm volatile("trap;");
#endif
  }
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__host__ __device__ constexpr uint32_t umma_idesc_u8s8(int M, int N) {  // D = s32, A = u8, B = s8, both K-major
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// same, from the low descriptor words (the high word -- 1024-byte group stride, version 1, SWIZZLE_128B -- is a constant)
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | 0x10000u; }
__device__ __forceinline__ void umma_i8_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
#ifdef BOSIP_I8_DEBUG_NOMMA   // energy probe: the whole operand stream without the multiplications (results are garbage)
  return;
#endif
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %4};\nmov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %3, p;\n}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
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
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {  // 32 lanes x 32 columns <- 0
  const uint32_t z = 0u;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n"
      ::"r"(taddr), "r"(z) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ double i32_to_f64(uint32_t v) {  // exact: 2^52 + 2^31 + v as a bit pattern, minus the offset
  return __hiloint2double(0x43300000, (int)(v ^ 0x80000000u)) - 4503601774854144.0;
}

#ifndef BOSIP_I8_INT_TAIL
#define BOSIP_I8_INT_TAIL 0   // 1: exact integer row sums of squares (S <= 7), 0: FP64 row sums.  MEASURED (round 2, same box, ms per 4 C3 / 2 C5 / 14 C2
                              // launches): FP64 10.85 / 18.58 / 1.93, integer 10.88 / 19.31 / 2.39 -- the ~800 integer instructions per tile sit on the
                              // accumulator hand-over path of short tiles, the starved FP64 tail does not.  Build-time A/B switch, default off.
#endif
// W (192-bit unsigned, six 32-bit words) += y^2 << sh, exactly.  |y| < 2^63 (N < 8192, see i8_prepare), 0 <= sh <= 30.
// Integer pipe only (IMAD.WIDE.U32, funnel shifts, carry chains): FP64 instructions all but stop while UTCIMMA executes -- and slow the
// MMAs of BOTH CTAs of a pair down while they trickle through -- these do not.
__device__ __forceinline__ void sq_acc192(uint32_t (&W)[6], long long y, uint32_t sh) {
  const uint32_t yl = (uint32_t)y, yh = (uint32_t)((unsigned long long)y >> 32);
  const uint32_t sg = (uint32_t)((int)yh >> 31);
  uint32_t al, ah;
  asm("sub.cc.u32 %0, %2, %4;\n\tsubc.u32 %1, %3, %4;\n" : "=r"(al), "=r"(ah) : "r"(yl ^ sg), "r"(yh ^ sg), "r"(sg));   // |y|
  const unsigned long long lo = (unsigned long long)al * al, hi = (unsigned long long)ah * ah, mid = (unsigned long long)(ah << 1) * al;
  uint32_t p0 = (uint32_t)lo, p1 = (uint32_t)(lo >> 32), p2 = (uint32_t)hi, p3 = (uint32_t)(hi >> 32);
  asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, %4;\n\taddc.u32 %2, %2, 0;\n" : "+r"(p1), "+r"(p2), "+r"(p3) : "r"((uint32_t)mid), "r"((uint32_t)(mid >> 32)));
  const uint32_t t0 = p0 << sh, t1 = __funnelshift_l(p0, p1, sh), t2 = __funnelshift_l(p1, p2, sh), t3 = __funnelshift_l(p2, p3, sh),
                 t4 = __funnelshift_l(p3, 0u, sh);
  asm("add.cc.u32 %0, %0, %6;\n\taddc.cc.u32 %1, %1, %7;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %9;\n\taddc.cc.u32 %4, %4, %10;\n\taddc.u32 %5, %5, 0;\n"
      : "+r"(W[0]), "+r"(W[1]), "+r"(W[2]), "+r"(W[3]), "+r"(W[4]), "+r"(W[5]) : "r"(t0), "r"(t1), "r"(t2), "r"(t3), "r"(t4));
}
__device__ __forceinline__ double u192_to_f64(const uint32_t (&W)[6]) {   // relative error <= 2^-52
  const unsigned long long u0 = W[0] | ((unsigned long long)W[1] << 32), u1 = W[2] | ((unsigned long long)W[3] << 32), u2 = W[4] | ((unsigned long long)W[5] << 32);
  return fma(__ull2double_rn(u2), 340282366920938463463374607431768211456.0, fma(__ull2double_rn(u1), 18446744073709551616.0, __ull2double_rn(u0)));
}

// ------------------------------------------------------------------------------------------------ the GEMM
#ifdef BOSIP_I8_PROFILE  // dev build: cycles block 0 spends waiting vs issuing, printed by the host after every launch
__device__ unsigned long long i8_dbg[32];
__device__ long long i8_ts[16][512];
__device__ long long i8_tw[2][8][512];   // per epilogue warp: [0] past the t_full wait, [1] first group released   // per-tile timestamps of block 0 (plain stores: cheap)
#define I8_TS(ev, tile) do { if (blockIdx.x == 0 && (tile) < 512) i8_ts[ev][tile] = clock64(); } while (0)
#define I8_TW(ev, w, tile) do { if (blockIdx.x == 0 && (tile) < 512) i8_tw[ev][w][tile] = clock64(); } while (0)
#else
#define I8_TS(ev, tile) do {} while (0)
#define I8_TW(ev, w, tile) do {} while (0)
#endif
struct I8Args {
  const uint8_t* A;       // [p][mtA][KB][S][16 KB]     K* slices (kstar.cu, i8 mode)
  const uint8_t* B;       // [p][sets][S][8 KB]         alpha^2 L^{-1} slices, position = byte significance (descending b)
  const uint8_t* Bp;      // [p][sets][2][S x 8 KB]     the same slices as per-rank images of the paired kernel (i8_pair_images_kernel)
  const double* cs;       // [p][T64*64]                2^(f_n - 13), 0 for padding rows
  const uint32_t* nib;    // [p][T64*64/8]              s_n = f_n - f_base(output), 4 bits per row (S <= 7: integer row sums)
  const double* qscale;   // [p]                        value of one unit of the integer row sum: 4^(f_base + R - 13 - 8(S-1))
  double* q_part;         // [nsub][p][mc]
  long long mc;
  size_t b_per_out;       // bytes of B per output
  int KB, T64, mtA, m_tiles, p, nsub, n_items, N;
  int m_pairs;            // paired kernel: query-tile pairs per output, n_items = p * m_pairs * nsub
};


// One MMA-issuing thread.  A single thread sustains ~100 clk per tcgen05.mma (descriptor arithmetic, uniform-register moves,
// barrier polls), more than the small slice MMAs take to execute, so the slice rows of a k-block are dealt to two threads.
// Integer accumulation is order-independent and the accumulators are handed over zeroed, so the streams need no ordering.
// PH ("phased"): every slice row a issues at most two stacked MMAs per k-step -- #1 on B positions a .. a+3 (group accumulators
// S-1 .. S-4, the low-order ones) and #2 on positions a+4 .. S-1 (groups S-5 .. 0).  The unphased kernel issues both back to back, so
// ALL S accumulators of an n-tile complete together and the tensor pipe idles while the epilogue drains them (~13 % of a C3 tile:
// TMEM holds 448 of 512 columns, there is no second accumulator set).  The phased kernel walks the tile's k-blocks TWICE: phase L
// issues every #1 (22 of the 28 slice pairs at S = 7) and commits `t_full[0]`, phase H issues every #2 (rows a < S-4 only) and commits
// `t_full[1]`.  The epilogue drains the four low-order groups while phase H runs and the high-order groups while the NEXT tile's
// phase L runs, so the pipe never waits for a drain.  Price: the operand tiles phase H needs (A slices 0 .. S-5, B positions
// 4 .. S-1) are fetched a second time (+43 % L2 -> shared-memory traffic at S = 7).  Integer accumulation is exact, so the result
// is bit-identical to the unphased kernel.  MEASURED (round 2): slower -- see launch_i8; kept as an opt-in experiment.
// MD = 2 ("split last k-block") takes the idea without the re-fetch: only in the LAST k-block of a tile the #2 MMAs (rows a < S-4)
// are deferred until every #1 has been issued and `t_full[0]` committed; their A tiles simply stay in their ring slots a little
// longer.  The epilogue then starts draining the low-order groups while the tensor pipe still works on the deferred #2s (~870 clk
// at S = 7), and the next tile's first rows find their accumulators released when the pipe gets to them.
// MD = 3 additionally splits the FIRST k-block: all #1 MMAs of the new tile need the low-order groups only (which the epilogue
// releases first), the #2 MMAs -- the only ones that touch the high-order groups -- go to the end of the k-block.
// MD: 0 = plain, 1 = phased, 2 = split last k-block, 3 = split first and last k-block.
template <int S, int Q, int NI, int MD, int SA>
__device__ __forceinline__ void i8_issue(const I8Args& g, uint32_t sA_u, uint32_t sB_u, uint32_t tmem, uint64_t* a_full, uint64_t* a_empty,
                                         uint64_t* b_full, uint64_t* b_empty, uint64_t* t_full, uint64_t* t_empty) {
  using Cfg = I8Cfg<S>;
  uint32_t a_cnt = 0, b_cnt = 0, tile_cnt = 0;  // a_cnt: A tiles this thread has consumed
  const uint32_t a_lo0 = umma_desc_lo(sA_u), b_lo0 = umma_desc_lo(sB_u);
  for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
    const int sub = item % g.nsub;
    for (int t = g.T64 - 1; t >= 0; --t) {
      if (i8_subset(t, g.T64, g.nsub) != sub) continue;
      const int nkb = (t >> 1) + 1;
      // k < min(N, 64 (t + 1)) are the only non-zero columns of this n-tile: the last k-block may need fewer than 4 k-steps
      const int k_end = min(g.N, I8_BN * (t + 1));
      constexpr bool PH = (MD == 1);
#pragma unroll
      for (int phase = 0; phase < (PH ? 2 : 1); ++phase) {
        // groups >= waited have been claimed by this thread for this tile (phase H starts below the four low-order groups);
        // waited_h does the same for the high-order groups of a split first k-block (MD = 3)
        int waited = (PH && phase == 1) ? S - 4 : S;
        int waited_h = S - 4;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t bs = b_cnt & 1;
          mbar_wait_b(&b_full[bs], (b_cnt >> 1) & 1, 10 + Q);
          if (kb == 0 && phase == 0) I8_TS(14 + Q, tile_cnt);    // B set of the tile's first k-block has landed
          const uint32_t b_lo = b_lo0 + bs * (uint32_t)(Cfg::B_SET >> 4);
          const int nks = min(4, (k_end - kb * I8_BK + 31) >> 5);
          // split k-block: the #1 MMAs of every row first, the #2 MMAs (rows a < S-4) afterwards.  MD = 2: the last k-block of a
          // tile; MD = 3: the first one as well
          const bool split = (MD >= 2) && (kb == nkb - 1 || (MD == 3 && kb == 0));
          uint32_t hold_as[S], hold_lo[S];                      // ring slot / descriptor of the rows whose #2 MMA is deferred
#pragma unroll
          for (int a = SA - 1; a >= 0; --a) {   // A slices 0 .. SA-1 exist (SA < S: the low-order slices of K* are not formed at all)
            if (Cfg::owner(a, NI) != Q) continue;
            if (PH && phase == 1 && a >= S - 4) continue;     // rows without a second MMA
            // slice a meets groups a .. S-1.  The epilogue hands every group back zeroed, so all MMAs accumulate and the next
            // tile starts on the low-order groups while the previous one is still draining the high-order ones.
            if (kb == 0) {
              const bool low_only = (PH && phase == 0) || (MD == 3);     // this row's #1 MMA needs the low-order groups only
              const int lowg = low_only ? (a > S - 4 ? a : S - 4) : a;
#pragma unroll
              for (int gi = S - 1; gi >= 0; --gi)
                if (gi >= lowg && gi < waited) mbar_wait_b(&t_empty[gi], tile_cnt & 1, 100 + 10 * Q + gi);
              if (waited == S) I8_TS(10 + Q, tile_cnt);   // first group wait of the tile satisfied
              if (lowg < waited) waited = lowg;
            }
            const uint32_t as = (uint32_t)Cfg::ring_base(Q, NI) + a_cnt % Cfg::ring(Q, NI);   // this thread's own ring
            mbar_wait_b(&a_full[as], (a_cnt / Cfg::ring(Q, NI)) & 1, 20 + Q);
            ++a_cnt;
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t a_lo = a_lo0 + as * (uint32_t)(I8_A_TILE >> 4);
            if (kb == 0 && phase == 0 && waited == ((Q == 0) ? S - 1 : S - 3)) I8_TS(4 + Q, tile_cnt);   // this thread's first row of the tile is about to be issued
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              if (ks < nks) {
#pragma unroll
                for (int c0 = a; c0 < S; c0 += 4) {  // slice positions c0 .. c0+nb-1 (descending b) -> group columns 64 (c0 - a) ..
                  if (PH && ((phase == 0) != (c0 == a))) continue;
                  if (MD >= 2 && split && c0 != a) continue;
                  const int nb = (S - c0) < 4 ? (S - c0) : 4;
                  umma_i8_acc(tmem + (uint32_t)(64 * (c0 - a)), a_lo + 2 * ks, b_lo + (uint32_t)(c0 * (I8_B_TILE >> 4) + 2 * ks),
                              umma_idesc_u8s8(I8_BM, 64 * nb));
                }
              }
            }
            if (MD >= 2 && split && a < S - 4) { hold_as[a] = as; hold_lo[a] = a_lo; }
            else umma_commit(&a_empty[as]);
          }
          if (MD >= 2 && split) {
            if (kb == nkb - 1) umma_commit(&t_full[0]);        // the low-order groups are complete once these MMAs retire
#pragma unroll
            for (int a = S - 5; a >= 0; --a) {
              if (Cfg::owner(a, NI) != Q) continue;
              if (MD == 3 && kb == 0) {                        // the deferred #2 MMAs are the first to touch the high-order groups
#pragma unroll
                for (int gi = S - 5; gi >= 0; --gi)
                  if (gi >= a && gi < waited_h) mbar_wait_b(&t_empty[gi], tile_cnt & 1, 200 + 10 * Q + gi);
                if (a < waited_h) waited_h = a;
              }
              const int c0 = a + 4, nb = S - c0;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (ks < nks)
                  umma_i8_acc(tmem + 256u, hold_lo[a] + 2 * ks, b_lo + (uint32_t)(c0 * (I8_B_TILE >> 4) + 2 * ks), umma_idesc_u8s8(I8_BM, 64 * nb));
              umma_commit(&a_empty[hold_as[a]]);
            }
          }
          umma_commit(&b_empty[bs]);
          ++b_cnt;
          if (kb == 0 && phase == 0) I8_TS(6 + Q, tile_cnt);   // first k-block of the tile issued
        }
        umma_commit(&t_full[MD >= 2 ? 1 : phase]);
      }
      I8_TS(0 + Q, tile_cnt);                  // whole tile issued
      ++tile_cnt;
    }
  }
}

template <int S, int NI, int MD, int SA>
__global__ void __launch_bounds__(I8_THREADS, 1) vargemm_i8_kernel(const I8Args g) {
  constexpr bool PH = (MD == 1);
  using Cfg = I8Cfg<S>;
  constexpr int A_SLOTS = Cfg::A_SLOTS;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = base;
  unsigned char* sB = base + (size_t)A_SLOTS * I8_A_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * (size_t)Cfg::B_SET);
  uint64_t* a_full = bars;                       // [A_SLOTS]
  uint64_t* a_empty = bars + A_SLOTS;       // [A_SLOTS]
  uint64_t* b_full = bars + 2 * A_SLOTS;    // [2]
  uint64_t* b_empty = b_full + 2;                // [2]
  uint64_t* t_full = b_empty + 2;                // [2] accumulators complete -> epilogue ([1]: the high-order groups of the phased kernel)
  uint64_t* t_empty = t_full + 2;                // [S] group accumulator drained and zeroed -> MMA
  __shared__ uint32_t tmem_base_s;
  __shared__ double half_sum[I8_BM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef BOSIP_I8_PROFILE
  const long long kernel_t0 = clock64();
#endif

  if (tid == 0) {
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], NI); }
    mbar_init(&t_full[0], NI); mbar_init(&t_full[1], NI);
    for (int i = 0; i < S; ++i) mbar_init(&t_empty[i], I8_EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    // ================================================================ TMA producer
    if (lane == 0) {
      uint32_t a_cnt[2] = {0, 0}, b_cnt = 0;
      for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
        const int sub = item % g.nsub, grp = item / g.nsub, mt = grp % g.m_tiles, j = grp / g.m_tiles;
        const uint8_t* Aj = g.A + ((size_t)j * g.mtA + mt) * g.KB * SA * I8_A_TILE;
        const uint8_t* Bj = g.B + (size_t)j * g.b_per_out;
        for (int t = g.T64 - 1; t >= 0; --t) {
          if (i8_subset(t, g.T64, g.nsub) != sub) continue;
          const int nkb = (t >> 1) + 1;
          const uint8_t* Bt = Bj + (size_t)i8_tile_off(t) * Cfg::B_SET;
#pragma unroll
          for (int phase = 0; phase < (PH ? 2 : 1); ++phase) {
            // phase H of the phased kernel needs B positions 4 .. S-1 and A slices S-5 .. 0 only
            const uint32_t b_off = (PH && phase == 1) ? 4u * I8_B_TILE : 0u;
            const uint32_t b_bytes = (PH && phase == 1) ? (uint32_t)(S - 4) * I8_B_TILE : (uint32_t)Cfg::B_SET;
            const int a_top = (PH && phase == 1) ? S - 5 : SA - 1;
            for (int kb = 0; kb < nkb; ++kb) {
              const uint32_t bs = b_cnt & 1;
              mbar_wait_b(&b_empty[bs], ((b_cnt >> 1) & 1) ^ 1, 1);
#ifdef BOSIP_I8_DEBUG_NOLOAD  // bottleneck isolation: barriers only, operands stay whatever shared memory holds
              mbar_arrive(&b_full[bs]);
#else
              mbar_expect_tx(&b_full[bs], b_bytes);
              bulk_g2s(sB + (size_t)bs * Cfg::B_SET + b_off, Bt + (size_t)kb * Cfg::B_SET + b_off, b_bytes, &b_full[bs]);
#endif
              ++b_cnt;
              const uint8_t* Ak = Aj + (size_t)kb * SA * I8_A_TILE;
#pragma unroll 1
              for (int a = a_top; a >= 0; --a) {  // least significant slice first: it only touches the accumulators the epilogue frees first
                const int q = Cfg::owner(a, NI);
                const uint32_t rq = (uint32_t)Cfg::ring(q, NI);
                const uint32_t as = (uint32_t)Cfg::ring_base(q, NI) + a_cnt[q] % rq;
                mbar_wait_b(&a_empty[as], ((a_cnt[q] / rq) & 1) ^ 1, 2);
#ifdef BOSIP_I8_DEBUG_NOLOAD
                mbar_arrive(&a_full[as]);
#else
                mbar_expect_tx(&a_full[as], I8_A_TILE);
                bulk_g2s(sA + (size_t)as * I8_A_TILE, Ak + (size_t)a * I8_A_TILE, I8_A_TILE, &a_full[as]);
#endif
                ++a_cnt[q];
              }
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ================================================================ MMA issuers (one thread each, disjoint slice rows)
    if (lane == 0) {
      if (warp == 1) i8_issue<S, 0, NI, MD, SA>(g, smem_u32(sA), smem_u32(sB), tmem, a_full, a_empty, b_full, b_empty, t_full, t_empty);
      else if (NI > 1) i8_issue<S, 1, NI, MD, SA>(g, smem_u32(sA), smem_u32(sB), tmem, a_full, a_empty, b_full, b_empty, t_full, t_empty);
    }
  } else {
    // ================================================================ epilogue: 8 warps, thread = (row, 32-column half)
    const int ew = warp - 2;
    const int quad = warp & 3;   // TMEM lanes 32 quad .. 32 quad + 31 are the ones this warp may read
    const int half = ew >> 2;
    const int row = quad * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(32 * half);
    uint32_t tile_cnt = 0;
    // all accumulators start at zero; every later hand-over is (drain, zero, release) per group
#pragma unroll
    for (int gi = 0; gi < S; ++gi) tmem_zero32(taddr + (uint32_t)(64 * gi));
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int gi = 0; gi < S; ++gi) mbar_arrive(&t_empty[gi]);
    }
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
      const int sub = item % g.nsub, grp = item / g.nsub, mt = grp % g.m_tiles, j = grp / g.m_tiles;
      const double* csj = g.cs + (size_t)j * g.T64 * I8_BN;
      double rs = 0.0;
      uint32_t W[6] = {0u, 0u, 0u, 0u, 0u, 0u};   // integer tail: exact sum of squares of this thread's columns over all tiles of the item
      for (int t = g.T64 - 1; t >= 0; --t) {
        if (i8_subset(t, g.T64, g.nsub) != sub) continue;
        mbar_wait_b(&t_full[0], tile_cnt & 1, 30);   // phased kernel: the four low-order groups; otherwise all of them
        if (tid == 64) I8_TS(2, tile_cnt);       // accumulators complete
        if (lane == 0) I8_TW(0, ew, tile_cnt);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        // column scales of this thread's 32 columns: lane l fetches column l now (the load latency hides behind the Horner chain)
        // and the values are handed round by shuffles at the end -- 32 dependent global loads per thread took ~13 000 clk per tile
        double my_cs = 0.0;
        uint4 nb = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (S <= 7 && BOSIP_I8_INT_TAIL) nb = __ldg(reinterpret_cast<const uint4*>(g.nib + ((size_t)j * g.T64 * I8_BN + (size_t)t * I8_BN + 32 * half) / 8));
        else my_cs = __ldg(csj + (size_t)t * I8_BN + 32 * half + lane);
        // sum_g 2^-8g D_g per column, least significant group first (group gi lives at columns 64 (S-1-gi)); each group is zeroed
        // and handed back to the MMA threads as soon as it has been read.  FP64 instructions run ~17 x slower while the tensor
        // pipe is busy (tools/i8mma_probe contend: DFMA 2.2 -> 36 clk per warp instruction; IMAD.WIDE, shifts, adds, I2F unaffected),
        // and this code overlaps the next tile's MMAs.  So for S <= 7 the whole assembly is ONE int64 per column: the four low-order
        // groups are summed exactly (< 2^56) and rounded to units of 2^R -- R = 11 at S = 7 (rms 590 low-order units against the
        // ~630 of truncation the scheme has anyway: measured, errors grow 1.4 x), 8 at S = 6 -- and the remaining groups land on top at bits 8k - R; the single-
        // pair most significant group (|D_0| <= K 2^13, K < 8192) ends below 2^63.  One conversion, one scale and one FMA per column remain.
        uint32_t v[32];
        if constexpr (S <= 7) {
          constexpr int R = (S == 7) ? 11 : 8;
          long long acc[32];
#pragma unroll
          for (int gi = S - 1; gi >= 0; --gi) {
            if (MD != 0 && gi == S - 5) {          // the high-order groups complete with phase H / the deferred MMAs
              mbar_wait_b(&t_full[1], tile_cnt & 1, 31);
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            const uint32_t ta = taddr + (uint32_t)(64 * (S - 1 - gi));
            tmem_ld32(ta, v);
            tmem_zero32(ta);
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[gi]);
            if (tid == 64 && gi == S - 1) I8_TS(3, tile_cnt);
            if (lane == 0 && gi == S - 1) I8_TW(1, ew, tile_cnt);
            if (tid == 64 && gi == 0) I8_TS(8, tile_cnt);
            if (tid == 96 && gi == 0) I8_TS(12, tile_cnt);   // the same on a sub-partition without an MMA-issuing thread
            const int k = S - 1 - gi;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const long long d = (long long)(int)v[c];
              if (k == 0) acc[c] = d;
              else if (k <= 3) acc[c] += d * (long long)(1 << (8 * k));
              else if (8 * k - R < 31) acc[c] += d * (long long)(1 << (8 * k - R));
              else acc[c] += d << (8 * k - R);
              if (k == 3) acc[c] = (acc[c] + (1ll << (R - 1))) >> R;
            }
          }
          if constexpr (BOSIP_I8_INT_TAIL) {
            // Row sum of squares in INTEGER arithmetic: column c holds v = acc[c] 2^(f_c + R - 13 - 8(S-1)) with f_c = f_base + s_c
            // (s_c in 0..15, four bits per column in nb), so v^2 = acc[c]^2 4^(s_c) units of the output's qscale; the squares are added
            // exactly into a 192-bit integer that lives across all tiles of the item -- no FP64-pipe instruction in the tile loop.
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const uint32_t w = (c < 8) ? nb.x : (c < 16) ? nb.y : (c < 24) ? nb.z : nb.w;
              sq_acc192(W, acc[c], ((w >> (4 * (c & 7))) & 15u) * 2u);
            }
          } else {
            const double sc = my_cs * ldexp(1.0, R - 8 * (S - 1));   // acc is in units of 2^R low-order units
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const double x = __ll2double_rn(acc[c]) * __shfl_sync(0xffffffffu, sc, c);
              rs = fma(x, x, rs);
            }
          }
          ++tile_cnt;
        } else {
          double w[32];
#pragma unroll
          for (int gi = S - 1; gi >= 0; --gi) {
            if (MD != 0 && gi == S - 5) {
              mbar_wait_b(&t_full[1], tile_cnt & 1, 31);
              asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            }
            const uint32_t ta = taddr + (uint32_t)(64 * (S - 1 - gi));
            tmem_ld32(ta, v);
            tmem_zero32(ta);
            asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[gi]);
#pragma unroll
            for (int c = 0; c < 32; ++c) w[c] = (gi == S - 1) ? i32_to_f64(v[c]) : fma(w[c], 0.00390625, i32_to_f64(v[c]));
          }
          ++tile_cnt;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const double x = w[c] * __shfl_sync(0xffffffffu, my_cs, c);
            rs = fma(x, x, rs);
          }
        }
        if (tid == 64) I8_TS(9, tile_cnt - 1);   // epilogue of the tile finished
        if (tid == 96) I8_TS(13, tile_cnt - 1);
      }
      if constexpr (S <= 7 && BOSIP_I8_INT_TAIL) rs = u192_to_f64(W) * __ldg(g.qscale + j);
      // combine the two column halves of a row in a fixed order, one partial per work item
      if (half == 1) half_sum[row] = rs;
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      if (half == 0) g.q_part[((size_t)sub * g.p + j) * g.mc + (size_t)mt * I8_BM + row] = rs + half_sum[row];
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
#ifdef BOSIP_I8_PROFILE
  if (blockIdx.x == 0 && tid == 0) i8_dbg[24] += (unsigned long long)(clock64() - kernel_t0);
#endif
}

// ================================================================================================ paired kernel (cta_group::2)
// Two CTAs of a cluster (two SMs of one TPC) run ONE tcgen05.mma stream with M = 256: each CTA owns a 128-row query tile of the
// same (output, n-tile subset) and holds HALF of every stacked B operand, so the shared-memory operand traffic of the tensor core
// per SM drops from 96 KB to 76 KB per k-step.  Why that matters (round-2 measurements, BOSIP_I8_PROFILE timelines + ncu): the
// unpaired kernel keeps the tensor core's shared-memory read pipe 84 % busy while MMAs run (l1tex__data_pipe_tc_wavefronts_mem_shared)
// and the TMA fills (42 KB per k-step) compete for the same memory: k-steps take 1028 clk instead of the 960 the MMA sequence needs
// with resident operands, and the first B set of an n-tile lands up to 3500 clk after it was requested.
//
// Logical B of an MMA = the stacked slices, N rows; rank 0 supplies rows [0, N/2) (D columns [0, N/2)), rank 1 the rest, and ONE
// descriptor offset is applied in both CTAs.  With stacks of 4, 2 and 1 slices only (N = 256, 128, 64) the halves are whole slices
// or 32-row half slices, and the per-rank shared-memory images below make every MMA's offset valid in both CTAs at 56 KB per set:
//     rank 0: slots 0..5 = positions 0..5,                      slot 6 = [pos 6 rows 0-31  | pos 4 rows 0-31 ]
//     rank 1: slots 0..4 = positions 2..6, slot 5 = position 6, slot 6 = [pos 6 rows 32-63 | pos 4 rows 32-63]
// Row a (A slice a) of a k-step issues: a <= 3: N = 256 at slot a (positions a .. a+3); a in {0,1,4,5}: N = 128 at slot 5 (positions
// 5, 6); a in {0,4}: N = 64 at slot 6 + 4 KB (position 4); a in {2,6}: N = 64 at slot 6 (position 6) -- 12 MMAs for the 28 slice pairs
// (tools/i8mma2_probe: 944 clk per k-step with resident operands; every shape checked against a host GEMM).
// Roles per CTA: warp 0 TMA producer (own A tiles, own B image); warps 2..9 epilogue on the CTA's own TMEM; warps 1 and 10: in the
// leader (rank 0) the two MMA-issuing threads, in the peer two relay threads that forward "my tile has landed" to the leader's
// *_peer barriers (a plain cp.async.bulk cannot complete bytes on another CTA's mbarrier -- probed).  tcgen05.commit multicasts
// its arrival to both CTAs' a_empty / b_empty / t_full; the peer's epilogue warps release accumulators on the leader's t_empty.
#ifndef BOSIP_I8P_SPIN
#define BOSIP_I8P_SPIN 3   // bit 0: the peer's relay threads spin on test_wait, bit 1: the leader's issuers spin on the *_peer barriers
#endif
__device__ __forceinline__ uint32_t cl_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_clusterid() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_nclusters() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;\n" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cl_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void cl_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }
// arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope): what is handed over is TMEM whose
// tcgen05.st has been waited for, or a TMA tile whose completion the barrier itself reported -- `.release.cluster` here cost
// ~1000 clk per arrive (accumulator hand-over of a tile 8800 instead of 2450 clk; round-2 timeline)
__device__ __forceinline__ void mbar_arrive_cl(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_cl(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope: arrivals come from the peer CTA
  uint32_t ok;
#ifdef BOSIP_I8P_ACQ_CLUSTER
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#else
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
#if defined(BOSIP_I8P_SPIN) && (BOSIP_I8P_SPIN & 2)
  mbar_spin_b(bar, parity); return;
#endif
  if (mbar_try_cl(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_cl(bar, parity))
    if (clock64() - t0 > 8000000000ll) asm volatile("trap;");
}
__device__ __forceinline__ void umma2_i8_acc(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, 1, 0;\nmov.b64 da, {%1, %4};\nmov.b64 db, {%2, %4};\n"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], da, db, %3, p;\n}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(0x40004040u) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

constexpr size_t I8P_SMEM = I8Cfg<7>::SMEM + 256;   // 35 barriers

// one MMA-issuing thread of the leader CTA (S = 7)
template <int Q>
__device__ __forceinline__ void i8p_issue(const I8Args& g, uint32_t sA_u, uint32_t sB_u, uint32_t tmem, uint64_t* a_full, uint64_t* a_empty,
                                          uint64_t* a_peer, uint64_t* b_full, uint64_t* b_empty, uint64_t* b_peer, uint64_t* t_full,
                                          uint64_t* t_empty) {
  constexpr int S = 7, NI = 2;
  using Cfg = I8Cfg<S>;
  uint32_t a_cnt = 0, b_cnt = 0, tile_cnt = 0;
  const uint32_t a_lo0 = umma_desc_lo(sA_u), b_lo0 = umma_desc_lo(sB_u);
  constexpr uint32_t SLOT = I8_B_TILE >> 4;
  for (int item = (int)cl_clusterid(); item < g.n_items; item += (int)cl_nclusters()) {
    const int sub = item % g.nsub;
    for (int t = g.T64 - 1; t >= 0; --t) {
      if (i8_subset(t, g.T64, g.nsub) != sub) continue;
      const int nkb = (t >> 1) + 1;
      const int k_end = min(g.N, I8_BN * (t + 1));
      int waited = S;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t bs = b_cnt & 1;
        mbar_wait_b(&b_full[bs], (b_cnt >> 1) & 1, 10 + Q);
        mbar_wait_cl(&b_peer[bs], (b_cnt >> 1) & 1);
        if (kb == 0) I8_TS(14 + Q, tile_cnt);
        const uint32_t b_lo = b_lo0 + bs * (uint32_t)(Cfg::B_SET >> 4);
        const int nks = min(4, (k_end - kb * I8_BK + 31) >> 5);
#pragma unroll
        for (int a = S - 1; a >= 0; --a) {
          if (Cfg::owner(a, NI) != Q) continue;
          if (kb == 0) {   // slice a meets groups a .. S-1; both CTAs' epilogues have handed them back zeroed
#pragma unroll
            for (int gi = S - 1; gi >= 0; --gi)
              if (gi >= a && gi < waited) mbar_wait_cl(&t_empty[gi], tile_cnt & 1);
            if (waited == S) I8_TS(10 + Q, tile_cnt);
            if (a < waited) waited = a;
          }
          const uint32_t as = (uint32_t)Cfg::ring_base(Q, NI) + a_cnt % Cfg::ring(Q, NI);
          mbar_wait_b(&a_full[as], (a_cnt / Cfg::ring(Q, NI)) & 1, 20 + Q);
          mbar_wait_cl(&a_peer[as], (a_cnt / Cfg::ring(Q, NI)) & 1);
          ++a_cnt;
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t a_lo = a_lo0 + as * (uint32_t)(I8_A_TILE >> 4);
          if (kb == 0 && waited == ((Q == 0) ? S - 1 : S - 3)) I8_TS(4 + Q, tile_cnt);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            if (ks < nks) {
              if (a <= 3) umma2_i8_acc(tmem, a_lo + 2 * ks, b_lo + (uint32_t)a * SLOT + 2 * ks, umma_idesc_u8s8(256, 256));
              if (a == 0 || a == 1 || a == 4 || a == 5)
                umma2_i8_acc(tmem + (uint32_t)(64 * (5 - a)), a_lo + 2 * ks, b_lo + 5 * SLOT + 2 * ks, umma_idesc_u8s8(256, 128));
              if (a == 0 || a == 4)
                umma2_i8_acc(tmem + (uint32_t)(64 * (4 - a)), a_lo + 2 * ks, b_lo + 6 * SLOT + (4096 >> 4) + 2 * ks, umma_idesc_u8s8(256, 64));
              if (a == 2 || a == 6)
                umma2_i8_acc(tmem + (uint32_t)(64 * (6 - a)), a_lo + 2 * ks, b_lo + 6 * SLOT + 2 * ks, umma_idesc_u8s8(256, 64));
            }
          }
          umma2_commit(&a_empty[as]);
        }
        umma2_commit(&b_empty[bs]);
        ++b_cnt;
        if (kb == 0) I8_TS(6 + Q, tile_cnt);
      }
      umma2_commit(&t_full[0]);
      I8_TS(0 + Q, tile_cnt);
      ++tile_cnt;
    }
  }
}

// the peer CTA's stand-in for an issuing thread: forwards the landing of its CTA's operand tiles to the leader
#if (BOSIP_I8P_SPIN & 1)
#define I8P_RELAY_WAIT(bar, par) mbar_spin_b(bar, par)
#else
#define I8P_RELAY_WAIT(bar, par) mbar_wait_b(bar, par, 40)
#endif
template <int Q>
__device__ __forceinline__ void i8p_relay(const I8Args& g, uint64_t* a_full, uint64_t* a_peer, uint64_t* b_full, uint64_t* b_peer) {
  constexpr int S = 7, NI = 2;
  using Cfg = I8Cfg<S>;
  uint32_t a_cnt = 0, b_cnt = 0;
  for (int item = (int)cl_clusterid(); item < g.n_items; item += (int)cl_nclusters()) {
    const int sub = item % g.nsub;
    for (int t = g.T64 - 1; t >= 0; --t) {
      if (i8_subset(t, g.T64, g.nsub) != sub) continue;
      const int nkb = (t >> 1) + 1;
      for (int kb = 0; kb < nkb; ++kb) {
        if (Q == 0) {
          const uint32_t bs = b_cnt & 1;
          I8P_RELAY_WAIT(&b_full[bs], (b_cnt >> 1) & 1);
          mbar_arrive_cl(cl_mapa(smem_u32(&b_peer[bs]), 0));
          ++b_cnt;
        }
#pragma unroll
        for (int a = S - 1; a >= 0; --a) {
          if (Cfg::owner(a, NI) != Q) continue;
          const uint32_t as = (uint32_t)Cfg::ring_base(Q, NI) + a_cnt % Cfg::ring(Q, NI);
          I8P_RELAY_WAIT(&a_full[as], (a_cnt / Cfg::ring(Q, NI)) & 1);
          mbar_arrive_cl(cl_mapa(smem_u32(&a_peer[as]), 0));
          ++a_cnt;
        }
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I8_THREADS, 1) vargemm_i8p_kernel(const I8Args g) {
  constexpr int S = 7, NI = 2;
  using Cfg = I8Cfg<S>;
  constexpr int A_SLOTS = Cfg::A_SLOTS;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = base;
  unsigned char* sB = base + (size_t)A_SLOTS * I8_A_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * (size_t)Cfg::B_SET);
  uint64_t* a_full = bars;                  // [A_SLOTS] this CTA's A tile has landed (TMA)
  uint64_t* a_empty = a_full + A_SLOTS;     // [A_SLOTS] the MMAs that read the slot (in both CTAs) are complete (commit multicast)
  uint64_t* a_peer = a_empty + A_SLOTS;     // [A_SLOTS] leader only: the peer's A tile has landed (relay)
  uint64_t* b_full = a_peer + A_SLOTS;      // [2]
  uint64_t* b_empty = b_full + 2;           // [2]
  uint64_t* b_peer = b_empty + 2;           // [2] leader only
  uint64_t* t_full = b_peer + 2;            // [1] accumulators complete (commit multicast)
  uint64_t* t_empty = t_full + 1;           // [S] leader only: group accumulator drained and zeroed in BOTH CTAs
  __shared__ uint32_t tmem_base_s;
  __shared__ double half_sum[I8_BM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cl_ctarank();

  if (tid == 0) {
    for (int i = 0; i < A_SLOTS; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&a_peer[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], NI); mbar_init(&b_peer[i], 1); }
    mbar_init(&t_full[0], NI);
    for (int i = 0; i < S; ++i) mbar_init(&t_empty[i], 2 * I8_EPI_WARPS);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cl_sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int n_cl = (int)cl_nclusters(), cl_id = (int)cl_clusterid();

  if (warp == 0) {
    // ================================================================ TMA producer (own query tile, own B image)
    if (lane == 0) {
      uint32_t a_cnt[2] = {0, 0}, b_cnt = 0;
      for (int item = cl_id; item < g.n_items; item += n_cl) {
        const int sub = item % g.nsub, grp = item / g.nsub, mp = grp % g.m_pairs, j = grp / g.m_pairs;
        const int mt = min(2 * mp + (int)rank, g.m_tiles - 1);   // an odd last pair: the peer repeats the last tile (result dropped)
        const uint8_t* Aj = g.A + ((size_t)j * g.mtA + mt) * g.KB * S * I8_A_TILE;
        const uint8_t* Bj = g.Bp + (size_t)j * g.b_per_out * 2;
        for (int t = g.T64 - 1; t >= 0; --t) {
          if (i8_subset(t, g.T64, g.nsub) != sub) continue;
          const int nkb = (t >> 1) + 1;
          const uint8_t* Bt = Bj + ((size_t)i8_tile_off(t) * 2 + rank) * Cfg::B_SET;
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t bs = b_cnt & 1;
            mbar_wait_b(&b_empty[bs], ((b_cnt >> 1) & 1) ^ 1, 1);
            mbar_expect_tx(&b_full[bs], (uint32_t)Cfg::B_SET);
            bulk_g2s(sB + (size_t)bs * Cfg::B_SET, Bt + (size_t)kb * 2 * Cfg::B_SET, (uint32_t)Cfg::B_SET, &b_full[bs]);
            ++b_cnt;
            const uint8_t* Ak = Aj + (size_t)kb * S * I8_A_TILE;
#pragma unroll 1
            for (int a = S - 1; a >= 0; --a) {
              const int q = Cfg::owner(a, NI);
              const uint32_t rq = (uint32_t)Cfg::ring(q, NI);
              const uint32_t as = (uint32_t)Cfg::ring_base(q, NI) + a_cnt[q] % rq;
              mbar_wait_b(&a_empty[as], ((a_cnt[q] / rq) & 1) ^ 1, 2);
              mbar_expect_tx(&a_full[as], I8_A_TILE);
              bulk_g2s(sA + (size_t)as * I8_A_TILE, Ak + (size_t)a * I8_A_TILE, I8_A_TILE, &a_full[as]);
              ++a_cnt[q];
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    if (lane == 0) {
      if (rank == 0) {
        if (warp == 1) i8p_issue<0>(g, smem_u32(sA), smem_u32(sB), tmem, a_full, a_empty, a_peer, b_full, b_empty, b_peer, t_full, t_empty);
        else i8p_issue<1>(g, smem_u32(sA), smem_u32(sB), tmem, a_full, a_empty, a_peer, b_full, b_empty, b_peer, t_full, t_empty);
      } else {
        if (warp == 1) i8p_relay<0>(g, a_full, a_peer, b_full, b_peer);
        else i8p_relay<1>(g, a_full, a_peer, b_full, b_peer);
      }
    }
  } else {
    // ================================================================ epilogue on this CTA's TMEM: 8 warps, thread = (row, 32-column half)
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    const int row = quad * 32 + lane;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(32 * half);
    const uint32_t te0 = cl_mapa(smem_u32(&t_empty[0]), 0);   // the LEADER's t_empty (its own window for rank 0)
    uint32_t tile_cnt = 0;
#pragma unroll
    for (int gi = 0; gi < S; ++gi) tmem_zero32(taddr + (uint32_t)(64 * gi));
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int gi = 0; gi < S; ++gi) mbar_arrive_cl(te0 + 8u * gi);
    }
    for (int item = cl_id; item < g.n_items; item += n_cl) {
      const int sub = item % g.nsub, grp = item / g.nsub, mp = grp % g.m_pairs, j = grp / g.m_pairs;
      const int mt = 2 * mp + (int)rank;
      const double* csj = g.cs + (size_t)j * g.T64 * I8_BN;
      double rs = 0.0;
      uint32_t W[6] = {0u, 0u, 0u, 0u, 0u, 0u};
      for (int t = g.T64 - 1; t >= 0; --t) {
        if (i8_subset(t, g.T64, g.nsub) != sub) continue;
        mbar_wait_b(&t_full[0], tile_cnt & 1, 30);
        if (tid == 64) I8_TS(2, tile_cnt);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        double my_cs = 0.0;
        uint4 nb = make_uint4(0u, 0u, 0u, 0u);
        if constexpr (BOSIP_I8_INT_TAIL) nb = __ldg(reinterpret_cast<const uint4*>(g.nib + ((size_t)j * g.T64 * I8_BN + (size_t)t * I8_BN + 32 * half) / 8));
        else my_cs = __ldg(csj + (size_t)t * I8_BN + 32 * half + lane);
        uint32_t v[32];
        constexpr int R = 11;
        long long acc[32];
#pragma unroll
        for (int gi = S - 1; gi >= 0; --gi) {   // same int64 assembly as the unpaired kernel (bit-identical results)
          const uint32_t ta = taddr + (uint32_t)(64 * (S - 1 - gi));
          tmem_ld32(ta, v);
          tmem_zero32(ta);
          asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cl(te0 + 8u * gi);
          if (tid == 64 && gi == S - 1) I8_TS(3, tile_cnt);
          if (tid == 64 && gi == 0) I8_TS(8, tile_cnt);
          const int k = S - 1 - gi;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const long long d = (long long)(int)v[c];
            if (k == 0) acc[c] = d;
            else if (k <= 3) acc[c] += d * (long long)(1 << (8 * k));
            else if (8 * k - R < 31) acc[c] += d * (long long)(1 << (8 * k - R));
            else acc[c] += d << (8 * k - R);
            if (k == 3) acc[c] = (acc[c] + (1ll << (R - 1))) >> R;
          }
        }
        if constexpr (BOSIP_I8_INT_TAIL) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const uint32_t w = (c < 8) ? nb.x : (c < 16) ? nb.y : (c < 24) ? nb.z : nb.w;
            sq_acc192(W, acc[c], ((w >> (4 * (c & 7))) & 15u) * 2u);
          }
        } else {
          const double sc = my_cs * ldexp(1.0, R - 8 * (S - 1));
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const double x = __ll2double_rn(acc[c]) * __shfl_sync(0xffffffffu, sc, c);
            rs = fma(x, x, rs);
          }
        }
        if (tid == 64) I8_TS(9, tile_cnt);
        ++tile_cnt;
      }
      if constexpr (BOSIP_I8_INT_TAIL) rs = u192_to_f64(W) * __ldg(g.qscale + j);
      if (half == 1) half_sum[row] = rs;
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
      if (half == 0 && mt < g.m_tiles) g.q_part[((size_t)sub * g.p + j) * g.mc + (size_t)mt * I8_BM + row] = rs + half_sum[row];
      asm volatile("bar.sync 1, 256;\n" ::: "memory");
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  cl_sync();   // the peer's shared memory and TMEM are read / written by the leader's MMAs until here
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
}

// per-rank B images of the paired kernel from the position-major sets (see the layout table above); thread = 16 bytes
__global__ void i8_pair_images_kernel(const uint8_t* __restrict__ B, size_t n_sets_total, uint8_t* __restrict__ Bp) {
  constexpr int S = 7;
  constexpr size_t SET = (size_t)S * I8_B_TILE;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // 16-byte chunk of the output
  const size_t total = n_sets_total * 2 * (SET / 16);
  if (e >= total) return;
  const size_t off = (e % (SET / 16)) * 16, sr = e / (SET / 16);
  const int rank = (int)(sr & 1);
  const size_t set = sr >> 1;
  const int slot = (int)(off / I8_B_TILE);
  const size_t in_slot = off % I8_B_TILE;
  int pos; size_t src_off;
  if (slot < 6 && rank == 0) { pos = slot; src_off = in_slot; }
  else if (slot < 5) { pos = slot + 2; src_off = in_slot; }
  else if (slot == 5) { pos = 6; src_off = in_slot; }
  else {   // slot 6: [pos 6 half | pos 4 half], rows 0-31 for rank 0, rows 32-63 for rank 1
    pos = (in_slot < 4096) ? 6 : 4;
    src_off = (in_slot & 4095) + (rank ? 4096 : 0);
  }
  *reinterpret_cast<uint4*>(Bp + e * 16) = *reinterpret_cast<const uint4*>(B + set * SET + (size_t)pos * I8_B_TILE + src_off);
}

__global__ void i8_qsum_kernel(const double* __restrict__ part, int nsub, size_t n, double* __restrict__ q) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = part[i];
  for (int u = 1; u < nsub; ++u) s += part[(size_t)u * n + i];
  q[i] = s;
}

// ------------------------------------------------------------------------------------------------ slicing L^{-1} (once per fit)
// f_n = smallest exponent with max_k |alpha^2 Linv[n,k]| <= 2^f_n
__global__ void i8_rowscale_kernel(const double* __restrict__ Linv, const double* __restrict__ alpha2, int N, int NP, int* __restrict__ fexp,
                                   double* __restrict__ cs) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (n >= NP) return;
  if (n >= N) { fexp[(size_t)j * NP + n] = 0; cs[(size_t)j * NP + n] = 0.0; return; }
  const double* Lj = Linv + (size_t)j * N * N;
  double mx = 0.0;
  for (int k = 0; k <= n; ++k) mx = fmax(mx, fabs(Lj[(size_t)n + (size_t)k * N]));
  mx *= alpha2[j];
  int e = 0;
  if (mx > 0.0) { frexp(mx, &e); if (ldexp(1.0, e - 1) == mx) e -= 1; }  // mx = m 2^e, m in [0.5, 1); exact powers of two fit one lower
  fexp[(size_t)j * NP + n] = e;
  cs[(size_t)j * NP + n] = ldexp(1.0, e - 13);
}

// One block per output: f_base = max(min_n f_n, max_n f_n - 15); rows below it are raised to f_base (their absolute resolution
// then equals that of a row 2^15 times smaller than the output's largest one -- far below every other row's error), so that
// s_n = f_n - f_base fits four bits.  Writes the nibble table the integer row sums use and the value of one row-sum unit.
__global__ void i8_rowclamp_kernel(int N, int NP, int S, int* __restrict__ fexp, double* __restrict__ cs, uint32_t* __restrict__ nib,
                                   double* __restrict__ qscale) {
  const int j = blockIdx.x, tid = threadIdx.x;
  __shared__ int smax[256], smin[256];
  int mx = -100000, mn = 100000;
  for (int n = tid; n < N; n += 256) { const int f = fexp[(size_t)j * NP + n]; mx = max(mx, f); mn = min(mn, f); }
  smax[tid] = mx; smin[tid] = mn;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (tid < o) { smax[tid] = max(smax[tid], smax[tid + o]); smin[tid] = min(smin[tid], smin[tid + o]); } __syncthreads(); }
  const int base = max(smin[0], smax[0] - 15);
  for (int n = tid; n < N; n += 256)
    if (fexp[(size_t)j * NP + n] < base) { fexp[(size_t)j * NP + n] = base; cs[(size_t)j * NP + n] = ldexp(1.0, base - 13); }
  __syncthreads();
  for (int w = tid; w < NP / 8; w += 256) {
    uint32_t v = 0u;
    for (int i = 0; i < 8; ++i) { const int n = 8 * w + i; if (n < N) v |= (uint32_t)(fexp[(size_t)j * NP + n] - base) << (4 * i); }
    nib[(size_t)j * (NP / 8) + w] = v;
  }
  if (tid == 0) { const int R = (S == 7) ? 11 : 8; qscale[j] = ldexp(1.0, 2 * (base + R - 13 - 8 * (S - 1))); }
}

// thread = (set, row, 16-byte chunk): writes the S slice bytes of 16 consecutive k
template <int S>
__global__ void i8_slice_linv_kernel(const double* __restrict__ Linv, const double* __restrict__ alpha2, const int* __restrict__ fexp,
                                     int N, int NP, int T64, size_t b_per_out, uint8_t* __restrict__ B) {
  const int j = blockIdx.y;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = (int)(e & 7), r = (int)((e >> 3) & 63);
  const size_t set = e >> 9;
  const int sets = i8_tile_off(T64);
  if (set >= (size_t)sets) return;
  // invert the triangular set index: tile t owns sets [off(t), off(t) + (t >> 1) + 1)
  int t = 0;
  while (t + 1 < T64 && (size_t)i8_tile_off(t + 1) <= set) ++t;
  const int kb = (int)(set - i8_tile_off(t));
  const int n = t * I8_BN + r;
  const double* Lj = Linv + (size_t)j * N * N;
  const int f = (n < N) ? fexp[(size_t)j * NP + n] : 0;
  unsigned long long bias = 0;
#pragma unroll
  for (int b = 0; b < S - 1; ++b) bias |= 0x80ull << (8 * b);
  uint32_t out[S][4];
#pragma unroll
  for (int s = 0; s < S; ++s) { out[s][0] = out[s][1] = out[s][2] = out[s][3] = 0u; }
  for (int i = 0; i < 16; ++i) {
    const int k = kb * I8_BK + chunk * 16 + i;
    long long I = 0;
    if (n < N && k <= n) {
      const double v = alpha2[j] * Lj[(size_t)n + (size_t)k * N];
      I = __double2ll_rn(ldexp(v, 8 * S - 2 - f));
    }
    const unsigned long long u = (unsigned long long)I + bias;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      uint32_t byte = (uint32_t)(u >> (8 * s)) & 0xFFu;
      if (s < S - 1) byte ^= 0x80u;  // balanced digit in [-128, 127]; the top byte is signed as it stands
      out[s][i >> 2] |= byte << (8 * (i & 3));
    }
  }
  uint8_t* dst = B + (size_t)j * b_per_out + set * (size_t)(S * I8_B_TILE) + sw128(r, chunk * 16);
#pragma unroll
  for (int s = 0; s < S; ++s)
    *reinterpret_cast<uint4*>(dst + (size_t)s * I8_B_TILE) = make_uint4(out[s][0], out[s][1], out[s][2], out[s][3]);
}

// BOSIP_B200_I8_PAIR=1 selects the paired (cta_group::2) kernel for S = 7.  Opt-in: it is bit-identical to the unpaired kernel and its
// long n-tiles run at 890 clk per k-step instead of 1030, but over a whole C3 / C5 / C2 sweep it is 2-4 % / 6 % / 35 % SLOWER
// (round 2, same box: 11.29 vs 10.85, 19.71 vs 18.58, 2.62 vs 1.93 ms): every operand hand-over now waits for the later of two
// CTAs plus a relay hop, and short n-tiles (one 168 KB operand fetch for two k-steps of MMAs) dominate the loss.
static bool i8_pair_enabled() {
  static const bool on = getenv("BOSIP_B200_I8_PAIR") && atoi(getenv("BOSIP_B200_I8_PAIR")) == 1;
  return on;
}

int i8_prepare(Ctx* ctx, Gp* gp, int S, cudaStream_t st) {
  if (gp->i8_S == S && gp->LinvI8) return 0;
  if (S < 6 || S > 8) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "INT8 variance engine: slices must be 6, 7 or 8");
  // int32 accumulator bound: up to S pairs per group, each |sum| <= K * 255 * 128 (this also keeps the epilogue's int64 assembly,
  // which places the single-pair most significant group at bit 32, below 2^61)
  if ((double)gp->N * S * 255.0 * 128.0 >= 2147483647.0 || gp->N >= 8192)
    BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "INT8 variance engine: N too large for exact integer accumulation (N < 8192)");
  if (gp->LinvI8) { cudaFree(gp->LinvI8); gp->LinvI8 = nullptr; }
  if (gp->cscale) { cudaFree(gp->cscale); gp->cscale = nullptr; }
  gp->i8_S = 0;   // not prepared until everything below has succeeded
  const int N = gp->N, p = gp->p;
  const int T64 = (N + I8_BN - 1) / I8_BN, NP = T64 * I8_BN;
  const size_t sets = (size_t)i8_tile_off(T64);
  const size_t b_per_out = sets * S * I8_B_TILE;
  int* fexp = nullptr; double* dAlpha2 = nullptr;
  struct Tmps { int** f; double** a; ~Tmps() { if (*f) cudaFree(*f); if (*a) cudaFree(*a); } } tmps{&fexp, &dAlpha2};  // freed on every exit path
  BOSIP_CUDA(ctx, cudaMalloc(&gp->LinvI8, b_per_out * p));
  BOSIP_CUDA(ctx, cudaMalloc(&gp->cscale, (size_t)p * NP * 8 + (size_t)(p + (p & 1)) * 8 + (size_t)p * (NP / 8) * 4));   // cs | qscale | nibbles
  gp->i8_qscale = gp->cscale + (size_t)p * NP;
  gp->i8_nib = reinterpret_cast<uint32_t*>(gp->i8_qscale + p + (p & 1));   // 16-byte aligned: the epilogue loads 32 nibbles as one uint4
  BOSIP_CUDA(ctx, cudaMalloc(&fexp, (size_t)p * NP * 4));
  BOSIP_CUDA(ctx, cudaMalloc(&dAlpha2, (size_t)p * 8));
  BOSIP_CUDA(ctx, cudaMemcpyAsync(dAlpha2, gp->alpha2, (size_t)p * 8, cudaMemcpyHostToDevice, st));
  i8_rowscale_kernel<<<dim3((NP + 127) / 128, p), 128, 0, st>>>(gp->Linv, dAlpha2, N, NP, fexp, gp->cscale);
  i8_rowclamp_kernel<<<p, 256, 0, st>>>(N, NP, S, fexp, gp->cscale, gp->i8_nib, gp->i8_qscale);
  const size_t threads = sets * 64 * 8;
  const dim3 grid((unsigned)((threads + 255) / 256), (unsigned)p);
  switch (S) {
    case 6: i8_slice_linv_kernel<6><<<grid, 256, 0, st>>>(gp->Linv, dAlpha2, fexp, N, NP, T64, b_per_out, gp->LinvI8); break;
    case 7: i8_slice_linv_kernel<7><<<grid, 256, 0, st>>>(gp->Linv, dAlpha2, fexp, N, NP, T64, b_per_out, gp->LinvI8); break;
    default: i8_slice_linv_kernel<8><<<grid, 256, 0, st>>>(gp->Linv, dAlpha2, fexp, N, NP, T64, b_per_out, gp->LinvI8); break;
  }
  BOSIP_CUDA(ctx, cudaGetLastError());
  if (gp->LinvI8P) { cudaFree(gp->LinvI8P); gp->LinvI8P = nullptr; }
  if (S == 7 && i8_pair_enabled()) {   // per-rank operand images of the paired kernel (twice the bytes: C3 8 MB, C5 120 MB per output set)
    BOSIP_CUDA(ctx, cudaMalloc(&gp->LinvI8P, 2 * b_per_out * p));
    const size_t chunks = sets * p * 2 * ((size_t)S * I8_B_TILE / 16);
    i8_pair_images_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(gp->LinvI8, sets * p, gp->LinvI8P);
    BOSIP_CUDA(ctx, cudaGetLastError());
  }
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  gp->i8_S = S; gp->i8_T64 = T64; gp->i8_b_per_out = b_per_out;
  return 0;
}

int i8_nsub(const Gp* gp) {
  const int T64 = (gp->N + I8_BN - 1) / I8_BN;
  static const int forced = getenv("BOSIP_I8_NSUB") ? atoi(getenv("BOSIP_I8_NSUB")) : 0;  // dev A/B switch
  // measured (27-pair engine, full clock / power-capped bench): C3 (T64 = 16) nsub 1 / 2 / 3 / 4 / 8: 11.17 / 10.36 / 10.40 / 10.62 / 10.89 ms per
  // 4 launches, bench 1.971e7 vs 1.937e7 pts/s for 2 vs 4; C5 (T64 = 64) nsub 2 / 4 / 8: 19.56 / 18.39 / 17.96 ms per 2 launches
  int nsub = forced > 0 ? forced : T64 / 8;
  nsub = nsub < 1 ? 1 : (nsub > 8 ? 8 : nsub);
  return nsub > T64 ? T64 : nsub;
}

#ifdef BOSIP_I8_PROFILE
static void i8_profile_dump(cudaStream_t st) {
  {
    static long long ts[16][512];
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(ts, i8_ts, sizeof(ts));
    const int nt = 200;   // tiles 1 .. nt of block 0
    double d[12] = {0};
    for (int t = 1; t <= nt; ++t) {
      const long long issued = std::max(ts[0][t - 1], ts[1][t - 1]);       // previous tile fully issued
      d[0] += (double)(ts[2][t - 1] - issued);                             // -> its accumulators complete (pipe drain)
      d[1] += (double)(ts[3][t - 1] - ts[2][t - 1]);                       // -> first group released
      d[2] += (double)(ts[8][t - 1] - ts[2][t - 1]);                       // -> last group released
      d[3] += (double)(std::min(ts[4][t], ts[5][t]) - ts[2][t - 1]);       // -> first MMA of the next tile issued
      d[4] += (double)(std::max(ts[6][t], ts[7][t]) - ts[2][t - 1]);       // -> first k-block of the next tile issued
      d[5] += (double)(std::max(ts[0][t], ts[1][t]) - std::max(ts[0][t - 1], ts[1][t - 1]));   // tile period
      d[6] += (double)(ts[9][t - 1] - ts[2][t - 1]);                       // epilogue duration
      d[7] += (double)(ts[10][t] - ts[3][t - 1]);                          // first release -> issuer 0 past its group wait
      d[8] += (double)(ts[4][t] - ts[10][t]);                              // -> issuer 0 past its A-tile wait
      d[9] += (double)(ts[9][t - 1] - ts[8][t - 1]);                       // last release -> epilogue of the tile finished
    }
    fprintf(stderr, "[i8 timeline, block 0, mean clk over %d tiles] issue-done->acc-complete %.0f | acc-complete->first release %.0f, ->last release %.0f, "
                    "->next tile first MMA %.0f, ->next tile first k-block issued %.0f | epilogue %.0f | tile period %.0f | first release->issuer0 woke %.0f, then A-tile wait %.0f | last release->epilogue end %.0f\n",
            nt, d[0] / nt, d[1] / nt, d[2] / nt, d[3] / nt, d[4] / nt, d[6] / nt, d[5] / nt, d[7] / nt, d[8] / nt, d[9] / nt);
    {
      long long lo = 1ll << 60, hi = 0; int slow = 0;
      for (int t = 0; t < nt; ++t) { const long long e = ts[9][t] - ts[8][t]; lo = std::min(lo, e); hi = std::max(hi, e); slow += e > 2000; }
      fprintf(stderr, "[i8 timeline] last release->epilogue end: min %lld max %lld, %d of %d tiles above 2000 clk; first tiles:", lo, hi, slow, nt);
      for (int t = 0; t < 12; ++t) fprintf(stderr, " %lld", ts[9][t] - ts[8][t]);
      fprintf(stderr, "\n");
      static long long tw[2][8][512];
      cudaMemcpyFromSymbol(tw, i8_tw, sizeof(tw));
      double sk0 = 0, sk1 = 0; double late[8] = {0};
      for (int t = 1; t <= nt; ++t) {
        long long lo0 = 1ll << 62, hi0 = 0, lo1 = 1ll << 62, hi1 = 0;
        for (int w8 = 0; w8 < 8; ++w8) { lo0 = std::min(lo0, tw[0][w8][t]); hi0 = std::max(hi0, tw[0][w8][t]); lo1 = std::min(lo1, tw[1][w8][t]); hi1 = std::max(hi1, tw[1][w8][t]); }
        sk0 += (double)(hi0 - lo0); sk1 += (double)(hi1 - lo1);
        for (int w8 = 0; w8 < 8; ++w8) late[w8] += (double)(tw[1][w8][t] - lo1);
      }
      fprintf(stderr, "[i8 timeline] epilogue warps: spread of passing t_full %.0f clk, of the first release %.0f clk; mean lateness of the first release per warp:", sk0 / nt, sk1 / nt);
      for (int w8 = 0; w8 < 8; ++w8) fprintf(stderr, " %.0f", late[w8] / nt);
      fprintf(stderr, "\n");
      if (getenv("BOSIP_I8_PROFILE_RAW")) {   // raw per-tile rows of block 0, relative to the completion of tile 7
        const long long z = ts[2][7];
        fprintf(stderr, "[i8 raw] tile | issued0 issued1 | acc-complete first-rel last-rel epi-end | iss0 past group wait, iss1 | iss0 first MMA, iss1 | first k-block issued 0, 1 | first B set landed 0, 1\n");
        for (int t = 8; t < 28; ++t)
          fprintf(stderr, "[i8 raw] %3d | %7lld %7lld | %7lld %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld | %7lld %7lld | %7lld %7lld\n", t, ts[0][t] - z, ts[1][t] - z,
                  ts[2][t] - z, ts[3][t] - z, ts[8][t] - z, ts[9][t] - z, ts[10][t] - z, ts[11][t] - z, ts[4][t] - z, ts[5][t] - z, ts[6][t] - z, ts[7][t] - z,
                  ts[14][t] - z, ts[15][t] - z);
      }
      double m3 = 0; long long hi3 = 0;
      for (int t = 0; t < nt; ++t) { const long long e = ts[13][t] - ts[12][t]; m3 += (double)e; hi3 = std::max(hi3, e); }
      fprintf(stderr, "[i8 timeline] the same tail on warp 3 (no MMA-issuing thread on its sub-partition): mean %.0f max %lld\n", m3 / nt, hi3);
    }
  }
}
#endif

template <int S, int NI, int MD, int SA = S>
static int launch_i8n(Ctx* ctx, const I8Args& a, cudaStream_t st) {
  const int grid = a.n_items < ctx->sms ? a.n_items : ctx->sms;
  BOSIP_CUDA(ctx, cudaFuncSetAttribute(vargemm_i8_kernel<S, NI, MD, SA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Cfg<S>::SMEM));
  vargemm_i8_kernel<S, NI, MD, SA><<<grid, I8_THREADS, I8Cfg<S>::SMEM, st>>>(a);
  BOSIP_CUDA(ctx, cudaGetLastError());
#ifdef BOSIP_I8_PROFILE
  i8_profile_dump(st);
#endif
  return 0;
}

template <int S, int NI, int SA = S>
static int launch_i8m(Ctx* ctx, const I8Args& a, int mode, cudaStream_t st) {
  if (SA == S) {   // the alternative schedules are development switches of the symmetric slicings only
    if (mode == 1) return launch_i8n<S, NI, 1, S>(ctx, a, st);
    if (mode == 2) return launch_i8n<S, NI, 2, S>(ctx, a, st);
    if (mode == 3) return launch_i8n<S, NI, 3, S>(ctx, a, st);
  }
  return launch_i8n<S, NI, 0, SA>(ctx, a, st);
}

// bitwise comparison of the rows the GEMM wrote: part[(u * p + j) * mc + r], r < rows
__global__ void i8_cmp_kernel(const unsigned long long* __restrict__ x, const unsigned long long* __restrict__ y, long long mc, long long rows,
                              int nvec, unsigned int* __restrict__ mismatches) {
  const long long total = (long long)nvec * rows;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long v = e / rows, r = e % rows;
    if (x[v * mc + r] != y[v * mc + r]) atomicAdd(mismatches, 1u);
  }
}

// Two issuing threads accumulate into the same TMEM columns.  Exact integer arithmetic makes the result independent of their
// interleaving PROVIDED the tensor pipe serialises accumulates from different threads, which PTX does not promise (tcgen05.mma is
// ordered per issuing thread only).  So the two-issuer kernel is used only after it has reproduced the single-issuer kernel bit
// for bit on this context's first real launch (same operands, every row of the chunk); on a mismatch the context stays on one
// issuer (~10 % slower).  BOSIP_B200_I8_ISSUERS=1 / 2 forces the choice (2: no self-check; development only).
// The paired kernel: clusters of two CTAs, as many as can be co-resident (one CTA per SM)
static int launch_i8_pair(Ctx* ctx, const I8Args& a0, cudaStream_t st) {
  I8Args a = a0;
  a.m_pairs = (a.m_tiles + 1) / 2;
  a.n_items = a.m_pairs * a.p * a.nsub;
  static int max_clusters = -1;
  BOSIP_CUDA(ctx, cudaFuncSetAttribute(vargemm_i8p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8P_SMEM));
  if (max_clusters < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ctx->sms & ~1)); cfg.blockDim = dim3(I8_THREADS); cfg.dynamicSmemBytes = I8P_SMEM;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, vargemm_i8p_kernel, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = ctx->sms / 2; }
    max_clusters = n < ctx->sms / 2 ? n : ctx->sms / 2;
    if (getenv("BOSIP_I8_DEBUG")) fprintf(stderr, "[bosip_b200] paired INT8 kernel: cudaOccupancyMaxActiveClusters = %d, using %d clusters of 2 on %d SMs\n", n, max_clusters, ctx->sms);
  }
  const int clusters = a.n_items < max_clusters ? a.n_items : max_clusters;
  vargemm_i8p_kernel<<<2 * clusters, I8_THREADS, I8P_SMEM, st>>>(a);
  BOSIP_CUDA(ctx, cudaGetLastError());
#ifdef BOSIP_I8_PROFILE
  fprintf(stderr, "[i8 timeline] paired kernel:\n");
  i8_profile_dump(st);
#endif
  return 0;
}

template <int S, int SA = S>
static int launch_i8(Ctx* ctx, const I8Args& a, cudaStream_t st) {
  // S = 7: the paired kernel (cta_group::2), after it has reproduced the single-issuer unpaired kernel bit for bit on this context's
  // first real launch (same self-check as for the two issuing threads below; on a mismatch the context stays unpaired)
  if (S == 7 && SA == S && a.Bp != nullptr && ctx->i8_pair >= 0) {
    if (ctx->i8_pair == 0) {
      const size_t n = (size_t)a.nsub * a.p * (size_t)a.mc;
      double* ref = nullptr;
      struct Tmp { double** q; ~Tmp() { if (*q) cudaFree(*q); } } tmp{&ref};
      BOSIP_CUDA(ctx, cudaMalloc(&ref, n * 8 + 8));
      unsigned int* d_mis = reinterpret_cast<unsigned int*>(ref + n);
      BOSIP_CUDA(ctx, cudaMemsetAsync(d_mis, 0, 8, st));
      I8Args b = a; b.q_part = ref;
      BOSIP_TRY((launch_i8n<S, 1, 0>(ctx, b, st)));
      BOSIP_TRY(launch_i8_pair(ctx, a, st));
      i8_cmp_kernel<<<ctx->sms * 4, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(ref), reinterpret_cast<const unsigned long long*>(a.q_part),
                                                   a.mc, (long long)a.m_tiles * I8_BM, a.nsub * a.p, d_mis);
      BOSIP_CUDA(ctx, cudaGetLastError());
      unsigned int mis = 0;
      BOSIP_CUDA(ctx, cudaMemcpyAsync(&mis, d_mis, 4, cudaMemcpyDeviceToHost, st));
      BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
      ctx->launches += 2;
      if (mis != 0) {
        BOSIP_CUDA(ctx, cudaMemcpyAsync(a.q_part, ref, n * 8, cudaMemcpyDeviceToDevice, st));
        BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->i8_pair = -1;
        fprintf(stderr, "[bosip_b200] INT8 engine self-check: the paired kernel differs from the unpaired one in %u row sums; staying unpaired\n", mis);
      } else {
        ctx->i8_pair = 1;
      }
      return 0;
    }
    return launch_i8_pair(ctx, a, st);
  }
  // BOSIP_B200_I8_MODE selects the MMA schedule (development A/B switch; all three give bit-identical results): 0 plain, 1 phased,
  // 2 split last k-block.  The phased kernel is 23 % SLOWER on B200 at C3 (profiles/r02_mma_schedule_ab.json: 3.94 vs 3.20 ms per launch --
  // phase H streams 72 KB of operands per k-block for only ~870 clk of MMA work, 85 B/clk per SM against 44 B/clk, so the TMA feed,
  // not the drain it removes, becomes the limiter).
  static const int mode = getenv("BOSIP_B200_I8_MODE") ? atoi(getenv("BOSIP_B200_I8_MODE")) : I8_DEFAULT_MODE;
  if (ctx->i8_issuers == 0) {
    const char* e = getenv("BOSIP_B200_I8_ISSUERS");
    const int forced = e ? atoi(e) : 0;
    if (forced == 1 || forced == 2) {
      ctx->i8_issuers = forced;
    } else {
      const size_t n = (size_t)a.nsub * a.p * (size_t)a.mc;
      double* ref = nullptr;
      struct Tmp { double** q; ~Tmp() { if (*q) cudaFree(*q); } } tmp{&ref};
      BOSIP_CUDA(ctx, cudaMalloc(&ref, n * 8 + 8));
      unsigned int* d_mis = reinterpret_cast<unsigned int*>(ref + n);
      BOSIP_CUDA(ctx, cudaMemsetAsync(d_mis, 0, 8, st));
      I8Args b = a; b.q_part = ref;
      BOSIP_TRY((launch_i8n<S, 1, 0, SA>(ctx, b, st)));            // the reference: one issuer, plain schedule
      BOSIP_TRY((launch_i8m<S, I8_ISSUERS, SA>(ctx, a, mode, st)));
      i8_cmp_kernel<<<ctx->sms * 4, 256, 0, st>>>(reinterpret_cast<const unsigned long long*>(ref), reinterpret_cast<const unsigned long long*>(a.q_part),
                                                   a.mc, (long long)a.m_tiles * I8_BM, a.nsub * a.p, d_mis);
      BOSIP_CUDA(ctx, cudaGetLastError());
      unsigned int mis = 0;
      BOSIP_CUDA(ctx, cudaMemcpyAsync(&mis, d_mis, 4, cudaMemcpyDeviceToHost, st));
      BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
      ctx->launches += 2;
      if (mis != 0) {   // keep the single-issuer results of this launch and stay on one issuer
        BOSIP_CUDA(ctx, cudaMemcpyAsync(a.q_part, ref, n * 8, cudaMemcpyDeviceToDevice, st));
        BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
        ctx->i8_issuers = 1;
        fprintf(stderr, "[bosip_b200] INT8 engine self-check: two MMA issuers differ from one in %u row sums; using one issuer\n", mis);
      } else {
        ctx->i8_issuers = I8_ISSUERS;
      }
      return 0;
    }
  }
  return ctx->i8_issuers == 1 ? launch_i8m<S, 1, SA>(ctx, a, mode, st) : launch_i8m<S, I8_ISSUERS, SA>(ctx, a, mode, st);
}

// Ks8: [p][mc/128][KB][S][16 KB]; q_part: [nsub][p][mc] scratch; q: [p][mc]
// a_slices: byte slices of K* the cross-kernel wrote (= gp->i8_S, or 6 with 7 slices of L^-1: the default, see common.cuh)
int launch_vargemm_i8(Ctx* ctx, const Gp* gp, const uint8_t* Ks8, int a_slices, int64_t mc, int64_t m_tiles, double* q_part, double* q, cudaStream_t st) {
  I8Args a;
  a.A = Ks8; a.B = gp->LinvI8; a.Bp = i8_pair_enabled() ? gp->LinvI8P : nullptr; a.m_pairs = 0; a.cs = gp->cscale; a.nib = gp->i8_nib; a.qscale = gp->i8_qscale; a.q_part = q_part; a.mc = mc; a.b_per_out = gp->i8_b_per_out;
  a.KB = (gp->N + I8_BK - 1) / I8_BK; a.T64 = gp->i8_T64; a.mtA = (int)(mc / I8_BM); a.m_tiles = (int)m_tiles; a.p = gp->p;
  a.N = gp->N; a.nsub = i8_nsub(gp); a.n_items = (int)(m_tiles * gp->p * a.nsub);
  if (gp->i8_S == 7 && a_slices == 6) {
    BOSIP_TRY((launch_i8<7, 6>(ctx, a, st)));
  } else {
    if (a_slices != gp->i8_S) BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "INT8 variance engine: unsupported slice combination");
    switch (gp->i8_S) {
      case 6: BOSIP_TRY(launch_i8<6>(ctx, a, st)); break;
      case 7: BOSIP_TRY(launch_i8<7>(ctx, a, st)); break;
      case 8: BOSIP_TRY(launch_i8<8>(ctx, a, st)); break;
      default: BOSIP_FAIL(ctx, BOSIP_B200_EINVAL, "INT8 variance engine not prepared");
    }
  }
  const size_t n = (size_t)gp->p * mc;
  i8_qsum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(q_part, a.nsub, n, q);
  BOSIP_CUDA(ctx, cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------ INT8 roofline probe
// One CTA per SM issues `iters` x 4 tcgen05.mma (M=128, N=256, K=32) on operands that stay resident in shared memory.
// pattern 0: operand bytes 0..3 (few set bits); pattern 1: hashed pseudo-random bytes (what byte slices of real numbers look like)
__global__ void __launch_bounds__(128) i8_peak_kernel(int iters, int pattern) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* sA = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sB = sA + 4 * I8_A_TILE;
  __shared__ __align__(8) uint64_t done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (4 * I8_A_TILE + 4 * 256 * 128) / 4; i += 128) {
    uint32_t v = 0x01010101u * (i & 3);
    if (pattern) { v = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; v ^= v >> 15; v *= 2246822519u; v ^= v >> 13; v *= 3266489917u; v ^= v >> 16; }
    reinterpret_cast<uint32_t*>(sA)[i] = v;
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  if (tid == 0) { mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tb = tmem_base_s;
  if (tid == 0) {
    for (int it = 0; it < iters; ++it) {
      const int st = it & 3;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_i8(tb + (uint32_t)((it & 1) * 256), umma_desc_sw128(smem_u32(sA + st * I8_A_TILE) + k * 32),
                umma_desc_sw128(smem_u32(sB + st * 256 * 128) + k * 32), umma_idesc_u8s8(128, 256), 1u);
    }
    umma_commit(&done);
    mbar_wait_b(&done, 0, 0);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tb), "n"(512));
}

// sustained_random (optional): the same 2 s loop with pseudo-random operand bytes -- the issue rate the part holds under its power cap
// when the multipliers see operands like the engine's byte slices (the operand-dependent switching energy decides the clock)
int i8_peak(Ctx* ctx, double* burst, double* sustained, double* sustained_random) {
  const size_t smem = 4 * (size_t)I8_A_TILE + 4 * 256 * 128 + 1024;
  BOSIP_CUDA(ctx, cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = ctx->s_comp;
  const int iters = 10000;  // 2.7 ms per launch
  const double ops = 2.0 * (double)ctx->sms * iters * 4.0 * 128.0 * 256.0 * 32.0;
  i8_peak_kernel<<<ctx->sms, 128, smem, st>>>(200, 0);
  BOSIP_CUDA(ctx, cudaStreamSynchronize(st));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t0, st));
    i8_peak_kernel<<<ctx->sms, 128, smem, st>>>(iters, 0);
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t1, st));
    BOSIP_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    float ms = 0;
    BOSIP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    if (ms < best) best = ms;
  }
  *burst = ops / (best * 1e-3) / 1e12;
  for (int pattern = 0; pattern < 2; ++pattern) {  // back to back for ~2 s: what the part holds under its power cap; the rate of the last quarter is reported
    double* out = pattern ? sustained_random : sustained;
    if (!out) continue;
    const int launches = 800, tail = 200;
    for (int l = 0; l < launches; ++l) {
      if (l == launches - tail) BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t0, st));
      i8_peak_kernel<<<ctx->sms, 128, smem, st>>>(iters, pattern);
    }
    BOSIP_CUDA(ctx, cudaEventRecord(ctx->ev_t1, st));
    BOSIP_CUDA(ctx, cudaEventSynchronize(ctx->ev_t1));
    float ms = 0;
    BOSIP_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
    *out = ops * tail / (ms * 1e-3) / 1e12;
  }
  return 0;
}

}  // namespace bosip
