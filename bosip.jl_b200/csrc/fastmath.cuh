// Branch-free FP64 exp / sqrt for the kernel-evaluation hot loops (cross-kernel, K build, MMD).
// The CUDA library versions are accurate but spend ~50 integer/conversion/branch instructions per value around ~30 FP64
// ones (ncu: 94 warp instructions per Matern-5/2 value, FP64 pipe only 60 % busy); these keep the FP64 work and drop the rest.
// Accuracy: <= 2 ulp (checked against the library on the GPU by tests/test_gpu_parity.py::test_fast_math_accuracy).
#pragma once

namespace bosip {

// exp(-s) for s >= 0 (any finite s; flushes to 0 below 2^-1021).  Cody-Waite reduction s = n ln2 - r, |r| <= ln2/2,
// degree-11 Chebyshev-interpolated polynomial (max relative error 4.2e-18 before rounding), exponent patched by integer add.
__device__ __forceinline__ double fast_exp_neg(double s) {
  const double x = -s;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: adding it rounds to the nearest integer in the low mantissa bits
  const double t = fma(x, 1.4426950408889634, MAGIC);
  const double n = t - MAGIC;
  double r = fma(n, -6.93147180369123816490e-01, x);
  r = fma(n, -1.90821492927058770002e-10, r);
  double p = 2.5110049204818658e-08;
  p = fma(p, r, 2.763265472252779e-07);
  p = fma(p, r, 2.755724088722987e-06);
  p = fma(p, r, 2.4801485441561313e-05);
  p = fma(p, r, 0.00019841269890076403);
  p = fma(p, r, 0.0013888888952352863);
  p = fma(p, r, 0.008333333333319589);
  p = fma(p, r, 0.04166666666648795);
  p = fma(p, r, 0.1666666666666668);
  p = fma(p, r, 0.5000000000000019);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int ni = __double2loint(t);  // n as a two's-complement integer
  const double v = __hiloint2double(__double2hiint(p) + (int)((unsigned)ni << 20), __double2loint(p));
  return s > 708.0 ? 0.0 : v;
}

// sqrt(x) for x >= 0 (x in the normal range or 0; tiny inputs clamp to 1e-290 -> 1e-145, i.e. 0 for our purposes).
// Hardware reciprocal-square-root seed + two coupled Newton (Goldschmidt) steps + one residual correction.
__device__ __forceinline__ double fast_sqrt(double x) {
  const double xc = fmax(x, 1e-290);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xc));
  double g = xc * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  const double d = fma(-g, g, xc);
  g = fma(d, h, g);
  return x > 0.0 ? g : 0.0;
}

// k as a function of the squared scaled distance (the kernel constant 1/2, 3, 5 is already folded into r2)
template <int KERNEL>
__device__ __forceinline__ double kernel_from_r2(double r2) {
  if (KERNEL == 0) return fast_exp_neg(r2);                       // SE: exp(-r^2/2)
  const double s = fast_sqrt(r2);
  if (KERNEL == 1) return (1.0 + s) * fast_exp_neg(s);            // Matern-3/2: (1 + sqrt3 r) exp(-sqrt3 r)
  return (1.0 + s + r2 * (1.0 / 3.0)) * fast_exp_neg(s);           // Matern-5/2: (1 + sqrt5 r + 5 r^2/3) exp(-sqrt5 r)
}


// ------------------------------------------------------------------------------------------------ table-driven variants (cross-kernel)
// exp(-s) = 2^m * T[j] * e^r with n = rint(-s * 64/ln2) = 64 m + j, T[j] = 2^(j/64) (correctly rounded), |r| <= ln2/128: a degree-5
// polynomial is enough (truncation 3.5e-17).  9 FP64-pipe instructions instead of 16; the table (512 B) lives in shared memory.
// The argument reduction uses ln2/64 rounded to double (off by 3.6e-19): the result carries a relative error of up to
// 3.4e-17 * s on top of the usual ~1.5 ulp, i.e. an ABSOLUTE error <= 3.4e-17 * s * exp(-s) <= 1.3e-17.  Every consumer sums
// kernel values against O(1) ones (K* rows, MMD means), so the absolute bound is the one that matters.
__device__ const double EXP2_64TH[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0, 0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0,
    0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0, 0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0, 0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0,
    0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0, 0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0, 0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0,
    0x1.6247eb03a5585p+0, 0x1.6623882552225p+0, 0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0, 0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0,
    0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0, 0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0, 0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0,
    0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0, 0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};

// Shared-memory copy of the table, each entry replicated REP times: lane l reads copy (l & (REP - 1)), which removes bank
// conflicts between lanes (REP = 16: none).  Measured on B200 this buys nothing (cross-kernel 3.49 / 3.55 / 3.86 ms per 303 104 C3
// points at REP = 1 / 8 / 16; the 8 KB copy costs the fourth resident CTA): these loops are bound by the issue port, which an
// FP64 instruction holds for two cycles (time ~ 2 * FP64 + other instructions), not by shared-memory wavefronts.  So REP = 1.
template <int REP>
__device__ __forceinline__ void load_exp_table(double* tab /* shared, 64 * REP doubles */) {
  for (int i = threadIdx.x; i < 64 * REP; i += blockDim.x) tab[i] = EXP2_64TH[i / REP];
}
template <int REP>
__device__ __forceinline__ const double* exp_table_lane(const double* tab) { return tab + (threadIdx.x & (REP - 1)); }

// tab: the calling lane's view of the table (exp_table_lane).  CUT = false skips the underflow cut-off: only for callers whose
// argument is bounded well below 708.
template <int REP, bool CUT = true>
__device__ __forceinline__ double fast_exp_neg_t(double s, const double* __restrict__ tab) {
  const double x = -s;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 92.33248261689366, MAGIC);       // 64 / ln2
  const double n = t - MAGIC;
  const double r = fma(n, -0x1.62e42fefa39efp-7, x);       // ln2/64
  double q = fma(r, 8.333333333333333e-03, 4.1666666666666664e-02);
  q = fma(q, r, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  const double p = fma(q, r * r, r);                        // e^r - 1
  const int ni = __double2loint(t);                         // n as a two's-complement integer
  const double T = tab[(ni & 63) * REP];
  const double v0 = fma(T, p, T);
  const int hi = __double2hiint(v0) + (int)((unsigned)(ni >> 6) << 20);
  if (!CUT) return __hiloint2double(hi, __double2loint(v0));
  // s in (708.06, +Inf] -> (almost) 0, decided on the high word in the integer pipe (negative s and NaN fall through).  Only the
  // high word is cleared: the result is a denormal below 2^-1042 rather than an exact zero, one select instead of two.
  const unsigned hs = (unsigned)__double2hiint(s);
  return __hiloint2double((hs - 0x40862001u) <= (0x7ff00000u - 0x40862001u) ? 0 : hi, __double2loint(v0));
}

// sqrt(x + 1e-300) for x >= 0 with ONE coupled Newton step before the residual correction: the hardware seed is good to ~2^-22,
// one step gives 2^-44 and the correction squares that again.  The offset replaces the compare/select guards around x = 0
// (rsqrt(0) = Inf): it is invisible for x >= 1e-284, and x = 0 gives 1e-150, which every kernel formula maps to k = 1 exactly.
__device__ __forceinline__ double fast_sqrt1(double x) {
  const double xc = x + 1e-300;
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xc));
  double g = xc * y;
  double h = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));   // y / 2 (y is a normal number)
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  const double d = fma(-g, g, xc);
  return fma(d, h, g);
}

template <int KERNEL, int REP, bool CUT = true>
__device__ __forceinline__ double kernel_from_r2_t(double r2, const double* __restrict__ tab) {
  if (KERNEL == 0) return fast_exp_neg_t<REP, CUT>(r2, tab);
  const double s = fast_sqrt1(r2);
  if (KERNEL == 1) return (1.0 + s) * fast_exp_neg_t<REP, CUT>(s, tab);
  return (1.0 + s + r2 * (1.0 / 3.0)) * fast_exp_neg_t<REP, CUT>(s, tab);
}

}  // namespace bosip
