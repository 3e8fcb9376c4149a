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
// exp(-s) = 2^m * T[j] * e^r with n = rint(-s * 256/ln2) = 256 m + j, T[j] = 2^(j/256) (correctly rounded), |r| <= ln2/512: a degree-4
// polynomial is enough (truncation 3.8e-17).  8 FP64-pipe instructions instead of 16; the table (2 KB) lives in shared memory.
// The argument reduction uses ln2/256 rounded to double (off by 9.1e-20): the result carries a relative error of up to
// 3.4e-17 * s on top of the usual ~1.5 ulp, i.e. an ABSOLUTE error <= 3.4e-17 * s * exp(-s) <= 1.3e-17.  Every consumer sums
// kernel values against O(1) ones (K* rows, MMD means), so the absolute bound is the one that matters.
constexpr int EXP_TAB_N = 256;
__device__ const double EXP2_256TH[EXP_TAB_N] = {
    0x1.0000000000000p+0, 0x1.00b1afa5abcbfp+0, 0x1.0163da9fb3335p+0, 0x1.02168143b0281p+0, 0x1.02c9a3e778061p+0, 0x1.037d42e11bbccp+0,
    0x1.04315e86e7f85p+0, 0x1.04e5f72f654b1p+0, 0x1.059b0d3158574p+0, 0x1.0650a0e3c1f89p+0, 0x1.0706b29ddf6dep+0, 0x1.07bd42b72a836p+0,
    0x1.0874518759bc8p+0, 0x1.092bdf66607e0p+0, 0x1.09e3ecac6f383p+0, 0x1.0a9c79b1f3919p+0, 0x1.0b5586cf9890fp+0, 0x1.0c0f145e46c85p+0,
    0x1.0cc922b7247f7p+0, 0x1.0d83b23395decp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.0efa55fdfa9c5p+0, 0x1.0fb66affed31bp+0, 0x1.1073028d7233ep+0,
    0x1.11301d0125b51p+0, 0x1.11edbab5e2ab6p+0, 0x1.12abdc06c31ccp+0, 0x1.136a814f204abp+0, 0x1.1429aaea92de0p+0, 0x1.14e95934f312ep+0,
    0x1.15a98c8a58e51p+0, 0x1.166a45471c3c2p+0, 0x1.172b83c7d517bp+0, 0x1.17ed48695bbc0p+0, 0x1.18af9388c8deap+0, 0x1.1972658375d2fp+0,
    0x1.1a35beb6fcb75p+0, 0x1.1af99f8138a1cp+0, 0x1.1bbe084045cd4p+0, 0x1.1c82f95281c6bp+0, 0x1.1d4873168b9aap+0, 0x1.1e0e75eb44027p+0,
    0x1.1ed5022fcd91dp+0, 0x1.1f9c18438ce4dp+0, 0x1.2063b88628cd6p+0, 0x1.212be3578a819p+0, 0x1.21f49917ddc96p+0, 0x1.22bdda27912d1p+0,
    0x1.2387a6e756238p+0, 0x1.2451ffb82140ap+0, 0x1.251ce4fb2a63fp+0, 0x1.25e85711ece75p+0, 0x1.26b4565e27cddp+0, 0x1.2780e341ddf29p+0,
    0x1.284dfe1f56381p+0, 0x1.291ba7591bb70p+0, 0x1.29e9df51fdee1p+0, 0x1.2ab8a66d10f13p+0, 0x1.2b87fd0dad990p+0, 0x1.2c57e39771b2fp+0,
    0x1.2d285a6e4030bp+0, 0x1.2df961f641589p+0, 0x1.2ecafa93e2f56p+0, 0x1.2f9d24abd886bp+0, 0x1.306fe0a31b715p+0, 0x1.31432edeeb2fdp+0,
    0x1.32170fc4cd831p+0, 0x1.32eb83ba8ea32p+0, 0x1.33c08b26416ffp+0, 0x1.3496266e3fa2dp+0, 0x1.356c55f929ff1p+0, 0x1.36431a2de883bp+0,
    0x1.371a7373aa9cbp+0, 0x1.37f26231e754ap+0, 0x1.38cae6d05d866p+0, 0x1.39a401b7140efp+0, 0x1.3a7db34e59ff7p+0, 0x1.3b57fbfec6cf4p+0,
    0x1.3c32dc313a8e5p+0, 0x1.3d0e544ede173p+0, 0x1.3dea64c123422p+0, 0x1.3ec70df1c5175p+0, 0x1.3fa4504ac801cp+0, 0x1.40822c367a024p+0,
    0x1.4160a21f72e2ap+0, 0x1.423fb2709468ap+0, 0x1.431f5d950a897p+0, 0x1.43ffa3f84b9d4p+0, 0x1.44e086061892dp+0, 0x1.45c2042a7d232p+0,
    0x1.46a41ed1d0057p+0, 0x1.4786d668b3237p+0, 0x1.486a2b5c13cd0p+0, 0x1.494e1e192aed2p+0, 0x1.4a32af0d7d3dep+0, 0x1.4b17dea6db7d7p+0,
    0x1.4bfdad5362a27p+0, 0x1.4ce41b817c114p+0, 0x1.4dcb299fddd0dp+0, 0x1.4eb2d81d8abffp+0, 0x1.4f9b2769d2ca7p+0, 0x1.508417f4531eep+0,
    0x1.516daa2cf6642p+0, 0x1.5257de83f4eefp+0, 0x1.5342b569d4f82p+0, 0x1.542e2f4f6ad27p+0, 0x1.551a4ca5d920fp+0, 0x1.56070dde910d2p+0,
    0x1.56f4736b527dap+0, 0x1.57e27dbe2c4cfp+0, 0x1.58d12d497c7fdp+0, 0x1.59c0827ff07ccp+0, 0x1.5ab07dd485429p+0, 0x1.5ba11fba87a03p+0,
    0x1.5c9268a5946b7p+0, 0x1.5d84590998b93p+0, 0x1.5e76f15ad2148p+0, 0x1.5f6a320dceb71p+0, 0x1.605e1b976dc09p+0, 0x1.6152ae6cdf6f4p+0,
    0x1.6247eb03a5585p+0, 0x1.633dd1d1929fdp+0, 0x1.6434634ccc320p+0, 0x1.652b9febc8fb7p+0, 0x1.6623882552225p+0, 0x1.671c1c70833f6p+0,
    0x1.68155d44ca973p+0, 0x1.690f4b19e9538p+0, 0x1.6a09e667f3bcdp+0, 0x1.6b052fa75173ep+0, 0x1.6c012750bdabfp+0, 0x1.6cfdcddd47645p+0,
    0x1.6dfb23c651a2fp+0, 0x1.6ef9298593ae5p+0, 0x1.6ff7df9519484p+0, 0x1.70f7466f42e87p+0, 0x1.71f75e8ec5f74p+0, 0x1.72f8286ead08ap+0,
    0x1.73f9a48a58174p+0, 0x1.74fbd35d7cbfdp+0, 0x1.75feb564267c9p+0, 0x1.77024b1ab6e09p+0, 0x1.780694fde5d3fp+0, 0x1.790b938ac1cf6p+0,
    0x1.7a11473eb0187p+0, 0x1.7b17b0976cfdbp+0, 0x1.7c1ed0130c132p+0, 0x1.7d26a62ff86f0p+0, 0x1.7e2f336cf4e62p+0, 0x1.7f3878491c491p+0,
    0x1.80427543e1a12p+0, 0x1.814d2add106d9p+0, 0x1.82589994cce13p+0, 0x1.8364c1eb941f7p+0, 0x1.8471a4623c7adp+0, 0x1.857f4179f5b21p+0,
    0x1.868d99b4492edp+0, 0x1.879cad931a436p+0, 0x1.88ac7d98a6699p+0, 0x1.89bd0a478580fp+0, 0x1.8ace5422aa0dbp+0, 0x1.8be05bad61778p+0,
    0x1.8cf3216b5448cp+0, 0x1.8e06a5e0866d9p+0, 0x1.8f1ae99157736p+0, 0x1.902fed0282c8ap+0, 0x1.9145b0b91ffc6p+0, 0x1.925c353aa2fe2p+0,
    0x1.93737b0cdc5e5p+0, 0x1.948b82b5f98e5p+0, 0x1.95a44cbc8520fp+0, 0x1.96bdd9a7670b3p+0, 0x1.97d829fde4e50p+0, 0x1.98f33e47a22a2p+0,
    0x1.9a0f170ca07bap+0, 0x1.9b2bb4d53fe0dp+0, 0x1.9c49182a3f090p+0, 0x1.9d674194bb8d5p+0, 0x1.9e86319e32323p+0, 0x1.9fa5e8d07f29ep+0,
    0x1.a0c667b5de565p+0, 0x1.a1e7aed8eb8bbp+0, 0x1.a309bec4a2d33p+0, 0x1.a42c980460ad8p+0, 0x1.a5503b23e255dp+0, 0x1.a674a8af46052p+0,
    0x1.a799e1330b358p+0, 0x1.a8bfe53c12e59p+0, 0x1.a9e6b5579fdbfp+0, 0x1.ab0e521356ebap+0, 0x1.ac36bbfd3f37ap+0, 0x1.ad5ff3a3c2774p+0,
    0x1.ae89f995ad3adp+0, 0x1.afb4ce622f2ffp+0, 0x1.b0e07298db666p+0, 0x1.b20ce6c9a8952p+0, 0x1.b33a2b84f15fbp+0, 0x1.b468415b749b1p+0,
    0x1.b59728de5593ap+0, 0x1.b6c6e29f1c52ap+0, 0x1.b7f76f2fb5e47p+0, 0x1.b928cf22749e4p+0, 0x1.ba5b030a1064ap+0, 0x1.bb8e0b79a6f1fp+0,
    0x1.bcc1e904bc1d2p+0, 0x1.bdf69c3f3a207p+0, 0x1.bf2c25bd71e09p+0, 0x1.c06286141b33dp+0, 0x1.c199bdd85529cp+0, 0x1.c2d1cd9fa652cp+0,
    0x1.c40ab5fffd07ap+0, 0x1.c544778fafb22p+0, 0x1.c67f12e57d14bp+0, 0x1.c7ba88988c933p+0, 0x1.c8f6d9406e7b5p+0, 0x1.ca3405751c4dbp+0,
    0x1.cb720dcef9069p+0, 0x1.ccb0f2e6d1675p+0, 0x1.cdf0b555dc3fap+0, 0x1.cf3155b5bab74p+0, 0x1.d072d4a07897cp+0, 0x1.d1b532b08c968p+0,
    0x1.d2f87080d89f2p+0, 0x1.d43c8eacaa1d6p+0, 0x1.d5818dcfba487p+0, 0x1.d6c76e862e6d3p+0, 0x1.d80e316c98398p+0, 0x1.d955d71ff6075p+0,
    0x1.da9e603db3285p+0, 0x1.dbe7cd63a8315p+0, 0x1.dd321f301b460p+0, 0x1.de7d5641c0658p+0, 0x1.dfc97337b9b5fp+0, 0x1.e11676b197d17p+0,
    0x1.e264614f5a129p+0, 0x1.e3b333b16ee12p+0, 0x1.e502ee78b3ff6p+0, 0x1.e653924676d76p+0, 0x1.e7a51fbc74c83p+0, 0x1.e8f7977cdb740p+0,
    0x1.ea4afa2a490dap+0, 0x1.eb9f4867cca6ep+0, 0x1.ecf482d8e67f1p+0, 0x1.ee4aaa2188510p+0, 0x1.efa1bee615a27p+0, 0x1.f0f9c1cb6412ap+0,
    0x1.f252b376bba97p+0, 0x1.f3ac948dd7274p+0, 0x1.f50765b6e4540p+0, 0x1.f6632798844f8p+0, 0x1.f7bfdad9cbe14p+0, 0x1.f91d802243c89p+0,
    0x1.fa7c1819e90d8p+0, 0x1.fbdba3692d514p+0, 0x1.fd3c22b8f71f1p+0, 0x1.fe9d96b2a23d9p+0};

// Shared-memory copy of the table, each entry replicated REP times: lane l reads copy (l & (REP - 1)), which removes bank
// conflicts between lanes (REP = 16: none).  Measured on B200 this buys nothing (cross-kernel 3.49 / 3.55 / 3.86 ms per 303 104 C3
// points at REP = 1 / 8 / 16; the 8 KB copy costs the fourth resident CTA): these loops are bound by the issue port, which an
// FP64 instruction holds for two cycles (time ~ 2 * FP64 + other instructions), not by shared-memory wavefronts.  So REP = 1.
template <int REP>
__device__ __forceinline__ void load_exp_table(double* tab /* shared, EXP_TAB_N * REP doubles */) {
  for (int i = threadIdx.x; i < EXP_TAB_N * REP; i += blockDim.x) tab[i] = EXP2_256TH[i / REP];
}
template <int REP>
__device__ __forceinline__ const double* exp_table_lane(const double* tab) { return tab + (threadIdx.x & (REP - 1)); }

// tab: the calling lane's view of the table (exp_table_lane).  CUT = false skips the underflow cut-off: only for callers whose
// argument is bounded well below 708.
// E2: the result is scaled by 2^E2 for free (added to the exponent patch) -- the INT8 slice writer wants k * 2^(8 SA - 1)
template <int REP, bool CUT = true, int E2 = 0>
__device__ __forceinline__ double fast_exp_neg_t(double s, const double* __restrict__ tab) {
  const double x = -s;
  const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, 369.3299304675746, MAGIC);        // 256 / ln2
  const double n = t - MAGIC;
  const double r = fma(n, -0x1.62e42fefa39efp-9, x);       // ln2/256
  double q = fma(r, 4.1666666666666664e-02, 1.6666666666666666e-01);
  q = fma(q, r, 0.5);
  const double p = fma(q, r * r, r);                        // e^r - 1
  const int ni = __double2loint(t);                         // n as a two's-complement integer
  const double T = tab[(ni & (EXP_TAB_N - 1)) * REP];
  const double v0 = fma(T, p, T);
  const int hi = __double2hiint(v0) + (int)((unsigned)((ni >> 8) + E2) << 20);
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

template <int KERNEL, int REP, bool CUT = true, int E2 = 0>
__device__ __forceinline__ double kernel_from_r2_t(double r2, const double* __restrict__ tab) {
  if (KERNEL == 0) return fast_exp_neg_t<REP, CUT, E2>(r2, tab);
  const double s = fast_sqrt1(r2);
  if (KERNEL == 1) return (1.0 + s) * fast_exp_neg_t<REP, CUT, E2>(s, tab);
  return (1.0 + s + r2 * (1.0 / 3.0)) * fast_exp_neg_t<REP, CUT, E2>(s, tab);
}

}  // namespace bosip
