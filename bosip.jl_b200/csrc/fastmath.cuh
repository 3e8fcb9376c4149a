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

}  // namespace bosip
