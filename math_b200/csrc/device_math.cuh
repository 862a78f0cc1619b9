// FP64 scalar kernels used by the GLM link functions, following the
// reference's branches so results agree to a few ulp
// (reference: stan/math/prim/fun/log1p_exp.hpp L45-52, log1m_exp.hpp L47-57,
// multiply_log.hpp L49-56, lgamma.hpp L63-67, digamma.hpp L47-49).
#pragma once
#include <cuda_runtime.h>

#include <cmath>

namespace smc {

// exp(x) for x <= 0 without a branch: Cody-Waite reduction by ln 2, degree-13 Taylor
// polynomial in Horner form (|r| <= 0.347: truncation 4e-18), scaling through the
// exponent field.  Worst error over [-708, 0] against expl: 0.88 ulp (libdevice:
// ~1 ulp).  Below -708, where the result would be subnormal, it returns 0 -- every
// caller adds the value to something >= 1 or uses it as e in e / (1 + e).  Being
// straight-line code, the two independent evaluations of the ordered link interleave
// (libdevice's exp carries a range branch that keeps the compiler from overlapping
// two calls): 0.831 -> 0.813 ms on config 5b.  Where one exp per element sits in a
// rolled loop (the categorical kernels) libdevice's shorter sequence wins by 1-5 %,
// so those keep it.
__device__ __forceinline__ double exp_nonpos(double x) {
  const double xc = x < -708.0 ? -708.0 : x;  // (a NaN stays a NaN)
  const double n = rint(xc * 1.4426950408889634074);
  double r = fma(n, -6.93147180369123816490e-01, xc);
  r = fma(n, -1.90821492927058770002e-10, r);
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double s = __longlong_as_double((__double2ll_rn(n) + 1023ll) << 52);
  return x < -708.0 ? 0.0 : p * s;
}

// log1p(e) for e >= 0 on ONE code path: log(u) + (e - (u - 1)) / u with u = 1 + e
// (the second term restores what rounding u lost; it is below ulp(1), so a single-
// precision reciprocal is enough).  libdevice's log1p switches between two
// evaluation schemes on the size of its argument and calls the out-of-line FP64
// division: in a warp whose lanes fall on both sides -- the rule for
// e = exp(-|theta|) in (0, 1] -- both schemes run back to back.  Worst error over
// [0, 1] incl. tiny arguments: 1.5 ulp with a 0.5-ulp log (glibc's log1p: 0.83);
// exact 0 at e = 0.
__device__ __forceinline__ double log1p_nonneg(double e) {
  const double u = 1.0 + e;
  const double c = e - (u - 1.0);
  return log(u) + c * (double)__frcp_rn((float)u);
}

__device__ __forceinline__ double log1p_exp(double a) {
  if (a > 0.0) return a + log1p_nonneg(exp(-a));
  return log1p_nonneg(exp(a));
}

__device__ __forceinline__ double log1m_exp(double a) {
  if (a > 0.0) return __longlong_as_double(0x7ff8000000000000ll);
  if (a > -0.693147) return log(-expm1(a));
  return log1p(-exp(a));
}

__device__ __forceinline__ double multiply_log(double a, double b) {
  if (b == 0.0 && a == 0.0) return 0.0;
  return a * log(b);
}

// boost::math::digamma, double-precision branch (Boost 1.84.0): reflection for
// x <= -1, asymptotic series for x >= 10, otherwise recurrence into [1,2] and
// the minimax rational (x - root)(Y + P(x-1)/Q(x-1)).
__host__ __device__ inline double digamma(double x) {
  double result = 0.0;
  if (x <= -1.0) {
    x = 1.0 - x;
    double rem = x - floor(x);
    if (rem > 0.5) rem -= 1.0;
    if (rem == 0.0) return nan("");
    result = 3.14159265358979323846 / tan(3.14159265358979323846 * rem);
  }
  if (x == 0.0) return nan("");
  if (x >= 10.0) {
    const double xm = x - 1.0;
    double r = log(xm) + 1.0 / (2.0 * xm);
    const double z = 1.0 / (xm * xm);
    double p = -0.44325980392156862745098039215686274510;
    p = fma(p, z, 0.083333333333333333333333333333333333333);
    p = fma(p, z, -0.021092796092796092796092796092796092796);
    p = fma(p, z, 0.0075757575757575757575757575757575757576);
    p = fma(p, z, -0.0041666666666666666666666666666666666667);
    p = fma(p, z, 0.003968253968253968253968253968253968254);
    p = fma(p, z, -0.0083333333333333333333333333333333333333);
    p = fma(p, z, 0.083333333333333333333333333333333333333);
    r -= z * p;
    return result + r;
  }
  while (x > 2.0) {
    x -= 1.0;
    result += 1.0 / x;
  }
  while (x < 1.0) {
    result -= 1.0 / x;
    x += 1.0;
  }
  const double Y = 0.99558162689208984375;  // (float)0.99558162689208984
  const double root1 = 1569415565.0 / 1073741824.0;
  const double root2 = (381566830.0 / 1073741824.0) / 1073741824.0;
  const double root3 = 0.9016312093258695918615325266959189453125e-19;
  double g = x - root1;
  g -= root2;
  g -= root3;
  const double t = x - 1.0;
  double p = -0.0020713321167745952;
  p = fma(p, t, -0.045251321448739056);
  p = fma(p, t, -0.28919126444774784);
  p = fma(p, t, -0.65031853770896507);
  p = fma(p, t, -0.32555031186804491);
  p = fma(p, t, 0.25479851061131551);
  double q = -0.55789841321675513e-6;
  q = fma(q, t, 0.0021284987017821144);
  q = fma(q, t, 0.054151797245674225);
  q = fma(q, t, 0.43593529692665969);
  q = fma(q, t, 1.4606242909763515);
  q = fma(q, t, 2.0767117023730469);
  q = fma(q, t, 1.0);
  const double r = p / q;
  return result + (g * Y + g * r);
}

}  // namespace smc
