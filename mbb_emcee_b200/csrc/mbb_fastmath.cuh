// Lean FP64 exp / expm1 / reciprocal for the FAST arithmetic mode.
//
// Why not libdevice: on sm_100a its exp/expm1/division spend more issue slots on
// integer fix-ups, special-case branches and UMOVs that materialise the
// polynomial constants than on FP64 work (ncu, profiles/r01_*: 27% of the
// executed instructions of the first likelihood kernel were FP64, the issue
// stage -- not the FP64 pipe -- was the limiter).  These versions
//   * keep the polynomial coefficients in __constant__ memory, so every DFMA
//     takes its coefficient straight from the constant bank (no UMOV);
//   * have no branches: saturation is done by clamping the integer exponent
//     (integer pipe), valid for finite arguments -- callers screen NaN/inf;
//   * replace IEEE division by MUFU.RCP64H + two Newton steps + one residual
//     correction (<= 1 ulp).
// Accuracy (tests/test_device_logic_cpu.py::test_fastmath): exp, expm1 <= 1.5
// ulp over the ranges used; polynomial fits from tools/gen_poly.py (max
// relative fit error 1.6e-17 and 4.9e-18).
#pragma once
#include <cmath>
#include <cstring>

#include "mbb_model_defs.cuh"

namespace mbb {

#if defined(__CUDACC__)
__device__ __constant__ double kExpC_dev[12] = {
    1.0, 1.0, 0.50000000000000189, 0.1666666666666668, 0.041666666666487953, 0.008333333333319589,
    0.0013888888952352863, 0.00019841269890076403, 2.4801485441561313e-05, 2.7557240887229869e-06,
    2.763265472252779e-07, 2.5110049204818658e-08};
__device__ __constant__ double kEm1C_dev[11] = {
    0.5, 0.16666666666666671, 0.041666666666666671, 0.0083333333333261358, 0.0013888888888883748,
    0.00019841269874820627, 2.4801587325547743e-05, 2.7557255400206422e-06, 2.7557273643110297e-07,
    2.5105217004720745e-08, 2.0914686968086876e-09};
#endif

MBB_HD double exp_coef(int i) {
#if defined(__CUDA_ARCH__)
  return kExpC_dev[i];
#else
  const double c[12] = {1.0, 1.0, 0.50000000000000189, 0.1666666666666668, 0.041666666666487953,
                        0.008333333333319589, 0.0013888888952352863, 0.00019841269890076403,
                        2.4801485441561313e-05, 2.7557240887229869e-06, 2.763265472252779e-07,
                        2.5110049204818658e-08};
  return c[i];
#endif
}

MBB_HD double em1_coef(int i) {
#if defined(__CUDA_ARCH__)
  return kEm1C_dev[i];
#else
  const double c[11] = {0.5, 0.16666666666666671, 0.041666666666666671, 0.0083333333333261358,
                        0.0013888888888883748, 0.00019841269874820627, 2.4801587325547743e-05,
                        2.7557255400206422e-06, 2.7557273643110297e-07, 2.5105217004720745e-08,
                        2.0914686968086876e-09};
  return c[i];
#endif
}

// x = k ln2 + r, |r| <= ln2/2; k clamped to the normal exponent range.
MBB_HD void reduce_ln2(double x, int& k, double& r) {
  const double kMagic = 6755399441055744.0;          // 1.5 * 2^52
  const double kLog2e = 1.4426950408889634;
  const double kLn2Hi = 6.93147180369123816490e-01;  // fdlibm split of ln 2
  const double kLn2Lo = 1.90821492927058770002e-10;
  const double t = fma(x, kLog2e, kMagic);
  const double kf = t - kMagic;
#if defined(__CUDA_ARCH__)
  k = __double2loint(t);
#else
  k = (int)kf;
#endif
  r = fma(-kf, kLn2Hi, x);
  r = fma(-kf, kLn2Lo, r);
  k = k > 1023 ? 1023 : k;
  k = k < -1022 ? -1022 : k;
}

// p * 2^k for p in [0.5, 2), k in [-1022, 1023] (exponent-field addition)
MBB_HD double scale2(double p, int k) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
#else
  return ldexp(p, k);
#endif
}

MBB_HD double pow2i(int k) {   // 2^k, k in [-1022, 1023]
#if defined(__CUDA_ARCH__)
  return __hiloint2double((k + 1023) << 20, 0);
#else
  return ldexp(1.0, k);
#endif
}

MBB_HD double exp_fast(double x) {
  int k;
  double r;
  reduce_ln2(x, k, r);
  double p = exp_coef(11);
#pragma unroll
  for (int i = 10; i >= 0; --i) p = fma(p, r, exp_coef(i));
  return scale2(p, k);
}

MBB_HD double expm1_fast(double x) {
  int k;
  double r;
  reduce_ln2(x, k, r);
  double q = em1_coef(10);
#pragma unroll
  for (int i = 9; i >= 0; --i) q = fma(q, r, em1_coef(i));
  q = fma(r * r, q, r);                 // expm1(r)
  const double s = pow2i(k);
  return fma(s, q, s - 1.0);
}

// 1/b for normal b (no subnormal / zero handling)
MBB_HD double rcp_fast(double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / b;
#endif
}

// a / b with one residual correction (<= 1 ulp)
MBB_HD double div_fast(double a, double b) {
  const double r = rcp_fast(b);
  const double y = a * r;
  return fma(fma(-b, y, a), r, y);
}

// exp(b * (l_hi + l_lo)), product carried in double-double
MBB_HD double exp_prod_fast(double b, double l_hi, double l_lo) {
  const double y = b * l_hi;
  const double e = fma(b, l_lo, fma(b, l_hi, -y));
  const double r = exp_fast(y);
  return fma(r, e, r);
}

// ---------------------------------------------------------------------------
// Table-driven variants: x = (64 m + j - 32) ln2/64 + r, |r| <= ln2/128,
// T[j] = 2^((j-32)/64) in [0.71, 1.40), so that |x| < 0.34 always has m == 0:
//   exp(x)   = 2^m T[j] (1 + p),      p = expm1(r) = r h(r), h of degree 4
//   expm1(x) = 2^m T[j] p + (2^m T[j] - 1), with the separately rounded table
//              entry T[j]-1 when m == 0 (no cancellation for small |x|).
// 10-11 FP64 instructions instead of 17-18; the 1 KB table {T[j], T[j]-1}
// lives in global memory and stays L1-resident (one LDG.128 per call).
// ---------------------------------------------------------------------------
struct ExpTabEntry {
  double t, tm1;
};

#if defined(__CUDACC__)
__device__ const ExpTabEntry kExpTab_dev[64] = {
#include "mbb_exptab.inc"
};
__device__ __constant__ double kTabPoly_dev[5] = {1.0, 0.49999999999962602, 0.16666666666661323,
                                                 0.041666717628250034, 0.0083333406135586794};
#endif

MBB_HD ExpTabEntry exp_tab_entry(int j) {
#if defined(__CUDA_ARCH__)
  const double2 v = __ldg(reinterpret_cast<const double2*>(kExpTab_dev) + j);
  ExpTabEntry e;
  e.t = v.x;
  e.tm1 = v.y;
  return e;
#else
  static const ExpTabEntry tab[64] = {
#include "mbb_exptab.inc"
  };
  return tab[j];
#endif
}

MBB_HD double tab_poly_coef(int i) {
#if defined(__CUDA_ARCH__)
  return kTabPoly_dev[i];
#else
  const double c[5] = {1.0, 0.49999999999962602, 0.16666666666661323, 0.041666717628250034,
                       0.0083333406135586794};
  return c[i];
#endif
}

// x = (64 m + j - 32) ln2/64 + r; returns p = expm1(r)
template <int MMAX = 1023>
MBB_HD double reduce_tab(double x, int& m, int& j) {
  const double kMagic = 6755399441055744.0;
  const double k64Log2e = 92.332482616893657;                 // 64 / ln 2
  const double kLn2Hi64 = 6.93147180369123816490e-01 / 64.0;  // exact scalings of the fdlibm split
  const double kLn2Lo64 = 1.90821492927058770002e-10 / 64.0;
  const double t = fma(x, k64Log2e, kMagic);
  const double kf = t - kMagic;
#if defined(__CUDA_ARCH__)
  const int k = __double2loint(t);
#else
  const int k = (int)kf;
#endif
  double r = fma(-kf, kLn2Hi64, x);
  r = fma(-kf, kLn2Lo64, r);
  j = (k + 32) & 63;
  m = (k + 32) >> 6;
  m = m > MMAX ? MMAX : m;
  m = m < -1022 ? -1022 : m;
  double h = tab_poly_coef(4);
#pragma unroll
  for (int i = 3; i >= 0; --i) h = fma(h, r, tab_poly_coef(i));
  return h * r;
}

// MMAX caps the result at < 2^(MMAX+1): exp_tab<10> saturates near 2e3, which is
// what callers want when the value only feeds expm1(-t) (keeps its argument
// inside the range where the integer exponent extraction is valid).
template <int MMAX = 1023>
MBB_HD double exp_tab(double x) {
  int m, j;
  const double p = reduce_tab<MMAX>(x, m, j);
  const double T = exp_tab_entry(j).t;
  return scale2(fma(T, p, T), m);
}

MBB_HD double expm1_tab(double x) {
  int m, j;
  const double p = reduce_tab(x, m, j);
  const ExpTabEntry e = exp_tab_entry(j);
  const double sT = scale2(e.t, m);
  const double base = (m == 0) ? e.tm1 : sT - 1.0;
  return fma(sT, p, base);
}

template <int MMAX = 1023>
MBB_HD double exp_prod_tab(double b, double l_hi, double l_lo) {
  const double y = b * l_hi;
  const double e = fma(b, l_lo, fma(b, l_hi, -y));
  const double r = exp_tab<MMAX>(y);
  return fma(r, e, r);
}

// which lean exp family the FAST model code uses (A/B switch for measurements)
#ifndef MBB_USE_EXPTAB
#define MBB_USE_EXPTAB 1
#endif
#if MBB_USE_EXPTAB
MBB_HD double exp_l(double x) { return exp_tab(x); }
MBB_HD double expm1_l(double x) { return expm1_tab(x); }
MBB_HD double exp_prod_l(double b, double hi, double lo) { return exp_prod_tab(b, hi, lo); }
// saturating at ~2e3: for optical depths t that only feed expm1(-t)
MBB_HD double exp_tau(double x) { return exp_tab<10>(x); }
MBB_HD double exp_prod_tau(double b, double hi, double lo) { return exp_prod_tab<10>(b, hi, lo); }
#else
MBB_HD double exp_l(double x) { return exp_fast(x); }
MBB_HD double expm1_l(double x) { return expm1_fast(x); }
MBB_HD double exp_prod_l(double b, double hi, double lo) { return exp_prod_fast(b, hi, lo); }
MBB_HD double exp_tau(double x) { return fmin(exp_fast(x), 2000.0); }
MBB_HD double exp_prod_tau(double b, double hi, double lo) { return fmin(exp_prod_fast(b, hi, lo), 2000.0); }
#endif
// The lean exp family extracts the binary exponent from the low 32 bits of
// x*64/ln2 + 1.5*2^52: valid for |x| < 2^31 ln2/64 ~ 2.3e7.  fast_setup gates
// the parameters (T >= 1e-3 K, beta, alpha <= 300) so every argument formed
// from them stays far inside that range.
constexpr double kFastMaxIndex = 300.0;
constexpr double kFastMinT = 1e-3;

}  // namespace mbb
