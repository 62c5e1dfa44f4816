// Modified-blackbody SED model: per-walker setup, node evaluation, root solves.
//
// Everything here is `__host__ __device__` so that the exact same source can be
// compiled (a) into the sm_100a kernels of mbb_kernels.cu and (b) into the
// host-side emulation library the CPU test-suite uses to check the *logic*
// (tests/hostemu, never shipped, never a fallback).
//
// Two arithmetic modes, selected per context (mbb_set_math_mode):
//   FAITHFUL  - the reference's formulas in the reference's evaluation order
//               (fnu.pyx:9-108, modified_blackbody.py:228-337, 441-491):
//               pow / expm1 / exp per node.
//   FAST      - algebraically identical, restructured so that no pow() is
//               evaluated per node: every power becomes exp(b * L) with
//               L = log(wave_i / wavenorm) precomputed on the host in
//               double-double, and the normalisation is applied as ratios
//               (f/fnorm near 1).  ~2.5x fewer FP64 instructions per node;
//               agrees with FAITHFUL to a few 1e-16 (tests/test_parity_gpu.py).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MBB_HD __host__ __device__ __forceinline__
#define MBB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define MBB_HD inline
#define MBB_HD_NOINLINE inline
#endif

namespace mbb {

// Physical constants exactly as the reference hardwires them
// (modified_blackbody.py:15-18, fnu.pyx:14-15).
constexpr double kH = 6.6260693e-34;      // J s
constexpr double kK = 1.3806505e-23;      // J / K
constexpr double kCum = 299792458e6;      // um / s
constexpr double kUmToGHz = 299792458e-3; // um * GHz
constexpr double kInf = __builtin_huge_val();

// Per-evaluation status.  0 and 1 are normal outcomes; the rest map to the
// exceptions the reference raises (SURVEY.md 8b "Errors").
enum Status : int {
  ST_OK = 0,
  ST_BELOW_LOWLIM = 1,   // likelihood.py:806-807 -> -inf
  ST_BAD_ALPHA = 2,      // modified_blackbody.py:219-221 ValueError
  ST_BAD_BETA = 3,       // :222-224 ValueError
  ST_BRACKET_LOW = 4,    // :294-300 ValueError
  ST_BRACKET_HIGH = 5,   // :310-316 ValueError
  ST_NO_CONVERGE = 6,    // brentq RuntimeError
  ST_OVERFLOW = 7,       // :326-328 OverflowError
  ST_PEAK_BRACKET = 8,   // :612-630 Exception
  ST_NONFINITE = 9       // NaN/inf parameters or result
};

struct Sed {
  double T, beta, lambda0, alpha, fnorm;
  double hokt9;    // 1e9*h/(k*T): array (Cython) path, fnu.pyx:16
  double hokt_e9;  // (h/(k*T))*1e9: scalar (numpy) path, modified_blackbody.py:461-464
  double hcokt, xnorm, x0, normfac, xmerge, kappa;
  // FAST-mode amplitudes (see header comment)
  double amp_grey, amp_pow, q_hi, q_lo;
  int status;
};

// ---------------------------------------------------------------------------
// small math helpers
// ---------------------------------------------------------------------------
MBB_HD bool finite_d(double x) { return x - x == 0.0; }

// exp(b * (l_hi + l_lo)) with the product carried in double-double.
MBB_HD double exp_prod(double b, double l_hi, double l_lo) {
  double y = b * l_hi;
  double e = fma(b, l_hi, -y) + b * l_lo;
  double r = exp(y);
  return fma(r, e, r);
}

// Python's float ** float raises OverflowError when a finite base overflows;
// callers test the result with finite_d().
MBB_HD double ppow(double x, double y) { return pow(x, y); }

// ---------------------------------------------------------------------------
// Merge-point equation, modified_blackbody.py:122-151 (alpha_merge_eqn).
// Python's OverflowError from `**` or expm1 => bterm = 0.
// ---------------------------------------------------------------------------
MBB_HD double merge_residual(double x, double alpha, double beta, double x0) {
  double t = ppow(x / x0, beta);
  double bterm;
  if (!(t <= 709.782712893384) ) {          // expm1 overflow / inf / (NaN -> 0)
    bterm = 0.0;
  } else if (t == 0.0) {
    bterm = 1.0;                            // limit t/expm1(t); reference would divide by zero
  } else {
    bterm = t / expm1(t);
  }
  return x - (1.0 - exp(-x)) * (3.0 + alpha + beta * bterm);
}

// ---------------------------------------------------------------------------
// Brent's method with scipy.optimize.brentq's defaults (xtol=2e-12,
// rtol=4*eps, maxiter=100), used at modified_blackbody.py:321 and :633.
// The classic algorithm (Brent 1973, ch. 4): inverse quadratic interpolation /
// secant with a bisection safeguard.  Same tolerances and the same update
// rules => the iterates track the reference's to rounding, so the returned
// root agrees far inside its own 2e-12 tolerance.
// ---------------------------------------------------------------------------
template <class F>
MBB_HD double brent_root(F f, double xa, double xb, double fa, double fb, int& status) {
  const double xtol = 2e-12, rtol = 8.881784197001252e-16;
  double xpre = xa, xcur = xb, fpre = fa, fcur = fb;
  double xblk = 0.0, fblk = 0.0, spre = 0.0, scur = 0.0;
  if (fpre == 0.0) return xpre;
  if (fcur == 0.0) return xcur;
  for (int it = 0; it < 100; ++it) {
    if (fpre != 0.0 && fcur != 0.0 && ((fpre < 0.0) != (fcur < 0.0))) {
      xblk = xpre; fblk = fpre;
      spre = scur = xcur - xpre;
    }
    if (fabs(fblk) < fabs(fcur)) {
      xpre = xcur; xcur = xblk; xblk = xpre;
      fpre = fcur; fcur = fblk; fblk = fpre;
    }
    double delta = (xtol + rtol * fabs(xcur)) / 2.0;
    double sbis = (xblk - xcur) / 2.0;
    if (fcur == 0.0 || fabs(sbis) < delta) return xcur;
    if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
      double stry;
      if (xpre == xblk) {
        stry = -fcur * (xcur - xpre) / (fcur - fpre);
      } else {
        double dpre = (fpre - fcur) / (xpre - xcur);
        double dblk = (fblk - fcur) / (xblk - xcur);
        stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
      }
      double lim = fmin(fabs(spre), 3.0 * fabs(sbis) - delta);
      if (2.0 * fabs(stry) < lim) { spre = scur; scur = stry; }
      else { spre = sbis; scur = sbis; }
    } else {
      spre = sbis; scur = sbis;
    }
    xpre = xcur; fpre = fcur;
    if (fabs(scur) > delta) xcur += scur;
    else xcur += (sbis > 0.0 ? delta : -delta);
    fcur = f(xcur);
  }
  status = ST_NO_CONVERGE;
  return xcur;
}

// Root of  x - a (1 - e^-x) = 0,  a = 3 + alpha + beta  (optically thin merge
// point).  The reference writes the closed form  a + W0(-a e^-a)
// (modified_blackbody.py:253-254, scipy.special.lambertw); the same number is
// obtained here by Newton iteration from x = a (monotone, quadratic: the
// residual is convex and positive at a).
MBB_HD double thin_merge_root(double a) {
  double x = a;
  for (int it = 0; it < 12; ++it) {
    double e = a * exp(-x);
    double fx = (x - a) + e;
    double dx = fx / (1.0 - e);
    x -= dx;
    if (fabs(dx) <= 1.2e-16 * x) break;
  }
  return x;
}

// ---------------------------------------------------------------------------
// Per-walker setup: modified_blackbody.__init__ (modified_blackbody.py:200-337)
// ---------------------------------------------------------------------------
template <bool THIN, bool ALPHA>
MBB_HD_NOINLINE void sed_setup(Sed& s, double T, double beta, double lambda0, double alpha,
                               double fnorm, double wavenorm) {
  s.T = T; s.beta = beta; s.lambda0 = lambda0; s.alpha = alpha; s.fnorm = fnorm;
  s.status = ST_OK;
  s.x0 = 0.0; s.xmerge = kInf; s.kappa = 0.0;
  s.amp_grey = s.amp_pow = s.q_hi = s.q_lo = 0.0;
  if (ALPHA && !(alpha > 0.0)) { s.status = ST_BAD_ALPHA; }
  if (!(beta >= 0.0)) { s.status = ST_BAD_BETA; }
  if (!finite_d(T) || !finite_d(fnorm) || (!THIN && !finite_d(lambda0))) s.status = ST_NONFINITE;
  const double kT = kK * T;
  s.hokt9 = 1e9 * kH / kT;
  s.hokt_e9 = kH / kT * 1e9;
  s.hcokt = kH * kCum / kT;
  s.xnorm = s.hcokt / wavenorm;
  if (!THIN) s.x0 = s.hcokt / lambda0;
  if (s.status != ST_OK) { s.normfac = 0.0; return; }
  const double xn = s.xnorm;

  if (THIN) {
    if (!ALPHA) {
      s.normfac = fnorm * expm1(xn) / ppow(xn, 3.0 + beta);             // :240-241
    } else {
      const double a = 3.0 + alpha + beta;                              // :253
      s.xmerge = thin_merge_root(a);                                    // :254
      s.kappa = ppow(s.xmerge, 3.0 + alpha + beta) / expm1(s.xmerge);   // :259-261
      if (xn > s.xmerge) s.normfac = fnorm * ppow(xn, alpha) / s.kappa; // :264-266
      else s.normfac = fnorm * expm1(xn) / ppow(xn, 3.0 + beta);        // :268-269
    }
  } else {
    if (!ALPHA) {
      s.normfac = -fnorm * expm1(xn) /
                  (expm1(-ppow(xn / s.x0, beta)) * ppow(xn, 3.0));      // :274-276
    } else {
      // bracket (:286-317): a=0.1 halving while f(a)>=0, b=15 doubling while f(b)<=0
      const double x0 = s.x0;
      double a = 0.1, av = merge_residual(a, alpha, beta, x0);
      int it = 0;
      while (av >= 0.0) {
        a /= 2.0;
        av = merge_residual(a, alpha, beta, x0);
        if (it > 100) { s.status = ST_BRACKET_LOW; break; }
        ++it;
      }
      double b = 15.0, bv = merge_residual(b, alpha, beta, x0);
      it = 0;
      while (s.status == ST_OK && bv <= 0.0) {
        b *= 2.0;
        bv = merge_residual(b, alpha, beta, x0);
        if (it > 100) { s.status = ST_BRACKET_HIGH; break; }
        ++it;
      }
      if (!(av < 0.0) || !(bv > 0.0)) {                                 // NaN residuals
        if (s.status == ST_OK) s.status = ST_NONFINITE;
      }
      if (s.status != ST_OK) { s.normfac = 0.0; return; }
      int st = ST_OK;
      s.xmerge = brent_root([=](double x) { return merge_residual(x, alpha, beta, x0); },
                            a, b, av, bv, st);                          // :321
      if (st != ST_OK) { s.status = st; s.normfac = 0.0; return; }
      const double xm = s.xmerge;
      s.kappa = -ppow(xm, 3.0 + alpha) * expm1(-ppow(xm / x0, beta)) / expm1(xm);  // :326-328
      if (xn > xm) {
        s.normfac = fnorm * ppow(xn, alpha) / s.kappa;                  // :332-333
      } else {
        double ef = expm1(-ppow(xn / x0, beta));                        // :335
        s.normfac = -fnorm * expm1(xn) / (ppow(xn, 3.0) * ef);          // :336-337
      }
    }
  }
  if (ALPHA && !finite_d(s.kappa)) s.status = ST_OVERFLOW;
  if (!finite_d(s.normfac)) s.status = ST_OVERFLOW;
}

// FAST-mode amplitudes, derived from the same per-walker constants:
//   grey side : f = amp_grey * G_i            G_i see node_fnu_fast
//   power side: f = amp_pow  * exp(alpha * L_i)
// where L_i = log(wave_i / wavenorm)  (so cx_i / xnorm = exp(-L_i)).
template <bool THIN, bool ALPHA>
MBB_HD void sed_setup_fast(Sed& s, double wavenorm) {
  if (s.status != ST_OK) return;
  const double xn = s.xnorm;
  double tn_fac = 1.0;   // -expm1(-t_n), t_n = (xnorm/x0)^beta  (thick only)
  if (!THIN) {
    // q = log(xnorm / x0) = log(lambda0 / wavenorm), carried as hi + lo
    double r = s.lambda0 / wavenorm;
    double q = log(r);
    // one Newton correction: log(r) = q + (r*exp(-q) - 1) + O(eps^2)
    s.q_hi = q;
    s.q_lo = fma(r, exp(-q), -1.0);
    tn_fac = -expm1(-exp_prod(s.beta, s.q_hi, s.q_lo));
  }
  // amplitude of the grey-body side if the normalisation wavelength is on it
  const double grey_at_norm = THIN ? s.fnorm * expm1(xn) : s.fnorm * expm1(xn) / tn_fac;
  if (!ALPHA) { s.amp_grey = grey_at_norm; return; }
  // R = grey(xmerge) * xmerge^alpha / (grey(xnorm) * xnorm^alpha) * [expm1(xnorm) / tn_fac]^-1 ...
  // written as ratios so nothing over/underflows:
  const double xm = s.xmerge;
  const double lmn = log(xm / xn);
  double R;
  if (THIN) {
    R = exp((3.0 + s.beta + s.alpha) * lmn) / expm1(xm);
  } else {
    double tm = exp(s.beta * (lmn + s.q_hi));
    R = -expm1(-tm) * exp((3.0 + s.alpha) * lmn) / expm1(xm);
  }
  // R = kappa * xnorm^-(3+b+alpha) [thin] or kappa * xnorm^-(3+alpha) [thick]
  if (xn > xm) {
    s.amp_pow = s.fnorm;
    s.amp_grey = s.fnorm / R;
  } else {
    s.amp_grey = grey_at_norm;
    s.amp_pow = grey_at_norm * R;
  }
}

// ---------------------------------------------------------------------------
// One quadrature node, FAITHFUL arithmetic.  `cx` is hokt9*freq (array path)
// or hokt_e9*freq (scalar path); expression order as fnu.pyx:23-25, 46-53,
// 73-76, 100-108 (identical in the numpy twin, modified_blackbody.py:468-490).
// ---------------------------------------------------------------------------
template <bool THIN, bool ALPHA>
MBB_HD double node_fnu(const Sed& s, double cx) {
  if (THIN) {
    if (!ALPHA) return s.normfac * ppow(cx, s.beta + 3.0) / expm1(cx);
    double v = (cx > s.xmerge) ? s.kappa * ppow(cx, -s.alpha)
                               : ppow(cx, s.beta + 3.0) / expm1(cx);
    return s.normfac * v;
  }
  if (!ALPHA) {
    double x0b = ppow(cx / s.x0, s.beta);
    return -s.normfac * expm1(-x0b) * ppow(cx, 3.0) / expm1(cx);
  }
  double v;
  if (cx > s.xmerge) {
    v = s.kappa * ppow(cx, -s.alpha);
  } else {
    double x0b = ppow(cx / s.x0, s.beta);
    v = -expm1(-x0b) * ppow(cx, 3.0) / expm1(cx);
  }
  return s.normfac * v;
}

// One node, FAST arithmetic.  l_hi/l_lo = log(wave_i/wavenorm) (double-double),
// rcube = (wavenorm/wave_i)^3, cx as above.
template <bool THIN, bool ALPHA>
MBB_HD double node_fnu_fast(const Sed& s, double cx, double l_hi, double l_lo, double rcube) {
  if (ALPHA && cx > s.xmerge) return s.amp_pow * exp_prod(s.alpha, l_hi, l_lo);
  const double em = expm1(cx);
  if (THIN) return s.amp_grey * exp_prod(-(s.beta + 3.0), l_hi, l_lo) / em;
  // t = (cx/x0)^beta = exp(beta * (q - L_i))
  double d_hi = s.q_hi - l_hi;
  double d_lo = s.q_lo - l_lo;
  double t = exp_prod(s.beta, d_hi, d_lo);
  return s.amp_grey * (-expm1(-t)) * rcube / em;
}

// ---------------------------------------------------------------------------
// Peak wavelength: modified_blackbody._snudev / max_wave (:556-637)
// ---------------------------------------------------------------------------
template <bool THIN>
MBB_HD double snu_deriv(double x, double beta, double x0) {
  const double ef = expm1(x);
  if (THIN) {
    return ppow(x, 2.0 + beta) * (3.0 + beta) / ef - exp(x) * ppow(x, 3.0 + beta) / (ef * ef);
  }
  const double xx0b = ppow(x / x0, beta);
  if (!finite_d(xx0b)) {                                               // :577-579
    return 3.0 * (x * x) / ef - exp(x) * (x * x * x) / (ef * ef);
  }
  const double eb = -expm1(-xx0b);
  const double x2 = x * x, x3 = x2 * x;
  return 3.0 * x2 * eb / ef - exp(x) * x3 * eb / (ef * ef) +
         beta * x3 * exp(-xx0b) * xx0b / (x * ef);
}

template <bool THIN>
MBB_HD_NOINLINE double max_wave(double T, double beta, double x0, int& status) {
  const double xbb = 2.82144;
  if (THIN && beta == 0.0) return kCum / (xbb * kK * T / kH);          // :602-604
  double a = xbb / 2.0, av = snu_deriv<THIN>(a, beta, x0);
  int it = 0;
  while (av <= 0.0) {
    if (it > 20) { status = ST_PEAK_BRACKET; return 0.0; }
    a /= 2.0;
    av = snu_deriv<THIN>(a, beta, x0);
    ++it;
  }
  double b = xbb * 2.0, bv = snu_deriv<THIN>(b, beta, x0);
  it = 0;
  while (bv >= 0.0) {
    if (it > 20) { status = ST_PEAK_BRACKET; return 0.0; }
    b *= 2.0;
    bv = snu_deriv<THIN>(b, beta, x0);
    ++it;
  }
  if (!(av > 0.0) || !(bv < 0.0)) { status = ST_NONFINITE; return 0.0; }
  int st = ST_OK;
  double xmax = brent_root([=](double x) { return snu_deriv<THIN>(x, beta, x0); }, a, b, av, bv, st);
  if (st != ST_OK) { status = st; return 0.0; }
  return kCum / (xmax * kK * T / kH);                                   // :636-637
}

// ---------------------------------------------------------------------------
// Limits and priors: likelihood._check_lowlim / _uplim_prior / _gprior
// (likelihood.py:643-752).  Index 5 is the ghost parameter lambda_peak.
// ---------------------------------------------------------------------------
struct Priors {
  double lowlim[5];
  double uplim[6];
  double gmean[6];
  double givar[6];
  unsigned char has_uplim[6];
  unsigned char has_gprior[6];
  int any_gprior;
};

MBB_HD bool below_lowlim(const Priors& pr, const double p[5]) {
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 5; ++i) bad = bad || (p[i] < pr.lowlim[i]);
  return bad;
}

// Returns the soft-upper-limit penalty in `pen` and the Gaussian-prior term in
// `gp`; the caller adds them in the reference's order (likelihood.py:828-832).
template <bool THIN>
MBB_HD void prior_terms(const Priors& pr, const double p[5], const Sed& s, double& pen,
                        double& gp, int& status) {
  pen = 0.0;
  gp = 0.0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if (pr.has_uplim[i]) {
      double lim = pr.uplim[i];
      if (p[i] > lim) {
        double w = 0.02 * (lim - pr.lowlim[i]);
        double d = p[i] - lim;
        pen -= 0.5 * (d * d) / (w * w);
      }
    }
  }
  double peak = 0.0;
  const bool need_peak = pr.has_uplim[5] || (pr.any_gprior && pr.has_gprior[5]);
  if (need_peak) peak = max_wave<THIN>(s.T, s.beta, s.x0, status);
  if (pr.has_uplim[5]) {
    double lim = pr.uplim[5];
    if (peak > lim) {
      double w = 0.02 * lim;
      double d = peak - lim;
      pen -= 0.5 * (d * d) / (w * w);
    }
  }
  if (pr.any_gprior) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      if (pr.has_gprior[i]) {
        double d = p[i] - pr.gmean[i];
        gp -= 0.5 * pr.givar[i] * (d * d);
      }
    }
    if (pr.has_gprior[5]) {
      double d = peak - pr.gmean[5];
      gp -= 0.5 * pr.givar[5] * (d * d);
    }
  }
}

// ---------------------------------------------------------------------------
// One complete evaluation of likelihood.__call__ (likelihood.py:790-834) by a
// single thread: limits gate -> per-walker setup -> band fluxes (node loop) ->
// chi-square (diagonal or full inverse covariance) -> soft upper limits ->
// Gaussian priors.  `Tab` is any type with freq/w/lhi/llo/rcube/band_off/
// scalar_path members indexable by node / band (SmallTab in the kernel
// parameter block, TabView over plain arrays).
// ---------------------------------------------------------------------------
struct TabView {
  const double* freq;
  const double* w;
  const double* lhi;
  const double* llo;
  const double* rcube;
  const int* band_off;
  const unsigned char* scalar_path;
  int nb;
};

constexpr int kMaxBandsPerThread = 64;

template <bool THIN, bool ALPHA, bool FAST, class Tab>
MBB_HD double loglike_one(const double p[5], double wavenorm, const Priors& pr, const Tab& t,
                          const double* flux, const double* ivar, const double* cinv, int& st) {
  st = ST_OK;
  if (below_lowlim(pr, p)) {
    st = ST_BELOW_LOWLIM;
    return -kInf;
  }
  Sed s;
  sed_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], wavenorm);
  if (FAST) sed_setup_fast<THIN, ALPHA>(s, wavenorm);
  st = s.status;
  const double nan = kInf - kInf;
  if (st != ST_OK) return nan;
  const int nb = t.nb;
  double chi = 0.0;
  double diff[kMaxBandsPerThread];
  for (int b = 0; b < nb; ++b) {
    const double hk = t.scalar_path[b] ? s.hokt_e9 : s.hokt9;
    double acc = 0.0;
    for (int i = t.band_off[b]; i < t.band_off[b + 1]; ++i) {
      const double cx = hk * t.freq[i];
      double f;
      if (FAST) f = node_fnu_fast<THIN, ALPHA>(s, cx, t.lhi[i], t.llo[i], t.rcube[i]);
      else f = node_fnu<THIN, ALPHA>(s, cx);
      acc = fma(f, t.w[i], acc);
    }
    const double df = flux[b] - acc;
    if (cinv) diff[b] = df;
    else chi = fma(df * df, ivar[b], chi);
  }
  if (cinv) {
    for (int r = 0; r < nb; ++r) {
      double row = 0.0;
      for (int c = 0; c < nb; ++c) row = fma(cinv[r * nb + c], diff[c], row);
      chi = fma(diff[r], row, chi);
    }
  }
  double pen, gp;
  prior_terms<THIN>(pr, p, s, pen, gp, st);
  double lnl = -0.5 * chi;
  lnl += pen;
  if (pr.any_gprior) lnl += gp;
  if (st != ST_OK) return nan;
  if (lnl != lnl) st = ST_NONFINITE;
  return lnl;
}

}  // namespace mbb
