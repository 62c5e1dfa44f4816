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
#include "mbb_model_defs.cuh"
#include "mbb_fastmath.cuh"

namespace mbb {

struct Sed {
  double T, beta, lambda0, alpha, fnorm;
  double hokt9;    // 1e9*h/(k*T): array (Cython) path, fnu.pyx:16
  double hokt_e9;  // (h/(k*T))*1e9: scalar (numpy) path, modified_blackbody.py:461-464
  double hcokt, xnorm, x0, normfac, xmerge, kappa;
  int status;
};

// FAST-mode per-walker state: lives in registers, never in local memory.
//   grey side : f = amp_grey * G_i            (see node_fnu_fast)
//   power side: f = amp_pow  * exp(alpha * L_i),  L_i = log(wave_i / wavenorm)
struct FastSed {
  double T, beta, alpha;
  double hokt9;                 // 1e9*h/(k*T)
  double x0, xmerge;
  double amp_grey, amp_pow, q_hi, q_lo;
  int status;
};

// ---------------------------------------------------------------------------
// small math helpers
// ---------------------------------------------------------------------------
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

// ---------------------------------------------------------------------------
// FAST-mode per-walker setup.  Same quantities as sed_setup, but
//   * only what the FAST node formulas need (no normfac / kappa: the
//     normalisation is carried as ratios, so nothing over/underflows);
//   * one IEEE division (hokt9, formed exactly as fnu.pyx:16 does) -- every
//     other quotient goes through div_fast, every exp/expm1 through
//     mbb_fastmath.cuh;
//   * xnorm = hokt9 * (c/wavenorm) instead of (h c / k T) / wavenorm: the same
//     number to 1-2 ulp.
// `nu_norm` = 299792.458 / wavenorm [GHz], precomputed on the host.
// ---------------------------------------------------------------------------
MBB_HD double merge_residual_fast(double x, double alpha, double beta, double inv_x0) {
  const double t = exp_tau(beta * log(x * inv_x0));
  // t/expm1(t): -> 1 as t -> 0, -> 0 as t -> inf (saturating expm1_fast)
  const double bterm = t < 1e-280 ? 1.0 : t * rcp_fast(expm1_l(t));
  return x + expm1_l(-x) * (3.0 + alpha + beta * bterm);
}

MBB_HD double thin_merge_root_fast(double a) {
  double x = a;
  for (int it = 0; it < 12; ++it) {
    const double e = a * exp_l(-x);
    const double dx = div_fast((x - a) + e, 1.0 - e);
    x -= dx;
    if (fabs(dx) <= 1.2e-16 * x) break;
  }
  return x;
}

template <bool THIN, bool ALPHA>
MBB_HD void fast_setup(FastSed& f, double T, double beta, double lambda0, double alpha, double fnorm,
                       double wavenorm, double nu_norm) {
  f.T = T; f.beta = beta; f.alpha = alpha;
  f.status = ST_OK;
  f.x0 = 0.0; f.xmerge = kInf; f.amp_grey = f.amp_pow = f.q_hi = f.q_lo = 0.0;
  if (ALPHA && !(alpha > 0.0)) f.status = ST_BAD_ALPHA;
  if (!(beta >= 0.0)) f.status = ST_BAD_BETA;
  if (!finite_d(T) || !finite_d(fnorm) || (!THIN && !finite_d(lambda0)) || (ALPHA && !finite_d(alpha)) ||
      !finite_d(beta))
    f.status = ST_NONFINITE;
  // range of validity of the lean exp family (mbb_fastmath.cuh)
  if (f.status == ST_OK && (!(T >= kFastMinT) || beta > kFastMaxIndex || (ALPHA && alpha > kFastMaxIndex)))
    f.status = ST_OVERFLOW;
  f.hokt9 = 1e9 * kH / (kK * T);
  if (f.status != ST_OK) return;
  const double xn = f.hokt9 * nu_norm;
  double inv_x0 = 0.0, tn_fac = 1.0;
  if (!THIN) {
    // q = log(xnorm/x0) = log(lambda0/wavenorm)
    const double r = lambda0 / wavenorm;
    f.x0 = div_fast(xn, r);
    inv_x0 = rcp_fast(f.x0);
    f.q_hi = log(r);
    f.q_lo = 0.0;
    tn_fac = -expm1_l(-exp_prod_tau(beta, f.q_hi, f.q_lo));     // 1 - exp(-(xn/x0)^beta)
  }
  const double em_n = expm1_l(xn);
  const double grey_at_norm = THIN ? fnorm * em_n : div_fast(fnorm * em_n, tn_fac);
  if (!ALPHA) {
    f.amp_grey = grey_at_norm;
    if (!finite_d(f.amp_grey)) f.status = ST_OVERFLOW;
    return;
  }
  if (THIN) {
    f.xmerge = thin_merge_root_fast(3.0 + alpha + beta);               // modified_blackbody.py:253-254
  } else {
    // bracket + Brent exactly as sed_setup (modified_blackbody.py:286-321)
    double a = 0.1, av = merge_residual_fast(a, alpha, beta, inv_x0);
    int it = 0;
    while (av >= 0.0) {
      a /= 2.0;
      av = merge_residual_fast(a, alpha, beta, inv_x0);
      if (it > 100) { f.status = ST_BRACKET_LOW; break; }
      ++it;
    }
    double b = 15.0, bv = merge_residual_fast(b, alpha, beta, inv_x0);
    it = 0;
    while (f.status == ST_OK && bv <= 0.0) {
      b *= 2.0;
      bv = merge_residual_fast(b, alpha, beta, inv_x0);
      if (it > 100) { f.status = ST_BRACKET_HIGH; break; }
      ++it;
    }
    if (f.status == ST_OK && (!(av < 0.0) || !(bv > 0.0))) f.status = ST_NONFINITE;
    if (f.status != ST_OK) return;
    int st = ST_OK;
    f.xmerge = brent_root([=](double x) { return merge_residual_fast(x, alpha, beta, inv_x0); }, a, b,
                          av, bv, st);
    if (st != ST_OK) { f.status = st; return; }
  }
  // R = grey(xmerge) xmerge^alpha / (grey(xnorm) xnorm^alpha), built from ratios
  const double xm = f.xmerge;
  const double lmn = log(div_fast(xm, xn));
  const double inv_em_m = rcp_fast(expm1_l(xm));
  double R;
  if (THIN) {
    R = exp_l((3.0 + beta + alpha) * lmn) * inv_em_m * em_n;        // times grey_at_norm/fnorm
    // (amp_grey = fnorm*em_n when xn <= xm) -> amp_pow = fnorm * R
  } else {
    const double tm = exp_tau(beta * (lmn + f.q_hi));
    R = -expm1_l(-tm) * exp_l((3.0 + alpha) * lmn) * inv_em_m * div_fast(em_n, tn_fac);
  }
  // here R = amp_pow / fnorm when the normalisation wavelength sits on the grey side
  if (xn > xm) {
    f.amp_pow = fnorm;
    f.amp_grey = div_fast(fnorm * grey_at_norm, fnorm * R);
  } else {
    f.amp_grey = grey_at_norm;
    f.amp_pow = fnorm * R;
  }
  if (!finite_d(f.amp_grey) || !finite_d(f.amp_pow)) f.status = ST_OVERFLOW;
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
// rcube = (wavenorm/wave_i)^3, cx = hokt9 * freq_i.
template <bool THIN, bool ALPHA>
MBB_HD double node_fnu_fast(const FastSed& s, double cx, double l_hi, double l_lo, double rcube) {
  if (ALPHA && cx > s.xmerge) return s.amp_pow * exp_prod_l(s.alpha, l_hi, l_lo);
  const double em = expm1_l(cx);
  if (THIN) return div_fast(s.amp_grey * exp_prod_l(-(s.beta + 3.0), l_hi, l_lo), em);
  // t = (cx/x0)^beta = exp(beta * (q - L_i))
  const double t = exp_prod_tau(s.beta, s.q_hi - l_hi, s.q_lo - l_lo);
  return div_fast(s.amp_grey * (-expm1_l(-t)) * rcube, em);
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
// Frequency integral of f_nu: modified_blackbody.freq_integrate
// (modified_blackbody.py:639-674), the integrand of L_IR (results.py:669-673).
//
// The reference hands f_nu to scipy.integrate.quad (adaptive QUADPACK,
// ~150-360 integrand calls).  Here the integral is done in closed form where
// one exists and by a fixed rule elsewhere:
//   * above the merge point f_nu is a pure power law, kappa x^-alpha:
//     integrated analytically;
//   * the grey-body part is integrated in u = ln x (x = h nu / k T), where the
//     integrand x * grey(x) is analytic and bell-shaped, with 8 equal panels
//     of 16-point Gauss-Legendre over [x1, min(x_end, xcut)]; beyond
//     xcut = 50 + 3 (4 + beta) the integrand is < 1e-17 of its peak.
// Against a 40-digit mpmath integral this is accurate to ~1e-15; the
// reference's own quad() result is only good to ~1e-9 when alpha is on (kink
// at the merge inside the interval), so parity with it is checked at 5e-9 and
// parity with the truth at 1e-13 (tests).
// ---------------------------------------------------------------------------
constexpr int kLirPanels = 8;
constexpr int kLirNodes = 16 * kLirPanels;

MBB_HD void gl16(int j, double& t, double& w) {
  const double T[8] = {0.095012509837637441, 0.28160355077925892, 0.45801677765722737,
                       0.61787624440264377,  0.755404408355003,   0.86563120238783176,
                       0.9445750230732326,   0.98940093499164994};
  const double W[8] = {0.18945061045506864, 0.18260341504492364, 0.16915651939500265,
                       0.14959598881657671, 0.12462897125553407, 0.095158511682492605,
                       0.062253523938647456, 0.027152459411754176};
  const int k = j < 8 ? 7 - j : j - 8;
  t = j < 8 ? -T[k] : T[k];
  w = W[k];
}

// grey-body shape (before normfac) at x, numpy formulation (modified_blackbody.py:468-490)
template <bool THIN>
MBB_HD double grey_shape(double x, double beta, double x0) {
  if (THIN) return ppow(x, 3.0 + beta) / expm1(x);
  return -expm1(-ppow(x / x0, beta)) * (x * x * x) / expm1(x);
}

struct LirSpan {
  double u1, du;     // grey-body part: u in [u1, u1 + kLirPanels*du], 0 panels if du <= 0
  double pow_part;   // analytic integral of kappa x^-alpha over the power-law part
};

template <bool ALPHA>
MBB_HD LirSpan lir_span(double x1, double x2, double beta, double alpha, double xmerge, double kappa) {
  LirSpan sp;
  sp.pow_part = 0.0;
  double xg2 = x2;
  if (ALPHA) {
    if (x2 > xmerge) {
      const double xa = x1 > xmerge ? x1 : xmerge;
      const double om = 1.0 - alpha;
      if (fabs(om) > 1e-9) sp.pow_part = kappa * (ppow(x2, om) - ppow(xa, om)) / om;
      else sp.pow_part = kappa * log(x2 / xa);
      xg2 = xmerge;
    }
  }
  const double xcut = fmax(50.0 + 3.0 * (4.0 + beta), x1 + 40.0);
  const double xe = fmin(xg2, xcut);
  sp.u1 = log(x1);
  sp.du = xe > x1 ? (log(xe) - sp.u1) / kLirPanels : 0.0;
  return sp;
}

// contribution of quadrature node `node` (0..kLirNodes-1) to the u-integral
template <bool THIN>
MBB_HD double lir_node(const LirSpan& sp, int node, double beta, double x0) {
  if (!(sp.du > 0.0)) return 0.0;
  double t, w;
  gl16(node & 15, t, w);
  const double a = sp.u1 + sp.du * (node >> 4);
  const double u = a + 0.5 * sp.du * (t + 1.0);
  const double x = exp(u);
  return 0.5 * sp.du * w * grey_shape<THIN>(x, beta, x0) * x;
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
MBB_HD void prior_terms(const Priors& pr, const double p[5], double T, double beta, double x0,
                        double& pen, double& gp, int& status) {
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
  if (need_peak) peak = max_wave<THIN>(T, beta, x0, status);
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

// chi-square of one evaluation given a callable returning f_nu at node i
template <class Tab, class NodeFn>
MBB_HD double chi_square(const Tab& t, const double* flux, const double* ivar, const double* cinv,
                         NodeFn node) {
  const int nb = t.nb;
  double chi = 0.0;
  if (!cinv) {
    for (int b = 0; b < nb; ++b) {
      double acc = 0.0;
      for (int i = t.band_off[b]; i < t.band_off[b + 1]; ++i) acc = fma(node(b, i), t.w[i], acc);
      const double df = flux[b] - acc;
      chi = fma(df * df, ivar[b], chi);
    }
    return chi;
  }
  double diff[kMaxBandsPerThread];
  for (int b = 0; b < nb; ++b) {
    double acc = 0.0;
    for (int i = t.band_off[b]; i < t.band_off[b + 1]; ++i) acc = fma(node(b, i), t.w[i], acc);
    diff[b] = flux[b] - acc;
  }
  for (int r = 0; r < nb; ++r) {
    double row = 0.0;
    for (int c = 0; c < nb; ++c) row = fma(cinv[r * nb + c], diff[c], row);
    chi = fma(diff[r], row, chi);
  }
  return chi;
}

template <bool THIN, bool ALPHA, bool FAST, class Tab>
MBB_HD double loglike_one(const double p[5], double wavenorm, double nu_norm, const Priors& pr,
                          const Tab& t, const double* flux, const double* ivar, const double* cinv,
                          int& st) {
  st = ST_OK;
  if (below_lowlim(pr, p)) {
    st = ST_BELOW_LOWLIM;
    return -kInf;
  }
  const double nan = kInf - kInf;
  double chi, pen, gp;
  if (FAST) {
    FastSed s;
    fast_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], wavenorm, nu_norm);
    st = s.status;
    if (st != ST_OK) return nan;
    chi = chi_square(t, flux, ivar, cinv, [&](int, int i) {
      return node_fnu_fast<THIN, ALPHA>(s, s.hokt9 * t.freq[i], t.lhi[i], t.llo[i], t.rcube[i]);
    });
    prior_terms<THIN>(pr, p, s.T, s.beta, s.x0, pen, gp, st);
  } else {
    Sed s;
    sed_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], wavenorm);
    st = s.status;
    if (st != ST_OK) return nan;
    chi = chi_square(t, flux, ivar, cinv, [&](int b, int i) {
      return node_fnu<THIN, ALPHA>(s, (t.scalar_path[b] ? s.hokt_e9 : s.hokt9) * t.freq[i]);
    });
    prior_terms<THIN>(pr, p, s.T, s.beta, s.x0, pen, gp, st);
  }
  double lnl = -0.5 * chi;
  lnl += pen;
  if (pr.any_gprior) lnl += gp;
  if (st != ST_OK) return nan;
  if (lnl != lnl) st = ST_NONFINITE;
  return lnl;
}

}  // namespace mbb
