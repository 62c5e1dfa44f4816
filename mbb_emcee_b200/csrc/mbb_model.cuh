// Modified-blackbody SED model: per-walker setup, node evaluation, root solves.
//
// Everything here is `__host__ __device__` so that the exact same source can be
// compiled (a) into the sm_100a kernels of mbb_kernels.cu and (b) into the
// host-side emulation library the CPU test-suite uses to check the *logic*
// (tests/hostemu, never shipped, never a fallback).
//
// Arithmetic modes, selected per context (mbb_set_math_mode):
//   FAITHFUL   - the reference's formulas in the reference's evaluation order
//                (fnu.pyx:9-108, modified_blackbody.py:228-337, 441-491):
//                pow / expm1 / exp per node.
//   FAST       - algebraically identical, restructured so that no pow() is
//                evaluated per node: every power becomes exp(b * L) with
//                L' = log(wave_i / wavenorm) * 64/ln2 precomputed on the host
//                from the 80-bit logarithm, the normalisation is applied as
//                ratios (f/fnorm near 1), and exp/expm1/reciprocal come from
//                mbb_fastmath.cuh.  ~3.5x fewer FP64 instructions per node;
//                agrees with FAITHFUL to ~1e-14 (tests/test_parity_gpu.py).
//   FAST_GAUSS - FAST over the 32-point Gauss rules of the tabulated bands
//                where gauss_band_masks() allows it (mbb_gaussrule.h).
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
// Node tables hold L'_i = log(wave_i/wavenorm)*64/ln2 (rounded to double from the
// 80-bit value: its 2^-53 relative rounding moves exp(b L_i) by |b L_i| 1.1e-16) and the
// effective weight  weff_i = w_i (thin) or w_i (wavenorm/wave_i)^3 (thick);
// a node contributes  v_i * weff_i  with
//   grey side, thin : v = amp_grey exp(-(beta+3) L_i) / expm1(x_i)
//   grey side, thick: v = amp_grey (1 - exp(-t_i)) / expm1(x_i),  t_i = t0 exp(-beta L_i)
//   power side      : v = amp_pow exp(apow L_i),  apow = alpha (thin) or alpha+3 (thick)
// and x_i*64/ln2 = nu_i * (xk_hi + xk_lo)   (see mbb_fastmath.cuh).
struct FastSed {
  double T, beta, alpha;
  double hokt9;                 // 1e9*h/(k*T)
  double x0, xmerge;
  double amp_grey, amp_pow;
  double xk_hi, xk_lo;          // hokt9 * 64/ln2
  double nb;                    // -(beta+3) thin, -beta thick
  double apow;                  // alpha thin, alpha+3 thick
  double t0;                    // thick: (lambda0/wavenorm)^beta
  double t0c;                   // t0 * 64/ln2
  double uq_hi, uq_lo;          // thick: beta*log(lambda0/wavenorm)*64/ln2 (CLAMP path)
  double nu_merge;              // xmerge / hokt9 [GHz]; +inf without alpha
  int status;
  int safe;                     // every exponent of the node loop provably inside the double range
};

// Model constants shared by every evaluation of a launch.
struct ModelP {
  double wavenorm;
  double nu_norm;     // 299792.458 / wavenorm [GHz]
  double nu_max;      // largest node frequency of the band table [GHz]
  double lmax;        // largest |log(wave_i/wavenorm)| of the band table
};

// Host-side construction of one FAST node record (mbb_set_bands and the test
// emulation use the same code): frequency exactly as the reference forms it
// (modified_blackbody.py:551-554), L' from the x87 80-bit logl (64-bit mantissa).
struct FastNode {
  double freq, weff, lp, labs;
};
inline FastNode fast_node(double wave_um, double weight, double wavenorm, bool thin) {
  FastNode n;
  n.freq = kUmToGHz / wave_um;
  const long double l = logl((long double)wave_um) - logl((long double)wavenorm);
  const long double lp = l * (64.0L / logl(2.0L));
  n.lp = (double)lp;
  n.labs = fabs((double)l);
  const long double r = (long double)wavenorm / (long double)wave_um;
  n.weff = thin ? weight : (double)((long double)weight * (r * r * r));
  return n;
}

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

// The per-walker constants that are "in 1/64-octave units", converted for the node
// loops that use the 256-entry table (TS & kTab256): times 4, exact.
MBB_HD void fast_sed_rescale256(FastSed& f) {
  f.xk_hi *= 4.0;
  f.xk_lo *= 4.0;
  f.t0c *= 4.0;
  f.uq_hi *= 4.0;
  f.uq_lo *= 4.0;
}

// ---------------------------------------------------------------------------
// FAST-mode per-walker setup.  Same quantities as sed_setup, but
//   * only what the FAST node formulas need (no normfac / kappa: the
//     normalisation is carried as ratios, so nothing over/underflows);
//   * every quotient goes through div_fast (<= 1 ulp), every exp/expm1 through
//     mbb_fastmath.cuh;
//   * xnorm = hokt9 * (c/wavenorm) instead of (h c / k T) / wavenorm: the same
//     number to 1-2 ulp.
// Also decides `safe`: with m.nu_max / m.lmax bounding the node table, every
// exponent the node loop forms stays below kSafeExp in magnitude, so the loop
// may run the CLAMP=false instantiations.
// ---------------------------------------------------------------------------
template <int TS>
MBB_HD double thin_merge_root_fast(double a, const double* tab) {
  double x = a;                       // a <= 603 (gates of fast_setup): the exponent needs no clamp
  for (int it = 0; it < 12; ++it) {
    const double e = a * exp_red<TS, false>(red_x(-x), tab);
    const double dx = div_fast((x - a) + e, 1.0 - e);
    x -= dx;
    if (fabs(dx) <= 1.2e-16 * x) break;
  }
  return x;
}

// The thick merge equation in u = log x:  G(u) = g(e^u),  g(x) = x - (1 - e^-x)(3 + alpha +
// beta B(t)),  B(t) = t/(e^t - 1),  t = (x/x0)^beta = exp(beta (u - u0)).  In this variable a
// residual evaluation is four lean exponentials and no logarithm.  Every exponent is bounded
// (u <= log(603) by the gates of fast_setup; beta (u - u0) is clamped to [-700, log 700], where
// B(t) is 1 resp. 0 to far below rounding), so the exponentials run unclamped on whatever table
// the caller works with (TS: the replicated shared-memory copy in the delta kernels).
struct MergeEval {
  double x, t, e, E, em1t;      // e^u, t, e^-x, 1 - e^-x, e^t - 1 (em1t only where t >= 1e-3)
  double S, dB, rem;            // 3 + alpha + beta B(t), B'(t), 1/(e^t - 1)
  double G, dG;                 // residual and dG/du = x g'(x)
};
constexpr double kMergeTauMax = 6.5510803350434044;       // log(700)

template <int TS>
MBB_HD void merge_eval(double u, double a_lo, double beta, double u0, const double* tab, MergeEval& v) {
#if defined(MBB_COUNT_EVALS)
  ++g_f64_evals;
#endif
  v.x = exp_red<TS, false>(red_x(u), tab);
  double arg = beta * (u - u0);
  arg = arg < -700.0 ? -700.0 : arg;
  arg = arg > kMergeTauMax ? kMergeTauMax : arg;
  v.t = exp_red<TS, false>(red_x(arg), tab);
  double B, dB;                       // B(t), B'(t)
  if (v.t < 1e-3) {
    B = 1.0 - v.t * (0.5 - v.t * (1.0 / 12.0));
    dB = -0.5 + v.t * (1.0 / 6.0);
    v.em1t = v.t;
    v.rem = 0.0;
  } else {
    v.em1t = expm1_red<TS, false>(red_x(v.t), tab);
    const double r = rcp_cubic(v.em1t);
    B = v.t * r;
    dB = (v.em1t - v.t * (v.em1t + 1.0)) * r * r;
    v.rem = r;
  }
  v.dB = dB;
  v.e = exp_red<TS, false>(red_x(-v.x), tab);
  v.E = 1.0 - v.e;                    // x >= ~2.6: no cancellation
  const double S = a_lo + beta * B;
  v.S = S;
  v.G = v.x - v.E * S;
  v.dG = v.x * (1.0 - v.e * S) - v.E * (beta * beta) * dB * v.t;
}

// d2G/du2 at the point of `v` (for the Halley step): with B'' = e^t (t (e^t + 1) - 2 (e^t - 1)) /
// (e^t - 1)^3, factored so that nothing overflows for t up to 700.  The cancellation in B'' for
// small t costs ~1e-9 relative at t = 1e-3 -- immaterial, G'' only corrects a correction.
MBB_HD double merge_d2G(const MergeEval& v, double beta) {
  const double b2 = beta * beta;
  double Bpp;
  if (v.t < 1e-3) {
    Bpp = 1.0 / 6.0;
  } else {
    const double et = v.em1t + 1.0;
    Bpp = (et * v.rem) * ((v.t * (et + 1.0) - 2.0 * v.em1t) * v.rem) * v.rem;
  }
  const double xe = v.x * v.e;
  return v.x - xe * v.S * (1.0 - v.x) - 2.0 * xe * b2 * v.t * v.dB - v.E * b2 * beta * v.t * (v.dB + v.t * Bpp);
}

// Single-precision approximation of the root (u = log x) to ~1e-5: the same safeguarded Newton
// iteration on MUFU exponentials (an evaluation costs ~45 issue slots against ~300 for the
// double-precision one).  It replaces the first 3-4 double-precision evaluations; what it
// returns is only a starting point, so its accuracy does not enter the result.
MBB_HD float f32_exp(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.44269504f));
  return r;
#elif defined(MBB_F32_NOISE)
  static unsigned lcg = 12345u;
  lcg = lcg * 1664525u + 1013904223u;
  return expf(x) * (1.0f + ((lcg >> 16) & 1 ? 4e-7f : -4e-7f));
#else
  return expf(x);
#endif
}
MBB_HD float f32_log(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * 0.693147181f;
#else
  return logf(x);
#endif
}
MBB_HD float f32_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}
// Newton steps kept inside the bracket (a step that leaves it is replaced by the midpoint); the
// "fails to halve" test of the double-precision iteration is left to that iteration.
MBB_HD float thick_merge_seed(float a_lo, float beta, float u0, float ul, float uh) {
  float u = 0.5f * (ul + uh);
  const float bb = beta * beta;
  for (int it = 0; it < 8; ++it) {
#if defined(MBB_COUNT_EVALS)
    ++g_f32_iters;
#endif
    const float x = f32_exp(u);
    const float t = f32_exp(fminf(fmaxf(beta * (u - u0), -80.0f), 4.38f));     // t <= 80: e^t finite
    float B, dB;
    if (t < 0.03f) {
      B = 1.0f - t * (0.5f - t * (1.0f / 12.0f));
      dB = -0.5f + t * (1.0f / 6.0f);
    } else {
      const float em = f32_exp(t) - 1.0f;
      const float r = f32_rcp(em);
      B = t * r;
      dB = (em - t * (em + 1.0f)) * r * r;
    }
    const float e = f32_exp(-x);
    const float E = 1.0f - e;
    const float S = a_lo + beta * B;
    const float G = x - E * S;
    const float dG = x * (1.0f - e * S) - E * bb * dB * t;
    // (no bracket update inside the rounding noise of G, where its sign means nothing)
    if (G < -1e-3f) ul = u; else if (G > 1e-3f) uh = u;
    float un = u - G * f32_rcp(dG);
    if (!(un > ul && un < uh)) un = 0.5f * (ul + uh);
    const float dx = un - u;
    u = un;
    if (fabsf(dx) <= 3e-6f) break;
  }
  return u;
}

// Root of the thick merge equation.
//   bracket: g <= x - (3+alpha)(1 - e^-x) =: gL, whose root a + W(-a e^-a) >= a (1 - e^(1-a))
//   (W concave on [-1/e, 0]), so g(xl) <= 0 at xl = a (1 - e^(1-a)), a = 3 + alpha; and
//   g >= x - (3+alpha+beta)(1 - e^-x), which is positive at x = 3 + alpha + beta =: xh.  g < 0
//   on (0, root) and > 0 beyond, so the ends may be moved outwards freely: they are formed in
//   single precision (MUFU.LG2) and widened by 1e-3.
//   iteration: the single-precision seed, then Newton in u, bisecting (in u) whenever a step
//   would leave the bracket or fails to halve the previous one -- for steep t(x) (large beta) g is
//   nearly a step and plain Newton cycles between its flat sides.  A Newton step |du| <= 3e-9
//   ends it: the error left is ~ C du^2 <= 1e-17 C (host experiment over 2e5 walkers, cfg2 and
//   wide clouds: root unchanged to 1e-15), so from a 1e-5 seed two evaluations suffice.
//   From the seed (error <= ~3e-6) ONE evaluation is enough: its Newton step du = G/G' already
//   leaves ~0.3 du^2 ~ 1e-13, and the Halley step du / (1 - du G''/(2 G')) (G'' in closed form,
//   merge_d2G) is third order: root error <= 9e-16 over 2e5 walkers drawn from T 3-80 K, beta
//   0.1-9, lambda0 10-1500 um, alpha 0.5-10, and over the cfg2 cloud (host experiment).  Taken
//   when |du| <= 3e-6; otherwise the safeguarded iteration runs as described.
// Returns in `v` the LAST evaluation (at u_eval = u_root + du), x_root by a second-order update
// of v.x, and `dh`: the caller matches the power law to the grey body at u_eval, where every
// factor is already known, and multiplies the amplitude by 1 + dh.  With h(u) = log(grey(x)
// x^alpha): h' = -G/E (the merge equation IS h' = 0) and h'' = -G'/E + O(G), so
// h(u_root) - h(u_eval) = du (G - G' du / 2) / E + O(du^3).
template <int TS>
MBB_HD double thick_merge_root_fast(double alpha, double beta, double u0, const double* tab, int& status,
                                    double& u_eval, MergeEval& v, double& dh) {
  const double a_lo = 3.0 + alpha;
  const float af = (float)a_lo, bf = (float)beta;
  const float ulf = f32_log(af * (1.0f - f32_exp(1.0f - af))) - 1e-3f;
  const float uhf = f32_log(af + bf) + 1e-3f;
  double ul = (double)ulf, uh = (double)uhf;
  double u = (double)thick_merge_seed(af, bf, (float)u0, ulf, uhf);
  if (!(u > ul && u < uh)) u = 0.5 * (ul + uh);
  merge_eval<TS>(u, a_lo, beta, u0, tab, v);
  u_eval = u;
  const double rdG = rcp_fast(v.dG);
  double du = v.G * rdG;
  if (fabs(du) <= 3.0e-6) {
    du *= fma(0.5 * du, merge_d2G(v, beta) * rdG, 1.0);       // Halley, to first order in du G''/G'
    u -= du;
  } else {
    double dxold = uh - ul, dx = dxold;
    for (int it = 0; it < 100; ++it) {
      if (!(v.G == v.G)) { status = ST_NONFINITE; break; }
      if (v.G == 0.0) break;
      if (v.G < 0.0) ul = u; else uh = u;
      if (((u - uh) * v.dG - v.G) * ((u - ul) * v.dG - v.G) > 0.0 || fabs(2.0 * v.G) > fabs(dxold * v.dG)) {
        dxold = dx;
        dx = 0.5 * (uh - ul);
        u = ul + dx;
        if (fabs(dx) <= 3.0e-16) break;       // du = dx/x: relative accuracy of the root
      } else {
        dxold = dx;
        dx = v.G * rcp_fast(v.dG);
        u -= dx;
        if (fabs(dx) <= 3.0e-9) break;
      }
      if (it == 99) { status = ST_NO_CONVERGE; break; }
      merge_eval<TS>(u, a_lo, beta, u0, tab, v);
      u_eval = u;
    }
    du = u_eval - u;
  }
  dh = du * fma(-0.5 * du, v.dG, v.G) * rcp_fast(v.E);
  // x_root = e^(u_eval - du) = x_eval (1 - du + du^2/2)
  return fma(v.x, -du * fma(-0.5, du, 1.0), v.x);
}

// finite and not NaN, by the exponent field (2 integer instructions)
MBB_HD bool finite_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (unsigned)(__double2hiint(x) & 0x7fffffff) < 0x7ff00000u;
#else
  return finite_d(x);
#endif
}

// `tab` / TS: the exp table the caller works with -- the plain one in global memory (default), or
// the calling lane's copy of the replicated shared-memory table (TS = kTabRepShift).
template <bool THIN, bool ALPHA, int TS = 0>
MBB_HD void fast_setup(FastSed& f, double T, double beta, double lambda0, double alpha, double fnorm,
                       const ModelP& m, const double* tab = exp2_tab_default()) {
  f.T = T; f.beta = beta; f.alpha = alpha;
  f.status = ST_OK;
  f.safe = 0;
  f.x0 = 0.0; f.xmerge = kInf; f.nu_merge = kInf;
  f.amp_grey = f.amp_pow = f.t0 = f.t0c = f.uq_hi = f.uq_lo = f.apow = 0.0;
  if (ALPHA && !(alpha > 0.0)) f.status = ST_BAD_ALPHA;
  if (!(beta >= 0.0)) f.status = ST_BAD_BETA;
  if (!finite_bits(T) || !finite_bits(fnorm) || (!THIN && !finite_bits(lambda0)) ||
      (ALPHA && !finite_bits(alpha)) || !finite_bits(beta))
    f.status = ST_NONFINITE;
  // range of validity of the lean exp family (mbb_fastmath.cuh)
  if (f.status == ST_OK && (!(T >= kFastMinT) || beta > kFastMaxIndex || (ALPHA && alpha > kFastMaxIndex)))
    f.status = ST_OVERFLOW;
  f.hokt9 = div_fast(1e9 * kH, kK * T);
  f.xk_hi = f.hokt9 * kC64Hi;
  f.xk_lo = fma(f.hokt9, kC64Lo, fma(f.hokt9, kC64Hi, -f.xk_hi));
  f.nb = THIN ? -(beta + 3.0) : -beta;
  if (f.status != ST_OK) return;
  const double xn = f.hokt9 * m.nu_norm;
  bool safe = f.hokt9 * m.nu_max <= kSafeExp && -f.nb * m.lmax <= kSafeExp;
  double tn_fac = 1.0, q = 0.0;
  if (!THIN) {
    // q = log(xnorm/x0) = log(lambda0/wavenorm)
    const double r = div_fast(lambda0, m.wavenorm);
    f.x0 = div_fast(xn, r);
    q = log_l(r);
    const double qc_hi = q * kC64Hi;
    const double qc_lo = fma(q, kC64Lo, fma(q, kC64Hi, -qc_hi));
    f.uq_hi = beta * qc_hi;
    f.uq_lo = fma(beta, qc_lo, fma(beta, qc_hi, -f.uq_hi));
    f.t0 = exp_red<TS, true>(red_prod(beta, qc_hi, qc_lo), tab);             // (lambda0/wavenorm)^beta
    tn_fac = one_minus_exp_red<TS, false>(red_x(-clamp_pos<kHi700>(f.t0)), tab);   // 1 - exp(-(xn/x0)^beta)
    f.t0c = f.t0 * kC64Hi;
    safe = safe && fabs(beta * q) <= kSafeExp;
  }
  const double em_n = expm1_red<TS, true>(red_prod(m.nu_norm, f.xk_hi, f.xk_lo), tab);
  const double grey_at_norm = THIN ? fnorm * em_n : div_fast(fnorm * em_n, tn_fac);
  if (!ALPHA) {
    f.amp_grey = grey_at_norm;
    f.safe = safe;
    if (!finite_bits(f.amp_grey)) f.status = ST_OVERFLOW;
    return;
  }
  f.apow = THIN ? alpha : alpha + 3.0;
  safe = safe && f.apow * m.lmax <= kSafeExp;
  double R;         // amp_pow / fnorm when the normalisation wavelength sits on the grey side:
                    // grey(xmerge) xmerge^alpha / (grey(xnorm) xnorm^alpha), built from ratios
  if (THIN) {
    f.xmerge = thin_merge_root_fast<TS>(3.0 + alpha + beta, tab);      // modified_blackbody.py:253-254
    const double xm = f.xmerge;
    const double lmn = log_l(div_fast(xm, xn));                        // log(xmerge/xnorm)
    const double inv_em_m = rcp_fast(expm1_red<TS, false>(red_x(xm), tab));
    R = exp_red<TS, true>(red_x((3.0 + beta + alpha) * lmn), tab) * inv_em_m * em_n;
  } else {
    // Root of the merge equation (modified_blackbody.py:122-151, 286-321).  The reference
    // brackets it by halving/doubling from [0.1, 15] and runs brentq (~14 residual
    // evaluations, each a log and three exponentials: three quarters of this setup); here a
    // closed-form bracket, a single-precision seed and a safeguarded Newton iteration in log x
    // (thick_merge_root_fast): typically 2 log-free double-precision residual evaluations.
    const double lnxn = log_l(xn);
    MergeEval v;
    double u_eval, dh;
    f.xmerge = thick_merge_root_fast<TS>(alpha, beta, lnxn - q, tab, f.status, u_eval, v, dh);
    if (f.status != ST_OK) return;
    // the power law is matched to the grey body at the last evaluation point (see above)
    double tau_fac;                                       // 1 - exp(-t)
    if (v.t < 1e-3) tau_fac = v.t * (1.0 - v.t * (0.5 - v.t * (1.0 / 6.0 - v.t * (1.0 / 24.0 - v.t * (1.0 / 120.0)))));
    else tau_fac = v.em1t * rcp_fast(v.em1t + 1.0);
    const double lmn = u_eval - lnxn;                     // log(x_eval/xnorm)
    R = tau_fac * exp_red<TS, true>(red_x((3.0 + alpha) * lmn), tab) * (v.e * rcp_fast(v.E)) * div_fast(em_n, tn_fac);
    R = fma(R, dh, R);
  }
  f.nu_merge = div_fast(f.xmerge, f.hokt9);
  const double xm = f.xmerge;
  // here R = amp_pow / fnorm when the normalisation wavelength sits on the grey side
  if (xn > xm) {
    f.amp_pow = fnorm;
    f.amp_grey = div_fast(fnorm * grey_at_norm, fnorm * R);
  } else {
    f.amp_grey = grey_at_norm;
    f.amp_pow = fnorm * R;
  }
  f.safe = safe;
  if (!finite_bits(f.amp_grey) || !finite_bits(f.amp_pow)) f.status = ST_OVERFLOW;
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

// a > b for non-negative doubles (and +inf) on the integer pipe
MBB_HD bool gt_pos(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(a) > __double_as_longlong(b);
#else
  return a > b;
#endif
}

// One node, FAST arithmetic: acc + f_nu(nu_i) w_i  (see FastSed for the
// formulas).  nu = node frequency [GHz], lp = L'_i, weff = effective weight.
// The reciprocal of expm1(x_i) is folded into the weight so that it runs
// beside the other exp chains; the last dependent step is a single FMA.
// TS: index-stride shift of `tab` (see scaled_T).
// node_pow / node_grey evaluate ONE branch regardless of the merge point (the
// Gauss-rule mode needs each branch's analytic continuation, see
// gauss_band_masks); node_acc picks the branch the reference picks.
template <bool CLAMP, int TS>
MBB_HD double node_pow(const FastSed& s, double lp, double weff, double acc, const double* tab) {
  return fma(exp_red<TS, CLAMP>(red_prod1(s.apow, lp), tab), s.amp_pow * weff, acc);
}

template <bool THIN, bool CLAMP, int TS>
MBB_HD double node_grey(const FastSed& s, double nu, double lp, double weff, double acc, const double* tab) {
  const double em = expm1_red<TS, CLAMP>(red_prod(nu, s.xk_hi, s.xk_lo), tab);
  const double wr = (s.amp_grey * weff) * rcp_cubic(em);
  if (THIN) return fma(exp_red<TS, CLAMP>(red_prod1(s.nb, lp), tab), wr, acc);
  // tc = t_i*64/ln2, t_i = (x_i/x0)^beta = t0 exp(-beta L_i); beyond ~700 only 1 - exp(-t) = 1 matters
  double tc;
  if (CLAMP) tc = exp_red_times<TS, true>(red_sum_prod(s.uq_hi, s.uq_lo, s.nb, lp), tab, kC64Hi * TabCfg<TS>::cscale);
  else tc = exp_red_times<TS, false>(red_prod1(s.nb, lp), tab, s.t0c);
  const double g = one_minus_exp_red<TS, CLAMP>(red_neg_scaled(clamp_pos<hi700c<TS>()>(tc)), tab);
  return fma(g, wr, acc);
}

template <bool THIN, bool ALPHA, bool CLAMP, int TS>
MBB_HD double node_acc(const FastSed& s, double nu, double lp, double weff, double acc, const double* tab) {
  if (ALPHA && gt_pos(nu, s.nu_merge)) return node_pow<CLAMP, TS>(s, lp, weff, acc, tab);
  return node_grey<THIN, CLAMP, TS>(s, nu, lp, weff, acc, tab);
}

// ---------------------------------------------------------------------------
// N grey-side nodes at once, `safe` walkers only (CLAMP=false).  Exactly the
// operations node_acc performs per node -- results are bit-identical -- but
// written breadth-first: step s of all 2N (thin) or 3N (thick) exp chains
// before step s+1 of any.  A warp issues in order and a dependent DFMA waits
// ~8 cycles (tools/fp64_probe.cu), so a thread must carry >= 4 independent
// chains to keep the half-rate FP64 pipe fed without relying on other warps;
// left to itself the compiler emits each chain serially under the register cap.
//   acc[i] += f_nu(nu_i) w_i   for i < N
// ---------------------------------------------------------------------------
template <int N, int TS>
MBB_HD void lean_p_n(const double (&f)[N], double (&p)[N]) {
  double g[N];
#pragma unroll
  for (int i = 0; i < N; ++i) g[i] = lean_g_coef<TS>(TabCfg<TS>::deg);
#pragma unroll
  for (int j = TabCfg<TS>::deg - 1; j >= 0; --j) {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = fma(g[i], f[i], lean_g_coef<TS>(j));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) p[i] = g[i] * f[i];
}

template <bool THIN, int N, int TS>
MBB_HD void grey_nodes_n(const FastSed& s, const double (&nu)[N], const double (&lp)[N],
                         const double (&weff)[N], double (&acc)[N], const double* tab) {
  // chains A: x_i = nu_i * xk, chains B: -b L_i
  double t[2 * N], f[2 * N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    t[i] = fma(nu[i], s.xk_hi, kMagic);
    t[N + i] = fma(s.nb, lp[i], kMagic);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    f[i] = fma(nu[i], s.xk_hi, -(t[i] - kMagic));
    f[N + i] = fma(s.nb, lp[i], -(t[N + i] - kMagic));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = fma(nu[i], s.xk_lo, f[i]);
  double sT[2 * N];
#pragma unroll
  for (int i = 0; i < 2 * N; ++i) sT[i] = scaled_T<TS, false>(tab, lo32_of(t[i]));
  double p[2 * N];
  lean_p_n<2 * N, TS>(f, p);
  double em[N], aw[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    em[i] = fma(sT[i], p[i], sT[i] - 1.0);
    aw[i] = s.amp_grey * weff[i];
  }
  if (THIN) {
    double E[N];
#pragma unroll
    for (int i = 0; i < N; ++i) E[i] = fma(sT[N + i], p[N + i], sT[N + i]);
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = fma(E[i], aw[i] * rcp_cubic(em[i]), acc[i]);
    return;
  }
  // thick: third chain on tc = t0c * exp(-beta L), beside the reciprocals
  double tc[N], f3[N], sT3[N], p3[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double s2 = sT[N + i] * s.t0c;
    tc[i] = clamp_pos<hi700c<TS>()>(fma(s2, p[N + i], s2));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const Red r = red_neg_scaled(tc[i]);
    f3[i] = r.f;
    sT3[i] = scaled_T<TS, false>(tab, r.k);
  }
  lean_p_n<N, TS>(f3, p3);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double g = fma(-sT3[i], p3[i], 1.0 - sT3[i]);
    acc[i] = fma(g, aw[i] * rcp_cubic(em[i]), acc[i]);
  }
}

// ---------------------------------------------------------------------------
// MBB_MATH_FAST_GAUSS support (rules built on the host: mbb_gaussrule.h)
// ---------------------------------------------------------------------------
// Per band: hull of its node frequencies, for the per-walker test whether the
// compressed rule is exact to rounding for this (walker, band) pair.
struct BandMeta {
  double nu_lo, nu_hi;    // [GHz]
  double dl;              // log(nu_hi / nu_lo)
  double has_rule;        // 0: no compressed rule; 1: rule; 2: rule, and the table's frequencies descend
                          // strictly (what the kink split's binary search needs)
};
constexpr double kGaussMaxType = 10.0;   // bound on the integrand's exponential type over the half-band
constexpr double kGaussMaxBetaDl = 4.0;  // thick: bound on beta * log(nu_hi/nu_lo), see gauss_band_masks
constexpr int kGaussPoints = 32;

// Which bands of this walker may use their compressed rule.
//   plain bit b: the band lies entirely on one side of the merge point.  Over the
//     half-band the factors nu^(3+beta) e^-x / (1 - e^-x) of the integrand are of
//     exponential type <= tau = (x-range + (3+beta) dL)/2 (power-law side: alpha dL/2);
//     the 32-point rule's error for such a function is ~ tau^64/64! (5e-26 at
//     tau = 10), and the Planck poles at x = 2 pi i k stay > 1.2 half-bands away.
//     The optically thick factor 1 - exp(-t), t ~ nu^beta, is the delicate one: as a
//     function of u = log nu it grows doubly exponentially beyond |Im u| = pi/(2 beta),
//     so the rule converges like rho^-64 with rho = y + sqrt(y^2+1), y ~ pi/(beta dL);
//     beta dL <= 4 gives rho >= 1.9, rho^-64 < 1e-17.
//   kink bit b: the merge point lies inside the band (f_nu has a kink there, the
//     rule cannot be applied to f_nu itself), but BOTH branches pass their bound over
//     the whole band.  Then  sum_i w_i f(nu_i) = sum_i w_i A(nu_i) + sum_{i in S} w_i
//     (B - A)(nu_i)  with A one branch continued analytically over the band (rule) and
//     S the table nodes on the other branch's side (full table, but only those).
// Nothing is compressed for a walker that is not `safe`.
struct GaussMasks {
  unsigned long long plain, kink;
};

template <bool THIN, bool ALPHA>
MBB_HD GaussMasks gauss_band_masks(const FastSed& s, const BandMeta* bm, int nb) {
  GaussMasks m;
  m.plain = m.kink = 0;
  for (int b = 0; b < nb; ++b) {
    const double lo = bm[b].nu_lo, hi = bm[b].nu_hi, dl = bm[b].dl;
    if (bm[b].has_rule == 0.0) continue;
    const bool grey_ok = 0.5 * (s.hokt9 * (hi - lo) + (3.0 + s.beta) * dl) <= kGaussMaxType &&
                         (THIN || s.beta * dl <= kGaussMaxBetaDl);
    const bool pow_ok = ALPHA && 0.5 * s.alpha * dl <= kGaussMaxType;
    if (ALPHA && lo > s.nu_merge) {
      if (pow_ok) m.plain |= 1ull << b;                                  // all power law
    } else if (ALPHA && hi > s.nu_merge) {
      if (pow_ok && grey_ok && bm[b].has_rule == 2.0) m.kink |= 1ull << b;   // the kink is inside
    } else if (grey_ok) {
      m.plain |= 1ull << b;
    }
  }
  return m;
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
  double uplim[6];       // +inf where has_uplim is 0 (mbb_set_priors enforces it)
  double gmean[6];
  double givar[6];
  unsigned char has_uplim[6];
  unsigned char has_gprior[6];
  int any_gprior;
  int always_terms;      // lambda_peak limit/prior or any Gaussian prior set: prior_terms must run
  // derived by priors_finalize():
  double inv_w2[5];      // 1/(0.02 (uplim - lowlim))^2 of the soft upper limits (0 where there is none)
  int peak_terms;        // a lambda_peak limit or prior is set: the peak solve (max_wave) is needed
};

// Fills the derived members from lowlim / has_uplim / uplim / has_gprior (host side: mbb_set_priors,
// the test emulation).  uplim becomes +inf where no limit is set.
inline void priors_finalize(Priors& p) {
  p.any_gprior = 0;
  for (int i = 0; i < 6; ++i) {
    if (!p.has_uplim[i]) p.uplim[i] = kInf;
    if (p.has_gprior[i]) p.any_gprior = 1;
  }
  for (int i = 0; i < 5; ++i) {
    const double w = 0.02 * (p.uplim[i] - p.lowlim[i]);
    p.inv_w2[i] = p.has_uplim[i] ? 1.0 / (w * w) : 0.0;
  }
  p.peak_terms = (p.has_uplim[5] || p.has_gprior[5]) ? 1 : 0;
  p.always_terms = (p.any_gprior || p.has_uplim[5]) ? 1 : 0;
}

MBB_HD bool below_lowlim(const Priors& pr, const double p[5]) {
  bool bad = false;
#pragma unroll
  for (int i = 0; i < 5; ++i) bad = bad || (p[i] < pr.lowlim[i]);
  return bad;
}

// true when no soft upper limit is exceeded and no other prior term is active:
// the common case, in which prior_terms() would return pen = gp = 0
MBB_HD bool priors_trivial(const Priors& pr, const double p[5]) {
  bool over = false;
#pragma unroll
  for (int i = 0; i < 5; ++i) over = over || (p[i] > pr.uplim[i]);
  return !over && !pr.always_terms;
}

// Soft upper limits and Gaussian priors of the five parameters proper (likelihood.py:672-752
// without the lambda_peak slot), for inlining into the specialised kernels: in a real fit a
// nuisance parameter sits above its soft limit for a good part of the walkers (an optically thin
// fit lets lambda0 wander up to its automatic limit, likelihood.py:187-204), so this is not a cold
// path.  Same terms as prior_terms(), added in the reference's order; the division by the limit
// width is a multiplication by the precomputed inv_w2.  Requires !pr.peak_terms.
MBB_HD double add_simple_priors(const Priors& pr, const double p[5], double lnl) {
  bool over = false;
#pragma unroll
  for (int i = 0; i < 5; ++i) over = over || (p[i] > pr.uplim[i]);      // +inf where no limit is set
  if (over) {
    double pen = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double d = p[i] - pr.uplim[i];
      if (d > 0.0) pen -= (0.5 * (d * d)) * pr.inv_w2[i];
    }
    lnl += pen;
  }
  if (pr.any_gprior) {
    double gp = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      if (pr.has_gprior[i]) {
        const double d = p[i] - pr.gmean[i];
        gp -= 0.5 * pr.givar[i] * (d * d);
      }
    }
    lnl += gp;
  }
  return lnl;
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
// Gaussian priors.  `Tab` is any type with freq/w/weff/lp/band_off/
// scalar_path members indexable by node / band (SmallTab in the kernel
// parameter block, TabView over plain arrays): w = passband weight (FAITHFUL),
// weff / lp = the FAST node constants (see FastSed).
// ---------------------------------------------------------------------------
struct TabView {
  const double* freq;
  const double* w;
  const double* weff;
  const double* lp;
  const int* band_off;
  const unsigned char* scalar_path;
  int nb;
};

constexpr int kMaxBandsPerThread = 64;

// chi-square of one evaluation given a callable returning the weighted node term
// Quadratic form of a full covariance: diff' inv(C) diff with the explicit inverse (what the
// reference does, likelihood.py:356, 823), or |L^-1 diff|^2 by forward substitution when the
// caller staged the Cholesky factor C = L L' (mbb_set_data_chol: strictly lower triangle = L,
// diagonal = 1/L_rr; `diff` is overwritten with L^-1 diff).
MBB_HD double quad_form(const double* m, double* diff, int nb, bool chol) {
  double chi = 0.0;
  if (chol) {
    for (int r = 0; r < nb; ++r) {
      double acc = diff[r];
      for (int c = 0; c < r; ++c) acc = fma(-m[r * nb + c], diff[c], acc);
      acc *= m[r * nb + r];
      diff[r] = acc;
      chi = fma(acc, acc, chi);
    }
    return chi;
  }
  for (int r = 0; r < nb; ++r) {
    double row = 0.0;
    for (int c = 0; c < nb; ++c) row = fma(m[r * nb + c], diff[c], row);
    chi = fma(diff[r], row, chi);
  }
  return chi;
}

template <class Tab, class NodeFn>
MBB_HD double chi_square(const Tab& t, const double* flux, const double* ivar, const double* cinv,
                         NodeFn node, bool chol = false) {
  const int nb = t.nb;
  double chi = 0.0;
  if (!cinv) {
    for (int b = 0; b < nb; ++b) {
      double acc = 0.0;
      for (int i = t.band_off[b]; i < t.band_off[b + 1]; ++i) acc = node(b, i, acc);
      const double df = flux[b] - acc;
      chi = fma(df * df, ivar[b], chi);
    }
    return chi;
  }
  double diff[kMaxBandsPerThread];
  for (int b = 0; b < nb; ++b) {
    double acc = 0.0;
    for (int i = t.band_off[b]; i < t.band_off[b + 1]; ++i) acc = node(b, i, acc);
    diff[b] = flux[b] - acc;
  }
  return quad_form(cinv, diff, nb, chol);
}

// FAST evaluations here always run the saturating (CLAMP) node code: this is
// the generic path (few-node tables that are not all-delta, out-of-range
// walkers handed over by the specialised kernels).
template <bool THIN, bool ALPHA, bool FAST, class Tab>
MBB_HD double loglike_one(const double p[5], const ModelP& m, const Priors& pr, const Tab& t,
                          const double* flux, const double* ivar, const double* cinv, int& st,
                          bool chol = false) {
  st = ST_OK;
  if (below_lowlim(pr, p)) {
    st = ST_BELOW_LOWLIM;
    return -kInf;
  }
  const double nan = kInf - kInf;
  double chi, pen, gp;
  if (FAST) {
    const double* tab = exp2_tab_default();
    FastSed s;
    fast_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], m);
    st = s.status;
    if (st != ST_OK) return nan;
    chi = chi_square(t, flux, ivar, cinv, [&](int, int i, double acc) {
      return node_acc<THIN, ALPHA, true, 0>(s, t.freq[i], t.lp[i], t.weff[i], acc, tab);
    }, chol);
    prior_terms<THIN>(pr, p, s.T, s.beta, s.x0, pen, gp, st);
  } else {
    Sed s;
    sed_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], m.wavenorm);
    st = s.status;
    if (st != ST_OK) return nan;
    chi = chi_square(t, flux, ivar, cinv, [&](int b, int i, double acc) {
      return fma(node_fnu<THIN, ALPHA>(s, (t.scalar_path[b] ? s.hokt_e9 : s.hokt9) * t.freq[i]), t.w[i], acc);
    }, chol);
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
