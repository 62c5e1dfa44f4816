// QUADPACK QAGS replay (dqagse + dqk21 + dqpsrt + dqelg), per thread.
//
// The reference computes L_IR with scipy.integrate.quad(f_nu, nu_min, nu_max)
// (modified_blackbody.py:671), i.e. QUADPACK's dqagse with epsabs = epsrel =
// 1.49e-8 and limit = 50.  Its result is a deterministic function of the
// integrand values -- but, with the merge-point kink inside the interval, only
// good to ~1e-8.  To reproduce the reference's NUMBER (not merely the
// integral) this header restates the published algorithm (Piessens, de
// Doncker-Kapenga, Ueberhuber, Kahaner, "QUADPACK", Springer 1983; netlib
// quadpack dqagse.f, dqk21.f, dqpsrt.f, dqelg.f) operation for operation: the
// same 21-point Gauss-Kronrod rule, error heuristic, bisection order, roundoff
// counters and epsilon-algorithm extrapolation.  Given integrand values equal to
// a few ulp, the subdivision sequence and therefore the result agree with
// scipy's to ~1e-15 (tests/test_device_logic_cpu.py::test_qags_matches_scipy
// also compares `neval`).
//
// Host+device; all state in local arrays sized for limit = 50.
#pragma once
#include "mbb_model_defs.cuh"

namespace mbb {

constexpr int kQagsLimit = 50;

struct Gk21 {
  double result, abserr, resabs, resasc;
};

// Products and sums of the rule with separate roundings, as the Fortran original compiled
// without fused multiply-add evaluates them (the host build of this header uses
// -ffp-contract=off for the same reason): the result then carries dqk21's own rounding errors.
MBB_HD double q_mul(double x, double y) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(x, y);
#else
  return x * y;
#endif
}
MBB_HD double q_add(double x, double y) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(x, y);
#else
  return x + y;
#endif
}

template <class F>
MBB_HD_NOINLINE Gk21 qk21(F f, double a, double b) {
  const double wg[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
                        0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
                        0.295524224714752870173815619188769};
  const double xgk[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
                          0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
                          0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                          0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
                          0.294392862701460198131126603103866, 0.148874338981631210884826001129720,
                          0.0};
  const double wgk[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
                          0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
                          0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                          0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
                          0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
                          0.149445554002916905664936468389821};
  const double epmach = 2.220446049250313e-16, uflow = 2.2250738585072014e-308;
  const double centr = q_mul(0.5, q_add(a, b));
  const double hlgth = q_mul(0.5, q_add(b, -a));
  const double dhlgth = fabs(hlgth);
  // All 21 integrand values first, through ONE call site in a rolled loop: the integrand (pow,
  // expm1, the merge-point branch) is a few hundred instructions, and 21 inlined copies per rule
  // times three inlined rules made a kernel that did not fit the instruction cache (ncu:
  // `no_instruction` 6 warps per issue; 94 -> 47 ms for the 3.5e6 unique samples of cfg4).  The
  // sums below then run in dqk21's own order.
  double fv1[10], fv2[10];
  double fc = 0.0;
#pragma unroll 1
  for (int i = 0; i < 21; ++i) {          // i = 0: centre; then node (i - 1) / 2, left for odd i, right for even
    const int jj = i == 0 ? 10 : (i - 1) >> 1;            // xgk[10] = 0: the centre
    const double absc = q_mul(hlgth, xgk[jj]);
    const double v = f(q_add(centr, (i & 1) ? -absc : absc));
    if (i == 0) fc = v;
    else if (i & 1) fv1[jj] = v;
    else fv2[jj] = v;
  }
  double resg = 0.0;
  double resk = q_mul(wgk[10], fc);
  double resabs = fabs(resk);
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    const int jtw = 2 * j + 1;
    const double fval1 = fv1[jtw], fval2 = fv2[jtw];
    const double fsum = q_add(fval1, fval2);
    resg = q_add(resg, q_mul(wg[j], fsum));
    resk = q_add(resk, q_mul(wgk[jtw], fsum));
    resabs = q_add(resabs, q_mul(wgk[jtw], q_add(fabs(fval1), fabs(fval2))));
  }
#pragma unroll 1
  for (int j = 0; j < 5; ++j) {
    const int jtwm1 = 2 * j;
    const double fval1 = fv1[jtwm1], fval2 = fv2[jtwm1];
    const double fsum = q_add(fval1, fval2);
    resk = q_add(resk, q_mul(wgk[jtwm1], fsum));
    resabs = q_add(resabs, q_mul(wgk[jtwm1], q_add(fabs(fval1), fabs(fval2))));
  }
  const double reskh = q_mul(resk, 0.5);
  double resasc = q_mul(wgk[10], fabs(q_add(fc, -reskh)));
#pragma unroll 1
  for (int j = 0; j < 10; ++j)
    resasc = q_add(resasc, q_mul(wgk[j], q_add(fabs(q_add(fv1[j], -reskh)), fabs(q_add(fv2[j], -reskh)))));
  Gk21 r;
  r.result = q_mul(resk, hlgth);
  r.resabs = q_mul(resabs, dhlgth);
  r.resasc = q_mul(resasc, dhlgth);
  r.abserr = fabs(q_mul(q_add(resk, -resg), hlgth));
  if (r.resasc != 0.0 && r.abserr != 0.0)
    r.abserr = r.resasc * fmin(1.0, pow(200.0 * r.abserr / r.resasc, 1.5));
  if (r.resabs > uflow / (50.0 * epmach)) r.abserr = fmax((epmach * 50.0) * r.resabs, r.abserr);
  return r;
}

// dqpsrt: keep iord[] ordering the error estimates in descending order.
// Arrays are 1-based like the original (index 0 unused).
MBB_HD void qpsrt(int limit, int last, int& maxerr, double& ermax, const double* elist, int* iord,
                  int& nrmax) {
  if (last <= 2) {
    iord[1] = 1;
    iord[2] = 2;
  } else {
    const double errmax = elist[maxerr];
    if (nrmax != 1) {
      const int ido = nrmax - 1;
      for (int i = 1; i <= ido; ++i) {
        const int isucc = iord[nrmax - 1];
        if (errmax <= elist[isucc]) break;
        iord[nrmax] = isucc;
        nrmax = nrmax - 1;
      }
    }
    int jupbn = last;
    if (last > (limit / 2 + 2)) jupbn = limit + 3 - last;
    const double errmin = elist[last];
    const int jbnd = jupbn - 1;
    const int ibeg = nrmax + 1;
    bool inserted = false;
    int i = ibeg;
    if (ibeg <= jbnd) {
      for (i = ibeg; i <= jbnd; ++i) {
        const int isucc = iord[i];
        if (errmax >= elist[isucc]) { inserted = true; break; }
        iord[i - 1] = isucc;
      }
    }
    if (!inserted) {
      iord[jbnd] = maxerr;
      iord[jupbn] = last;
    } else {
      iord[i - 1] = maxerr;
      int k = jbnd;
      bool placed = false;
      for (int j = i; j <= jbnd; ++j) {
        const int isucc = iord[k];
        if (errmin < elist[isucc]) { iord[k + 1] = last; placed = true; break; }
        iord[k + 1] = isucc;
        k = k - 1;
      }
      if (!placed) iord[i] = last;
    }
  }
  maxerr = iord[nrmax];
  ermax = elist[maxerr];
}

// dqelg: epsilon algorithm.  epstab is 1-based with 52 usable entries.
MBB_HD void qelg(int& n, double* epstab, double& result, double& abserr, double* res3la, int& nres) {
  const double epmach = 2.220446049250313e-16, oflow = 1.7976931348623157e308;
  nres = nres + 1;
  abserr = oflow;
  result = epstab[n];
  if (n >= 3) {
    const int limexp = 50;
    epstab[n + 2] = epstab[n];
    const int newelm = (n - 1) / 2;
    epstab[n] = oflow;
    const int num = n;
    int k1 = n;
    bool converged = false;
    for (int i = 1; i <= newelm; ++i) {
      const int k2 = k1 - 1;
      const int k3 = k1 - 2;
      double res = epstab[k1 + 2];
      const double e0 = epstab[k3];
      const double e1 = epstab[k2];
      const double e2 = res;
      const double e1abs = fabs(e1);
      const double delta2 = e2 - e1;
      const double err2 = fabs(delta2);
      const double tol2 = fmax(fabs(e2), e1abs) * epmach;
      const double delta3 = e1 - e0;
      const double err3 = fabs(delta3);
      const double tol3 = fmax(e1abs, fabs(e0)) * epmach;
      if (!(err2 > tol2 || err3 > tol3)) {
        // e0, e1, e2 equal to within machine accuracy: convergence
        result = res;
        abserr = err2 + err3;
        abserr = fmax(abserr, 5.0 * epmach * fabs(result));
        converged = true;
        break;
      }
      const double e3 = epstab[k1];
      epstab[k1] = e1;
      const double delta1 = e1 - e3;
      const double err1 = fabs(delta1);
      const double tol1 = fmax(e1abs, fabs(e3)) * epmach;
      bool omit = (err1 <= tol1 || err2 <= tol2 || err3 <= tol3);
      double ss = 0.0;
      if (!omit) {
        ss = 1.0 / delta1 + 1.0 / delta2 - 1.0 / delta3;
        const double epsinf = fabs(ss * e1);
        omit = !(epsinf > 1e-4);
      }
      if (omit) {
        n = i + i - 1;
        break;
      }
      res = e1 + 1.0 / ss;
      epstab[k1] = res;
      k1 = k1 - 2;
      const double error = err2 + fabs(res - e2) + err3;
      if (!(error > abserr)) {
        abserr = error;
        result = res;
      }
    }
    if (converged) {
      abserr = fmax(abserr, 5.0 * epmach * fabs(result));
      return;
    }
    // shift the table
    if (n == limexp) n = 2 * (limexp / 2) - 1;
    int ib = ((num / 2) * 2 == num) ? 2 : 1;
    const int ie = newelm + 1;
    for (int i = 1; i <= ie; ++i) {
      const int ib2 = ib + 2;
      epstab[ib] = epstab[ib2];
      ib = ib2;
    }
    if (num != n) {
      int indx = num - n + 1;
      for (int i = 1; i <= n; ++i) {
        epstab[i] = epstab[indx];
        indx = indx + 1;
      }
    }
    if (nres < 4) {
      res3la[nres] = result;
      abserr = oflow;
    } else {
      abserr = fabs(result - res3la[3]) + fabs(result - res3la[2]) + fabs(result - res3la[1]);
      res3la[1] = res3la[2];
      res3la[2] = res3la[3];
      res3la[3] = result;
    }
  }
  abserr = fmax(abserr, 5.0 * epmach * fabs(result));
}

struct QagsOut {
  double result, abserr;
  int neval, ier;
};

template <class F>
MBB_HD QagsOut qagse(F f, double a, double b, double epsabs, double epsrel) {
  const int limit = kQagsLimit;
  const double epmach = 2.220446049250313e-16, uflow = 2.2250738585072014e-308,
               oflow = 1.7976931348623157e308;
  double alist[kQagsLimit + 1], blist[kQagsLimit + 1], rlist[kQagsLimit + 1], elist[kQagsLimit + 1];
  int iord[kQagsLimit + 1];
  double rlist2[53], res3la[4];
  QagsOut o;
  int ier = 0, last = 0;
  double result = 0.0, abserr = 0.0;
  alist[1] = a; blist[1] = b; rlist[1] = 0.0; elist[1] = 0.0;
  int ierro = 0;
  Gk21 g = qk21(f, a, b);
  result = g.result;
  abserr = g.abserr;
  const double defabs = g.resabs;
  double resabs = g.resasc;
  double dres = fabs(result);
  double errbnd = fmax(epsabs, epsrel * dres);
  last = 1;
  rlist[1] = result;
  elist[1] = abserr;
  iord[1] = 1;
  if (abserr <= 100.0 * epmach * defabs && abserr > errbnd) ier = 2;
  if (limit == 1) ier = 1;
  if (ier != 0 || (abserr <= errbnd && abserr != resabs) || abserr == 0.0) {
    o.result = result; o.abserr = abserr; o.neval = 42 * last - 21; o.ier = ier;
    return o;
  }
  rlist2[1] = result;
  double errmax = abserr;
  int maxerr = 1;
  double area = result;
  double errsum = abserr;
  abserr = oflow;
  int nrmax = 1, nres = 0, numrl2 = 2, ktmin = 0;
  bool extrap = false, noext = false;
  int iroff1 = 0, iroff2 = 0, iroff3 = 0;
  int ksgn = -1;
  if (dres >= (1.0 - 50.0 * epmach) * defabs) ksgn = 1;
  double small = 0.0, erlarg = 0.0, ertest = 0.0, correc = 0.0, erlast;
  double reseps = 0.0, abseps = 0.0;
  bool sum_up = false;     // "go to 115"
  for (last = 2; last <= limit; ++last) {
    const double a1 = alist[maxerr];
    const double b1 = 0.5 * (alist[maxerr] + blist[maxerr]);
    const double a2 = b1;
    const double b2 = blist[maxerr];
    erlast = errmax;
    // (the two halves through one call site: one copy of the rule and its integrand in the code)
    Gk21 gh[2];
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) gh[hf] = qk21(f, hf ? a2 : a1, hf ? b2 : b1);
    const Gk21 g1 = gh[0], g2 = gh[1];
    const double area1 = g1.result, error1 = g1.abserr, defab1 = g1.resasc;
    const double area2 = g2.result, error2 = g2.abserr, defab2 = g2.resasc;
    const double area12 = area1 + area2;
    const double erro12 = error1 + error2;
    errsum = errsum + erro12 - errmax;
    area = area + area12 - rlist[maxerr];
    if (!(defab1 == error1 || defab2 == error2)) {
      if (!(fabs(rlist[maxerr] - area12) > 1e-5 * fabs(area12) || erro12 < 0.99 * errmax)) {
        if (extrap) iroff2 = iroff2 + 1;
        else iroff1 = iroff1 + 1;
      }
      if (last > 10 && erro12 > errmax) iroff3 = iroff3 + 1;
    }
    rlist[maxerr] = area1;
    rlist[last] = area2;
    errbnd = fmax(epsabs, epsrel * fabs(area));
    if (iroff1 + iroff2 >= 10 || iroff3 >= 20) ier = 2;
    if (iroff2 >= 5) ierro = 3;
    if (last == limit) ier = 1;
    if (fmax(fabs(a1), fabs(b2)) <= (1.0 + 100.0 * epmach) * (fabs(a2) + 1000.0 * uflow)) ier = 4;
    if (error2 > error1) {
      alist[maxerr] = a2;
      alist[last] = a1;
      blist[last] = b1;
      rlist[maxerr] = area2;
      rlist[last] = area1;
      elist[maxerr] = error2;
      elist[last] = error1;
    } else {
      alist[last] = a2;
      blist[maxerr] = b1;
      blist[last] = b2;
      elist[maxerr] = error1;
      elist[last] = error2;
    }
    qpsrt(limit, last, maxerr, errmax, elist, iord, nrmax);
    if (errsum <= errbnd) { sum_up = true; break; }
    if (ier != 0) break;
    if (last == 2) {
      small = fabs(b - a) * 0.375;
      erlarg = errsum;
      ertest = errbnd;
      rlist2[2] = area;
      continue;
    }
    if (noext) continue;
    erlarg = erlarg - erlast;
    if (fabs(b1 - a1) > small) erlarg = erlarg + erro12;
    if (!extrap) {
      // is the interval to be bisected next the smallest one?
      if (fabs(blist[maxerr] - alist[maxerr]) > small) continue;
      extrap = true;
      nrmax = 2;
    }
    bool do_extrap = true;
    if (!(ierro == 3 || erlarg <= ertest)) {
      // the smallest interval has the largest error: before bisecting, decrease the
      // sum of the errors over the larger intervals and extrapolate
      const int id = nrmax;
      int jupbnd = last;
      if (last > (2 + limit / 2)) jupbnd = limit + 3 - last;
      for (int k = id; k <= jupbnd; ++k) {
        maxerr = iord[nrmax];
        errmax = elist[maxerr];
        if (fabs(blist[maxerr] - alist[maxerr]) > small) { do_extrap = false; break; }
        nrmax = nrmax + 1;
      }
    }
    if (!do_extrap) continue;
    numrl2 = numrl2 + 1;
    rlist2[numrl2] = area;
    qelg(numrl2, rlist2, reseps, abseps, res3la, nres);
    ktmin = ktmin + 1;
    if (ktmin > 5 && abserr < 1e-3 * errsum) ier = 5;
    if (!(abseps >= abserr)) {
      ktmin = 0;
      abserr = abseps;
      result = reseps;
      correc = erlarg;
      ertest = fmax(epsabs, epsrel * fabs(reseps));
      if (abserr <= ertest) break;
    }
    if (numrl2 == 1) noext = true;
    if (ier == 5) break;
    maxerr = iord[1];
    errmax = elist[maxerr];
    nrmax = 1;
    extrap = false;
    small = small * 0.5;
    erlarg = errsum;
  }
  if (last > limit) last = limit;      // the do-loop ran to completion
  // label 100: set final result and error estimate
  if (!sum_up) {
    bool to115 = false, to130 = false;
    if (abserr == oflow) {
      to115 = true;
    } else if (ier + ierro != 0) {
      if (ierro == 3) abserr = abserr + correc;
      if (ier == 0) ier = 3;
      if (result != 0.0 && area != 0.0) {
        if (abserr / fabs(result) > errsum / fabs(area)) to115 = true;
      } else if (abserr > errsum) {
        to115 = true;
      } else if (area == 0.0) {
        to130 = true;
      }
    }
    if (!to115 && !to130) {
      // label 110: test on divergence
      if (!(ksgn == -1 && fmax(fabs(result), fabs(area)) <= defabs * 0.01)) {
        if (0.01 > (result / area) || (result / area) > 100.0 || errsum > fabs(area)) ier = 6;
      }
    }
    sum_up = to115;
  }
  if (sum_up) {
    result = 0.0;
    for (int k = 1; k <= last; ++k) result = result + rlist[k];
    abserr = errsum;
  }
  if (ier > 2) ier = ier - 1;
  o.result = result; o.abserr = abserr; o.neval = 42 * last - 21; o.ier = ier;
  return o;
}

}  // namespace mbb
