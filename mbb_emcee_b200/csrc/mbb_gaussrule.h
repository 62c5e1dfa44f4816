// Gauss quadrature rule of a discrete positive measure -- host-only, plain C++.
//
// A tabulated passband turns the band flux into a weighted sum over its N
// table nodes, sum_i w_i f_nu(nu_i) (reference response.py:544-576), N = 130-1108
// for the shipped filters.  f_nu is analytic in nu across a band, so the same
// sum is reproduced to rounding by the n-point Gauss rule of the discrete
// measure {nu_i, w_i}: n nodes nu*_k and weights W_k with
//     sum_k W_k p(nu*_k) = sum_i w_i p(nu_i)   for every polynomial p, deg p <= 2n-1.
// The error for an integrand of exponential type tau over the band's half-width
// is ~ tau^(2n)/(2n)!; the kernels use the compressed rule only where a
// per-walker bound on tau makes that < 1e-17 of the band flux, and never for a
// band that contains the walker's merge point (f_nu has a kink there); every
// other (walker, band) pair takes the full table.  See MBB_MATH_FAST_GAUSS.
//
// Construction: the recursion coefficients of the measure's orthogonal
// polynomials by the Gragg-Harrod RKPW update (numerically stable for discrete
// measures; W. Gautschi, "Orthogonal Polynomials: Computation and
// Approximation", 2004, sec. 2.2.3.2), then Golub-Welsch: nodes = eigenvalues
// of the Jacobi matrix, weights = mass * (first eigenvector components)^2, by
// the implicit QL iteration.  All in long double (x87 80-bit) on the abscissa
// mapped to [-1, 1].
#pragma once
#include <cmath>
#include <vector>

#include "mbb_model.cuh"

namespace mbb {

// recursion coefficients alpha_k (a) and beta_k (b; b[0] = total mass) of the
// discrete measure {x_i, w_i}, w_i > 0, for k < ncap <= N
inline void rkpw_coefficients(const std::vector<long double>& x, const std::vector<long double>& w, int ncap,
                              std::vector<long double>& a, std::vector<long double>& b) {
  const int N = (int)x.size();
  std::vector<long double> p0(x), p1(N, 0.0L);
  p1[0] = w[0];
  for (int n = 0; n < N - 1; ++n) {
    long double pn = w[n + 1], gam = 1.0L, sig = 0.0L, t = 0.0L;
    const long double xlam = x[n + 1];
    for (int k = 0; k <= n + 1; ++k) {
      const long double rho = p1[k] + pn;
      const long double tmp = gam * rho;
      long double tsig = sig;
      if (rho <= 0.0L) {
        gam = 1.0L;
        sig = 0.0L;
      } else {
        gam = p1[k] / rho;
        sig = pn / rho;
      }
      const long double tk = sig * (p0[k] - xlam) - gam * t;
      p0[k] -= (tk - t);
      t = tk;
      if (sig <= 0.0L) pn = tsig * p1[k];
      else pn = (t * t) / sig;
      tsig = sig;
      p1[k] = tmp;
    }
  }
  a.assign(p0.begin(), p0.begin() + ncap);
  b.assign(p1.begin(), p1.begin() + ncap);
}

// eigenvalues d[] and squared first eigenvector components z[]^2 of the symmetric
// tridiagonal matrix (diagonal d, off-diagonal e[1..n-1]) by implicit QL
inline bool tridiagonal_ql_first_row(std::vector<long double>& d, std::vector<long double>& e,
                                     std::vector<long double>& z) {
  const int n = (int)d.size();
  z.assign(n, 0.0L);
  z[0] = 1.0L;
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0L;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const long double dd = fabsl(d[m]) + fabsl(d[m + 1]);
        if (fabsl(e[m]) <= 1e-19L * dd) break;
      }
      if (m != l) {
        if (iter++ == 200) return false;
        long double g = (d[l + 1] - d[l]) / (2.0L * e[l]);
        long double r = hypotl(g, 1.0L);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0L ? fabsl(r) : -fabsl(r)));
        long double s = 1.0L, c = 1.0L, p = 0.0L;
        int i;
        for (i = m - 1; i >= l; --i) {
          long double f = s * e[i];
          const long double bb = c * e[i];
          e[i + 1] = (r = hypotl(f, g));
          if (r == 0.0L) {
            d[i + 1] -= p;
            e[m] = 0.0L;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0L * c * bb;
          d[i + 1] = g + (p = s * r);
          g = c * r - bb;
          f = z[i + 1];
          z[i + 1] = s * z[i] + c * f;
          z[i] = c * z[i] - s * f;
        }
        if (r == 0.0L && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0L;
      }
    } while (m != l);
  }
  return true;
}

// n-point Gauss rule {xs, ws} of the discrete measure {x_i, w_i}.  Returns false
// (caller keeps the full table) when the measure is not positive, has fewer
// than 2n points, or the eigen-iteration fails.
inline bool discrete_gauss_rule(const std::vector<double>& x_in, const std::vector<double>& w_in, int n,
                                std::vector<double>& xs, std::vector<double>& ws) {
  std::vector<long double> x, w;
  long double lo = 0.0L, hi = 0.0L;
  bool first = true;
  for (size_t i = 0; i < x_in.size(); ++i) {
    if (w_in[i] < 0.0 || !(w_in[i] == w_in[i])) return false;
    if (w_in[i] == 0.0) continue;
    if (first || x_in[i] < lo) lo = x_in[i];
    if (first || x_in[i] > hi) hi = x_in[i];
    first = false;
    x.push_back(x_in[i]);
    w.push_back(w_in[i]);
  }
  if ((int)x.size() < 2 * n || !(hi > lo)) return false;
  const long double mid = 0.5L * (hi + lo), half = 0.5L * (hi - lo);
  for (auto& v : x) v = (v - mid) / half;
  std::vector<long double> a, b;
  rkpw_coefficients(x, w, n, a, b);
  for (int k = 1; k < n; ++k)
    if (!(b[k] > 0.0L)) return false;
  std::vector<long double> d(a), e(n, 0.0L), z;
  for (int k = 1; k < n; ++k) e[k] = sqrtl(b[k]);
  if (!tridiagonal_ql_first_row(d, e, z)) return false;
  xs.resize(n);
  ws.resize(n);
  for (int k = 0; k < n; ++k) {
    xs[k] = (double)(mid + half * d[k]);
    ws[k] = (double)(b[0] * z[k] * z[k]);
    if (!(ws[k] > 0.0) || xs[k] < (double)lo || xs[k] > (double)hi) return false;
  }
  return true;
}

// The compressed tables of a band set: for every tabulated band with more than
// 2*kGaussPoints nodes the kGaussPoints-point rule of {freq_i, w_i}, its nodes
// turned into FAST node records by the same fast_node() the full tables use.
struct GaussTables {
  std::vector<double> freq, weff, lp;   // compressed nodes, band after band
  std::vector<int> off;                 // nb + 1
  std::vector<BandMeta> meta;           // nb
};

inline GaussTables build_gauss_tables(int nb, const int* band_off, const double* wave_um, const double* weight,
                                      const unsigned char* scalar_path, double wavenorm, bool thin) {
  GaussTables g;
  g.off.assign(nb + 1, 0);
  g.meta.resize(nb);
  for (int b = 0; b < nb; ++b) {
    const int i0 = band_off[b], i1 = band_off[b + 1];
    std::vector<double> x, w, xs, ws;
    double lo = 0.0, hi = 0.0;
    for (int i = i0; i < i1; ++i) {
      const double f = kUmToGHz / wave_um[i];
      x.push_back(f);
      w.push_back(weight[i]);
      if (i == i0 || f < lo) lo = f;
      if (i == i0 || f > hi) hi = f;
    }
    g.meta[b].nu_lo = lo;
    g.meta[b].nu_hi = hi;
    g.meta[b].dl = log(hi / lo);
    g.meta[b].has_rule = 0.0;
    if (i1 - i0 > 2 * kGaussPoints && !(scalar_path && scalar_path[b]) &&
        discrete_gauss_rule(x, w, kGaussPoints, xs, ws)) {
      bool descending = true;
      for (int i = i0 + 1; i < i1; ++i) descending = descending && x[i - i0] < x[i - i0 - 1];
      g.meta[b].has_rule = descending ? 2.0 : 1.0;
      for (int k = 0; k < kGaussPoints; ++k) {
        const FastNode n = fast_node(kUmToGHz / xs[k], ws[k], wavenorm, thin);
        g.freq.push_back(xs[k]);
        g.weff.push_back(n.weff);
        g.lp.push_back(n.lp);
      }
    }
    g.off[b + 1] = (int)g.freq.size();
  }
  return g;
}

}  // namespace mbb
