// Host experiment behind the merge-point solve of fast_setup (mbb_model.cuh): accuracy of the
// single-precision seed and of ONE double-precision evaluation followed by a Newton resp. Halley
// step, against the converged root, over the cfg2 walker cloud (default) or a wide cloud
// (any argument: T 3-80 K, beta 0.1-9, lambda0 10-1500 um, alpha 0.5-10).
//   g++ -O2 -std=c++17 -DMBB_F32_NOISE -o /tmp/merge_exp tools/merge_solve_experiment.cpp && /tmp/merge_exp [wide]
// MBB_F32_NOISE perturbs the host expf by +-4e-7 relative, the error level of MUFU.EX2.
#include <cstdio>
#include <random>
#include <cmath>
#include <algorithm>
#include <vector>
static int g_f32_iters = 0, g_f64_evals = 0;
#define MBB_COUNT_EVALS 1
#include "../mbb_emcee_b200/csrc/mbb_model.cuh"
using namespace mbb;
int main(int argc, char** argv) {
  const bool wide = argc > 1;
  std::mt19937_64 rng(1);
  std::normal_distribution<double> g(0, 1);
  std::uniform_real_distribution<double> U(0, 1);
  const int N = 200000;
  double worstN = 0, worstH = 0, worstA = 0, maxdu = 0; int nacc = 0;
  for (int i = 0; i < N; ++i) {
    double T, beta, l0, al;
    if (!wide) { T = std::max(14 + 2 * g(rng), 2.1); beta = std::max(1.8 + 0.2 * g(rng), 0.3); l0 = std::max(400 + 100 * g(rng), 2.1); al = std::max(3 + 0.3 * g(rng), 0.3); }
    else { T = 3 + 77 * U(rng); beta = 0.1 + 8.9 * U(rng); l0 = 10 + 1490 * U(rng); al = 0.5 + 9.5 * U(rng); }
    double hokt9 = 1e9 * kH / (kK * T), xn = hokt9 * kUmToGHz / 500, q = log(l0 / 500), u0 = log(xn) - q;
    double a_lo = 3 + al;
    float af = a_lo, bf = beta;
    float ulf = f32_log(af * (1.0f - f32_exp(1.0f - af))) - 1e-3f, uhf = f32_log(af + bf) + 1e-3f;
    double u = thick_merge_seed(af, bf, (float)u0, ulf, uhf);
    MergeEval v;
    const double* tab = exp2_tab_default();
    long double uu = u;
    for (int k = 0; k < 12; ++k) { merge_eval<0>((double)uu, a_lo, beta, u0, tab, v); uu -= (long double)v.G / v.dG; }
    merge_eval<0>(u, a_lo, beta, u0, tab, v);
    const double du = v.G / v.dG;
    // G''
    const double t = v.t, em = v.em1t, et = em + 1;
    double Bp, Bpp;
    if (t < 1e-3) { Bp = -0.5 + t / 6; Bpp = 1.0 / 6; }
    else { const double r = 1 / em; Bp = (em - t * et) * r * r; Bpp = (et * r) * ((t * (et + 1) - 2 * em) * r) * r; }
    const double S = a_lo + beta * (t < 1e-3 ? 1 - t * (0.5 - t / 12) : t / em);
    const double b2 = beta * beta;
    const double G2 = v.x - v.x * v.e * S * (1 - v.x) - 2 * v.x * v.e * b2 * t * Bp - v.E * b2 * beta * t * (Bp + t * Bpp);
    const double duh = du / (1 - 0.5 * du * G2 / v.dG);
    if (fabs(du) <= 1e-5) {
      ++nacc;
      worstN = std::max(worstN, (double)fabsl((u - du) - uu));
      { double eh = (double)fabsl((u - duh) - uu); if (eh > worstH) { worstH = eh; printf("T=%g beta=%g l0=%g al=%g t=%g x=%g du=%g duh=%g G2=%g dG=%g errN=%g\n", T, beta, l0, al, v.t, v.x, du, duh, G2, v.dG, (double)fabsl((u-du)-uu)); } }
      maxdu = std::max(maxdu, fabs(du));
    }
  }
  printf("accepted %d of %d; max |du| %.2e; worst root error: newton %.2e halley %.2e\n", nacc, N, maxdu, worstN, worstH);
}
