// TEST INFRASTRUCTURE -- host emulation of the device model code.
//
// Compiles mbb_emcee_b200/csrc/mbb_model.cuh (the exact source the CUDA
// kernels are built from) with g++ and runs it in plain host loops, so the CPU
// test-suite can check the *logic* of the device code (setup, node formulas in
// both arithmetic modes, root solves, limits/priors, chi-square) against the
// oracle without a GPU.  libm replaces libdevice, so agreement with the GPU is
// to rounding, not bit for bit.  Built into tests/_hostemu/ (git-ignored);
// NEVER loaded by the product package -- the product has no CPU path.
#include <cmath>
#include <cstdint>
#include <vector>

#include "../../mbb_emcee_b200/csrc/mbb_model.cuh"
#include "../../mbb_emcee_b200/csrc/mbb_philox.cuh"
#include "../../mbb_emcee_b200/csrc/mbb_quadpack.cuh"
#include "../../mbb_emcee_b200/csrc/mbb_gaussrule.h"
#include "../../mbb_emcee_b200/csrc/mbb_hostutil.h"

using namespace mbb;

namespace {
// Table flavour of the specialised node code being emulated: 0 = 64 entries (delta kernels),
// kTab256 = 256 entries (nodes kernel, Gauss-rule kernels: per-walker constants and L' times 4).
int g_ts = 0;

template <bool THIN, bool ALPHA, bool FAST, int TS = 0>
void run_loglike(long long n, const double* pars, const ModelP& m, const Priors& pr, const TabView& t0,
                 const double* flux, const double* ivar, const double* cinv, long long wps, int unclamped,
                 double* out, int* status, const GaussTables* gt0 = nullptr, long long* ncompressed = nullptr) {
  if (TS == 0 && g_ts != 0 && FAST && unclamped) {
    run_loglike<THIN, ALPHA, FAST, kTab256>(n, pars, m, pr, t0, flux, ivar, cinv, wps, unclamped, out, status, gt0,
                                            ncompressed);
    return;
  }
  // the 256 flavour reads node tables scaled like the device ones (mbb_capi.cu: L' * 4)
  TabView t = t0;
  std::vector<double> lp_scaled;
  GaussTables gts;
  const GaussTables* gt = gt0;
  if (TS != 0) {
    lp_scaled.assign(t0.lp, t0.lp + t0.band_off[t0.nb]);
    for (auto& v : lp_scaled) v *= 4.0;
    if (gt0) {
      gts = *gt0;
      for (auto& v : gts.lp) v *= 4.0;
      gt = &gts;
    }
  }
  const double* lpn = TS != 0 ? lp_scaled.data() : t0.lp;
  const int nb = t.nb;
  for (long long e = 0; e < n; ++e) {
    const long long src = e / wps;
    const double* fl = flux + src * nb;
    const double* iv = ivar ? ivar + src * nb : nullptr;
    const double* ci = cinv ? cinv + src * nb * nb : nullptr;
    int st;
    if (FAST && unclamped) {
      // the arithmetic of the specialised kernels (loglike_delta_kernel / nodes kernel):
      // CLAMP=false node code for `safe` walkers, the generic path otherwise
      const double* p = pars + 5 * e;
      const double* tab = TS != 0 ? exp2_tab256_default() : exp2_tab_default();
      FastSed s;
      st = ST_OK;
      if (below_lowlim(pr, p)) { out[e] = -kInf; status[e] = ST_BELOW_LOWLIM; continue; }
      fast_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], m);
      if (TS != 0) fast_sed_rescale256(s);
      if (s.status == ST_OK && s.safe) {
        double chi;
        if (gt) {
          // MBB_MATH_FAST_GAUSS: the band's 32-point rule where gauss_band_masks allows it
          // (serial mirror of band_partial_fast / band_partial_kink of the nodes kernel)
          const GaussMasks gm = gauss_band_masks<THIN, ALPHA>(s, gt->meta.data(), nb);
          double diff[kMaxBandsPerThread];
          for (int b = 0; b < nb; ++b) {
            double acc = 0.0;
            const int i0 = t.band_off[b], i1 = t.band_off[b + 1];
            if ((gm.plain >> b) & 1ull) {
              if (ncompressed) ++*ncompressed;
              for (int i = gt->off[b]; i < gt->off[b + 1]; ++i)
                acc = node_acc<THIN, ALPHA, false, TS>(s, gt->freq[i], gt->lp[i], gt->weff[i], acc, tab);
            } else if (ALPHA && ((gm.kink >> b) & 1ull)) {
              if (ncompressed) ++*ncompressed;
              int k = i0;
              while (k < i1 && t.freq[k] > s.nu_merge) ++k;
              const bool grey_rule = k - i0 <= i1 - k;
              for (int i = gt->off[b]; i < gt->off[b + 1]; ++i)
                acc = grey_rule ? node_grey<THIN, false, TS>(s, gt->freq[i], gt->lp[i], gt->weff[i], acc, tab)
                                : node_pow<false, TS>(s, gt->lp[i], gt->weff[i], acc, tab);
              if (grey_rule) {
                for (int i = i0; i < k; ++i) {
                  acc = node_pow<false, TS>(s, lpn[i], t.weff[i], acc, tab);
                  acc = node_grey<THIN, false, TS>(s, t.freq[i], lpn[i], -t.weff[i], acc, tab);
                }
              } else {
                for (int i = k; i < i1; ++i) {
                  acc = node_grey<THIN, false, TS>(s, t.freq[i], lpn[i], t.weff[i], acc, tab);
                  acc = node_pow<false, TS>(s, lpn[i], -t.weff[i], acc, tab);
                }
              }
            } else {
              for (int i = t.band_off[b]; i < t.band_off[b + 1]; ++i)
                acc = node_acc<THIN, ALPHA, false, TS>(s, t.freq[i], lpn[i], t.weff[i], acc, tab);
            }
            diff[b] = fl[b] - acc;
          }
          chi = 0.0;
          if (!ci) {
            for (int b = 0; b < nb; ++b) chi = fma(diff[b] * diff[b], iv[b], chi);
          } else {
            for (int r = 0; r < nb; ++r) {
              double row = 0.0;
              for (int c = 0; c < nb; ++c) row = fma(ci[r * nb + c], diff[c], row);
              chi = fma(diff[r], row, chi);
            }
          }
        } else
        chi = chi_square(t, fl, iv, ci, [&](int, int i, double acc) {
          return node_acc<THIN, ALPHA, false, TS>(s, t.freq[i], lpn[i], t.weff[i], acc, tab);
        });
        double lnl = -0.5 * chi;
        if (pr.peak_terms) {           // as delta_eval / the Gauss-rule thread kernel do
          double pen, gp;
          prior_terms<THIN>(pr, p, s.T, s.beta, s.x0, pen, gp, st);
          lnl += pen;
          if (pr.any_gprior) lnl += gp;
        } else {
          lnl = add_simple_priors(pr, p, lnl);
        }
        if (st != ST_OK) lnl = kInf - kInf;
        else if (lnl != lnl) st = ST_NONFINITE;
        out[e] = lnl;
        status[e] = st;
        continue;
      }
    }
    out[e] = loglike_one<THIN, ALPHA, FAST>(pars + 5 * e, m, pr, t0, fl, iv, ci, st);
    status[e] = st;
  }
}

template <bool THIN, bool ALPHA>
void run_consts(long long n, const double* pars, double wavenorm, int want_peak, double* out, int* status) {
  for (long long e = 0; e < n; ++e) {
    const double* p = pars + 5 * e;
    Sed s;
    sed_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], wavenorm);
    int st = s.status;
    double peak = 0.0;
    if (want_peak && st == ST_OK) peak = max_wave<THIN>(s.T, s.beta, s.x0, st);
    double* o = out + 6 * e;
    o[0] = s.normfac; o[1] = s.xmerge; o[2] = s.kappa; o[3] = s.x0; o[4] = s.xnorm; o[5] = peak;
    status[e] = st;
  }
}

template <bool THIN, bool ALPHA>
void run_fnu(long long n, const double* pars, double wavenorm, int nfreq, const double* freq, int scalar_path,
             int fast, double* out) {
  // fast: 1 = CLAMP node code, 2 = CLAMP=false node code where the walker is `safe`
  std::vector<FastNode> nodes(nfreq);
  ModelP m{wavenorm, kUmToGHz / wavenorm, 0.0, 0.0};
  for (int i = 0; i < nfreq; ++i) {
    nodes[i] = fast_node(kUmToGHz / freq[i], 1.0, wavenorm, THIN);
    nodes[i].freq = freq[i];
    if (freq[i] > m.nu_max) m.nu_max = freq[i];
    if (nodes[i].labs > m.lmax) m.lmax = nodes[i].labs;
  }
  const double* tab = exp2_tab_default();
  for (long long e = 0; e < n; ++e) {
    const double* p = pars + 5 * e;
    Sed s;
    FastSed fs;
    int status;
    if (fast) {
      fast_setup<THIN, ALPHA>(fs, p[0], p[1], p[2], p[3], p[4], m);
      status = fs.status;
    } else {
      sed_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], wavenorm);
      status = s.status;
    }
    for (int i = 0; i < nfreq; ++i) {
      double v;
      if (status != ST_OK) {
        v = kInf - kInf;
      } else if (fast) {
        const FastNode& nd = nodes[i];
        if (fast == 2 && fs.safe) v = node_acc<THIN, ALPHA, false, 0>(fs, nd.freq, nd.lp, nd.weff, 0.0, tab);
        else v = node_acc<THIN, ALPHA, true, 0>(fs, nd.freq, nd.lp, nd.weff, 0.0, tab);
      } else {
        v = node_fnu<THIN, ALPHA>(s, (scalar_path ? s.hokt_e9 : s.hokt9) * freq[i]);
      }
      out[e * nfreq + i] = v;
    }
  }
}

template <bool THIN, bool ALPHA>
void run_lir(long long n, const double* pars, double wavenorm, double fmin, double fmax, double prefac,
             double* out, int* status) {
  for (long long e = 0; e < n; ++e) {
    const double* p = pars + 5 * e;
    Sed s;
    sed_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], wavenorm);
    status[e] = s.status;
    if (s.status != ST_OK) { out[e] = kInf - kInf; continue; }
    const LirSpan sp = lir_span<ALPHA>(s.hokt_e9 * fmin, s.hokt_e9 * fmax, s.beta, s.alpha, s.xmerge, s.kappa);
    double acc = 0.0;
    for (int node = 0; node < kLirNodes; ++node) acc += lir_node<THIN>(sp, node, s.beta, s.x0);
    out[e] = prefac * (1e-17 * (s.normfac * (acc + sp.pow_part) / s.hokt_e9));
  }
}
template <bool THIN, bool ALPHA>
void run_qags(long long n, const double* pars, double wavenorm, double fmin, double fmax, double prefac,
              double* out, int* neval, int* status) {
  for (long long e = 0; e < n; ++e) {
    const double* p = pars + 5 * e;
    Sed s;
    sed_setup<THIN, ALPHA>(s, p[0], p[1], p[2], p[3], p[4], wavenorm);
    status[e] = s.status;
    if (s.status != ST_OK) { out[e] = kInf - kInf; neval[e] = 0; continue; }
    const QagsOut q = qagse([&](double nu) { return node_fnu<THIN, ALPHA>(s, s.hokt_e9 * nu); }, fmin, fmax,
                            1.49e-8, 1.49e-8);
    out[e] = prefac * (1e-17 * q.result);
    neval[e] = q.neval;
  }
}
}  // namespace

#define DISPATCH2(fn, thin, alpha, ...)                         \
  do {                                                          \
    if (thin) { if (alpha) fn<true, true>(__VA_ARGS__); else fn<true, false>(__VA_ARGS__); } \
    else { if (alpha) fn<false, true>(__VA_ARGS__); else fn<false, false>(__VA_ARGS__); }    \
  } while (0)

// grey_nodes_n (the breadth-first N-node form the delta kernel uses) against node_acc, node by
// node: both return acc = 0 + f_nu(nu_i) w_i for 6 nodes per parameter vector (two groups of 3).
template <bool THIN>
static void run_grey(long long n, const double* pars, double wavenorm, const double* wave, const double* weight,
                     double* out_single, double* out_group, int* safe) {
  FastNode nd[6];
  ModelP m{wavenorm, kUmToGHz / wavenorm, 0.0, 0.0};
  for (int i = 0; i < 6; ++i) {
    nd[i] = fast_node(wave[i], weight[i], wavenorm, THIN);
    if (nd[i].freq > m.nu_max) m.nu_max = nd[i].freq;
    if (nd[i].labs > m.lmax) m.lmax = nd[i].labs;
  }
  const double* tab = exp2_tab_default();
  for (long long e = 0; e < n; ++e) {
    const double* p = pars + 5 * e;
    FastSed s;
    fast_setup<THIN, false>(s, p[0], p[1], p[2], p[3], p[4], m);
    safe[e] = (s.status == ST_OK && s.safe) ? 1 : 0;
    if (!safe[e]) continue;
    for (int i = 0; i < 6; ++i)
      out_single[e * 6 + i] = node_acc<THIN, false, false, 0>(s, nd[i].freq, nd[i].lp, nd[i].weff, 0.0, tab);
    for (int g0 = 0; g0 < 6; g0 += 3) {
      double nu[3], lp[3], we[3], acc[3] = {0.0, 0.0, 0.0};
      for (int i = 0; i < 3; ++i) { nu[i] = nd[g0 + i].freq; lp[i] = nd[g0 + i].lp; we[i] = nd[g0 + i].weff; }
      grey_nodes_n<THIN, 3, 0>(s, nu, lp, we, acc, tab);
      for (int i = 0; i < 3; ++i) out_group[e * 6 + g0 + i] = acc[i];
    }
  }
}

static long long g_last_compressed = 0;

extern "C" {

struct EmuPriors {
  double lowlim[5], uplim[6], gmean[6], givar[6];
  unsigned char has_uplim[6], has_gprior[6];
};

void emu_loglike(int thin, int alpha, int fast, long long n, const double* pars, double wavenorm,
                 const EmuPriors* ep, int nb, const int* band_off, const double* wave, const double* weight,
                 const unsigned char* scalar_path, const double* flux, const double* ivar, const double* cinv,
                 long long wps, double* out, int* status) {
  Priors pr;
  for (int i = 0; i < 5; ++i) pr.lowlim[i] = ep->lowlim[i];
  for (int i = 0; i < 6; ++i) {
    pr.uplim[i] = ep->uplim[i]; pr.gmean[i] = ep->gmean[i]; pr.givar[i] = ep->givar[i];
    pr.has_uplim[i] = ep->has_uplim[i]; pr.has_gprior[i] = ep->has_gprior[i];
  }
  priors_finalize(pr);
  const int nn = band_off[nb];
  std::vector<double> freq(nn), lp(nn), weff(nn);
  ModelP m{wavenorm, kUmToGHz / wavenorm, 0.0, 0.0};
  for (int i = 0; i < nn; ++i) {       // same table construction as mbb_set_bands (mbb_capi.cu)
    const FastNode nd = fast_node(wave[i], weight[i], wavenorm, thin != 0);
    freq[i] = nd.freq; lp[i] = nd.lp; weff[i] = nd.weff;
    if (nd.freq > m.nu_max) m.nu_max = nd.freq;
    if (nd.labs > m.lmax) m.lmax = nd.labs;
  }
  TabView t{freq.data(), weight, weff.data(), lp.data(), band_off, scalar_path, nb};
  // fast: 1 = saturating node code, 2 = unclamped node code for `safe` walkers (the specialised
  // kernels' arithmetic), 3 = 2 + compressed Gauss rules (MBB_MATH_FAST_GAUSS)
  const int unclamped = fast >= 2;
  GaussTables gtab;
  if (fast == 3) gtab = build_gauss_tables(nb, band_off, wave, weight, scalar_path, wavenorm, thin != 0);
  const GaussTables* gt = fast == 3 ? &gtab : nullptr;
  long long ncomp = 0;
#define GO(T, A, F) run_loglike<T, A, F>(n, pars, m, pr, t, flux, ivar, cinv, wps, unclamped, out, status, gt, &ncomp)
  if (thin) {
    if (alpha) { if (fast) GO(true, true, true); else GO(true, true, false); }
    else { if (fast) GO(true, false, true); else GO(true, false, false); }
  } else {
    if (alpha) { if (fast) GO(false, true, true); else GO(false, true, false); }
    else { if (fast) GO(false, false, true); else GO(false, false, false); }
  }
#undef GO
  g_last_compressed = ncomp;
}

// number of (walker, band) pairs the last fast=3 call took through a compressed rule
long long emu_last_compressed() { return g_last_compressed; }

void emu_consts(int thin, int alpha, long long n, const double* pars, double wavenorm, int want_peak,
                double* out, int* status) {
  DISPATCH2(run_consts, thin, alpha, n, pars, wavenorm, want_peak, out, status);
}

void emu_lir(int thin, int alpha, long long n, const double* pars, double wavenorm, double fmin, double fmax,
             double prefac, double* out, int* status) {
  DISPATCH2(run_lir, thin, alpha, n, pars, wavenorm, fmin, fmax, prefac, out, status);
}

// table flavour of the emulated specialised node code (see g_ts), its size and its N/ln2
void emu_set_tab_bits(int bits) { g_ts = bits == 8 ? kTab256 : 0; }
int emu_tab_bits() { return g_ts ? 8 : 6; }
double emu_c64_hi() { return kC64Hi * (g_ts ? 4.0 : 1.0); }
}  // extern "C"

namespace {
template <int TS>
void run_fastmath(int mode, long long n, const double* x, double* out) {
  const double* tab = TS != 0 ? exp2_tab256_default() : exp2_tab_default();
  constexpr double cs = TabCfg<TS>::cscale;
  for (long long i = 0; i < n; ++i) {
    if (mode == 0) out[i] = exp_l(x[i]);
    else if (mode == 1) out[i] = expm1_l(x[i]);
    else if (mode == 2) out[i] = rcp_fast(x[i]);
    else if (mode == 4) out[i] = exp_red<TS, false>(red_prod(x[i], kC64Hi * cs, kC64Lo * cs), tab);
    else if (mode == 5) out[i] = expm1_red<TS, false>(red_prod(x[i], kC64Hi * cs, kC64Lo * cs), tab);
    else if (mode == 6)
      out[i] = one_minus_exp_red<TS, false>(red_neg_scaled(clamp_pos<hi700c<TS>()>(x[i] * (kC64Hi * cs))), tab);
    else if (mode == 7) {
      const long double b = 0.7L * ((long double)TabCfg<TS>::n / logl(2.0L));
      const double bh = (double)b, bl = (double)(b - (long double)bh);
      out[i] = exp_red<TS, false>(red_prod(x[i], bh, bl), tab);
    } else if (mode == 8) out[i] = rcp_cubic(x[i]);
    else if (mode == 9) out[i] = log_l(x[i]);
    else out[i] = div_fast(x[i], x[i] + 1.0);
  }
}
}  // namespace

extern "C" {
// element-wise checks of the lean math: mode 0 exp (CLAMP), 1 expm1 (CLAMP), 2 reciprocal,
// 3 a/b with b = x+1, 4 exp (no clamp), 5 expm1 (no clamp), 6 1-exp(-x) through the scaled path,
// 7 exp(x*y) as a product reduction with y = 0.7 (double-double of 0.7*N/ln2 formed here),
// 8 1/x by rcp_cubic, 9 log (table + series).  Modes 4-7 use the table flavour set by emu_set_tab_bits; 0-1 are the
// per-walker forms (always 64 entries).
void emu_fastmath(int mode, long long n, const double* x, double* out) {
  if (g_ts) run_fastmath<kTab256>(mode, n, x, out);
  else run_fastmath<0>(mode, n, x, out);
}

// chi-square of a full covariance from the staged matrix (explicit inverse, or the Cholesky form
// of mbb_set_data_chol: strictly lower triangle = L, diagonal = 1/L_rr) -- the device's quad_form
double emu_quad_form(const double* m, const double* diff, int nb, int chol) {
  std::vector<double> d(diff, diff + nb);
  return quad_form(m, d.data(), nb, chol != 0);
}

void emu_grey_nodes(int thin, long long n, const double* pars, double wavenorm, const double* wave,
                    const double* weight, double* out_single, double* out_group, int* safe) {
  if (thin) run_grey<true>(n, pars, wavenorm, wave, weight, out_single, out_group, safe);
  else run_grey<false>(n, pars, wavenorm, wave, weight, out_single, out_group, safe);
}

// n-point Gauss rule of the discrete measure {x, w}; returns 1 on success
int emu_gauss_rule(int N, const double* x, const double* w, int n, double* xs, double* ws) {
  std::vector<double> xv(x, x + N), wv(w, w + N), xo, wo;
  if (!discrete_gauss_rule(xv, wv, n, xo, wo)) return 0;
  for (int k = 0; k < n; ++k) { xs[k] = xo[k]; ws[k] = wo[k]; }
  return 1;
}

// floor(g / wps) by the multiply-shift constants the launcher computes (mbb_hostutil.h)
void emu_wps_division(long long n, const unsigned* g, unsigned wps, unsigned* q) {
  const WpsDivision d = wps_division(wps);
  for (long long i = 0; i < n; ++i) q[i] = d.mul ? wps_divide(d, g[i]) : g[i] / wps;
}

// FAST per-walker setup: xmerge, amp_grey, amp_pow, safe, status (thin = 0/1, alpha = 0/1)
void emu_fast_setup(int thin, int alpha, long long n, const double* pars, double wavenorm, double* out,
                    int* status) {
  ModelP m{wavenorm, kUmToGHz / wavenorm, kUmToGHz / 50.0, 3.0};
  for (long long e = 0; e < n; ++e) {
    const double* p = pars + 5 * e;
    FastSed s;
    if (thin) { if (alpha) fast_setup<true, true>(s, p[0], p[1], p[2], p[3], p[4], m); else fast_setup<true, false>(s, p[0], p[1], p[2], p[3], p[4], m); }
    else { if (alpha) fast_setup<false, true>(s, p[0], p[1], p[2], p[3], p[4], m); else fast_setup<false, false>(s, p[0], p[1], p[2], p[3], p[4], m); }
    out[4 * e] = s.xmerge; out[4 * e + 1] = s.amp_grey; out[4 * e + 2] = s.amp_pow; out[4 * e + 3] = s.safe;
    status[e] = s.status;
  }
}

void emu_qags(int thin, int alpha, long long n, const double* pars, double wavenorm, double fmin, double fmax,
              double prefac, double* out, int* neval, int* status) {
  DISPATCH2(run_qags, thin, alpha, n, pars, wavenorm, fmin, fmax, prefac, out, neval, status);
}

// QAGS on a few textbook integrands (kind), to compare with scipy.integrate.quad
void emu_qags_test(int kind, double a, double b, double epsabs, double epsrel, double* result, double* abserr,
                   int* neval, int* ier) {
  auto f = [kind](double x) -> double {
    switch (kind) {
      case 0: return sqrt(x);
      case 1: return log(x);
      case 2: return 1.0 / sqrt(x);
      case 3: return sqrt(fabs(x - 0.3));
      case 4: return cos(50.0 * x);
      case 5: return exp(-x * x);
      case 6: return fabs(x - 1.0 / 3.0);
      case 7: return 1.0 / (1.0 + 1000.0 * (x - 0.5) * (x - 0.5));
      default: return x;
    }
  };
  const QagsOut q = qagse(f, a, b, epsabs, epsrel);
  *result = q.result; *abserr = q.abserr; *neval = q.neval; *ier = q.ier;
}

void emu_philox(const unsigned* ctr, const unsigned* key, unsigned* out, double* u) {
  // the shipped form (round keys formed beforehand) and the textbook one must agree
  const Philox r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3],
                                 philox_keys(((unsigned long long)key[1] << 32) | key[0]));
  const Philox r2 = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = r.c[i] == r2.c[i] ? r.c[i] : ~r2.c[i];
  *u = u53(r.c[0], r.c[1]);
}

// the two uniforms the sampler cuts out of one block (bit-assembled on the device)
void emu_philox_uniforms(unsigned w0, unsigned w1, unsigned w3, double* uz, double* ua) {
  *uz = u52w(w0, w1 >> 12);
  *ua = u43(w3, w1 & 0x7ffu);
}

void emu_fnu(int thin, int alpha, long long n, const double* pars, double wavenorm, int nfreq,
             const double* freq, int scalar_path, int fast, double* out) {
  DISPATCH2(run_fnu, thin, alpha, n, pars, wavenorm, nfreq, freq, scalar_path, fast, out);
}
}
