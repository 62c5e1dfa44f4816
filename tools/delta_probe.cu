// Development probe for the delta-band likelihood kernel: one model variant on the 6 delta
// bands of BASELINE cfg1/cfg5, timed with CUDA events and compared against the FAITHFUL
// thread-per-evaluation kernel (reference formulas, libdevice) on the same inputs.
// Builds in seconds (two kernel instantiations instead of the whole library), so SASS and
// timing of a change can be looked at before the library is rebuilt.
//   tools/_build/delta_probe [thin 0|1] [alpha 0|1] [nsrc] [reps]
// One JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include <algorithm>

#include "../mbb_emcee_b200/csrc/mbb_kernels.cuh"

using namespace mbb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

template <bool THIN, bool ALPHA>
static int run(int nsrc, int reps) {
  constexpr int NB = 6;
  const int nw = 512;
  const long long n = (long long)nsrc * nw;
  const double waves[NB] = {70, 100, 160, 250, 350, 500};
  const double wavenorm = 500.0;
  SmallTab t;
  memset(&t, 0, sizeof(t));
  t.nb = NB;
  ModelP m;
  m.wavenorm = wavenorm;
  m.nu_norm = kUmToGHz / wavenorm;
  m.nu_max = 0;
  m.lmax = 0;
  for (int i = 0; i < NB; ++i) {
    const FastNode nd = fast_node(waves[i], 1.0, wavenorm, THIN);
    t.freq[i] = nd.freq; t.w[i] = 1.0; t.weff[i] = nd.weff; t.lp[i] = nd.lp;
    t.band_off[i] = i;
    t.scalar_path[i] = 0;
    m.nu_max = std::max(m.nu_max, nd.freq);
    m.lmax = std::max(m.lmax, nd.labs);
  }
  t.band_off[NB] = NB;
  Priors pr;
  memset(&pr, 0, sizeof(pr));
  const double low[5] = {1, 0.1, 1, 0.1, 1e-3};
  for (int i = 0; i < 5; ++i) pr.lowlim[i] = low[i];
  pr.has_uplim[2] = 1;
  pr.uplim[2] = 1500.0;          // likelihood.set_phot: 3 x the longest wavelength
  for (int i = 0; i < 6; ++i) pr.givar[i] = 1.0;
  priors_finalize(pr);

  std::mt19937_64 rng(12345);
  std::normal_distribution<double> g(0.0, 1.0);
  std::uniform_real_distribution<double> u(0.0, 1.0);
  std::vector<double> flux((size_t)nsrc * NB + 2, 0.0), ivar((size_t)nsrc * NB + 2, 0.0), P((size_t)n * 5);
  const double sig[5] = {2, 0.2, 100, 0.3, 5};
  for (int s = 0; s < nsrc; ++s) {
    const double truth[5] = {8 + 17 * u(rng), 1.2 + 1.2 * u(rng), 400, 3.0, exp(log(5.0) + log(20.0) * u(rng))};
    for (int b = 0; b < NB; ++b) {
      const double f = 10 + 50 * u(rng);
      flux[(size_t)s * NB + b] = f;
      const double sg = std::max(0.1 * f, 1.0);
      ivar[(size_t)s * NB + b] = 1.0 / (sg * sg);
    }
    for (int w = 0; w < nw; ++w) {
      double* p = &P[((size_t)s * nw + w) * 5];
      for (int i = 0; i < 5; ++i) p[i] = std::max(truth[i] + sig[i] * g(rng), low[i] * 2 + 0.1);
    }
  }
  double *d_flux, *d_ivar, *d_P, *d_out, *d_ref;
  int *d_st, *d_st2;
  ColdArgs* d_cold;
  CK(cudaMalloc(&d_flux, flux.size() * 8));
  CK(cudaMalloc(&d_ivar, ivar.size() * 8));
  CK(cudaMalloc(&d_P, P.size() * 8));
  CK(cudaMalloc(&d_out, n * 8));
  CK(cudaMalloc(&d_ref, n * 8));
  CK(cudaMalloc(&d_st, n * 4));
  CK(cudaMalloc(&d_st2, n * 4));
  CK(cudaMalloc(&d_cold, sizeof(ColdArgs)));
  CK(cudaMemcpy(d_flux, flux.data(), flux.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ivar, ivar.data(), ivar.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_P, P.data(), P.size() * 8, cudaMemcpyHostToDevice));
  ColdArgs h;
  h.t = t; h.pr = pr; h.m = m;
  CK(cudaMemcpy(d_cold, &h, sizeof(h), cudaMemcpyHostToDevice));

  EvalArgs a;
  memset(&a, 0, sizeof(a));
  a.pars = d_P; a.out = d_out; a.status = d_st; a.n = n; a.wps = nw; a.layout = 0;
  set_wps_division(a);
  DataRef d;
  d.flux = d_flux; d.ivar = d_ivar; d.cinv = nullptr; d.chol = 0; d.nsrc = nsrc; d.nb = NB;

  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  constexpr size_t kTabBytes = sizeof(double) * kTabRepDoubles;
  CK(cudaFuncSetAttribute(loglike_delta_kernel<THIN, ALPHA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)kTabBytes));
  const long long ntiles = (n + kDeltaTile - 1) / kDeltaTile;
  const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)sms * MBB_DELTA_MINB);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  std::vector<float> ms;
  for (int r = 0; r < reps + 3; ++r) {
    CK(cudaEventRecord(e0));
    loglike_delta_kernel<THIN, ALPHA, NB><<<grid, MBB_DELTA_BLOCK, kTabBytes>>>(a, m, pr, d, t, d_cold, 1, 1);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float x;
    CK(cudaEventElapsedTime(&x, e0, e1));
    if (r >= 3) ms.push_back(x);
  }
  CK(cudaGetLastError());
  std::sort(ms.begin(), ms.end());
  // reference formulation
  EvalArgs b = a;
  b.out = d_ref; b.status = d_st2;
  loglike_thread_kernel<THIN, ALPHA, false><<<(unsigned)((n + 255) / 256), 256>>>(b, m, pr, d, t);
  CK(cudaDeviceSynchronize());
  std::vector<double> o(n), r(n);
  std::vector<int> st(n);
  CK(cudaMemcpy(o.data(), d_out, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(r.data(), d_ref, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st.data(), d_st, n * 4, cudaMemcpyDeviceToHost));
  double worst = 0;
  long long bad = 0, nerr = 0, worst_i = -1;
  for (long long i = 0; i < n; ++i) {
    if (st[i] > 1) ++nerr;
    if (std::isinf(r[i]) || std::isnan(r[i])) { if (!(o[i] == r[i]) && !(std::isnan(o[i]) && std::isnan(r[i]))) ++bad; continue; }
    const double e = fabs(o[i] - r[i]) / fabs(r[i]);
    if (e > worst) { worst = e; worst_i = i; }
  }
  const double med = ms[ms.size() / 2];
  printf("{\"kernel\": \"loglike_delta_kernel<%d,%d,6>\", \"evals\": %lld, \"ms_median\": %.4f, \"ms_min\": %.4f, "
         "\"evals_per_s\": %.4e, \"max_rel_vs_faithful\": %.3e, \"mismatched_nonfinite\": %lld, \"status_errors\": %lld}\n",
         (int)THIN, (int)ALPHA, n, med, ms.front(), n / (med * 1e-3), worst, bad, nerr);
  if (worst_i >= 0 && worst > 1e-12) {
    const double* p = &P[(size_t)worst_i * 5];
    fprintf(stderr, "worst at %lld: T=%.17g beta=%.17g l0=%.17g alpha=%.17g fn=%.17g got %.17g want %.17g\n", worst_i,
            p[0], p[1], p[2], p[3], p[4], o[worst_i], r[worst_i]);
  }
  return worst < 1e-12 && bad == 0 ? 0 : 2;
}

int main(int argc, char** argv) {
  const int thin = argc > 1 ? atoi(argv[1]) : 1;
  const int alpha = argc > 2 ? atoi(argv[2]) : 0;
  const int nsrc = argc > 3 ? atoi(argv[3]) : 20000;
  const int reps = argc > 4 ? atoi(argv[4]) : 20;
  if (thin && !alpha) return run<true, false>(nsrc, reps);
  if (!thin && alpha) return run<false, true>(nsrc, reps);
  if (thin && alpha) return run<true, true>(nsrc, reps);
  return run<false, false>(nsrc, reps);
}
