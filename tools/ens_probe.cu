// Development probe for the source-resident sampler kernel (mbb_ensemble.cuh): one model
// variant on the 6 delta bands of BASELINE cfg5, 512 walkers per source, timed with CUDA events.
// Builds in seconds (two kernel instantiations), prints ms per iteration and a hash of the final
// ensembles / acceptance counts, so that a change can be checked for bit-identical chains and
// timed before the library is rebuilt.
//   tools/_build/ens_probe [thin 0|1] [alpha 0|1] [nsrc] [iterations] [stats 0|1] [nwalkers]
// One JSON object on stdout.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <random>
#include <algorithm>

#include "../mbb_emcee_b200/csrc/mbb_ensemble.cuh"

using namespace mbb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static unsigned long long fnv(const void* p, size_t bytes, unsigned long long h = 1469598103934665603ull) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

template <bool THIN, bool ALPHA>
static int run(int nsrc, int iters, int with_stats, int nw) {
  constexpr int NB = 6;
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
  pr.uplim[2] = 1500.0;
  for (int i = 0; i < 6; ++i) pr.givar[i] = 1.0;
  priors_finalize(pr);

  // photometry of a real greybody per source (so that the posterior is compact and the
  // acceptance fraction is the one of a fit), walkers in a ball around the truth
  std::mt19937_64 rng(12345);
  std::normal_distribution<double> g(0.0, 1.0);
  std::uniform_real_distribution<double> u(0.0, 1.0);
  std::vector<double> flux((size_t)nsrc * NB + 2, 0.0), ivar((size_t)nsrc * NB + 2, 0.0), P((size_t)n * 5);
  const double sig[5] = {1.0, 0.1, 20, 0.15, 2};
  for (int s = 0; s < nsrc; ++s) {
    const double T = 8 + 17 * u(rng), beta = 1.2 + 1.2 * u(rng), fn = exp(log(5.0) + log(20.0) * u(rng));
    const double truth[5] = {T, beta, 400, 3.0, fn};
    for (int b = 0; b < NB; ++b) {
      // thin greybody shape, normalised at 500 um (good enough to give the likelihood a peak)
      const double x = 14387.769 / (waves[b] * T), xn = 14387.769 / (wavenorm * T);
      const double f = fn * pow(wavenorm / waves[b], 3.0 + beta) * expm1(xn) / expm1(x);
      const double sg = std::max(0.1 * f, 0.05 * fn);
      flux[(size_t)s * NB + b] = f + sg * g(rng);
      ivar[(size_t)s * NB + b] = 1.0 / (sg * sg);
    }
    for (int w = 0; w < nw; ++w) {
      double* p = &P[((size_t)s * nw + w) * 5];
      for (int i = 0; i < 5; ++i) p[i] = std::max(truth[i] + sig[i] * g(rng), low[i] * 2 + 0.1);
    }
  }
  double *d_flux, *d_ivar, *d_P, *d_lnp, *d_stats, *d_scratch;
  int *d_st, *d_nacc;
  ColdArgs* d_cold;
  CK(cudaMalloc(&d_flux, flux.size() * 8));
  CK(cudaMalloc(&d_ivar, ivar.size() * 8));
  CK(cudaMalloc(&d_P, P.size() * 8));
  CK(cudaMalloc(&d_lnp, n * 8));
  CK(cudaMalloc(&d_st, n * 4));
  CK(cudaMalloc(&d_nacc, n * 4));
  CK(cudaMalloc(&d_stats, (size_t)nsrc * kFitStats * 8));
  CK(cudaMalloc(&d_cold, sizeof(ColdArgs)));
  CK(cudaMemcpy(d_flux, flux.data(), flux.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ivar, ivar.data(), ivar.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_P, P.data(), P.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemset(d_st, 0, n * 4));
  CK(cudaMemset(d_nacc, 0, n * 4));
  ColdArgs hc;
  hc.t = t; hc.pr = pr; hc.m = m;
  CK(cudaMemcpy(d_cold, &hc, sizeof(hc), cudaMemcpyHostToDevice));

  DataRef d;
  d.flux = d_flux; d.ivar = d_ivar; d.cinv = nullptr; d.chol = 0; d.nsrc = nsrc; d.nb = NB;
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  {  // starting log-probabilities
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.pars = d_P; a.out = d_lnp; a.status = d_st; a.n = n; a.wps = nw; a.layout = 0;
    set_wps_division(a);
    constexpr size_t kTabBytes = sizeof(double) * kTabRepDoubles;
    CK(cudaFuncSetAttribute(loglike_delta_kernel<THIN, ALPHA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)kTabBytes));
    const long long ntiles = (n + kDeltaTile - 1) / kDeltaTile;
    const unsigned grid = (unsigned)std::min<long long>(ntiles, (long long)sms * MBB_DELTA_MINB);
    loglike_delta_kernel<THIN, ALPHA, NB><<<grid, MBB_DELTA_BLOCK, kTabBytes>>>(a, m, pr, d, t, d_cold, 1, 1);
    CK(cudaDeviceSynchronize());
  }
  const int h = nw / 2;
  const int G = h <= kEnsThreads ? kEnsThreads / h : 1;
  const size_t smem = ens_resident_smem(G, nw, NB, with_stats != 0);
  auto kern = (h == kEnsThreads && G == 1) ? ens_resident_kernel<THIN, ALPHA, NB, true>
                                           : ens_resident_kernel<THIN, ALPHA, NB, false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEnsThreads, smem));
  const long long ngroups = (nsrc + G - 1) / G;
  const unsigned grid = (unsigned)std::min<long long>(ngroups, (long long)sms * per_sm);
  CK(cudaMalloc(&d_scratch, (size_t)grid * 5 * kEnsThreads * 8));
  EnsFit f;
  memset(&f, 0, sizeof(f));
  f.pos = d_P; f.lnp = d_lnp; f.nacc = d_nacc; f.status = d_st; f.stats = with_stats ? d_stats : nullptr;
  f.scratch = d_scratch;
  f.nsrc = nsrc; f.src0 = 0; f.dsrc0 = 0; f.chain_nsrc = nsrc; f.nw = nw; f.h = h; f.G = G;
  f.niter = iters; f.main_from = 0; f.main_done = 0; f.thin = 10; f.nrec = iters / 10; f.merge = 0;
  f.keys = philox_keys(7); f.step0 = 0; f.sc = stretch_scale(2.0);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  unsigned long long hash = 0;
  for (int r = 0; r < 3; ++r) {
    // every repetition restarts from the same state, so the hash is that of ONE launch
    CK(cudaMemcpy(d_P, P.data(), P.size() * 8, cudaMemcpyHostToDevice));
    {
      EvalArgs a;
      memset(&a, 0, sizeof(a));
      a.pars = d_P; a.out = d_lnp; a.status = d_st; a.n = n; a.wps = nw; a.layout = 0;
      set_wps_division(a);
      const long long ntiles = (n + kDeltaTile - 1) / kDeltaTile;
      const unsigned g2 = (unsigned)std::min<long long>(ntiles, (long long)sms * MBB_DELTA_MINB);
      loglike_delta_kernel<THIN, ALPHA, NB><<<g2, MBB_DELTA_BLOCK, sizeof(double) * kTabRepDoubles>>>(a, m, pr, d, t,
                                                                                                    d_cold, 1, 1);
    }
    CK(cudaMemset(d_nacc, 0, n * 4));
    CK(cudaEventRecord(e0));
    kern<<<grid, kEnsThreads, smem>>>(f, m, pr, d, t, d_cold);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float x;
    CK(cudaEventElapsedTime(&x, e0, e1));
    best = std::min(best, x);
  }
  std::vector<double> o(P.size()), l(n), s((size_t)nsrc * kFitStats);
  std::vector<int> na(n);
  CK(cudaMemcpy(o.data(), d_P, P.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(l.data(), d_lnp, n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(na.data(), d_nacc, n * 4, cudaMemcpyDeviceToHost));
  hash = fnv(o.data(), o.size() * 8);
  hash = fnv(l.data(), l.size() * 8, hash);
  hash = fnv(na.data(), na.size() * 4, hash);
  unsigned long long shash = 0;
  if (with_stats) {
    CK(cudaMemcpy(s.data(), d_stats, s.size() * 8, cudaMemcpyDeviceToHost));
    shash = fnv(s.data(), s.size() * 8);
  }
  double acc = 0;
  for (long long i = 0; i < n; ++i) acc += na[i];
  long long nonfinite = 0;
  for (long long i = 0; i < n; ++i) nonfinite += !std::isfinite(l[i]);
  printf("{\"kernel\": \"ens_resident_kernel<%d,%d,6>\", \"nsrc\": %d, \"nwalkers\": %d, \"iterations\": %d, \"stats\": %d, "
         "\"ctas_per_sm\": %d, \"smem\": %zu, \"ms\": %.4f, \"ms_per_iteration\": %.4f, \"proposals_per_s\": %.4e, "
         "\"acceptance\": %.4f, \"nonfinite_lnp\": %lld, \"hash\": \"%016llx\", \"stats_hash\": \"%016llx\"}\n",
         (int)THIN, (int)ALPHA, nsrc, nw, iters, with_stats, per_sm, smem, best, best / iters,
         (double)n * iters / (best * 1e-3), acc / ((double)n * iters), nonfinite, hash, shash);
  return 0;
}

int main(int argc, char** argv) {
  const int thin = argc > 1 ? atoi(argv[1]) : 1;
  const int alpha = argc > 2 ? atoi(argv[2]) : 0;
  const int nsrc = argc > 3 ? atoi(argv[3]) : 20000;
  const int iters = argc > 4 ? atoi(argv[4]) : 20;
  const int st = argc > 5 ? atoi(argv[5]) : 1;
  const int nw = argc > 6 ? atoi(argv[6]) : 512;
  if (thin && !alpha) return run<true, false>(nsrc, iters, st, nw);
  if (!thin && alpha) return run<false, true>(nsrc, iters, st, nw);
  if (thin && alpha) return run<true, true>(nsrc, iters, st, nw);
  return run<false, false>(nsrc, iters, st, nw);
}
