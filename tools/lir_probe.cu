// Development probe for the L_IR kernels (QUADPACK replay / fixed rule) on a synthetic set of
// unique chain samples: time per launch and a hash of the results.
//   tools/_build/lir_probe [nsamples] [reps] [dump.bin]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <random>
#include <algorithm>

#include "../mbb_emcee_b200/csrc/mbb_kernels.cuh"

using namespace mbb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static unsigned long long fnv(const void* p, size_t bytes) {
  unsigned long long h = 1469598103934665603ull;
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < bytes; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000000;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  // random-walk samples around the cfg4 truth, walker-major like a chain's work list
  std::mt19937_64 rng(7);
  std::normal_distribution<double> g(0.0, 1.0);
  std::vector<double> P((size_t)n * 5);
  const double truth[5] = {14.0, 1.8, 400.0, 3.0, 30.0}, step[5] = {0.15, 0.02, 8.0, 0.03, 0.4};
  const double lo[5] = {3, 0.3, 30, 0.6, 1};
  double cur[5];
  for (int i = 0; i < n; ++i) {
    if (i % 7000 == 0) for (int k = 0; k < 5; ++k) cur[k] = truth[k] * (1.0 + 0.2 * g(rng));
    for (int k = 0; k < 5; ++k) { cur[k] += step[k] * g(rng); if (cur[k] < lo[k]) cur[k] = lo[k]; }
    for (int k = 0; k < 5; ++k) P[(size_t)i * 5 + k] = cur[k];
  }
  std::vector<int> work(n);
  for (int i = 0; i < n; ++i) work[i] = i;
  double *d_P, *d_out;
  int *d_work, *d_st;
  unsigned* d_n;
  CK(cudaMalloc(&d_P, P.size() * 8));
  CK(cudaMalloc(&d_out, (size_t)n * 8));
  CK(cudaMalloc(&d_work, (size_t)n * 4));
  CK(cudaMalloc(&d_st, (size_t)n * 4));
  CK(cudaMalloc(&d_n, 4));
  CK(cudaMemcpy(d_P, P.data(), P.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_work, work.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  const unsigned un = (unsigned)n;
  CK(cudaMemcpy(d_n, &un, 4, cudaMemcpyHostToDevice));
  const double opz = 3.0, fmin = kUmToGHz / (1000.0 * opz), fmax = kUmToGHz / (8.0 * opz);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    chain_lir_qags_kernel<false, true><<<(unsigned)((n + 127) / 128), 128>>>(d_P, d_work, d_n, 500.0, fmin, fmax, 1.0,
                                                                             d_out, d_st);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float x;
    CK(cudaEventElapsedTime(&x, e0, e1));
    best = std::min(best, x);
  }
  float best_g = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    chain_lir_kernel<false, true><<<(unsigned)((n + 127) / 128), 128>>>(d_P, d_work, d_n, 500.0, fmin, fmax, 1.0, d_out,
                                                                        d_st);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float x;
    CK(cudaEventElapsedTime(&x, e0, e1));
    best_g = std::min(best_g, x);
  }
  {
    std::vector<double> og(n);
    CK(cudaMemcpy(og.data(), d_out, (size_t)n * 8, cudaMemcpyDeviceToHost));
    printf("{\"kernel\": \"chain_lir_kernel<0,1>\", \"samples\": %d, \"ms\": %.3f, \"hash\": \"%016llx\"}\n", n, best_g,
           fnv(og.data(), (size_t)n * 8));
  }
  chain_lir_qags_kernel<false, true><<<(unsigned)((n + 127) / 128), 128>>>(d_P, d_work, d_n, 500.0, fmin, fmax, 1.0,
                                                                           d_out, d_st);
  CK(cudaDeviceSynchronize());
  std::vector<double> o(n);
  std::vector<int> st(n);
  CK(cudaMemcpy(o.data(), d_out, (size_t)n * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(st.data(), d_st, (size_t)n * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int i = 0; i < n; ++i) bad += (st[i] != 0) || !std::isfinite(o[i]);
  printf("{\"kernel\": \"chain_lir_qags_kernel<0,1>\", \"samples\": %d, \"ms\": %.3f, \"samples_per_s\": %.4e, "
         "\"bad\": %lld, \"hash\": \"%016llx\"}\n", n, best, n / (best * 1e-3), bad, fnv(o.data(), (size_t)n * 8));
  if (argc > 3) {                       // dump the results for a comparison between builds
    FILE* fh = fopen(argv[3], "wb");
    if (fh) { fwrite(o.data(), 8, (size_t)n, fh); fclose(fh); }
    if (argc > 4) {                     // and the parameter vectors
      fh = fopen(argv[4], "wb");
      if (fh) { fwrite(P.data(), 8, P.size(), fh); fclose(fh); }
    }
  }
  return 0;
}
