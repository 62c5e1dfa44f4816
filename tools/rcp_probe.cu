// Measures the seed accuracy of MUFU.RCP64H (PTX rcp.approx.ftz.f64) on the
// GPU at hand: max |1 - b*r0| over a dense sample of mantissas and a range of
// exponents, plus the resulting error of a * rcp_cubic(b) (mbb_fastmath.cuh)
// against IEEE division.  Justifies the single cubic correction step of rcp_cubic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/rcp_probe tools/rcp_probe.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../mbb_emcee_b200/csrc/mbb_fastmath.cuh"

__global__ void probe(double* out_seed, double* out_div, int n_per_thread) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  double worst_seed = 0.0, worst_div = 0.0;
  unsigned long long s = 0x9E3779B97F4A7C15ull * (tid + 1);
  for (int i = 0; i < n_per_thread; ++i) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const unsigned long long mant = s & 0x000fffffffffffffull;
    const int ex = 1023 + (int)((s >> 52) % 1400) - 700;        // 2^-700 .. 2^699
    const double b = __longlong_as_double(((unsigned long long)ex << 52) | mant);
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));
    const double e = fabs(fma(-b, r0, 1.0));
    if (e > worst_seed) worst_seed = e;
    const double a = 1.0 + (double)(s >> 40) * 1e-7;
    const double q = a * mbb::rcp_cubic(b), qt = a / b;
    const double de = fabs(q - qt) / fabs(qt);
    if (de > worst_div) worst_div = de;
  }
  out_seed[tid] = worst_seed;
  out_div[tid] = worst_div;
}

int main() {
  const int blocks = 296, threads = 256, n = blocks * threads;
  double *d_seed, *d_div;
  cudaMalloc(&d_seed, n * sizeof(double));
  cudaMalloc(&d_div, n * sizeof(double));
  probe<<<blocks, threads>>>(d_seed, d_div, 20000);
  double* h = (double*)malloc(2 * n * sizeof(double));
  cudaMemcpy(h, d_seed, n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaMemcpy(h + n, d_div, n * sizeof(double), cudaMemcpyDeviceToHost);
  if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  double ws = 0, wd = 0;
  for (int i = 0; i < n; ++i) { if (h[i] > ws) ws = h[i]; if (h[n + i] > wd) wd = h[n + i]; }
  printf("{\"samples\": %.3g, \"rcp64h_max_rel_err\": %.6g, \"rcp64h_log2\": %.3f, "
         "\"mul_rcp_cubic_max_rel_err\": %.6g, \"mul_rcp_cubic_ulps\": %.3f}\n",
         (double)n * 20000, ws, log2(ws), wd, wd / 1.1102230246251565e-16);
  return 0;
}
