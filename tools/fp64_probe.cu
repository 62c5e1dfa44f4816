// FP64 pipe characteristics of the GPU at hand (feeds the scheduling decisions
// documented in DESIGN.md): dependent-issue latency of DFMA, and DFMA throughput
// per SM as a function of resident warps per SM sub-partition and of the
// number of independent chains per thread (ILP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/fp64_probe tools/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cycles, int iters, double m, double c) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = 1.0 + i + threadIdx.x * 1e-9;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int ILP>
double run(int warps_per_sm, int sms, int iters, long long* d_cyc, double* d_out, long long* cyc_out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = warps_per_sm * 32;       // one CTA per SM
  chain<ILP><<<sms, threads>>>(d_out, d_cyc, 100, 0.999999, 1e-9);
  cudaEventRecord(e0);
  chain<ILP><<<sms, threads>>>(d_out, d_cyc, iters, 0.999999, 1e-9);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(cyc_out, d_cyc, sizeof(long long), cudaMemcpyDeviceToHost);
  return 2.0 * ILP * (double)iters * threads * sms / (ms * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  long long* d_cyc; double* d_out; long long cyc;
  cudaMalloc(&d_cyc, 8); cudaMalloc(&d_out, 8);
  const int iters = 20000;
  printf("{\"sms\": %d", sms);
  run<1>(1, 1, iters, d_cyc, d_out, &cyc);
  printf(", \"dfma_dependent_latency_cycles\": %.2f", (double)cyc / iters);
  run<2>(1, 1, iters, d_cyc, d_out, &cyc);
  printf(", \"dfma_issue_interval_one_warp_ilp2\": %.2f", (double)cyc / (2.0 * iters));
  run<8>(1, 1, iters, d_cyc, d_out, &cyc);
  printf(", \"dfma_issue_interval_one_warp_ilp8\": %.2f", (double)cyc / (8.0 * iters));
  printf(", \"tflops_by_warps_per_smsp\": {");
  const int wl[] = {1, 2, 4, 6, 8, 12, 16};
  for (int i = 0; i < 7; ++i) {
    const int w = wl[i] * 4;
    if (w * 32 > 1024) { // two CTAs worth: use ILP to emulate? skip beyond 1024 threads per CTA
      continue;
    }
    double t1 = run<1>(w, sms, iters, d_cyc, d_out, &cyc);
    double t2 = run<2>(w, sms, iters, d_cyc, d_out, &cyc);
    double t4 = run<4>(w, sms, iters, d_cyc, d_out, &cyc);
    printf("%s\"%d\": {\"ilp1\": %.2f, \"ilp2\": %.2f, \"ilp4\": %.2f}", i ? ", " : "", wl[i], t1, t2, t4);
  }
  printf("}}\n");
  return cudaGetLastError() != cudaSuccess;
}
