// What FP64-pipe utilisation can exp-dominated code reach on this GPU?
// Each thread runs ILP independent chains of the lean exp of mbb_fastmath.cuh
// (x <- exp(-c x): product reduction, 4-step Horner, replicated shared-memory
// table lookup, integer scaling: 9 FP64 + 4 integer/LDS instructions per exp
// and link), with W warps resident per SM sub-partition.  Reported: exps per
// second and the FP64 pipe share they imply (2 pipe cycles per FP64 warp
// instruction).  This is the ceiling the likelihood kernels are measured against
// in DESIGN.md: same instruction mix, no memory traffic, no per-evaluation work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/exp_probe tools/exp_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../mbb_emcee_b200/csrc/mbb_fastmath.cuh"

using namespace mbb;

// Ablations of one exp link (MODE): 0 = the full lean exp; 1 = no table lookup / integer scaling
// (sT = 1: 10 FP64 instructions only); 2 = polynomial and final FMA only (6 FP64); 3 = reduction
// only (4 FP64 + the low-word extraction).
template <int MODE>
__device__ __forceinline__ double link(double x, double c_hi, double c_lo, const double* tab) {
  if (MODE == 0) return exp_red<kTabRepShift, false>(red_prod(x, c_hi, c_lo), tab);
  if (MODE == 1) {
    const Red r = red_prod(x, c_hi, c_lo);
    return fma(1.0, lean_p<0>(r.f), 1.0 + 1e-300 * r.k);
  }
  if (MODE == 2) {
    const double f = x * 0.25;
    return fma(0.5, lean_p<0>(f), 0.5);
  }
  const Red r = red_prod(x, c_hi, c_lo);
  return 0.6 + r.f + 1e-300 * r.k;
}

template <int ILP, int EXTRA, int MODE = 0>
__global__ void chains(double* out, int iters, double c_hi, double c_lo) {
  __shared__ __align__(16) double s_tab[kTabRepDoubles];
  const double* g = reinterpret_cast<const double*>(kExp2Tab_dev);
  for (int i = threadIdx.x; i < kTabRepDoubles; i += blockDim.x) s_tab[i] = g[i >> kTabRepShift];
  __syncthreads();
  const double* tab = s_tab + (threadIdx.x & 15);
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = 0.3 + 0.01 * i + 1e-6 * threadIdx.x;
  unsigned junk = threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      x[i] = link<MODE>(x[i], c_hi, c_lo, tab);
    }
  }
  double s = (junk == 0xdeadbeefu) ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 12345.678) out[0] = s;
}

template <int ILP, int EXTRA = 0, int MODE = 0>
double run(int warps_per_sm, int sms, int iters, double* d_out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = warps_per_sm * 32;
  const double c_hi = -kC64Hi, c_lo = -kC64Lo;           // x <- exp(-x): stays in (0.4, 0.8)
  chains<ILP, EXTRA, MODE><<<sms, threads>>>(d_out, 100, c_hi, c_lo);
  cudaEventRecord(e0);
  chains<ILP, EXTRA, MODE><<<sms, threads>>>(d_out, iters, c_hi, c_lo);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return (double)ILP * iters * threads * sms / (ms * 1e-3);      // exps per second
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* d_out; cudaMalloc(&d_out, 8);
  const int iters = 4000;
  const double clk = p.clockRate * 1e3;
  // FP64 warp instructions per exp in this loop: red_prod 4 + Horner 4 + g*f 1 + fma 1 = 10
  const double pipe_exps = sms * 4.0 * clk / (10 * 2.0) * 32.0;
  printf("{\"sms\": %d, \"clock_hz\": %.4g, \"fp64_per_exp\": 10, \"pipe_bound_exps_per_s\": %.4g, \"rows\": [", sms, clk, pipe_exps);
  const int wl[] = {2, 4, 6, 8};
  for (int i = 0; i < 4; ++i) {
    const int w = wl[i] * 4;
    const double r1 = run<1>(w, sms, iters, d_out), r2 = run<2>(w, sms, iters, d_out), r4 = run<4>(w, sms, iters, d_out);
    printf("%s{\"warps_per_smsp\": %d, \"ilp1\": %.4g, \"ilp2\": %.4g, \"ilp4\": %.4g, \"fp64_share\": [%.3f, %.3f, %.3f]}",
           i ? ", " : "", wl[i], r1, r2, r4, r1 / pipe_exps, r2 / pipe_exps, r4 / pipe_exps);
  }
  printf("]}\n");
  const double e0 = run<2, 0>(32, sms, iters, d_out);
  const double m1 = run<2, 0, 1>(32, sms, iters, d_out), m2 = run<2, 0, 2>(32, sms, iters, d_out),
               m3 = run<2, 0, 3>(32, sms, iters, d_out);
  printf("{\"ablation_cycles_per_link\": {\"full_exp\": %.2f, \"no_table_10_fp64\": %.2f, \"poly_only_6_fp64\": %.2f, "
         "\"reduction_only_4_fp64\": %.2f}}\n", sms * 4.0 * clk * 32.0 / e0, sms * 4.0 * clk * 32.0 / m1,
         sms * 4.0 * clk * 32.0 / m2, sms * 4.0 * clk * 32.0 / m3);
  return cudaGetLastError() != cudaSuccess;
}
