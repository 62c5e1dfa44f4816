// Chain statistics on the device: count, sum and exact order statistics of the columns of a
// row-major [n][ncols] block, optionally clipped to [lowlim, uplim] per column -- what
// mbb_results._parcen_internal / par_lowlim / par_uplim get from numpy.mean and
// numpy.percentile over 10^7-sample chains (reference results.py:314-431, 946-985).
//
// Order statistics by MSD radix selection on the order-preserving 64-bit image of a double:
// eight passes of one byte; a pass histograms the byte below the prefix each wanted rank has
// fixed so far (<= kStatMaxSel ranks per column, ranks with the same prefix share a histogram),
// the host picks the bucket holding the rank and descends.  A column is a grid row
// (blockIdx.y) whose threads stride the block's rows; the columns' blocks walk the same rows at
// the same time, so a pass reads the block from DRAM about once (the rest are L2 hits).
// Histograms live in shared memory and same-bucket hits of a warp are merged with
// __match_any_sync before the atomic, because chain values share their leading bytes.
// HBM-bound: 9 sweeps x 8 B x n x ncols (10^7 x 5: 3.6 GB, ~0.6 ms at the measured 6.5 TB/s).
#pragma once
#include <cuda_runtime.h>

namespace mbb {

constexpr int kStatMaxCols = 8;
constexpr int kStatMaxSel = 8;        // wanted ranks per column (4 quantiles x 2 neighbours)

__device__ __forceinline__ unsigned long long stat_key(double x) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);      // ascending doubles -> ascending keys
}
inline double stat_unkey(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  double x;
  memcpy(&x, &u, 8);
  return x;
}

struct StatCols {
  double lo[kStatMaxCols], hi[kStatMaxCols];     // keep lo <= x <= hi
  int ncols;
};

// per column (blockIdx.y): count of kept values, of NaNs, and one partial sum per block (summed
// on the host in extended precision).  Threads stride the rows; the five columns' blocks walk
// the same rows at the same time, so the block is read from DRAM about once (L2).
__global__ void __launch_bounds__(256)
chain_moments_kernel(const double* __restrict__ x, long long total, StatCols sc,
                     unsigned long long* __restrict__ count, unsigned long long* __restrict__ nnan,
                     double* __restrict__ partial /*[ncols][gridDim.x]*/) {
  __shared__ double s_sum[8];
  __shared__ unsigned s_cnt[8], s_nan[8];
  const int col = blockIdx.y, nc = sc.ncols;
  const long long n = total / nc;
  const double lo = sc.lo[col], hi = sc.hi[col];
  double sum = 0.0, comp = 0.0;
  unsigned cnt = 0, nn = 0;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
    const double v = __ldg(x + r * nc + col);
    if (v != v) { ++nn; continue; }
    if (v < lo || v > hi) continue;
    ++cnt;
    const double y = v - comp;          // Kahan within the thread's run of ~n/(grid*256) terms
    const double t = sum + y;
    comp = (t - sum) - y;
    sum = t;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {    // fixed tree: the result does not depend on scheduling
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { s_sum[w] = sum; s_cnt[w] = cnt; s_nan[w] = nn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double acc = 0.0;
    unsigned c = 0, m = 0;
    for (int i = 0; i < 8; ++i) { acc += s_sum[i]; c += s_cnt[i]; m += s_nan[i]; }
    partial[(long long)col * gridDim.x + blockIdx.x] = acc;
    if (c) atomicAdd(&count[col], (unsigned long long)c);
    if (m) atomicAdd(&nnan[col], (unsigned long long)m);
  }
}

// one radix pass: group g of column c histograms byte (key >> shift) & 255 of the kept values
// whose bits above that byte equal prefix[c][g]
struct StatPass {
  unsigned long long prefix[kStatMaxCols][kStatMaxSel];
  int ngroups[kStatMaxCols];
  int shift;                           // 56, 48, ..., 0
};

__global__ void __launch_bounds__(256)
chain_select_hist_kernel(const double* __restrict__ x, long long total, StatCols sc, StatPass ps,
                         unsigned* __restrict__ hist /*[ncols][kStatMaxSel][256]*/) {
  __shared__ unsigned s_h[kStatMaxSel * 256];      // this block's column (blockIdx.y): <= 8 groups
  const int col = blockIdx.y;
  const int nc = sc.ncols;
  const int ng = ps.ngroups[col];
  for (int i = threadIdx.x; i < ng * 256; i += blockDim.x) s_h[i] = 0;
  __syncthreads();
  const double lo = sc.lo[col], hi = sc.hi[col];
  const long long n = total / nc;
  const int shift = ps.shift;
  const unsigned lane_mask = 0xffffffffu;
  for (long long r0 = (long long)blockIdx.x * blockDim.x; r0 < n; r0 += (long long)gridDim.x * blockDim.x) {
    const long long r = r0 + threadIdx.x;
    bool keep = false;
    unsigned long long key = 0;
    if (r < n) {
      const double v = __ldg(x + r * nc + col);
      keep = (v == v) && !(v < lo) && !(v > hi);
      key = stat_key(v);
    }
    const unsigned long long above = shift == 56 ? 0ull : (key >> (shift + 8));
    const int byte = (int)((key >> shift) & 255ull);
    for (int g = 0; g < ng; ++g) {
      const bool hit = keep && (shift == 56 || above == ps.prefix[col][g]);
      const int bin = hit ? byte : -1;
      const unsigned peers = __match_any_sync(lane_mask, bin);
      if (hit && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&s_h[g * 256 + bin], __popc(peers));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ng * 256; i += blockDim.x)
    if (s_h[i]) atomicAdd(&hist[(col * kStatMaxSel) * 256 + i], s_h[i]);
}

}  // namespace mbb
