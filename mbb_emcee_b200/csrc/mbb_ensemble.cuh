// Device-resident affine-invariant ensemble sampler for many sources at once
// (BASELINE configs[4]).  The stretch move of Goodman &
// Weare (2010) with the scheduling of emcee 2.2 (two half-ensembles per
// iteration; reference mbb_fit.py:80-81, 525-542 drives exactly this through
// emcee), with the proposals, the log-probability, the accept/reject step AND
// the posterior summaries on the GPU: walker positions never cross PCIe, a fit
// returns per-source statistics (what results.py:314-431 derives from a chain)
// and, optionally, a thinned chain.
//
// Two device forms, bit-identical in what they compute:
//   ens_resident_kernel  delta-band FAST configurations.  A CTA owns whole
//       sources: their ensembles (positions, log-probabilities, photometry) are
//       loaded into shared memory once, ALL iterations of the launch run there
//       -- the partner gather of the stretch move is a shared-memory read, one
//       CTA barrier per half-step, no HBM traffic inside the loop -- and the
//       per-thread posterior accumulators are reduced once at the end.
//   ens_propose_kernel + likelihood kernels + ens_accept_kernel (+
//       ens_stats_kernel per recorded iteration)  any band set / math mode.
//
// Randomness is counter-based (Philox4x32-10, Salmon et al. 2011): the draws
// of walker `k` of the updating half of GLOBAL source `s` at half-step `hs`
// depend only on (seed, s*nw/2 + k, hs), so results are independent of launch
// geometry and of how sources are sharded over GPUs, and can be replayed on
// the host (tests/test_ensemble_gpu.py does exactly that with a numpy Philox
// and the CPU oracle).  Statistically equivalent to, not bit-identical with,
// emcee's Mersenne-Twister stream.  One Philox block per proposal: 52 bits for
// the stretch factor, 32 for the partner, 43 for the acceptance uniform.
#pragma once
#include "mbb_kernels.cuh"
#include "mbb_philox.cuh"

namespace mbb {

// ---- per-source posterior summary row (doubles) -----------------------------
// over the recorded samples (every thin-th iteration of the main run, all walkers)
constexpr int kFitStats = 28;
enum FitStat {
  FS_N = 0,         // number of samples
  FS_MEAN = 1,      // [5] mean
  FS_M2 = 6,        // [5] sum of squared deviations from the mean (variance = M2/(N-1))
  FS_MIN = 11,      // [5]
  FS_MAX = 16,      // [5]
  FS_BESTLNP = 21,  // largest log-probability among the samples
  FS_BEST = 22,     // [5] the sample that has it (first in (thread, time) order on ties)
  FS_ACC = 27       // mean acceptance fraction of the main run so far
};

struct Draw {
  double z;      // stretch factor ((a-1)u+1)^2/a
  double u;      // the acceptance uniform
  int partner;   // index into the complementary half
};

// emcee 2.2 accepts where (dim-1) ln z + lnp(q) - lnp(s) > ln u.  The same test
// without logarithms:  u < z^4 exp(lnp(q) - lnp(s))  -- one lean exp instead of two
// libdevice logs.  The difference is clamped to [-745, 700] first: -inf (proposal
// below a lower limit) and NaN reject, +inf accepts, exactly as the logarithmic form
// does (tests/test_ensemble_gpu.py checks the two forms against each other).
__device__ __forceinline__ bool stretch_accept(double z, double u, double newlnp, double oldlnp) {
  const double dl = fmin(fmax(newlnp - oldlnp, -745.0), 700.0);
  const double z2 = z * z;
  return u < (z2 * z2) * exp_l(dl);
}

// One Philox block -> (z, partner, u).  emcee: zz = ((a - 1.) * rand + 1) ** 2. / a with
// separate roundings; for a power of two (the default a = 2) the division is an exact
// multiplication by 1/a.
struct StretchScale {
  double a, am1, inv_a;
  int pow2;
};
inline StretchScale stretch_scale(double a) {
  StretchScale s;
  s.a = a;
  s.am1 = a - 1.0;
  s.inv_a = 1.0 / a;
  int e;
  s.pow2 = frexp(a, &e) == 0.5;
  return s;
}

__device__ __forceinline__ Draw stretch_draw(const PhiloxKeys& keys, unsigned long long widx,
                                             unsigned long long hstep, const StretchScale& sc, int ncomp) {
  const Philox r = philox4x32_10((unsigned)widx, (unsigned)(widx >> 32), (unsigned)hstep,
                                 (unsigned)(hstep >> 32), keys);
  Draw d;
  const double uz = u52w(r.c[0], r.c[1] >> 12);
  const double t = __dadd_rn(__dmul_rn(sc.am1, uz), 1.0);
  const double tt = __dmul_rn(t, t);
  d.z = sc.pow2 ? __dmul_rn(tt, sc.inv_a) : __ddiv_rn(tt, sc.a);
  d.partner = (int)__umulhi(r.c[2], (unsigned)ncomp);
  d.u = u43(r.c[3], r.c[1] & 0x7ffu);
  return d;
}

struct EnsArgs {
  double* pos;              // [nsrc][nw][5]
  double* lnp;              // [nsrc][nw]
  int* nacc;                // [nsrc][nw]
  int* status;              // [nsrc][nw], first non-trivial status seen (sticky)
  double* q;                // [nsrc*h][5] proposal scratch
  double* qlnp;             // [nsrc*h]
  int* qst;                 // [nsrc*h]
  long long nsrc, src0;     // src0: global index of source 0 (sharded runs), enters the RNG counter
  int nw, h, half;          // h = nw/2; half 0 updates walkers [0,h) against [h,nw)
  int count;                // != 0: accepted moves are counted (main run)
  unsigned long long hstep;   // hstep = 2*iteration + half
  PhiloxKeys keys;            // philox_keys(seed)
  StretchScale sc;
};

// q = c[j] - z (c[j] - s), in emcee's operation order
__global__ void __launch_bounds__(256) ens_propose_kernel(const EnsArgs g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nsrc * g.h) return;
  const long long src = i / g.h;
  const int k = (int)(i - src * g.h);
  const Draw d = stretch_draw(g.keys, (unsigned long long)((g.src0 + src) * g.h + k), g.hstep, g.sc, g.h);
  const int own = g.half == 0 ? k : g.h + k;
  const int oth = (g.half == 0 ? g.h : 0) + d.partner;
  const double* s = g.pos + (src * g.nw + own) * 5;
  const double* c = g.pos + (src * g.nw + oth) * 5;
  double* q = g.q + i * 5;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const double cj = c[j];
    q[j] = __dsub_rn(cj, __dmul_rn(d.z, __dsub_rn(cj, s[j])));
  }
}

__global__ void __launch_bounds__(256) ens_accept_kernel(const EnsArgs g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nsrc * g.h) return;
  const long long src = i / g.h;
  const int k = (int)(i - src * g.h);
  const Draw d = stretch_draw(g.keys, (unsigned long long)((g.src0 + src) * g.h + k), g.hstep, g.sc, g.h);
  const int own = g.half == 0 ? k : g.h + k;
  const long long w = src * g.nw + own;
  const double newlnp = g.qlnp[i];
  const int st = g.qst[i];
  if (st > ST_BELOW_LOWLIM) {          // the reference would have raised here
    if (g.status[w] <= ST_BELOW_LOWLIM) g.status[w] = st;
    return;
  }
  if (stretch_accept(d.z, d.u, newlnp, g.lnp[w])) {
    const double* q = g.q + i * 5;
    double* p = g.pos + w * 5;
#pragma unroll
    for (int j = 0; j < 5; ++j) p[j] = q[j];
    g.lnp[w] = newlnp;
    if (g.count) g.nacc[w] += 1;
  }
}

// Tabulated passbands, small ensembles (a single source's 125-walker half-ensemble): the half-step
// in TWO launches instead of four -- the proposal is formed by the thread that then does the
// per-walker setup of the warp path, and the CTA that has summed an evaluation's nodes
// (nodes_small_eval) accepts or rejects it on the spot.  Same draws, same arithmetic: bit-identical
// with ens_propose_kernel / loglike_setup_kernel / loglike_nodes_small_kernel / ens_accept_kernel.
template <bool THIN, bool ALPHA>
__global__ void __launch_bounds__(128)
ens_propose_setup_kernel(const EnsArgs g, const ModelP m, const Priors pr, double* __restrict__ scratch,
                         int* __restrict__ sst, const BandMeta* __restrict__ band_meta, const int nb_gauss,
                         const double2* __restrict__ node_a, const int* __restrict__ band_off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nsrc * g.h) return;
  const long long src = i / g.h;
  const int k = (int)(i - src * g.h);
  const Draw d = stretch_draw(g.keys, (unsigned long long)((g.src0 + src) * g.h + k), g.hstep, g.sc, g.h);
  const int own = g.half == 0 ? k : g.h + k;
  const int oth = (g.half == 0 ? g.h : 0) + d.partner;
  const double* s = g.pos + (src * g.nw + own) * 5;
  const double* c = g.pos + (src * g.nw + oth) * 5;
  double* q = g.q + i * 5;
  double p[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const double cj = c[j];
    p[j] = __dsub_rn(cj, __dmul_rn(d.z, __dsub_rn(cj, s[j])));
    q[j] = p[j];
  }
  setup_eval<THIN, ALPHA, true>(p, i, m, pr, scratch, sst, band_meta, nb_gauss, node_a, band_off);
}

template <bool THIN, bool ALPHA>
__global__ void __launch_bounds__(kSmallNodesThreads)
ens_nodes_accept_kernel(const EnsArgs g, const EvalArgs a, const int any_gprior, const DataRef d, const NodeTab t,
                        const double* __restrict__ scratch, const int* __restrict__ sst) {
  const long long i = blockIdx.x;
  double newlnp;
  int st;
  if (!nodes_small_eval<THIN, ALPHA>(a, i, any_gprior, d, t, scratch, sst, newlnp, st)) return;
  const long long src = i / g.h;
  const int k = (int)(i - src * g.h);
  const Draw dr = stretch_draw(g.keys, (unsigned long long)((g.src0 + src) * g.h + k), g.hstep, g.sc, g.h);
  const int own = g.half == 0 ? k : g.h + k;
  const long long w = src * g.nw + own;
  if (st > ST_BELOW_LOWLIM) {          // the reference would have raised here
    if (g.status[w] <= ST_BELOW_LOWLIM) g.status[w] = st;
    return;
  }
  if (stretch_accept(dr.z, dr.u, newlnp, g.lnp[w])) {
    const double* q = g.q + i * 5;
    double* p = g.pos + w * 5;
#pragma unroll
    for (int j = 0; j < 5; ++j) p[j] = q[j];
    g.lnp[w] = newlnp;
    if (g.count) g.nacc[w] += 1;
  }
}

// ---------------------------------------------------------------------------
// Posterior accumulators.  The hs threads that serve one source split its 5 x nw
// sample values by COMPONENT: thread r accumulates component j = r % 5 over the
// walkers slot, slot + nslots, ... (slot = r / 5, nslots = hs / 5), so a record is
// one conflict-free linear sweep of the ensemble in shared memory and the running
// S1 = sum (x - c), S2 = sum (x - c)^2, min and max of a thread are four REGISTERS
// for the whole run -- no read-modify-write of any memory per record.  (c = the
// source's walker 0 at the first record: a posterior width away from the mean,
// so S2 - S1^2/n loses no digits.)  The best sample is tracked per walker owner.
// fit_reduce_component / fit_write_row turn the per-thread partials of one source
// into its summary row, merging with what the row already holds (an earlier
// segment of the same run) by the pairwise update of Chan, Golub & LeVeque (1979).
// Deterministic: fixed serial order over the threads of the source.
// ---------------------------------------------------------------------------
struct StatAcc {
  double s1, s2, mn, mx, shift;
};

__device__ __forceinline__ void stat_reset(StatAcc& a) {
  a.s1 = 0.0;
  a.s2 = 0.0;
  a.mn = kInf;
  a.mx = -kInf;
  a.shift = 0.0;
}

// one record of one source: P = its ensemble [nw][5]
__device__ __forceinline__ void stat_record(StatAcc& a, const double* P, int nw, int cj, int slot, int nslots) {
  for (int w = slot; w < nw; w += nslots) {
    const double v = P[w * 5 + cj], dv = v - a.shift;
    a.s1 += dv;
    a.s2 = fma(dv, dv, a.s2);
    a.mn = fmin(a.mn, v);
    a.mx = fmax(a.mx, v);
  }
}

struct FitPartials {
  const double* part;     // [4][stride]: S1, S2, min, max of each thread (component = thread-in-source % 5)
  const double* bestpos;  // [5][stride]
  const double* bestlnp;  // [stride]
  int stride;
};

// component c (0..20) of the source served by threads [t0, t0 + hs): 0-4 S1, 5-9 S2, 10-14 min,
// 15-19 max; c = 20 returns the winning thread index (as a double) of the best-sample search
__device__ __forceinline__ double fit_reduce_component(const FitPartials& p, int c, int t0, int hs) {
  if (c == 20) {
    int win = t0;
    double best = p.bestlnp[t0];
    for (int t = t0 + 1; t < t0 + hs; ++t) {
      const double v = p.bestlnp[t];
      if (v > best) { best = v; win = t; }
    }
    return (double)win;
  }
  const int kind = c / 5, j = c - kind * 5, nslots = hs / 5;
  const double* q = p.part + kind * p.stride + t0 + j;
  double v = kind < 2 ? 0.0 : (kind == 2 ? kInf : -kInf);
  for (int sl = 0; sl < nslots; ++sl) {
    const double x = q[sl * 5];
    v = kind < 2 ? v + x : (kind == 2 ? fmin(v, x) : fmax(v, x));
  }
  return v;
}

// red[0..20] of one source -> its row; n_seg samples in this segment, shift[5]
__device__ __forceinline__ void fit_write_row(double* row, const double* red, const double* shift,
                                              const FitPartials& p, double n_seg, double acc_frac, int merge) {
  const int win = (int)red[20];
  const double blnp = p.bestlnp[win];
  if (n_seg > 0.0) {
    const double nA = merge ? row[FS_N] : 0.0;
    const double n = nA + n_seg;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const double s1 = red[j], s2 = red[5 + j];
      double mean = shift[j] + s1 / n_seg;
      double m2 = fmax(s2 - s1 * s1 / n_seg, 0.0);
      double mn = red[10 + j], mx = red[15 + j];
      if (nA > 0.0) {
        const double meanA = row[FS_MEAN + j], dlt = mean - meanA;
        m2 = row[FS_M2 + j] + m2 + dlt * dlt * (nA * n_seg / n);
        mean = meanA + dlt * (n_seg / n);
        mn = fmin(mn, row[FS_MIN + j]);
        mx = fmax(mx, row[FS_MAX + j]);
      }
      row[FS_MEAN + j] = mean;
      row[FS_M2 + j] = m2;
      row[FS_MIN + j] = mn;
      row[FS_MAX + j] = mx;
    }
    if (!(nA > 0.0) || blnp > row[FS_BESTLNP]) {
      row[FS_BESTLNP] = blnp;
#pragma unroll
      for (int j = 0; j < 5; ++j) row[FS_BEST + j] = p.bestpos[j * p.stride + win];
    }
    row[FS_N] = n;
  } else if (!merge) {
    for (int j = 0; j < kFitStats; ++j) row[j] = 0.0;
    row[FS_BESTLNP] = -kInf;
  }
  row[FS_ACC] = acc_frac;
}

// Summary update from ensembles in global memory (the propose / evaluate / accept path): one
// CTA per source, called once per recorded iteration; each call merges nw samples into the row.
constexpr int kStatsThreads = 128;
__global__ void __launch_bounds__(kStatsThreads)
ens_stats_kernel(const double* __restrict__ pos, const double* __restrict__ lnp, const int* __restrict__ nacc,
                 double* __restrict__ stats, int nw, int merge, double main_iters) {
  __shared__ double s_part[4 * kStatsThreads];
  __shared__ double s_bp[5 * kStatsThreads];
  __shared__ double s_bl[kStatsThreads];
  __shared__ double s_red[24];
  __shared__ double s_shift[5];
  __shared__ int s_acc[kStatsThreads / 32];
  const int tid = threadIdx.x;
  const long long src = blockIdx.x;
  const double* P = pos + src * nw * 5;
  const int hs = nw / 2 < kStatsThreads ? nw / 2 : kStatsThreads;      // >= 6
  const int nslots = hs / 5, cj = tid % 5, slot = tid / 5;
  StatAcc a;
  stat_reset(a);
  if (tid < hs && slot < nslots) {
    a.shift = P[cj];
    stat_record(a, P, nw, cj, slot, nslots);
  }
  if (tid < 5) s_shift[tid] = P[tid];
  s_part[tid] = a.s1;
  s_part[kStatsThreads + tid] = a.s2;
  s_part[2 * kStatsThreads + tid] = a.mn;
  s_part[3 * kStatsThreads + tid] = a.mx;
  double best = -kInf;
  int acc = 0;
  for (int w = tid; w < nw; w += hs) {
    if (tid >= hs) break;
    const double l = lnp[src * nw + w];
    if (l > best) {
      best = l;
#pragma unroll
      for (int j = 0; j < 5; ++j) s_bp[j * kStatsThreads + tid] = P[w * 5 + j];
    }
    acc += nacc[src * nw + w];
  }
  s_bl[tid] = best;
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) s_acc[tid >> 5] = acc;
  __syncthreads();
  const FitPartials p{s_part, s_bp, s_bl, kStatsThreads};
  if (tid < 21) s_red[tid] = fit_reduce_component(p, tid, 0, hs);
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int i = 0; i < kStatsThreads / 32; ++i) tot += s_acc[i];
    fit_write_row(stats + src * kFitStats, s_red, s_shift, p, (double)nw,
                  main_iters > 0.0 ? (double)tot / ((double)nw * main_iters) : 0.0, merge);
  }
}

// ---------------------------------------------------------------------------
// The source-resident sampler.
// ---------------------------------------------------------------------------
struct EnsFit {
  double* pos;              // [nsrc][nw][5], updated in place
  double* lnp;              // [nsrc][nw]
  int* nacc;                // [nsrc][nw], accepted moves of the main run (accumulated)
  int* status;              // [nsrc][nw], sticky
  double* stats;            // [nsrc][kFitStats] or null
  double* chain;            // this launch's records: [nrec][chain_nsrc][nw][5], or null
  double* chain_lnp;        // [nrec][chain_nsrc][nw], or null
  double* scratch;          // [gridDim.x][5][256]: per-thread best sample
  long long nsrc, src0;     // src0: global index of source 0 (RNG counter)
  long long dsrc0;          // index of source 0 in the photometry arrays of mbb_set_data
  long long chain_nsrc;     // sources per record of the chain arrays (>= nsrc: a shard writes into a larger array)
  int nw, h, G;             // G sources per CTA
  int niter;                // iterations in this launch
  int main_from;            // iterations >= main_from belong to the main run (counted, recorded)
  long long main_done;      // main iterations before this launch (a multiple of thin)
  int thin;
  int nrec;                 // recorded iterations in this launch
  int merge;                // the stats rows already hold earlier segments of this main run
  unsigned long long step0;   // step0: global iteration index of this launch's iteration 0
  PhiloxKeys keys;            // philox_keys(seed)
  StretchScale sc;
};

constexpr int kEnsThreads = 256;
#ifndef MBB_ENS_MINB
#define MBB_ENS_MINB 3
#endif

// dynamic shared memory of ens_resident_kernel, in doubles (then ints):
//   exp table | pos [G*nw*5] | lnp [G*nw] | flux,ivar [2*G*NB (even)] | shift [G*5 (even)] |
//   per-thread accumulators [6][256] + reduction [G*22 (even)]  (only with stats) | nacc ints [G*nw]
// The accumulators (S1, S2, min, max, shift, best log-probability of a thread) live in shared
// memory between records on purpose: as long-lived registers they end up in local memory, and the
// compiler then stores them back on every iteration of the record loop.
__host__ __device__ inline size_t ens_even(size_t n) { return (n + 1) & ~(size_t)1; }
__host__ __device__ inline size_t ens_resident_smem(int G, int nw, int nb, bool stats) {
  size_t d = kTabRepDoubles + (size_t)G * nw * 6 + ens_even((size_t)2 * G * nb) + ens_even((size_t)G * 5);
  if (stats) d += 6 * kEnsThreads + ens_even((size_t)G * 22);
  return d * 8 + (size_t)G * nw * 4;
}

// HT: the half-ensemble has exactly kEnsThreads walkers (BASELINE cfg5: 512 walkers) -- one
// source per CTA, one walker per thread and half-step; every index of the loop below is then a
// compile-time function of threadIdx.x (the generic form re-derives them from the launch
// arguments under register pressure: 140 of its 830 instructions per proposal).
template <bool THIN, bool ALPHA, int NB, bool HT>
__global__ void __launch_bounds__(kEnsThreads, MBB_ENS_MINB)
ens_resident_kernel(const EnsFit g, const ModelP m, const Priors pr, const DataRef d, const SmallTab t,
                    const ColdArgs* __restrict__ cold) {
  extern __shared__ __align__(16) double smem[];
  const int nw = HT ? 2 * kEnsThreads : g.nw, h = HT ? kEnsThreads : g.h, G = HT ? 1 : g.G;
  const bool stats = g.stats != nullptr;
  double* const s_tab = smem;
  double* const s_pos = s_tab + kTabRepDoubles;
  double* const s_lnp = s_pos + (size_t)G * nw * 5;
  double* const s_dat = s_lnp + (size_t)G * nw;
  double* const s_shift = s_dat + ens_even((size_t)2 * G * NB);
  double* const s_acc = s_shift + ens_even((size_t)G * 5);          // [6][256]: S1, S2, min, max, shift, best lnp
  double* const s_red = s_acc + 6 * kEnsThreads;
  int* const s_nacc = reinterpret_cast<int*>(stats ? s_red + ens_even((size_t)G * 22) : s_acc);
  const int tid = threadIdx.x;
  stage_exp_table(s_tab);
  const double* tab = lane_exp_table(s_tab);
  // thread -> (source of the group, first walker of the half, stride)
  const int sg = HT ? 0 : (h <= kEnsThreads ? tid / h : 0);
  const int kfirst = HT ? tid : (h <= kEnsThreads ? tid - sg * h : tid);
  const int kstride = HT ? kEnsThreads : (h <= kEnsThreads ? h : kEnsThreads);
  // summaries: thread r of the hs threads serving a source owns component r % 5, walkers r / 5 + i * nslots
  const int hs = h <= kEnsThreads ? h : kEnsThreads;
  const int nslots = hs / 5, cj = kfirst % 5, slot = kfirst / 5;
  const long long ngroups = (g.nsrc + G - 1) / G;
  double* const bp = g.scratch + (size_t)blockIdx.x * 5 * kEnsThreads;
  const bool use_cinv = d.cinv != nullptr;
  const int thin = g.thin;

  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long s_first = grp * G;
    const int ga = HT ? 1 : (int)((g.nsrc - s_first) < G ? (g.nsrc - s_first) : G);
    const bool active = HT || sg < ga;
    const long long src = s_first + sg;                  // this thread's source (index within the call)
    __syncthreads();      // the previous group's epilogue is done with the shared arrays (first pass: table staged)
    {
      const double2* gp2 = reinterpret_cast<const double2*>(g.pos + s_first * nw * 5);
      double2* sp2 = reinterpret_cast<double2*>(s_pos);
      for (int i = tid; i < ga * nw * 5 / 2; i += kEnsThreads) sp2[i] = gp2[i];
      const double2* gl2 = reinterpret_cast<const double2*>(g.lnp + s_first * nw);
      double2* sl2 = reinterpret_cast<double2*>(s_lnp);
      for (int i = tid; i < ga * nw / 2; i += kEnsThreads) sl2[i] = gl2[i];
      for (int i = tid; i < ga * nw; i += kEnsThreads) s_nacc[i] = g.nacc[s_first * nw + i];
      for (int i = tid; i < ga * NB; i += kEnsThreads) {
        s_dat[i] = d.flux[(g.dsrc0 + s_first) * NB + i];
        s_dat[G * NB + i] = use_cinv ? 0.0 : d.ivar[(g.dsrc0 + s_first) * NB + i];
      }
    }
    if (stats) {
      s_acc[tid] = 0.0;
      s_acc[kEnsThreads + tid] = 0.0;
      s_acc[2 * kEnsThreads + tid] = kInf;
      s_acc[3 * kEnsThreads + tid] = -kInf;
      s_acc[4 * kEnsThreads + tid] = 0.0;
      s_acc[5 * kEnsThreads + tid] = -kInf;
    }
    __syncthreads();

    // the next iteration of this launch that ends on a record (main iteration j records when
    // (main_done + j + 1) % thin == 0), and its record index: counters instead of a 64-bit
    // remainder per iteration
    int next_rec = g.main_from + (int)((thin - 1) - g.main_done % thin);
    int rec = 0;
    unsigned long long hstep = 2ull * g.step0;
    const unsigned long long wbase = (unsigned long long)((g.src0 + src) * h);

    // (Measured and rejected: a split barrier -- mbarrier arrive, the Philox draw of the next
    // half-step, then the wait -- so that an early warp has 11 % of its next proposal to do
    // before it waits: 0.3365 against 0.3337 ms per iteration of 2e4 sources, same chains.)
    for (int it = 0; it < g.niter; ++it) {
      const bool is_main = it >= g.main_from;
#pragma unroll 1
      for (int half = 0; half < 2; ++half, ++hstep) {
        if (active) {
          const int own0 = sg * nw + (half == 0 ? 0 : h);      // first walker of the updating half
          const int oth0 = sg * nw + (half == 0 ? h : 0);
#pragma unroll 1
          for (int k = kfirst; k < h; k += kstride) {
            const Draw dr = stretch_draw(g.keys, wbase + (unsigned)k, hstep, g.sc, h);
            const int own = own0 + k;
            double* const sj = s_pos + own * 5;
            const double* const cj5 = s_pos + (oth0 + dr.partner) * 5;
            double q[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
              const double c = cj5[j];
              q[j] = __dsub_rn(c, __dmul_rn(dr.z, __dsub_rn(c, sj[j])));
            }
            double diff[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) diff[b] = s_dat[sg * NB + b];
            int st;
            const double newlnp = delta_eval<THIN, ALPHA, NB>(q, g.dsrc0 + src, diff, m, pr, d, t, cold, tab,
                                                              use_cinv ? nullptr : s_dat + (G + sg) * NB, st);
            if (st > ST_BELOW_LOWLIM) {
              int* gs = g.status + (s_first * nw + own);
              if (*gs <= ST_BELOW_LOWLIM) *gs = st;
            } else if (stretch_accept(dr.z, dr.u, newlnp, s_lnp[own])) {
#pragma unroll
              for (int j = 0; j < 5; ++j) sj[j] = q[j];
              s_lnp[own] = newlnp;
              if (is_main) s_nacc[own] += 1;
            }
            if (HT) break;
          }
        }
        __syncthreads();
      }
      if (it == next_rec) {
        next_rec += thin;
        {
          if (stats && active) {
            const double* P = s_pos + (size_t)sg * nw * 5;
            if (slot < nslots) {
              StatAcc a;
              a.s1 = s_acc[tid];
              a.s2 = s_acc[kEnsThreads + tid];
              a.mn = s_acc[2 * kEnsThreads + tid];
              a.mx = s_acc[3 * kEnsThreads + tid];
              a.shift = rec == 0 ? P[cj] : s_acc[4 * kEnsThreads + tid];
              stat_record(a, P, nw, cj, slot, nslots);
              s_acc[tid] = a.s1;
              s_acc[kEnsThreads + tid] = a.s2;
              s_acc[2 * kEnsThreads + tid] = a.mn;
              s_acc[3 * kEnsThreads + tid] = a.mx;
              if (rec == 0) s_acc[4 * kEnsThreads + tid] = a.shift;
            }
            double best = s_acc[5 * kEnsThreads + tid];
            bool better = false;
            for (int k = kfirst; k < h; k += kstride) {           // best sample: by the walkers' owner
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int w = sg * nw + hh * h + k;
                const double l = s_lnp[w];
                if (l > best) {
                  best = l;
                  better = true;
#pragma unroll
                  for (int j = 0; j < 5; ++j) bp[j * kEnsThreads + tid] = s_pos[(size_t)w * 5 + j];
                }
              }
            }
            if (better) s_acc[5 * kEnsThreads + tid] = best;
          }
          if (g.chain || g.chain_lnp) {
            if (g.chain) {
              double2* dst = reinterpret_cast<double2*>(g.chain + ((size_t)rec * g.chain_nsrc + s_first) * nw * 5);
              const double2* sp2 = reinterpret_cast<const double2*>(s_pos);
              for (int i = tid; i < ga * nw * 5 / 2; i += kEnsThreads) dst[i] = sp2[i];
            }
            if (g.chain_lnp) {
              double2* dst = reinterpret_cast<double2*>(g.chain_lnp + ((size_t)rec * g.chain_nsrc + s_first) * nw);
              const double2* sl2 = reinterpret_cast<const double2*>(s_lnp);
              for (int i = tid; i < ga * nw / 2; i += kEnsThreads) dst[i] = sl2[i];
            }
          }
          ++rec;
          // a record reads rows that other threads own: they must not move before everyone has read
          if (stats || g.chain || g.chain_lnp) __syncthreads();
        }
      }
    }

    // ---- epilogue: ensembles back to HBM, per-thread partials -> summary rows
    {
      double2* gp2 = reinterpret_cast<double2*>(g.pos + s_first * nw * 5);
      const double2* sp2 = reinterpret_cast<const double2*>(s_pos);
      for (int i = tid; i < ga * nw * 5 / 2; i += kEnsThreads) gp2[i] = sp2[i];
      double2* gl2 = reinterpret_cast<double2*>(g.lnp + s_first * nw);
      const double2* sl2 = reinterpret_cast<const double2*>(s_lnp);
      for (int i = tid; i < ga * nw / 2; i += kEnsThreads) gl2[i] = sl2[i];
      for (int i = tid; i < ga * nw; i += kEnsThreads) g.nacc[s_first * nw + i] = s_nacc[i];
    }
    if (stats) {
      if (active && slot == 0) s_shift[sg * 5 + cj] = s_acc[4 * kEnsThreads + tid];
      __syncthreads();           // (also makes the per-thread global scratch visible CTA-wide)
      const FitPartials p{s_acc, bp, s_acc + 5 * kEnsThreads, kEnsThreads};
      for (int pi = tid; pi < ga * 22; pi += kEnsThreads) {
        const int s2 = pi / 22, c = pi - s2 * 22;
        const int t0 = h <= kEnsThreads ? s2 * h : 0;
        double v;
        if (c < 21) {
          v = fit_reduce_component(p, c, t0, hs);
        } else {
          int tot = 0;
          for (int w = 0; w < nw; ++w) tot += s_nacc[s2 * nw + w];
          v = (double)tot;
        }
        s_red[pi] = v;
      }
      __syncthreads();
      if (tid < ga) {
        const int main_iters_here = g.niter - (g.main_from > 0 ? g.main_from : 0);
        const double iters = (double)(g.main_done + (main_iters_here > 0 ? main_iters_here : 0));
        const double frac = iters > 0.0 ? s_red[tid * 22 + 21] / ((double)nw * iters) : 0.0;
        fit_write_row(g.stats + (s_first + tid) * kFitStats, s_red + tid * 22, s_shift + tid * 5, p,
                      (double)g.nrec * (double)nw, frac, g.merge);
      }
    }
  }
}

}  // namespace mbb
