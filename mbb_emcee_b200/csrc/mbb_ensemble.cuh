// Device-resident affine-invariant ensemble sampler for many sources at once
// (SURVEY.md 8f row 1).  The stretch move of Goodman & Weare (2010) with the
// scheduling of emcee 2.2 (two half-ensembles per iteration, reference
// mbb_fit.py:80-81, 533-542 drives exactly this through emcee), but with the
// proposals, the log-probability and the accept/reject step all on the GPU, so
// walker positions never cross PCIe.
//
// Randomness is counter-based (Philox4x32-10, Salmon et al. 2011): the draws
// of walker `w` of source `s` at half-step `h` of iteration `t` depend only on
// (seed, s*nw/2 + w, 2t+h), so results are independent of launch geometry and
// can be replayed on the host (tests/test_ensemble_gpu.py does exactly that
// with a numpy Philox and the CPU oracle).  Statistically equivalent to, not
// bit-identical with, emcee's Mersenne-Twister stream.
#pragma once
#include "mbb_kernels.cuh"
#include "mbb_philox.cuh"

namespace mbb {

struct Draw {
  double z;      // stretch factor ((a-1)u+1)^2/a
  double u;      // the acceptance uniform
  int partner;   // index into the complementary half
};

// emcee 2.2 accepts where (dim-1) ln z + lnp(q) - lnp(s) > ln u.  The same test
// without logarithms:  u < z^4 exp(lnp(q) - lnp(s))  -- one lean exp instead of two
// libdevice logs (~240 of the fused kernel's ~1400 issue cycles per warp).  The
// difference is clamped to [-745, 700] first: -inf (proposal below a lower limit)
// and NaN reject, +inf accepts, exactly as the logarithmic form does.
__device__ __forceinline__ bool stretch_accept(double z, double u, double newlnp, double oldlnp) {
  const double dl = fmin(fmax(newlnp - oldlnp, -745.0), 700.0);
  const double z2 = z * z;
  return u < (z2 * z2) * exp_l(dl);
}

__device__ __forceinline__ Draw stretch_draw(unsigned long long seed, unsigned long long widx,
                                             unsigned long long hstep, double a, int ncomp) {
  const unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
  const Philox r = philox4x32_10((unsigned)widx, (unsigned)(widx >> 32), (unsigned)hstep,
                                 (unsigned)(hstep >> 32) << 1, k0, k1);
  const Philox s = philox4x32_10((unsigned)widx, (unsigned)(widx >> 32), (unsigned)hstep,
                                 ((unsigned)(hstep >> 32) << 1) | 1u, k0, k1);
  Draw d;
  // emcee: zz = ((a - 1.) * rand + 1) ** 2. / a   -- separate roundings, no FMA
  const double t = __dadd_rn(__dmul_rn(a - 1.0, u53(r.c[0], r.c[1])), 1.0);
  d.z = __ddiv_rn(__dmul_rn(t, t), a);
  d.partner = (int)(((unsigned long long)r.c[2] * (unsigned long long)ncomp) >> 32);
  d.u = u53(s.c[0], s.c[1]);
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
  long long nsrc;
  int nw, h, half;          // h = nw/2; half 0 updates walkers [0,h) against [h,nw)
  unsigned long long seed, hstep;   // hstep = 2*iteration + half
  double a;
};

// q = c[j] - z (c[j] - s), in emcee's operation order
__global__ void __launch_bounds__(256) ens_propose_kernel(const EnsArgs g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nsrc * g.h) return;
  const long long src = i / g.h;
  const int k = (int)(i - src * g.h);
  const Draw d = stretch_draw(g.seed, (unsigned long long)i, g.hstep, g.a, g.h);
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

// accept where (dim-1) ln z + lnp(q) - lnp(s) > ln u   (emcee 2.2 _propose_stretch), see stretch_accept
__global__ void __launch_bounds__(256) ens_accept_kernel(const EnsArgs g) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nsrc * g.h) return;
  const long long src = i / g.h;
  const int k = (int)(i - src * g.h);
  const Draw d = stretch_draw(g.seed, (unsigned long long)i, g.hstep, g.a, g.h);
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
    g.nacc[w] += 1;
  }
}

// Fused half-step for delta-band configurations: draw, propose, evaluate and
// accept in one kernel; the proposal never leaves registers.  Same draws, same
// arithmetic, hence the same chain as the three-kernel pipeline above.
// this kernel waits on gathers (ncu: long_scoreboard 3.9 of ~9 stalled warps per issue), so a
// fourth resident CTA at 64 registers pays for its spills: 2.55 vs 2.60 ms per iteration (2 CTAs
// at 116 registers: 3.20 ms).
// Measured and rejected: a persistent form that stages whole source ensembles (20 KB of
// positions + log-probabilities + photometry rows per tile) by cp.async.bulk two tiles ahead,
// so that the partner gather becomes a shared-memory read -- bit-identical chains, but 2.57-2.61
// ms per iteration: the 60 KB of staging allow 3 CTAs/SM, 80 registers still spill (q, z, u and
// the old log-probability stay live across the evaluation), and the CTA barrier per tile costs
// what the gathers did (17 % of the stall samples; the accept/reject branch makes the warps
// uneven).  Releasing the ring stage with a per-warp count instead of that barrier ("last warp
// refills") was slower again, here (2.62 ms) and in loglike_delta_kernel (1.29 vs 1.20 ms).
#ifndef MBB_ENS_MINB
#define MBB_ENS_MINB 4
#endif
template <bool THIN, bool ALPHA, int NB>
__global__ void __launch_bounds__(MBB_DELTA_BLOCK, MBB_ENS_MINB)
ens_delta_kernel(const EnsArgs g, const ModelP m, const Priors pr, const DataRef d, const SmallTab t,
                 const ColdArgs* __restrict__ cold) {
  __shared__ __align__(16) double s_tab[kTabRepDoubles];
  stage_exp_table(s_tab);
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.nsrc * g.h) return;
  const unsigned ih = (unsigned)g.h;
  const long long src = (g.nsrc * g.h < (1LL << 32)) ? (long long)((unsigned)i / ih) : i / g.h;
  const int k = (int)(i - src * g.h);
  // Same draws as stretch_draw(), issued in the order that hides the gathers: the first Philox
  // block gives the partner index, the loads (own row, partner row, old log-probability,
  // photometry) go out, and the second block and the division run under them.
  const unsigned k0 = (unsigned)g.seed, k1 = (unsigned)(g.seed >> 32);
  const unsigned long long widx = (unsigned long long)i;
  const Philox r = philox4x32_10((unsigned)widx, (unsigned)(widx >> 32), (unsigned)g.hstep,
                                 (unsigned)(g.hstep >> 32) << 1, k0, k1);
  const int partner = (int)(((unsigned long long)r.c[2] * (unsigned long long)g.h) >> 32);
  const int own = g.half == 0 ? k : g.h + k;
  const int oth = (g.half == 0 ? g.h : 0) + partner;
  const long long w = src * g.nw + own;
  const double* __restrict__ s = g.pos + w * 5;
  const double* __restrict__ c = g.pos + (src * g.nw + oth) * 5;
  double sj[5], cj[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    sj[j] = s[j];
    cj[j] = c[j];
  }
  const double old = g.lnp[w];
  double diff[NB];
  delta_load_data<NB>(d, src, diff);
  const Philox r2 = philox4x32_10((unsigned)widx, (unsigned)(widx >> 32), (unsigned)g.hstep,
                                  ((unsigned)(g.hstep >> 32) << 1) | 1u, k0, k1);
  Draw dr;
  {
    const double t1 = __dadd_rn(__dmul_rn(g.a - 1.0, u53(r.c[0], r.c[1])), 1.0);
    dr.z = __ddiv_rn(__dmul_rn(t1, t1), g.a);
    dr.partner = partner;
    dr.u = u53(r2.c[0], r2.c[1]);
  }
  double q[5];
#pragma unroll
  for (int j = 0; j < 5; ++j) q[j] = __dsub_rn(cj[j], __dmul_rn(dr.z, __dsub_rn(cj[j], sj[j])));
  int st;
  const double newlnp = delta_eval<THIN, ALPHA, NB>(q, src, diff, m, pr, d, t, cold, lane_exp_table(s_tab), nullptr, st);
  if (st > ST_BELOW_LOWLIM) {
    if (g.status[w] <= ST_BELOW_LOWLIM) g.status[w] = st;
    return;
  }
  if (stretch_accept(dr.z, dr.u, newlnp, old)) {
    double* p = g.pos + w * 5;
#pragma unroll
    for (int j = 0; j < 5; ++j) p[j] = q[j];
    g.lnp[w] = newlnp;
    g.nacc[w] += 1;
  }
}

}  // namespace mbb
