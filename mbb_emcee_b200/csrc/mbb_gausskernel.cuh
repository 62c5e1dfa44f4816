// MBB_MATH_FAST_GAUSS, thread-per-evaluation kernel.
//
// With the tabulated bands reduced to 32-point Gauss rules an evaluation is a
// few hundred node evaluations, not thousands, and the warp-per-evaluation
// mapping of loglike_nodes_kernel spends more on per-evaluation overhead (a
// separate setup kernel and its scratch round trip, a shuffle reduction per
// band, the same scalar bookkeeping in all 32 lanes: ~740 of 1770 warp
// instructions per evaluation, ncu) than on nodes.  Here one THREAD owns an
// evaluation, exactly like the delta-band kernel: setup, the masks deciding
// rule / kink split / full table per band, the node loops and the chi-square
// all stay in its registers; a warp instruction advances 32 evaluations.
//
//   * persistent CTAs; parameter tiles arrive by cp.async.bulk + mbarrier into
//     a 3-stage shared-memory ring (same pipeline as loglike_delta_kernel).
//     (Measured and rejected: every warp its own pipeline of 32-evaluation tiles
//     with per-warp mbarriers and no CTA barrier -- 1.58 vs 1.53 ms for cfg2, 2.74
//     vs 2.65 ms for cfg5p; rule groups of 1 or 4 nodes and 2 CTAs/SM at 120
//     registers: all within 2 %.);
//   * the compressed rules (nb x 32 nodes) and the replicated exp table live in
//     shared memory; every lane reads the SAME rule node at the same time
//     (broadcast, conflict-free);
//   * a band's grey-side rule nodes are evaluated two at a time with their six
//     exp chains written breadth-first (grey_nodes_n), which gives each thread
//     the independent work the 8-cycle DFMA latency asks for;
//   * the full tables stay in global memory (read-only path): only bands that
//     fail the per-walker bound touch them, at warp-uniform addresses;
//   * with the power-law join a band that contains the walker's merge point needs
//     table corrections of walker-dependent length (band_partial_kink).  A thread
//     doing them serially makes its 31 neighbours wait (measured: 2.88 ms for cfg2
//     against 2.11 ms in the warp kernel), so they are done by the WARP: after the
//     per-thread bands, each lane with a kink band broadcasts its eight per-walker
//     constants by shuffle and all 32 lanes stride that band's table nodes;
//   * diagonal errors only (a full covariance needs all band residuals at once:
//     warp kernel as well).
#pragma once
#include "mbb_kernels.cuh"

namespace mbb {

struct GaussThreadTab {
  const double2* a;        // full table {freq, weff}   (global)
  const double* b;         // full table L'             (global)
  const int* band_off;     // nb + 1                    (global)
  const double2* ca;       // compressed {freq, weff}   (global; staged to smem)
  const double* cb;
  const int* comp_off;     // nb + 1
  const BandMeta* meta;    // nb
  int nb, nc;
};

__host__ __device__ inline size_t gauss_thread_smem(int nb, int nc) {
  return (size_t)kNodesTabDoubles * 8 + (size_t)3 * kDeltaTile * 40 + 3 * 8 + 8 +
         (size_t)nc * 16 + nodes_b_bytes(nc) + (size_t)(nb + 1) * 8 + (size_t)nb * sizeof(BandMeta) + 64;
}

// full-table band for one thread
template <bool THIN, bool ALPHA, bool CLAMP>
__device__ __forceinline__ double band_table_thread(const FastSed& fs, const double2* __restrict__ ga,
                                                    const double* __restrict__ gb, int i0, int i1,
                                                    const double* tab) {
  double acc = 0.0;
  for (int i = i0; i < i1; ++i) {
    const double2 fw = __ldg(ga + i);
    acc = node_acc<THIN, ALPHA, CLAMP, kNodesTS>(fs, fw.x, __ldg(gb + i), fw.y, acc, tab);
  }
  return acc;
}

// grey branch over a compressed rule in shared memory, two nodes at a time
template <bool THIN>
__device__ __forceinline__ double rule_grey_thread(const FastSed& fs, const double2* __restrict__ ca,
                                                   const double* __restrict__ cb, int c0, int c1,
                                                   const double* tab) {
  double acc2[2] = {0.0, 0.0};
  int i = c0;
  for (; i + 1 < c1; i += 2) {
    const double2 f0 = ca[i], f1 = ca[i + 1];
    const double nu[2] = {f0.x, f1.x}, lp[2] = {cb[i], cb[i + 1]}, we[2] = {f0.y, f1.y};
    grey_nodes_n<THIN, 2, kNodesTS>(fs, nu, lp, we, acc2, tab);
  }
  double acc = acc2[0] + acc2[1];
  if (i < c1) {
    const double2 fw = ca[i];
    acc = node_grey<THIN, false, kNodesTS>(fs, fw.x, cb[i], fw.y, acc, tab);
  }
  return acc;
}

__device__ __forceinline__ double rule_pow_thread(const FastSed& fs, const double2* __restrict__ ca,
                                                  const double* __restrict__ cb, int c0, int c1,
                                                  const double* tab) {
  double acc = 0.0;
  for (int i = c0; i < c1; ++i) acc = node_pow<false, kNodesTS>(fs, cb[i], ca[i].y, acc, tab);
  return acc;
}

__device__ __forceinline__ double shfl_d(double v, int src_lane) {
  return __shfl_sync(0xffffffffu, v, src_lane);
}

constexpr int kThreadKinkSlots = 2;   // kink bands one evaluation can hand to the warp phase

template <bool THIN, bool ALPHA>
__global__ void __launch_bounds__(MBB_DELTA_BLOCK, MBB_DELTA_MINB)
loglike_gauss_thread_kernel(const EvalArgs a, const ModelP m, const Priors pr, const DataRef d,
                            const GaussThreadTab t, const ColdArgs* __restrict__ cold, const int use_tma) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_tab = reinterpret_cast<double*>(smem_raw);
  double* s_par = s_tab + kNodesTabDoubles;                                   // [3][kDeltaTile*5]
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_par + 3 * kDeltaTile * 5);   // [3] (+1 pad)
  double2* s_ca = reinterpret_cast<double2*>(s_bar + 4);
  double* s_cb = reinterpret_cast<double*>(s_ca + t.nc);
  BandMeta* s_meta = reinterpret_cast<BandMeta*>(reinterpret_cast<unsigned char*>(s_cb) + nodes_b_bytes(t.nc));
  int* s_off = reinterpret_cast<int*>(s_meta + t.nb);
  int* s_coff = s_off + t.nb + 1;
  const int tid = threadIdx.x, nb = t.nb;
  const long long sd = a.soa_stride ? a.soa_stride : a.n;
  stage_exp_table256(s_tab);
  for (int i = tid; i < t.nc; i += blockDim.x) {
    s_ca[i] = t.ca[i];
    s_cb[i] = t.cb[i];
  }
  for (int i = tid; i <= nb; i += blockDim.x) {
    s_off[i] = t.band_off[i];
    s_coff[i] = t.comp_off[i];
  }
  for (int i = tid; i < nb; i += blockDim.x) s_meta[i] = t.meta[i];
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) mbar_init(&s_bar[i], 1);
  }
  __syncthreads();
  auto issue = [&](unsigned tl, int stage) {
    const long long e0 = (long long)tl * kDeltaTile;
    if (!use_tma || a.n - e0 < kDeltaTile) return;
    double* dst = s_par + stage * (kDeltaTile * 5);
    mbar_expect_tx(&s_bar[stage], kDeltaTile * 40u);
    if (a.layout == 0) {
      bulk_g2s(dst, a.pars + e0 * 5, kDeltaTile * 40u, &s_bar[stage]);
    } else {
#pragma unroll
      for (int i = 0; i < 5; ++i)
        bulk_g2s(dst + i * kDeltaTile, a.pars + (long long)i * sd + e0, kDeltaTile * 8u, &s_bar[stage]);
    }
  };
  const double* tab = lane_exp_table(s_tab);
  const unsigned nt = (unsigned)((a.n + kDeltaTile - 1) / kDeltaTile);
  unsigned tile = blockIdx.x;
  if (tid == 0 && tile < nt) issue(tile, 0);
  int stage = 0;
  unsigned par = 0;
  for (; tile < nt; tile += gridDim.x) {
    const long long e0 = (long long)tile * kDeltaTile, e = e0 + tid;
    const bool via_tma = use_tma && a.n - e0 >= kDeltaTile;
    if (tid == 0 && tile + gridDim.x < nt) issue(tile + gridDim.x, stage + 1 == 3 ? 0 : stage + 1);
    const bool active = e < a.n;
    double p[5];
    if (via_tma) {
      mbar_wait(&s_bar[stage], par);
      const double* sp = s_par + stage * (kDeltaTile * 5);
#pragma unroll
      for (int i = 0; i < 5; ++i) p[i] = a.layout == 0 ? sp[tid * 5 + i] : sp[i * kDeltaTile + tid];
    } else if (active) {
      load_pars(a, e, p);
    }
    __syncthreads();
    // ---- per-thread part: setup, masks, every band that needs no warp cooperation ----
    int st = ST_OK;
    double lnl = qnan(), chi = 0.0;
    bool live = false;                     // this lane carries an evaluation through to the chi-square
    FastSed fs;
    int kband[kThreadKinkSlots], ksplit[kThreadKinkSlots], nkink = 0;
    const double* __restrict__ fl = nullptr;
    const double* __restrict__ ivp = nullptr;
    if (active) {
      if (below_lowlim(pr, p)) {
        st = ST_BELOW_LOWLIM;
        lnl = -kInf;
      } else {
        fast_setup<THIN, ALPHA>(fs, p[0], p[1], p[2], p[3], p[4], m);
        fast_sed_rescale256(fs);           // node loops below: 1/256-octave units (kNodesTS)
        st = fs.status;
        live = st == ST_OK;
      }
    }
    if (live) {
      GaussMasks gm;
      gm.plain = gm.kink = 0;
      if (fs.safe) gm = gauss_band_masks<THIN, ALPHA>(fs, s_meta, nb);
      const long long src = source_of(a, e);
      fl = d.flux + src * nb;
      ivp = d.ivar + src * nb;
      for (int b = 0; b < nb; ++b) {
        const int i0 = s_off[b], i1 = s_off[b + 1], c0 = s_coff[b], c1 = s_coff[b + 1];
        double acc;
        if ((gm.plain >> b) & 1ull) {
          if (ALPHA && gt_pos(s_meta[b].nu_lo, fs.nu_merge)) acc = rule_pow_thread(fs, s_ca, s_cb, c0, c1, tab);
          else acc = rule_grey_thread<THIN>(fs, s_ca, s_cb, c0, c1, tab);
        } else if (ALPHA && ((gm.kink >> b) & 1ull) && nkink < kThreadKinkSlots) {
          // split index of the (frequency-descending) table at the merge point; the band itself
          // is evaluated by the whole warp below
          int lo = i0, hi = i1;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (gt_pos(__ldg(&t.a[mid].x), fs.nu_merge)) lo = mid + 1;
            else hi = mid;
          }
          kband[nkink] = b;
          ksplit[nkink] = lo;
          ++nkink;
          continue;
        } else if (fs.safe) {
          acc = band_table_thread<THIN, ALPHA, false>(fs, t.a, t.b, i0, i1, tab);
        } else {
          acc = band_table_thread<THIN, ALPHA, true>(fs, t.a, t.b, i0, i1, tab);
        }
        const double df = __ldg(fl + b) - acc;
        chi = fma(df * df, __ldg(ivp + b), chi);
      }
    }
    // ---- warp part: the kink bands, one (evaluation, band) at a time, nodes over the lanes ----
    if (ALPHA) {
      const int lane = tid & 31;
#pragma unroll
      for (int slot = 0; slot < kThreadKinkSlots; ++slot) {
        unsigned pending = __ballot_sync(0xffffffffu, live && nkink > slot);
        while (pending) {
          const int L = __ffs(pending) - 1;
          pending &= pending - 1;
          FastSed g;
          g.xk_hi = shfl_d(fs.xk_hi, L); g.xk_lo = shfl_d(fs.xk_lo, L);
          g.nb = shfl_d(fs.nb, L); g.apow = shfl_d(fs.apow, L);
          g.t0c = shfl_d(fs.t0c, L); g.nu_merge = shfl_d(fs.nu_merge, L);
          g.amp_grey = shfl_d(fs.amp_grey, L); g.amp_pow = shfl_d(fs.amp_pow, L);
          const int b = __shfl_sync(0xffffffffu, kband[slot], L);
          const int k = __shfl_sync(0xffffffffu, ksplit[slot], L);
          double acc = band_partial_kink<THIN>(g, t.a, t.b, s_off[b], s_off[b + 1], s_ca, s_cb, s_coff[b],
                                               s_coff[b + 1], k, lane, tab);
          acc = warp_sum(acc);
          if (lane == L) {
            const double df = __ldg(fl + b) - acc;
            chi = fma(df * df, __ldg(ivp + b), chi);
          }
        }
      }
    }
    if (active) {
      if (live) {
        lnl = -0.5 * chi;
        if (pr.peak_terms) {
          lnl = add_priors_cold<THIN>(lnl, p[0], p[1], p[2], p[3], p[4], fs.x0, cold, &st);
          if (st != ST_OK) lnl = qnan();
        } else {
          lnl = add_simple_priors(pr, p, lnl);
        }
        if (st == ST_OK && lnl != lnl) st = ST_NONFINITE;
      }
      a.out[e] = lnl;
      if (a.status) a.status[e] = st;
    }
    if (++stage == 3) {
      stage = 0;
      par ^= 1u;
    }
  }
}

}  // namespace mbb
