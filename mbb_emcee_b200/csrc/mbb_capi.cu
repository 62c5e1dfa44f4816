// C ABI of libmbb_b200.so -- see include/mbb_b200.h for the contract.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mbb_b200.h"
#include "mbb_ensemble.cuh"
#include "mbb_gaussrule.h"
#include "mbb_gausskernel.cuh"
#include "mbb_kernels.cuh"
#include "mbb_chainstats.cuh"

using namespace mbb;

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess)                                                             \
      return fail(std::string(#call) + ": " + cudaGetErrorString(_e));                 \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template <class T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct mbb_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  int64_t launches = 0;

  // model
  double wavenorm = 500.0;
  int opthin = 0, noalpha = 0;
  int math_mode = MBB_MATH_FAST;
  int lir_method = MBB_LIR_QUADPACK;

  // passbands
  int nb = 0, nn = 0;
  std::vector<int> h_off;
  std::vector<unsigned char> h_scalar;
  std::vector<double> h_wave, h_weight;   // as given to mbb_set_bands
  DevBuf<double2> d_node_fw;      // {freq, passband weight}: FAITHFUL kernels, chain_flux
  DevBuf<double2> d_node_fast_a;  // {freq, weff}  FAST (depends on opthin)
  DevBuf<double> d_node_fast_b;   // L' = log(wave/wavenorm)*256/ln2, FAST (depends on wavenorm)
  // MBB_MATH_FAST_GAUSS: 32-point Gauss rules of the tabulated bands
  DevBuf<double2> d_comp_a;
  DevBuf<double> d_comp_b;
  DevBuf<int> d_comp_off;
  DevBuf<BandMeta> d_band_meta;
  int nc = 0;                     // compressed nodes in total (0: no band has a rule)
  double nu_max = 0.0, lmax = 0.0;
  bool fast_tables_ok = false;    // FAST tables match the current (wavenorm, opthin)
  uint64_t delta_attr_set = 0;    // delta-kernel instantiations whose shared-memory attribute is set
  DevBuf<ColdArgs> d_cold;        // SmallTab + priors + model for the kernels' cold paths
  bool cold_ok = false;
  DevBuf<int> d_off;
  DevBuf<unsigned char> d_scalar;
  SmallTab small;
  bool bands_set = false;

  // data
  int nsrc = 0, data_nb = 0;
  DevBuf<double> d_flux, d_ivar, d_cinv;
  bool has_ivar = false, has_cinv = false;
  bool cinv_is_chol = false;      // d_cinv holds Cholesky factors (mbb_set_data_chol)

  Priors pri;

  // pipelined staging for MBB_HOST mbb_loglike calls (see mbb_loglike)
  struct Slot {
    cudaStream_t s = nullptr;
    cudaEvent_t done = nullptr;
    PinBuf<double> hin, hout;
    PinBuf<int> hst, hsrc;
    DevBuf<double> din, dout;
    DevBuf<int> dst, dsrc;
    int64_t pend_e0 = -1, pend_n = 0;   // chunk whose outputs still sit in hout/hst
    // fixed columns (mbb_set_fixed_params) this slot's din already holds: [5][fill_stride]
    uint64_t fill_epoch = 0;
    int64_t fill_stride = 0;
    const double* fill_base = nullptr;
    // source chunks of mbb_ensemble_fit(MBB_HOST)
    DevBuf<double> epos, elnp, estats, escratch;
    DevBuf<int> enacc, est;
  };
  Slot slots[3];
  bool slots_ready = false;
  // mbb_set_fixed_params: columns the caller promises to be constant
  unsigned fixed_mask = 0;
  double fixed_val[5] = {0, 0, 0, 0, 0};
  uint64_t fixed_epoch = 1;

  // staging for the other MBB_HOST calls
  PinBuf<double> h_in, h_out;
  PinBuf<int> h_st, h_src;
  DevBuf<double> d_in, d_out, d_aux0, d_aux1;
  DevBuf<int> d_st, d_src, d_owner, d_work;
  DevBuf<unsigned> d_count;
  DevBuf<unsigned long long> d_stat_cnt;     // mbb_chain_stats: counts, NaN counts
  DevBuf<double> d_stat_part;
  DevBuf<unsigned> d_stat_hist;
  // ensemble sampler scratch
  // per-stream scratch of the split warp path (main stream + 3 pipeline slots)
  struct Scratch {
    cudaStream_t st = nullptr;
    bool used = false;
    DevBuf<double> c;
    DevBuf<int> s;
  };
  Scratch scratch[5];
  cudaError_t scratch_for(cudaStream_t st, size_t cap, double** c_out, int** s_out) {
    Scratch* sc = nullptr;
    for (auto& x : scratch)
      if (x.used && x.st == st) sc = &x;
    if (!sc)
      for (auto& x : scratch)
        if (!x.used) { sc = &x; x.used = true; x.st = st; break; }
    if (!sc) return cudaErrorMemoryAllocation;
    cudaError_t e = sc->c.reserve(cap * kScratchStride);
    if (e != cudaSuccess) return e;
    e = sc->s.reserve(cap);
    if (e != cudaSuccess) return e;
    *c_out = sc->c.p;
    *s_out = sc->s.p;
    return cudaSuccess;
  }
  // small host-buffer calls of mbb_loglike (the 125-walker half-steps of a single-source fit, reference
  // mbb_fit.py:533-542): upload, kernels and downloads replayed as ONE CUDA graph launch per call
  struct SmallGraph {
    int64_t n = 0;
    long long wps = 0;
    int layout = 0, with_status = 0, seen = 0;
    uint64_t gen = 0;
    cudaGraphExec_t exec = nullptr;
  };
  static constexpr int64_t kGraphMaxN = 8192;
  SmallGraph graphs[4];
  int graph_next = 0;
  uint64_t gen = 1;                      // bumped by every set_* call: what a captured graph has baked in
  cudaStream_t gstream = nullptr;
  PinBuf<double> g_hin, g_hout;
  PinBuf<int> g_hst;
  DevBuf<double> g_din, g_dout;
  DevBuf<int> g_dst;
  int64_t graph_launches = 0;

  DevBuf<double> d_epos, d_elnp, d_eq, d_eqlnp, d_escratch, d_estats, d_echain[2], d_echainl[2];
  DevBuf<int> d_enacc, d_est, d_eqst;
  cudaStream_t copy_stream = nullptr;          // D2H of chain segments behind the sampler (mbb_ensemble_fit)
  cudaEvent_t seg_done[2] = {nullptr, nullptr}, seg_copied[2] = {nullptr, nullptr};
};

namespace {

void default_priors(Priors& p) {
  // likelihood.py:73, 83-92
  const double low[5] = {1, 0.1, 1, 0.1, 1e-3};
  for (int i = 0; i < 5; ++i) p.lowlim[i] = low[i];
  for (int i = 0; i < 6; ++i) {
    p.uplim[i] = kInf;
    p.has_uplim[i] = 0;
    p.has_gprior[i] = 0;
    p.gmean[i] = 0.0;
    p.givar[i] = 1.0;
  }
  p.has_uplim[1] = p.has_uplim[3] = 1;
  p.uplim[1] = p.uplim[3] = 20.0;
  priors_finalize(p);
}

struct Use {
  mbb_ctx* c;
  int prev = -1;
  explicit Use(mbb_ctx* ctx) : c(ctx) {
    cudaGetDevice(&prev);
    if (prev != c->device) cudaSetDevice(c->device);
  }
  ~Use() {
    if (prev >= 0 && prev != c->device) cudaSetDevice(prev);
  }
};

ModelP model_of(const mbb_ctx* c) {
  ModelP m;
  m.wavenorm = c->wavenorm;
  m.nu_norm = kUmToGHz / c->wavenorm;
  m.nu_max = c->nu_max;
  m.lmax = c->lmax;
  return m;
}

// (Re)build the mode-dependent node tables: the FAST constants depend on
// wavenorm (L') and on opthin (weff), which mbb_set_model may change after
// mbb_set_bands.
cudaError_t upload_cold(mbb_ctx* c);

cudaError_t ensure_tables(mbb_ctx* c) {
  if (c->fast_tables_ok) return c->cold_ok ? cudaSuccess : upload_cold(c);
  const int nn = c->nn;
  const bool thin = c->opthin != 0;
  std::vector<double2> fw((size_t)nn), fa((size_t)nn);
  std::vector<double> fb((size_t)nn + 1, 0.0);   // +1: padded to a multiple of 16 bytes for the bulk copy
  c->nu_max = 0.0;
  c->lmax = 0.0;
  SmallTab& s = c->small;
  memset(&s, 0, sizeof(s));
  s.nb = c->nb;
  for (int i = 0; i < nn; ++i) {
    const FastNode n = fast_node(c->h_wave[i], c->h_weight[i], c->wavenorm, thin);
    fw[i] = make_double2(n.freq, c->h_weight[i]);
    fa[i] = make_double2(n.freq, n.weff);
    fb[i] = 4.0 * n.lp;          // the nodes / Gauss-rule kernels work in 1/256-octave units (kNodesTS)
    if (n.freq > c->nu_max) c->nu_max = n.freq;
    if (n.labs > c->lmax) c->lmax = n.labs;
    if (nn <= kSmallMaxNodes) {
      s.freq[i] = n.freq; s.w[i] = c->h_weight[i]; s.weff[i] = n.weff; s.lp[i] = n.lp;
    }
  }
  if (nn <= kSmallMaxNodes) {
    for (int b = 0; b <= c->nb; ++b) s.band_off[b] = c->h_off[b];
    for (int b = 0; b < c->nb; ++b) s.scalar_path[b] = c->h_scalar[b];
  }
  cudaError_t e;
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;
  if ((e = c->d_node_fw.reserve((size_t)nn)) != cudaSuccess) return e;
  if ((e = c->d_node_fast_a.reserve((size_t)nn)) != cudaSuccess) return e;
  if ((e = c->d_node_fast_b.reserve((size_t)nn + 1)) != cudaSuccess) return e;
  const size_t bytes = (size_t)nn * sizeof(double2);
  if ((e = cudaMemcpy(c->d_node_fw.p, fw.data(), bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(c->d_node_fast_a.p, fa.data(), bytes, cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(c->d_node_fast_b.p, fb.data(), ((size_t)nn + 1) * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
    return e;
  // compressed rules (MBB_MATH_FAST_GAUSS), see mbb_gaussrule.h
  {
    const GaussTables g = build_gauss_tables(c->nb, c->h_off.data(), c->h_wave.data(), c->h_weight.data(),
                                             c->h_scalar.data(), c->wavenorm, thin);
    c->nc = (int)g.freq.size();
    std::vector<double2> ca((size_t)c->nc + 1, make_double2(0.0, 0.0));
    std::vector<double> cb((size_t)c->nc + 2, 0.0);       // padded to a multiple of 16 bytes
    for (int k = 0; k < c->nc; ++k) {
      ca[k] = make_double2(g.freq[k], g.weff[k]);
      cb[k] = 4.0 * g.lp[k];
    }
    if ((e = c->d_comp_a.reserve(ca.size())) != cudaSuccess) return e;
    if ((e = c->d_comp_b.reserve(cb.size())) != cudaSuccess) return e;
    if ((e = c->d_comp_off.reserve(g.off.size())) != cudaSuccess) return e;
    if ((e = c->d_band_meta.reserve(g.meta.size())) != cudaSuccess) return e;
    if ((e = cudaMemcpy(c->d_comp_a.p, ca.data(), ca.size() * sizeof(double2), cudaMemcpyHostToDevice)) != cudaSuccess)
      return e;
    if ((e = cudaMemcpy(c->d_comp_b.p, cb.data(), cb.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess)
      return e;
    if ((e = cudaMemcpy(c->d_comp_off.p, g.off.data(), g.off.size() * sizeof(int), cudaMemcpyHostToDevice)) !=
        cudaSuccess)
      return e;
    if ((e = cudaMemcpy(c->d_band_meta.p, g.meta.data(), g.meta.size() * sizeof(BandMeta),
                        cudaMemcpyHostToDevice)) != cudaSuccess)
      return e;
  }
  c->fast_tables_ok = true;
  return upload_cold(c);
}

cudaError_t upload_cold(mbb_ctx* c) {
  ColdArgs h;
  h.t = c->small;
  h.pr = c->pri;
  h.m = model_of(c);
  cudaError_t e;
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return e;
  for (auto& sl : c->slots)
    if (sl.s && (e = cudaStreamSynchronize(sl.s)) != cudaSuccess) return e;
  if ((e = c->d_cold.reserve(1)) != cudaSuccess) return e;
  if ((e = cudaMemcpy(c->d_cold.p, &h, sizeof(h), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
  c->cold_ok = true;
  return cudaSuccess;
}

void begin_timing(mbb_ctx* c) { cudaEventRecord(c->ev0, c->stream); }
void end_timing(mbb_ctx* c) {
  cudaEventRecord(c->ev1, c->stream);
  c->timed = true;
}

template <template <bool, bool, bool> class L, class... A>
void dispatch3(bool thin, bool alpha, bool fast, A&&... args) {
  if (thin) {
    if (alpha) {
      if (fast) L<true, true, true>::run(args...); else L<true, true, false>::run(args...);
    } else {
      if (fast) L<true, false, true>::run(args...); else L<true, false, false>::run(args...);
    }
  } else {
    if (alpha) {
      if (fast) L<false, true, true>::run(args...); else L<false, true, false>::run(args...);
    } else {
      if (fast) L<false, false, true>::run(args...); else L<false, false, false>::run(args...);
    }
  }
}

template <bool THIN, bool ALPHA, bool FAST>
struct LaunchThread {
  static void run(mbb_ctx* c, cudaStream_t st, const EvalArgs& a, const DataRef& d, cudaError_t* err) {
    const unsigned grid = (unsigned)((a.n + 255) / 256);
    const ModelP m = model_of(c);
    loglike_thread_kernel<THIN, ALPHA, FAST><<<grid, 256, 0, st>>>(a, m, c->pri, d, c->small);
    *err = cudaSuccess;
  }
};

constexpr int kMaxDeltaNB = 8;
constexpr long long kSmallNodesMaxN = 4096;    // up to here a CTA per evaluation beats a warp per evaluation

template <bool THIN, bool ALPHA, int NB>
struct DeltaLauncher {
  static void go(mbb_ctx* c, cudaStream_t st, const EvalArgs& a, const DataRef& d, int nb) {
    if (nb == NB) {
      // persistent: one wave of CTAs walks the tiles; parameter tiles arrive by TMA when the
      // buffer allows it (16-byte aligned; SoA rows 16-byte aligned too)
      const long long ntiles = (a.n + kDeltaTile - 1) / kDeltaTile;
      const long long resident = (long long)c->sm_count * MBB_DELTA_MINB;
      const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);
      const long long sd = a.soa_stride ? a.soa_stride : a.n;
      static const bool no_tma = getenv("MBB_B200_NO_TMA") != nullptr;
      const int use_tma = !no_tma && ((uintptr_t)a.pars % 16 == 0) && (a.layout == MBB_AOS || sd % 2 == 0);
      // photometry rows of a tile travel with it when sources are implicit (walkers_per_source)
      // and the errors are diagonal; d_flux / d_ivar are cudaMalloc'ed (256-byte aligned) and padded
      static const bool no_stage = getenv("MBB_B200_NO_STAGE_DATA") != nullptr;
      // (walkers_per_source below 2^31 with multiply-shift constants: the kernel finds a thread's
      // source relative to its tile's first one in 32-bit arithmetic)
      const int stage_data = use_tma && !no_stage && !a.src_index && d.ivar && !d.cinv && a.wps_mul != 0 &&
                             a.wps < (1LL << 31);
      const ModelP m = model_of(c);
      // the replicated exp table is the kernel's dynamic shared memory (beside ~33 KB static)
      constexpr size_t kTabBytes = sizeof(double) * kTabRepDoubles;
      const uint64_t bit = 1ull << ((THIN ? 1 : 0) | (ALPHA ? 2 : 0) | (NB << 2));
      if (!(c->delta_attr_set & bit)) {      // once per context (the attribute is per device)
        cudaFuncSetAttribute(loglike_delta_kernel<THIN, ALPHA, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kTabBytes);
        c->delta_attr_set |= bit;
      }
      loglike_delta_kernel<THIN, ALPHA, NB><<<grid, MBB_DELTA_BLOCK, kTabBytes, st>>>(a, m, c->pri, d, c->small,
                                                                                   c->d_cold.p, use_tma, stage_data);
    } else {
      DeltaLauncher<THIN, ALPHA, NB - 1>::go(c, st, a, d, nb);
    }
  }
};
template <bool THIN, bool ALPHA>
struct DeltaLauncher<THIN, ALPHA, 0> {
  static void go(mbb_ctx*, cudaStream_t, const EvalArgs&, const DataRef&, int) {}
};

// source-resident sampler (mbb_ensemble.cuh): grid = one wave of CTAs, each walking groups of G sources
struct EnsResidentPlan {
  int G = 0;
  size_t smem = 0;
  bool ok = false;
};

EnsResidentPlan ens_resident_plan(const mbb_ctx* c, int nwalkers, bool stats) {
  EnsResidentPlan p;
  const int h = nwalkers / 2;
  p.G = h <= kEnsThreads ? kEnsThreads / h : 1;
  p.smem = ens_resident_smem(p.G, nwalkers, c->nb, stats);
  // one CTA must leave room for a second one's static needs: cap at the opt-in limit
  while (p.G > 1 && p.smem > c->smem_optin) {
    --p.G;
    p.smem = ens_resident_smem(p.G, nwalkers, c->nb, stats);
  }
  p.ok = p.smem <= c->smem_optin;
  return p;
}

template <bool THIN, bool ALPHA, int NB>
struct EnsResidentLauncher {
  static void go(mbb_ctx* c, EnsFit& g, const DataRef& d, int nb, size_t smem, cudaStream_t st,
                 DevBuf<double>* scratch, cudaError_t* err) {
    if (nb == NB) {
      // half-ensembles of exactly kEnsThreads walkers: the instantiation with compile-time indices
      auto kern = g.h == kEnsThreads && g.G == 1 ? ens_resident_kernel<THIN, ALPHA, NB, true>
                                                 : ens_resident_kernel<THIN, ALPHA, NB, false>;
      *err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (*err != cudaSuccess) return;
      int per_sm = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEnsThreads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
      const long long ngroups = (g.nsrc + g.G - 1) / g.G;
      const long long resident = (long long)c->sm_count * per_sm;
      const unsigned grid = (unsigned)(ngroups < resident ? ngroups : resident);
      *err = scratch->reserve((size_t)grid * 5 * kEnsThreads);
      if (*err != cudaSuccess) return;
      g.scratch = scratch->p;
      const ModelP m = model_of(c);
      kern<<<grid, kEnsThreads, smem, st>>>(g, m, c->pri, d, c->small, c->d_cold.p);
    } else {
      EnsResidentLauncher<THIN, ALPHA, NB - 1>::go(c, g, d, nb, smem, st, scratch, err);
    }
  }
};
template <bool THIN, bool ALPHA>
struct EnsResidentLauncher<THIN, ALPHA, 0> {
  static void go(mbb_ctx*, EnsFit&, const DataRef&, int, size_t, cudaStream_t, DevBuf<double>*, cudaError_t*) {}
};

template <bool THIN, bool ALPHA, bool FAST>
struct LaunchEnsResident {
  static void run(mbb_ctx* c, EnsFit& g, const DataRef& d, size_t smem, cudaStream_t st, DevBuf<double>* scratch,
                  cudaError_t* err) {
    EnsResidentLauncher<THIN, ALPHA, kMaxDeltaNB>::go(c, g, d, c->nb, smem, st, scratch, err);
  }
};

template <bool THIN, bool ALPHA, bool FAST>
struct LaunchDelta {
  static void run(mbb_ctx* c, cudaStream_t st, const EvalArgs& a, const DataRef& d, cudaError_t* err) {
    DeltaLauncher<THIN, ALPHA, kMaxDeltaNB>::go(c, st, a, d, c->nb);
    *err = cudaSuccess;
  }
};

// split warp path: setup kernel + lean nodes kernel, in chunks that bound the scratch
template <bool THIN, bool ALPHA, bool FAST>
struct LaunchSplit {
  static void run(mbb_ctx* c, cudaStream_t st, const EvalArgs& a0, const DataRef& d, cudaError_t* err) {
    const long long kChunk = 1LL << 22;
    const long long cap = a0.n < kChunk ? a0.n : kChunk;
    *err = cudaSuccess;
    NodeTab t;
    t.a = FAST ? c->d_node_fast_a.p : c->d_node_fw.p;
    t.b = c->d_node_fast_b.p;
    t.band_off = c->d_off.p;
    t.scalar_path = c->d_scalar.p;
    t.nb = c->nb;
    t.nn = c->nn;
    const bool gauss = FAST && c->math_mode == MBB_MATH_FAST_GAUSS && c->nc > 0;
    t.ca = c->d_comp_a.p;
    t.cb = c->d_comp_b.p;
    t.comp_off = c->d_comp_off.p;
    t.nc = gauss ? c->nc : 0;
    size_t smem = nodes_kernel_smem(c->nn, true, FAST, t.nc);
    const bool in_smem = smem <= c->smem_optin;
    if (!in_smem) smem = nodes_kernel_smem(c->nn, false, FAST, t.nc);
    auto nodes = gauss ? (in_smem ? loglike_nodes_kernel<THIN, ALPHA, FAST, true, FAST>
                                  : loglike_nodes_kernel<THIN, ALPHA, FAST, false, FAST>)
                       : (in_smem ? loglike_nodes_kernel<THIN, ALPHA, FAST, true, false>
                                  : loglike_nodes_kernel<THIN, ALPHA, FAST, false, false>);
    *err = cudaFuncSetAttribute(nodes, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (*err != cudaSuccess) return;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nodes, kNodesThreads, smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    const ModelP m = model_of(c);
    for (long long off = 0; off < a0.n; off += cap) {
      EvalArgs a = a0;
      a.n = (a0.n - off) < cap ? (a0.n - off) : cap;
      a.e0 = a0.e0 + off;
      a.soa_stride = a0.soa_stride ? a0.soa_stride : a0.n;
      a.pars = a0.layout == MBB_AOS ? a0.pars + off * 5 : a0.pars + off;
      a.out = a0.out + off;
      a.status = a0.status ? a0.status + off : nullptr;
      a.src_index = a0.src_index ? a0.src_index + off : nullptr;
      double* scratch = nullptr;
      int* sst = nullptr;
      *err = c->scratch_for(st, (size_t)cap, &scratch, &sst);
      if (*err != cudaSuccess) return;
      loglike_setup_kernel<THIN, ALPHA, FAST><<<(unsigned)((a.n + 127) / 128), 128, 0, st>>>(
          a, m, c->pri, scratch, sst, c->d_band_meta.p, gauss ? c->nb : 0, c->d_node_fast_a.p, c->d_off.p);
      const long long want = (a.n + kNodesWarps - 1) / kNodesWarps;
      const long long resident = (long long)c->sm_count * per_sm;
      const unsigned grid = (unsigned)(want < resident ? want : resident);
      // small batches (a single source's half-ensemble): one CTA per evaluation instead of one warp
      static const bool no_small = getenv("MBB_B200_NO_SMALL_NODES") != nullptr;
      if (FAST && !gauss && !no_small && a.n <= kSmallNodesMaxN) {
        loglike_nodes_small_kernel<THIN, ALPHA><<<(unsigned)a.n, kSmallNodesThreads, 0, st>>>(
            a, c->pri.any_gprior, d, t, scratch, sst);
      } else {
        nodes<<<grid, kNodesThreads, smem, st>>>(a, c->pri.any_gprior, d, t, scratch, sst);
      }
      c->launches += 1;   // the caller counts one launch per call; add the second kernel
    }
  }
};

// Half-step of the propose / evaluate / accept sampler over tabulated passbands for small ensembles, in
// two launches (mbb_ensemble.cuh: ens_propose_setup_kernel, ens_nodes_accept_kernel); FAST, full tables.
template <bool THIN, bool ALPHA, bool FAST>
struct LaunchEnsHalfStepFused {
  static void run(mbb_ctx* c, cudaStream_t st, const EnsArgs& g, const EvalArgs& a, const DataRef& d,
                  cudaError_t* err) {
    NodeTab t;
    t.a = c->d_node_fast_a.p;
    t.b = c->d_node_fast_b.p;
    t.band_off = c->d_off.p;
    t.scalar_path = c->d_scalar.p;
    t.nb = c->nb;
    t.nn = c->nn;
    t.ca = c->d_comp_a.p;
    t.cb = c->d_comp_b.p;
    t.comp_off = c->d_comp_off.p;
    t.nc = 0;
    double* scratch = nullptr;
    int* sst = nullptr;
    *err = c->scratch_for(st, (size_t)a.n, &scratch, &sst);
    if (*err != cudaSuccess) return;
    const ModelP m = model_of(c);
    ens_propose_setup_kernel<THIN, ALPHA><<<(unsigned)((a.n + 127) / 128), 128, 0, st>>>(
        g, m, c->pri, scratch, sst, c->d_band_meta.p, 0, c->d_node_fast_a.p, c->d_off.p);
    ens_nodes_accept_kernel<THIN, ALPHA><<<(unsigned)a.n, kSmallNodesThreads, 0, st>>>(g, a, c->pri.any_gprior, d, t,
                                                                                      scratch, sst);
  }
};

// MBB_MATH_FAST_GAUSS with diagonal errors: thread-per-evaluation kernel (mbb_gausskernel.cuh)
template <bool THIN, bool ALPHA, bool UNUSED>
struct LaunchGaussThread {
  static void run(mbb_ctx* c, cudaStream_t st, const EvalArgs& a, const DataRef& d, cudaError_t* err) {
    GaussThreadTab t;
    t.a = c->d_node_fast_a.p;
    t.b = c->d_node_fast_b.p;
    t.band_off = c->d_off.p;
    t.ca = c->d_comp_a.p;
    t.cb = c->d_comp_b.p;
    t.comp_off = c->d_comp_off.p;
    t.meta = c->d_band_meta.p;
    t.nb = c->nb;
    t.nc = c->nc;
    const size_t smem = gauss_thread_smem(c->nb, c->nc);
    auto kern = loglike_gauss_thread_kernel<THIN, ALPHA>;
    *err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (*err != cudaSuccess) return;
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MBB_DELTA_BLOCK, smem) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    const long long ntiles = (a.n + kDeltaTile - 1) / kDeltaTile;
    const long long resident = (long long)c->sm_count * per_sm;
    const unsigned grid = (unsigned)(ntiles < resident ? ntiles : resident);
    const long long sd = a.soa_stride ? a.soa_stride : a.n;
    static const bool no_tma = getenv("MBB_B200_NO_TMA") != nullptr;
    const int use_tma = !no_tma && ((uintptr_t)a.pars % 16 == 0) && (a.layout == MBB_AOS || sd % 2 == 0);
    const ModelP m = model_of(c);
    kern<<<grid, MBB_DELTA_BLOCK, smem, st>>>(a, m, c->pri, d, t, c->d_cold.p, use_tma);
  }
};

template <bool THIN, bool ALPHA, bool UNUSED>
struct LaunchFnu {
  static void run(mbb_ctx* c, const EvalArgs& a, const double* freq, int nfreq, int scalar_path) {
    dim3 grid((unsigned)((nfreq + 255) / 256), (unsigned)a.n);
    const ModelP m = model_of(c);
    fnu_kernel<THIN, ALPHA><<<grid, 256, 0, c->stream>>>(a, m, freq, nfreq, scalar_path);
  }
};

template <bool THIN, bool ALPHA, bool UNUSED>
struct LaunchFnuFast {
  static void run(mbb_ctx* c, const EvalArgs& a, const double* freq, const double* weff, const double* lp,
                  int nfreq) {
    dim3 grid((unsigned)((nfreq + 255) / 256), (unsigned)a.n);
    const ModelP m = model_of(c);
    fnu_fast_kernel<THIN, ALPHA><<<grid, 256, 0, c->stream>>>(a, m, freq, weff, lp, nfreq);
  }
};

template <bool THIN, bool ALPHA, bool UNUSED>
struct LaunchConsts {
  static void run(mbb_ctx* c, const EvalArgs& a, int want_peak) {
    const unsigned grid = (unsigned)((a.n + 127) / 128);
    const ModelP m = model_of(c);
    sed_consts_kernel<THIN, ALPHA><<<grid, 128, 0, c->stream>>>(a, m, want_peak);
  }
};

}  // namespace

extern "C" {

int mbb_version(void) { return 100; }

const char* mbb_last_error(void) { return g_err.c_str(); }

int mbb_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
    return -1;
  }
  return n;
}

int mbb_ctx_create(int device_ordinal, mbb_ctx** out) {
  if (!out) return fail("mbb_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(std::string("no usable CUDA device (there is no CPU fallback): ") +
                (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device_ordinal < 0 || device_ordinal >= n) return fail("mbb_ctx_create: bad device ordinal");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device_ordinal));
  if (prop.major < 10)
    return fail("mbb_ctx_create: kernels are built for sm_100a only; device is sm_" +
                std::to_string(prop.major) + std::to_string(prop.minor));
  mbb_ctx* c = new mbb_ctx();
  c->device = device_ordinal;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  default_priors(c->pri);
  Use u(c);
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
    delete c;
    return fail("mbb_ctx_create: stream/event creation failed");
  }
  *out = c;
  return 0;
}

int mbb_ctx_destroy(mbb_ctx* c) {
  if (!c) return 0;
  Use u(c);
  cudaStreamSynchronize(c->stream);
  c->d_cold.release(); c->d_comp_a.release(); c->d_comp_b.release(); c->d_comp_off.release();
  c->d_band_meta.release(); c->d_node_fw.release(); c->d_node_fast_a.release(); c->d_node_fast_b.release(); c->d_off.release(); c->d_scalar.release();
  c->d_flux.release(); c->d_ivar.release(); c->d_cinv.release();
  c->h_in.release(); c->h_out.release(); c->h_st.release(); c->h_src.release();
  c->d_in.release(); c->d_out.release(); c->d_aux0.release(); c->d_aux1.release();
  c->d_st.release(); c->d_src.release(); c->d_owner.release(); c->d_work.release();
  c->d_count.release();
  for (auto& x : c->scratch) { x.c.release(); x.s.release(); }
  c->d_epos.release(); c->d_elnp.release(); c->d_eq.release(); c->d_eqlnp.release();
  c->d_enacc.release(); c->d_est.release(); c->d_eqst.release();
  c->d_escratch.release(); c->d_estats.release();
  for (int i = 0; i < 2; ++i) {
    c->d_echain[i].release(); c->d_echainl[i].release();
    if (c->seg_done[i]) cudaEventDestroy(c->seg_done[i]);
    if (c->seg_copied[i]) cudaEventDestroy(c->seg_copied[i]);
  }
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  for (auto& g : c->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (c->gstream) cudaStreamDestroy(c->gstream);
  c->g_hin.release(); c->g_hout.release(); c->g_hst.release();
  c->g_din.release(); c->g_dout.release(); c->g_dst.release();
  for (auto& sl : c->slots) {
    sl.hin.release(); sl.hout.release(); sl.hst.release(); sl.hsrc.release();
    sl.din.release(); sl.dout.release(); sl.dst.release(); sl.dsrc.release();
    sl.epos.release(); sl.elnp.release(); sl.estats.release(); sl.escratch.release();
    sl.enacc.release(); sl.est.release();
    if (sl.done) cudaEventDestroy(sl.done);
    if (sl.s) cudaStreamDestroy(sl.s);
  }
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int mbb_sync(mbb_ctx* c) {
  if (!c) return fail("null context");
  Use u(c);
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int64_t mbb_launch_count(const mbb_ctx* c) { return c ? c->launches : 0; }

uint64_t mbb_stream_handle(const mbb_ctx* c) { return c ? (uint64_t)(uintptr_t)c->stream : 0; }

int mbb_last_kernel_ms(mbb_ctx* c, float* ms) {
  if (!c || !ms) return fail("null argument");
  if (!c->timed) return fail("no timed call yet");
  Use u(c);
  CK(cudaEventSynchronize(c->ev1));
  CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return 0;
}

int mbb_host_alloc(size_t bytes, void** out) {
  if (!out) return fail("mbb_host_alloc: out is NULL");
  *out = nullptr;
  CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return 0;
}

int mbb_host_register(void* p, size_t bytes) {
  if (!p || bytes == 0) return fail("mbb_host_register: empty range");
  CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return 0;
}

int mbb_host_unregister(void* p) {
  if (p) CK(cudaHostUnregister(p));
  return 0;
}

int mbb_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return 0;
}

int mbb_set_model(mbb_ctx* c, double wavenorm, int opthin, int noalpha) {
  if (!c) return fail("null context");
  if (!(wavenorm > 0.0)) return fail("wavenorm must be positive");
  // the FAST node tables hold log(lambda/wavenorm) and (thick) weights scaled by
  // (wavenorm/lambda)^3: rebuilt lazily at the next launch
  if (wavenorm != c->wavenorm || (opthin ? 1 : 0) != c->opthin) c->fast_tables_ok = false;
  c->gen++;
  c->wavenorm = wavenorm;
  c->opthin = opthin ? 1 : 0;
  c->noalpha = noalpha ? 1 : 0;
  return 0;
}

int mbb_set_math_mode(mbb_ctx* c, int mode) {
  if (!c) return fail("null context");
  if (mode != MBB_MATH_FAITHFUL && mode != MBB_MATH_FAST && mode != MBB_MATH_FAST_GAUSS)
    return fail("unknown math mode");
  c->math_mode = mode;
  c->gen++;
  return 0;
}

int mbb_set_lir_method(mbb_ctx* c, int method) {
  if (!c) return fail("null context");
  if (method != MBB_LIR_QUADPACK && method != MBB_LIR_GAUSS) return fail("unknown L_IR method");
  c->lir_method = method;
  return 0;
}

int mbb_set_bands(mbb_ctx* c, int nbands, const int32_t* band_off, const double* node_wave_um,
                  const double* node_weight, const uint8_t* scalar_path) {
  if (!c) return fail("null context");
  if (nbands <= 0 || nbands > kMaxBands) return fail("nbands must be in 1.." + std::to_string(kMaxBands));
  if (!band_off || !node_wave_um || !node_weight) return fail("null table pointer");
  if (band_off[0] != 0) return fail("band_off[0] must be 0");
  for (int b = 0; b < nbands; ++b)
    if (band_off[b + 1] <= band_off[b]) return fail("band_off must be strictly increasing");
  const int nn = band_off[nbands];
  for (int i = 0; i < nn; ++i)
    if (!(node_wave_um[i] > 0.0)) return fail("node wavelengths must be positive");
  Use u(c);
  CK(cudaStreamSynchronize(c->stream));
  c->nb = nbands;
  c->nn = nn;
  c->h_off.assign(band_off, band_off + nbands + 1);
  c->h_scalar.assign(nbands, 0);
  if (scalar_path)
    for (int b = 0; b < nbands; ++b) c->h_scalar[b] = scalar_path[b] ? 1 : 0;
  c->h_wave.assign(node_wave_um, node_wave_um + nn);
  c->h_weight.assign(node_weight, node_weight + nn);
  CK(c->d_off.reserve(nbands + 1));
  CK(c->d_scalar.reserve(nbands));
  CK(cudaMemcpyAsync(c->d_off.p, c->h_off.data(), (nbands + 1) * sizeof(int), cudaMemcpyHostToDevice,
                     c->stream));
  CK(cudaMemcpyAsync(c->d_scalar.p, c->h_scalar.data(), nbands, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->fast_tables_ok = false;
  c->gen++;
  CK(ensure_tables(c));
  c->bands_set = true;
  return 0;
}

int mbb_set_data(mbb_ctx* c, int nsrc, int nbands, const double* flux, const double* ivar,
                 const double* cinv) {
  if (!c) return fail("null context");
  if (nsrc <= 0 || nbands <= 0) return fail("nsrc and nbands must be positive");
  if (!flux) return fail("flux is NULL");
  if ((ivar == nullptr) == (cinv == nullptr)) return fail("give exactly one of ivar / cinv");
  Use u(c);
  CK(cudaStreamSynchronize(c->stream));
  const size_t n1 = (size_t)nsrc * nbands;
  CK(c->d_flux.reserve(n1 + 2));      // +2: the delta kernel's bulk copies round up to 16 bytes
  CK(cudaMemcpy(c->d_flux.p, flux, n1 * sizeof(double), cudaMemcpyHostToDevice));
  c->has_ivar = c->has_cinv = c->cinv_is_chol = false;
  if (ivar) {
    CK(c->d_ivar.reserve(n1 + 2));
    CK(cudaMemcpy(c->d_ivar.p, ivar, n1 * sizeof(double), cudaMemcpyHostToDevice));
    c->has_ivar = true;
  } else {
    const size_t n2 = n1 * nbands;
    CK(c->d_cinv.reserve(n2));
    CK(cudaMemcpy(c->d_cinv.p, cinv, n2 * sizeof(double), cudaMemcpyHostToDevice));
    c->has_cinv = true;
  }
  c->nsrc = nsrc;
  c->data_nb = nbands;
  c->gen++;
  return 0;
}

int mbb_set_data_chol(mbb_ctx* c, int nsrc, int nbands, const double* flux, const double* chol) {
  if (!c) return fail("null context");
  if (nsrc <= 0 || nbands <= 0) return fail("nsrc and nbands must be positive");
  if (!flux || !chol) return fail("flux / chol is NULL");
  // staged form: strictly lower triangle = L, diagonal = 1/L_rr, upper triangle zero
  const size_t nb = (size_t)nbands, n2 = (size_t)nsrc * nb * nb;
  std::vector<double> lhat(n2, 0.0);
  for (size_t s = 0; s < (size_t)nsrc; ++s) {
    const double* L = chol + s * nb * nb;
    double* H = lhat.data() + s * nb * nb;
    for (size_t r = 0; r < nb; ++r) {
      const double d = L[r * nb + r];
      if (!(d > 0.0) || !std::isfinite(d)) return fail("Cholesky factor with a non-positive or non-finite diagonal");
      H[r * nb + r] = 1.0 / d;
      for (size_t cc = 0; cc < r; ++cc) {
        if (!std::isfinite(L[r * nb + cc])) return fail("non-finite entry in a Cholesky factor");
        H[r * nb + cc] = L[r * nb + cc];
      }
    }
  }
  if (mbb_set_data(c, nsrc, nbands, flux, nullptr, lhat.data()) != 0) return -1;
  c->cinv_is_chol = true;
  return 0;
}

int mbb_set_priors(mbb_ctx* c, const double lowlim[5], const uint8_t has_uplim[6],
                   const double uplim[6], const uint8_t has_gprior[6], const double gmean[6],
                   const double givar[6]) {
  if (!c) return fail("null context");
  if (!lowlim || !has_uplim || !uplim || !has_gprior || !gmean || !givar) return fail("null argument");
  Priors& p = c->pri;
  for (int i = 0; i < 5; ++i) p.lowlim[i] = lowlim[i];
  for (int i = 0; i < 6; ++i) {
    p.has_uplim[i] = has_uplim[i] ? 1 : 0;
    p.uplim[i] = uplim[i];
    p.has_gprior[i] = has_gprior[i] ? 1 : 0;
    p.gmean[i] = gmean[i];
    p.givar[i] = givar[i];
  }
  priors_finalize(p);
  c->gen++;
  c->cold_ok = false;
  return 0;
}

namespace {

bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

int launch_loglike(mbb_ctx* c, cudaStream_t st, const EvalArgs& a_in) {
  EvalArgs a = a_in;
  set_wps_division(a);
  DataRef d;
  d.flux = c->d_flux.p;
  d.ivar = c->has_ivar ? c->d_ivar.p : nullptr;
  d.cinv = c->has_cinv ? c->d_cinv.p : nullptr;
  d.chol = c->has_cinv && c->cinv_is_chol ? 1 : 0;
  d.nsrc = c->nsrc;
  d.nb = c->nb;
  const bool thin = c->opthin != 0, alpha = c->noalpha == 0, fast = c->math_mode != MBB_MATH_FAITHFUL;
  CK(ensure_tables(c));
  cudaError_t err = cudaSuccess;
  if (fast && c->nn == c->nb && c->nb <= kMaxDeltaNB) dispatch3<LaunchDelta>(thin, alpha, true, c, st, a, d, &err);
  else if (c->nn <= kSmallMaxNodes) dispatch3<LaunchThread>(thin, alpha, fast, c, st, a, d, &err);
  else {
    // Gauss rules + diagonal errors: one thread per evaluation, kink bands by the warp
    // (mbb_gausskernel.cuh: cfg5p 5.29 -> 2.64 ms, cfg2 2.11 -> 1.53 ms against the warp path);
    // full covariances keep the warp path.
    static const bool gauss_warp = getenv("MBB_B200_GAUSS_WARP") != nullptr;
    // (a small batch -- e.g. the 125 walkers of one emcee half-step -- cannot fill the GPU with one
    // thread per evaluation; there the warp path's 32 lanes per evaluation give the lower latency)
    const bool gthread = c->math_mode == MBB_MATH_FAST_GAUSS && c->nc > 0 && !d.cinv && !gauss_warp &&
                         a.n >= (long long)c->sm_count * 384 &&
                         gauss_thread_smem(c->nb, c->nc) <= c->smem_optin;
    if (gthread) dispatch3<LaunchGaussThread>(thin, alpha, true, c, st, a, d, &err);
    else dispatch3<LaunchSplit>(thin, alpha, fast, c, st, a, d, &err);
  }
  CK(err);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

int ensure_slots(mbb_ctx* c) {
  if (c->slots_ready) return 0;
  for (auto& sl : c->slots) {
    CK(cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
  }
  c->slots_ready = true;
  return 0;
}

// Evaluations per pipeline chunk of the host path (env MBB_B200_CHUNK overrides).
int64_t host_chunk() {
  static int64_t v = 0;
  if (v == 0) {
    const char* e = getenv("MBB_B200_CHUNK");
    v = e ? atoll(e) : (int64_t)1 << 21;
    if (v < 1024) v = 1024;
  }
  return v;
}

}  // namespace

namespace {

// Small host-buffer batch through a captured graph.  Returns -1 when the call was not handled (first
// sighting of this shape: the plain path runs once and sizes every buffer), else 0 / 1.
int loglike_small_graph(mbb_ctx* c, int64_t n, const double* pars, int layout, long long wps, double* out_lnlike,
                        int32_t* out_status) {
  static const bool off = getenv("MBB_B200_NO_GRAPH") != nullptr;
  if (off) return -1;
  mbb_ctx::SmallGraph* g = nullptr;
  for (auto& x : c->graphs)
    if (x.n == n && x.layout == layout && x.wps == wps && x.with_status == (out_status ? 1 : 0)) g = &x;
  if (!g) {
    g = &c->graphs[c->graph_next];
    c->graph_next = (c->graph_next + 1) % 4;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    *g = mbb_ctx::SmallGraph();
    g->n = n; g->layout = layout; g->wps = wps; g->with_status = out_status ? 1 : 0;
  }
  if (g->exec && g->gen != c->gen) {
    cudaGraphExecDestroy(g->exec);
    g->exec = nullptr;
  }
  if (!g->exec) {
    if (g->seen++ == 0) return -1;
    // capture: fixed staging buffers, fixed scratch (sized for the largest graphed batch)
    if (!c->gstream) CK(cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking));
    const size_t cap = (size_t)mbb_ctx::kGraphMaxN;
    CK(c->g_hin.reserve(cap * 5)); CK(c->g_din.reserve(cap * 5));
    CK(c->g_hout.reserve(cap)); CK(c->g_dout.reserve(cap));
    CK(c->g_hst.reserve(cap)); CK(c->g_dst.reserve(cap));
    double* sc = nullptr;
    int* ss = nullptr;
    CK(c->scratch_for(c->gstream, cap, &sc, &ss));
    CK(ensure_tables(c));
    EvalArgs a{};
    a.n = n; a.e0 = 0; a.wps = wps; a.layout = layout;
    a.pars = c->g_din.p; a.src_index = nullptr; a.out = c->g_dout.p; a.status = c->g_dst.p;
    const int64_t before = c->launches;
    // (an uncaptured run on this stream first: function attributes, lazily loaded modules)
    if (launch_loglike(c, c->gstream, a)) return 1;
    CK(cudaStreamSynchronize(c->gstream));
    c->launches = before;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(c->gstream, cudaStreamCaptureModeRelaxed));
    cudaMemcpyAsync(c->g_din.p, c->g_hin.p, (size_t)n * 5 * sizeof(double), cudaMemcpyHostToDevice, c->gstream);
    const int rc = launch_loglike(c, c->gstream, a);
    cudaMemcpyAsync(c->g_hout.p, c->g_dout.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->gstream);
    if (out_status)
      cudaMemcpyAsync(c->g_hst.p, c->g_dst.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->gstream);
    const cudaError_t ce = cudaStreamEndCapture(c->gstream, &graph);
    g->gen = c->gen;
    c->graph_launches = c->launches - before;      // kernels per replay
    c->launches = before;
    if (rc != 0 || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      g->seen = 0;
      return -1;
    }
    const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) {
      g->exec = nullptr;
      cudaGetLastError();
      g->seen = 0;
      return -1;
    }
  }
  memcpy(c->g_hin.p, pars, (size_t)n * 5 * sizeof(double));
  CK(cudaEventRecord(c->ev0, c->gstream));
  CK(cudaGraphLaunch(g->exec, c->gstream));
  CK(cudaEventRecord(c->ev1, c->gstream));
  c->timed = true;
  c->launches += c->graph_launches;
  CK(cudaStreamSynchronize(c->gstream));
  memcpy(out_lnlike, c->g_hout.p, (size_t)n * sizeof(double));
  if (out_status) memcpy(out_status, c->g_hst.p, (size_t)n * sizeof(int));
  return 0;
}

}  // namespace

__global__ void __launch_bounds__(256) fill_column_kernel(double* __restrict__ p, long long n, double v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}

int mbb_set_fixed_params(mbb_ctx* c, const int32_t* fixed, const double* values) {
  if (!c) return fail("null context");
  unsigned mask = 0;
  double val[5] = {0, 0, 0, 0, 0};
  if (fixed) {
    if (!values) return fail("null values");
    for (int i = 0; i < 5; ++i)
      if (fixed[i]) {
        if (!std::isfinite(values[i])) return fail("fixed parameter value is not finite");
        mask |= 1u << i;
        val[i] = values[i];
      }
  }
  Use u(c);
  c->fixed_mask = mask;
  for (int i = 0; i < 5; ++i) c->fixed_val[i] = val[i];
  c->fixed_epoch += 1;
  return 0;
}

// Host path: the batch is cut into chunks that flow through three slots, each
// with its own stream: H2D copy, kernel and D2H copy of consecutive chunks
// overlap (two copy engines + SMs).  Caller memory that is already pinned
// (cudaHostAlloc / cudaHostRegister / torch pin_memory) is copied from and to
// directly; pageable memory is staged through the slot's pinned buffers.
int mbb_loglike(mbb_ctx* c, int64_t n, const double* pars, int layout, const int32_t* src_index,
                int64_t walkers_per_source, double* out_lnlike, int32_t* out_status, int mem) {
  if (!c) return fail("null context");
  if (n < 0) return fail("n is negative");
  if (n == 0) return 0;
  if (!pars || !out_lnlike) return fail("null pars/out pointer");
  if (!c->bands_set) return fail("mbb_set_bands has not been called");
  if (!c->has_ivar && !c->has_cinv) return fail("mbb_set_data has not been called");
  if (c->data_nb != c->nb) return fail("data has a different number of bands than the band table");
  if (layout != MBB_AOS && layout != MBB_SOA) return fail("bad layout");
  if (!src_index) {
    if (walkers_per_source <= 0) return fail("walkers_per_source must be positive");
    if ((n + walkers_per_source - 1) / walkers_per_source > c->nsrc)
      return fail("n / walkers_per_source exceeds the number of sources set");
  }
  Use u(c);
  const long long wps = walkers_per_source > 0 ? walkers_per_source : 1;
  if (mem == MBB_DEVICE) {
    EvalArgs a{};
    a.n = n; a.e0 = 0; a.wps = wps; a.layout = layout;
    a.pars = pars; a.src_index = src_index; a.out = out_lnlike; a.status = out_status;
    begin_timing(c);
    int rc = launch_loglike(c, c->stream, a);
    end_timing(c);
    return rc;
  }
  if (src_index)
    for (int64_t i = 0; i < n; ++i)
      if (src_index[i] < 0 || src_index[i] >= c->nsrc) return fail("src_index out of range");
  if (!src_index && n <= mbb_ctx::kGraphMaxN) {
    const int rc = loglike_small_graph(c, n, pars, layout, wps, out_lnlike, out_status);
    if (rc >= 0) return rc;
  }
  if (ensure_slots(c)) return 1;
  const bool pin_in = is_pinned(pars) && is_pinned(src_index);
  const bool pin_out = is_pinned(out_lnlike) && is_pinned(out_status);
  // chunk = at most host_chunk() evaluations, but at least ~8 chunks per call once the batch is
  // worth pipelining (>= 2^17), so that copies and kernels of neighbouring chunks overlap also for
  // node-heavy configurations with few evaluations; a multiple of the 256-evaluation tile
  int64_t CH = host_chunk();
  if (n < CH * 8 && n >= ((int64_t)1 << 17)) {
    CH = ((n + 7) / 8 + 255) & ~(int64_t)255;
    if (CH < ((int64_t)1 << 14)) CH = (int64_t)1 << 14;
  }
  if (CH > n) CH = n;
  const int64_t nchunks = (n + CH - 1) / CH;
  for (auto& sl : c->slots) sl.pend_e0 = -1;
  auto drain = [&](mbb_ctx::Slot& sl) -> int {
    CK(cudaStreamSynchronize(sl.s));
    if (sl.pend_e0 >= 0 && !pin_out) {
      memcpy(out_lnlike + sl.pend_e0, sl.hout.p, (size_t)sl.pend_n * sizeof(double));
      if (out_status) memcpy(out_status + sl.pend_e0, sl.hst.p, (size_t)sl.pend_n * sizeof(int));
    }
    sl.pend_e0 = -1;
    return 0;
  };
  begin_timing(c);
  for (int64_t k = 0; k < nchunks; ++k) {
    mbb_ctx::Slot& sl = c->slots[k % 3];
    if (drain(sl)) return 1;
    const int64_t e0 = k * CH;
    const int64_t m = (n - e0) < CH ? (n - e0) : CH;
    CK(sl.din.reserve((size_t)CH * 5));
    CK(sl.dout.reserve((size_t)CH));
    CK(sl.dst.reserve((size_t)CH));
    if (k == 0) CK(cudaStreamWaitEvent(sl.s, c->ev0, 0));
    // ---- inputs
    int64_t col_stride = m;
    if (layout == MBB_AOS) {
      sl.fill_epoch = 0;          // overwrites whatever fixed columns the block held
      const double* src = pars + e0 * 5;
      if (!pin_in) {
        CK(sl.hin.reserve((size_t)CH * 5));
        memcpy(sl.hin.p, src, (size_t)m * 5 * sizeof(double));
        src = sl.hin.p;
      }
      CK(cudaMemcpyAsync(sl.din.p, src, (size_t)m * 5 * sizeof(double), cudaMemcpyHostToDevice, sl.s));
    } else {
      if (!pin_in) CK(sl.hin.reserve((size_t)CH * 5));
      // with fixed columns the device block keeps the stride CH for every chunk, so that a
      // column filled once stays valid for all chunks this slot serves
      col_stride = c->fixed_mask ? CH : m;
      const bool filled = c->fixed_mask && sl.fill_epoch == c->fixed_epoch && sl.fill_stride == CH &&
                          sl.fill_base == sl.din.p;
      for (int i = 0; i < 5; ++i) {
        double* dcol = sl.din.p + (size_t)i * col_stride;
        if ((c->fixed_mask >> i) & 1u) {
          if (!filled) {
            fill_column_kernel<<<(unsigned)((CH + 2047) / 2048), 256, 0, sl.s>>>(dcol, CH, c->fixed_val[i]);
            c->launches += 1;
          }
          continue;
        }
        const double* src = pars + (size_t)i * n + e0;
        if (!pin_in) {
          memcpy(sl.hin.p + (size_t)i * m, src, (size_t)m * sizeof(double));
          src = sl.hin.p + (size_t)i * m;
        }
        CK(cudaMemcpyAsync(dcol, src, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, sl.s));
      }
      sl.fill_epoch = c->fixed_mask ? c->fixed_epoch : 0;
      sl.fill_stride = CH;
      sl.fill_base = sl.din.p;
    }
    EvalArgs a{};
    a.n = m; a.e0 = e0; a.wps = wps; a.layout = layout;
    a.soa_stride = layout == MBB_SOA ? col_stride : 0;
    a.pars = sl.din.p; a.src_index = nullptr; a.out = sl.dout.p; a.status = sl.dst.p;
    if (src_index) {
      CK(sl.dsrc.reserve((size_t)CH));
      const int* src = src_index + e0;
      if (!pin_in) {
        CK(sl.hsrc.reserve((size_t)CH));
        memcpy(sl.hsrc.p, src, (size_t)m * sizeof(int));
        src = sl.hsrc.p;
      }
      CK(cudaMemcpyAsync(sl.dsrc.p, src, (size_t)m * sizeof(int), cudaMemcpyHostToDevice, sl.s));
      a.src_index = sl.dsrc.p;
    }
    // ---- kernel
    if (launch_loglike(c, sl.s, a)) return 1;
    // ---- outputs
    double* dsto = out_lnlike + e0;
    int* dsts = out_status ? out_status + e0 : nullptr;
    if (!pin_out) {
      CK(sl.hout.reserve((size_t)CH));
      CK(sl.hst.reserve((size_t)CH));
      dsto = sl.hout.p;
      dsts = out_status ? sl.hst.p : nullptr;
      sl.pend_e0 = e0;
      sl.pend_n = m;
    }
    CK(cudaMemcpyAsync(dsto, sl.dout.p, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, sl.s));
    if (dsts) CK(cudaMemcpyAsync(dsts, sl.dst.p, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost, sl.s));
    CK(cudaEventRecord(sl.done, sl.s));
  }
  for (auto& sl : c->slots) {
    if (drain(sl)) return 1;
  }
  end_timing(c);
  return 0;
}

int mbb_fnu(mbb_ctx* c, int64_t n, const double* pars, int layout, int nfreq, const double* freq_ghz,
            int scalar_path, int math_mode, double* out, int32_t* out_status, int mem) {
  if (!c) return fail("null context");
  if (n <= 0 || nfreq <= 0) return fail("n and nfreq must be positive");
  if (n > 65535) return fail("mbb_fnu: at most 65535 parameter vectors per call");
  if (!pars || !freq_ghz || !out) return fail("null pointer");
  if (layout != MBB_AOS && layout != MBB_SOA) return fail("bad layout");
  if (math_mode < 0) math_mode = c->math_mode;
  if (math_mode != MBB_MATH_FAITHFUL && math_mode != MBB_MATH_FAST && math_mode != MBB_MATH_FAST_GAUSS)
    return fail("bad math mode");
  const bool fast = math_mode != MBB_MATH_FAITHFUL;
  Use u(c);
  // FAST: node records of the frequency grid, built as mbb_set_bands builds a delta band's
  if (fast) {
    std::vector<double> hf((size_t)nfreq);
    if (mem == MBB_DEVICE) {
      CK(cudaMemcpyAsync(hf.data(), freq_ghz, (size_t)nfreq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    } else {
      memcpy(hf.data(), freq_ghz, (size_t)nfreq * sizeof(double));
    }
    std::vector<double> rec((size_t)nfreq * 2);
    double numax = 0.0, lmax = 0.0;
    for (int i = 0; i < nfreq; ++i) {
      if (!(hf[i] > 0.0) || !std::isfinite(hf[i])) return fail("mbb_fnu: frequencies must be positive and finite");
      const FastNode nd = fast_node(kUmToGHz / hf[i], 1.0, c->wavenorm, c->opthin != 0);
      rec[i] = nd.weff;
      rec[(size_t)nfreq + i] = nd.lp;
      if (hf[i] > numax) numax = hf[i];
      if (nd.labs > lmax) lmax = nd.labs;
    }
    CK(c->d_aux0.reserve(rec.size()));
    CK(cudaMemcpyAsync(c->d_aux0.p, rec.data(), rec.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));      // rec goes out of scope
  }
  EvalArgs a{};
  a.n = n;
  a.wps = 1;
  a.layout = layout;
  a.src_index = nullptr;
  const double* dfreq = freq_ghz;
  const size_t no = (size_t)n * nfreq;
  if (mem == MBB_DEVICE) {
    a.pars = pars;
    a.out = out;
    a.status = out_status;
  } else {
    CK(c->h_in.reserve((size_t)n * 5 + nfreq));
    CK(c->d_in.reserve((size_t)n * 5 + nfreq));
    CK(c->h_out.reserve(no));
    CK(c->d_out.reserve(no));
    CK(c->h_st.reserve((size_t)n));
    CK(c->d_st.reserve((size_t)n));
    memcpy(c->h_in.p, pars, (size_t)n * 5 * sizeof(double));
    memcpy(c->h_in.p + n * 5, freq_ghz, (size_t)nfreq * sizeof(double));
    CK(cudaMemcpyAsync(c->d_in.p, c->h_in.p, ((size_t)n * 5 + nfreq) * sizeof(double),
                       cudaMemcpyHostToDevice, c->stream));
    a.pars = c->d_in.p;
    dfreq = c->d_in.p + n * 5;
    a.out = c->d_out.p;
    a.status = c->d_st.p;
  }
  begin_timing(c);
  if (fast)
    dispatch3<LaunchFnuFast>(c->opthin != 0, c->noalpha == 0, false, c, a, dfreq, (const double*)c->d_aux0.p,
                             (const double*)(c->d_aux0.p + nfreq), nfreq);
  else
    dispatch3<LaunchFnu>(c->opthin != 0, c->noalpha == 0, false, c, a, dfreq, nfreq, scalar_path);
  end_timing(c);
  c->launches += 1;
  CK(cudaGetLastError());
  if (mem != MBB_DEVICE) {
    CK(cudaMemcpyAsync(c->h_out.p, c->d_out.p, no * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(c->h_st.p, c->d_st.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memcpy(out, c->h_out.p, no * sizeof(double));
    if (out_status) memcpy(out_status, c->h_st.p, (size_t)n * sizeof(int));
  }
  return 0;
}

int mbb_sed_consts(mbb_ctx* c, int64_t n, const double* pars, int layout, int want_peak,
                   double* out_consts, int32_t* out_status, int mem) {
  if (!c) return fail("null context");
  if (n <= 0) return fail("n must be positive");
  if (!pars || !out_consts) return fail("null pointer");
  if (layout != MBB_AOS && layout != MBB_SOA) return fail("bad layout");
  Use u(c);
  EvalArgs a{};
  a.n = n;
  a.wps = 1;
  a.layout = layout;
  a.src_index = nullptr;
  if (mem == MBB_DEVICE) {
    a.pars = pars;
    a.out = out_consts;
    a.status = out_status;
  } else {
    CK(c->h_in.reserve((size_t)n * 5));
    CK(c->d_in.reserve((size_t)n * 5));
    CK(c->h_out.reserve((size_t)n * 6));
    CK(c->d_out.reserve((size_t)n * 6));
    CK(c->h_st.reserve((size_t)n));
    CK(c->d_st.reserve((size_t)n));
    memcpy(c->h_in.p, pars, (size_t)n * 5 * sizeof(double));
    CK(cudaMemcpyAsync(c->d_in.p, c->h_in.p, (size_t)n * 5 * sizeof(double), cudaMemcpyHostToDevice,
                       c->stream));
    a.pars = c->d_in.p;
    a.out = c->d_out.p;
    a.status = c->d_st.p;
  }
  begin_timing(c);
  dispatch3<LaunchConsts>(c->opthin != 0, c->noalpha == 0, false, c, a, want_peak);
  end_timing(c);
  c->launches += 1;
  CK(cudaGetLastError());
  if (mem != MBB_DEVICE) {
    CK(cudaMemcpyAsync(c->h_out.p, c->d_out.p, (size_t)n * 6 * sizeof(double), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaMemcpyAsync(c->h_st.p, c->d_st.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    memcpy(out_consts, c->h_out.p, (size_t)n * 6 * sizeof(double));
    if (out_status) memcpy(out_status, c->h_st.p, (size_t)n * sizeof(int));
  }
  return 0;
}

int mbb_chain_post(mbb_ctx* c, int64_t nwalkers, int64_t nsteps, const double* chain, int which,
                   double z, double dl_mpc, double lir_min_um, double lir_max_um, double kappa,
                   double kappa_wave_um, double* out_peak, double* out_lir, double* out_dustmass,
                   int32_t* out_status, int mem) {
  if (!c) return fail("null context");
  if (nwalkers <= 0 || nsteps <= 0) return fail("empty chain");
  if (!chain) return fail("chain is NULL");
  const int64_t ns = nwalkers * nsteps;
  if (ns >= (int64_t)1 << 31) return fail("chain too long for one call (>= 2^31 samples); shard it");
  if ((which & 1) && !out_peak) return fail("out_peak is NULL");
  if ((which & 2) && !out_lir) return fail("out_lir is NULL");
  if ((which & 4) && !out_dustmass) return fail("out_dustmass is NULL");
  if ((which & 2) && (!(lir_min_um > 0.0) || !(lir_max_um > 0.0))) return fail("L_IR limits must be positive");
  Use u(c);
  const bool host = mem != MBB_DEVICE;
  const double* dchain = chain;
  double *dpk = out_peak, *ddm = out_dustmass, *dlir = out_lir;
  int* dst = out_status;
  // host chains of >= 2^20 samples go through the three slots in chunks of walker rows (the
  // dedupe rule never crosses a walker): the copies of one chunk run behind the kernels of its
  // neighbours
  static const bool no_pipe = getenv("MBB_B200_NO_CHAIN_PIPELINE") != nullptr;
  const int64_t nchunks = (host && !no_pipe && nwalkers >= 8 && ns >= ((int64_t)1 << 20)) ? 8 : 1;
  if (host) {
    CK(c->d_in.reserve((size_t)ns * 5));
    dchain = c->d_in.p;
    if (which & 1) { CK(c->d_aux0.reserve((size_t)ns)); dpk = c->d_aux0.p; }
    if (which & 4) { CK(c->d_aux1.reserve((size_t)ns)); ddm = c->d_aux1.p; }
    if (which & 2) { CK(c->d_out.reserve((size_t)ns)); dlir = c->d_out.p; }
    CK(c->d_st.reserve((size_t)ns));
    dst = c->d_st.p;
  }
  CK(c->d_owner.reserve((size_t)ns));
  CK(c->d_work.reserve((size_t)ns));
  CK(c->d_count.reserve((size_t)nchunks));
  CK(cudaMemsetAsync(c->d_count.p, 0, (size_t)nchunks * sizeof(unsigned), c->stream));
  // results.compute_dustmass precomputation (results.py:778-793)
  DustConsts dc;
  {
    const double dl = dl_mpc * 3.0856775814913673e24;
    dc.dl2 = dl * dl;
    dc.opz = 1.0 + z;
    const double wavenorm_rest = c->wavenorm / dc.opz;
    const double nunorm_rest = 299792458e6 / wavenorm_rest;
    dc.temp_fac = 6.6260693e-27 * nunorm_rest / 1.38065e-16;
    dc.bnu_fac = 2 * 6.6260693e-27 * pow(nunorm_rest, 3) / pow(299792458e2, 2);
    dc.knu_fac = wavenorm_rest / kappa_wave_um;
    dc.kappa = kappa;
    dc.wavenorm = c->wavenorm;
    dc.opthin = c->opthin;
  }
  // mbb_freqint (results.py:1310-1326) and freq_integrate (modified_blackbody.py:658-674)
  double lo = lir_min_um, hi = lir_max_um;
  if (lo > hi) { double t = lo; lo = hi; hi = t; }
  const double fmin = kUmToGHz / (hi * (1.0 + z)), fmax = kUmToGHz / (lo * (1.0 + z));
  const double prefac = dl_mpc > 0.0 ? 3.11749657e4 * (dl_mpc * dl_mpc) : 1.0;
  const bool thin = c->opthin != 0, alpha = c->noalpha == 0;
  // walker rows [r0, r1) on stream s: dedupe, unique samples, repeats filled in
  auto rows = [&](cudaStream_t s, int64_t r0, int64_t r1, unsigned* count) {
    const int64_t off = r0 * nsteps, m = (r1 - r0) * nsteps;
    const double* ch = dchain + off * 5;
    int *own = c->d_owner.p + off, *work = c->d_work.p + off, *st = dst ? dst + off : nullptr;
    double *pk = (which & 1) ? dpk + off : nullptr, *li = (which & 2) ? dlir + off : nullptr,
           *dm = (which & 4) ? ddm + off : nullptr;
    chain_dedupe_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(ch, r1 - r0, nsteps, own, work, count);
    const unsigned grid = (unsigned)((m + 127) / 128);
    chain_unique_kernel<<<grid, 128, 0, s>>>(ch, work, count, which, dc, pk, dm, st);
    if (which & 2) {
#define LIR(K, T, A) K<T, A><<<grid, 128, 0, s>>>(ch, work, count, c->wavenorm, fmin, fmax, prefac, li, st)
      if (c->lir_method == MBB_LIR_GAUSS) {
        if (thin) { if (alpha) LIR(chain_lir_kernel, true, true); else LIR(chain_lir_kernel, true, false); }
        else { if (alpha) LIR(chain_lir_kernel, false, true); else LIR(chain_lir_kernel, false, false); }
      } else {
        if (thin) { if (alpha) LIR(chain_lir_qags_kernel, true, true); else LIR(chain_lir_qags_kernel, true, false); }
        else { if (alpha) LIR(chain_lir_qags_kernel, false, true); else LIR(chain_lir_qags_kernel, false, false); }
      }
#undef LIR
      c->launches += 1;
    }
    chain_fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(own, r1 - r0, nsteps, pk, li, dm, st);
    c->launches += 3;
  };
  // host <-> device copies of rows [r0, r1) on stream s
  auto rows_in = [&](cudaStream_t s, int64_t r0, int64_t r1) -> int {
    const size_t off = (size_t)(r0 * nsteps), m = (size_t)((r1 - r0) * nsteps);
    CK(cudaMemcpyAsync(c->d_in.p + off * 5, chain + off * 5, m * 5 * sizeof(double), cudaMemcpyHostToDevice, s));
    return 0;
  };
  auto rows_out = [&](cudaStream_t s, int64_t r0, int64_t r1) -> int {
    const size_t off = (size_t)(r0 * nsteps), m = (size_t)((r1 - r0) * nsteps);
    if (which & 1) CK(cudaMemcpyAsync(out_peak + off, dpk + off, m * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (which & 2) CK(cudaMemcpyAsync(out_lir + off, dlir + off, m * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (which & 4) CK(cudaMemcpyAsync(out_dustmass + off, ddm + off, m * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (out_status) CK(cudaMemcpyAsync(out_status + off, dst + off, m * sizeof(int), cudaMemcpyDeviceToHost, s));
    return 0;
  };
  if (nchunks > 1) {
    if (ensure_slots(c)) return 1;
    begin_timing(c);
    for (int64_t k = 0; k < nchunks; ++k) {
      mbb_ctx::Slot& sl = c->slots[k % 3];
      const int64_t r0 = nwalkers * k / nchunks, r1 = nwalkers * (k + 1) / nchunks;
      if (k < 3) CK(cudaStreamWaitEvent(sl.s, c->ev0, 0));
      if (rows_in(sl.s, r0, r1)) return 1;
      rows(sl.s, r0, r1, c->d_count.p + k);
      if (rows_out(sl.s, r0, r1)) return 1;
    }
    CK(cudaGetLastError());
    for (auto& sl : c->slots) CK(cudaStreamSynchronize(sl.s));
    end_timing(c);
    return 0;
  }
  if (host && rows_in(c->stream, 0, nwalkers)) return 1;
  begin_timing(c);
  rows(c->stream, 0, nwalkers, c->d_count.p);
  end_timing(c);
  CK(cudaGetLastError());
  if (host) {
    if (rows_out(c->stream, 0, nwalkers)) return 1;
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int mbb_chain_flux(mbb_ctx* c, int64_t nwalkers, int64_t nsteps, const double* chain, int band,
                   double* out_flux, int32_t* out_status, int mem) {
  if (!c) return fail("null context");
  if (nwalkers <= 0 || nsteps <= 0) return fail("empty chain");
  if (!chain || !out_flux) return fail("null pointer");
  if (!c->bands_set) return fail("mbb_set_bands has not been called");
  if (band < 0 || band >= c->nb) return fail("band index out of range");
  const int64_t ns = nwalkers * nsteps;
  if (ns >= (int64_t)1 << 31) return fail("chain too long for one call (>= 2^31 samples); shard it");
  Use u(c);
  const double* dchain = chain;
  double* dout = out_flux;
  int* dst = out_status;
  if (mem != MBB_DEVICE) {
    CK(c->d_in.reserve((size_t)ns * 5));
    CK(cudaMemcpyAsync(c->d_in.p, chain, (size_t)ns * 5 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dchain = c->d_in.p;
    CK(c->d_out.reserve((size_t)ns));
    dout = c->d_out.p;
    CK(c->d_st.reserve((size_t)ns));
    dst = c->d_st.p;
  }
  CK(c->d_owner.reserve((size_t)ns));
  CK(c->d_work.reserve((size_t)ns));
  CK(c->d_count.reserve(1));
  CK(cudaMemsetAsync(c->d_count.p, 0, sizeof(unsigned), c->stream));
  begin_timing(c);
  chain_dedupe_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, c->stream>>>(
      dchain, nwalkers, nsteps, c->d_owner.p, c->d_work.p, c->d_count.p);
  const unsigned grid = (unsigned)((ns + 127) / 128);
  const int i0 = c->h_off[band], i1 = c->h_off[band + 1], sp = c->h_scalar[band];
  const bool thin = c->opthin != 0, alpha = c->noalpha == 0;
#define FLX(T, A) chain_flux_kernel<T, A><<<grid, 128, 0, c->stream>>>(dchain, c->d_work.p, c->d_count.p, \
                                    c->wavenorm, c->d_node_fw.p, i0, i1, sp, dout, dst)
  if (thin) { if (alpha) FLX(true, true); else FLX(true, false); }
  else { if (alpha) FLX(false, true); else FLX(false, false); }
#undef FLX
  chain_fill_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, c->stream>>>(c->d_owner.p, nwalkers, nsteps, dout,
                                                                         nullptr, nullptr, dst);
  end_timing(c);
  c->launches += 3;
  CK(cudaGetLastError());
  if (mem != MBB_DEVICE) {
    CK(cudaMemcpyAsync(out_flux, dout, (size_t)ns * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (out_status)
      CK(cudaMemcpyAsync(out_status, dst, (size_t)ns * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

int mbb_chain_stats(mbb_ctx* c, int64_t n, int ncols, const double* x, const double* lowlim,
                    const double* uplim, int nq, const double* q, double* mean, int64_t* count,
                    double* q_lo, double* q_hi, double* q_gamma, int mem) {
  if (!c) return fail("null context");
  if (n <= 0 || ncols <= 0 || ncols > kStatMaxCols) return fail("n must be positive and 1 <= ncols <= 8");
  if (nq < 0 || 2 * nq > kStatMaxSel) return fail("at most 4 quantiles per call");
  if (!x || !mean || !count || (nq > 0 && (!q || !q_lo || !q_hi || !q_gamma))) return fail("null pointer");
  if (n >= (int64_t)1 << 32) return fail("more than 2^32 rows; shard the chain");
  for (int i = 0; i < nq; ++i)
    if (!(q[i] >= 0.0 && q[i] <= 1.0)) return fail("quantiles must lie in [0, 1]");
  Use u(c);
  const long long total = (long long)n * ncols;
  const double* dx = x;
  if (mem != MBB_DEVICE) {
    CK(c->d_in.reserve((size_t)total));
    CK(cudaMemcpyAsync(c->d_in.p, x, (size_t)total * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dx = c->d_in.p;
  }
  StatCols sc;
  sc.ncols = ncols;
  for (int k = 0; k < kStatMaxCols; ++k) {
    sc.lo[k] = (lowlim && k < ncols) ? lowlim[k] : -kInf;
    sc.hi[k] = (uplim && k < ncols) ? uplim[k] : kInf;
  }
  const long long want = (n + 255) / 256, cap = (long long)c->sm_count * 8;
  const unsigned gx = (unsigned)(want < cap ? want : cap);
  const dim3 grid(gx, (unsigned)ncols);
  CK(c->d_stat_cnt.reserve(2 * kStatMaxCols));
  CK(c->d_stat_part.reserve((size_t)ncols * gx));
  CK(c->d_stat_hist.reserve((size_t)kStatMaxCols * kStatMaxSel * 256));
  CK(cudaMemsetAsync(c->d_stat_cnt.p, 0, 2 * kStatMaxCols * sizeof(unsigned long long), c->stream));
  begin_timing(c);
  chain_moments_kernel<<<grid, 256, 0, c->stream>>>(dx, total, sc, c->d_stat_cnt.p, c->d_stat_cnt.p + kStatMaxCols,
                                                    c->d_stat_part.p);
  c->launches += 1;
  unsigned long long h_cnt[2 * kStatMaxCols];
  std::vector<double> h_part((size_t)ncols * gx);
  CK(cudaMemcpyAsync(h_cnt, c->d_stat_cnt.p, sizeof(h_cnt), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(h_part.data(), c->d_stat_part.p, h_part.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const double nan = kInf - kInf;
  // wanted ranks: numpy's default 'linear' quantile -- virtual index (count - 1) q formed as
  // count q + (1 + q (-1)) - 1, neighbours floor and floor + 1, clipped to the ends
  struct Sel { unsigned long long rank, prefix, rem; };
  Sel sel[kStatMaxCols][kStatMaxSel];
  int nsel[kStatMaxCols];
  for (int col = 0; col < ncols; ++col) {
    const unsigned long long cnt = h_cnt[col], bad = h_cnt[kStatMaxCols + col];
    count[col] = (int64_t)cnt;
    long double acc = 0.0L;
    for (unsigned b = 0; b < gx; ++b) acc += (long double)h_part[(size_t)col * gx + b];
    mean[col] = (cnt && !bad) ? (double)(acc / (long double)cnt) : nan;
    nsel[col] = 0;
    for (int i = 0; i < nq; ++i) {
      q_lo[col * nq + i] = q_hi[col * nq + i] = q_gamma[col * nq + i] = nan;
      if (!cnt || bad) continue;                      // numpy: NaN in -> NaN out; empty -> caller's error
      const double vi = (double)cnt * q[i] + (1.0 + q[i] * (-1.0)) - 1.0;
      double prev = floor(vi);
      q_gamma[col * nq + i] = vi - prev;
      double next = prev + 1.0;
      if (vi >= (double)(cnt - 1)) prev = next = (double)(cnt - 1);
      if (vi < 0.0) prev = next = 0.0;
      sel[col][nsel[col]++] = Sel{(unsigned long long)prev, 0ull, (unsigned long long)prev};
      sel[col][nsel[col]++] = Sel{(unsigned long long)next, 0ull, (unsigned long long)next};
    }
  }
  std::vector<unsigned> h_hist((size_t)kStatMaxCols * kStatMaxSel * 256);
  for (int shift = 56; shift >= 0 && nq > 0; shift -= 8) {
    StatPass ps;
    memset(&ps, 0, sizeof(ps));
    ps.shift = shift;
    int group_of[kStatMaxCols][kStatMaxSel];
    bool any = false;
    for (int col = 0; col < ncols; ++col) {
      int ng = 0;
      for (int s = 0; s < nsel[col]; ++s) {
        int g = -1;
        for (int k = 0; k < ng; ++k)
          if (ps.prefix[col][k] == sel[col][s].prefix) g = k;
        if (g < 0) {
          g = ng++;
          ps.prefix[col][g] = sel[col][s].prefix;
        }
        group_of[col][s] = g;
      }
      ps.ngroups[col] = ng;
      any = any || ng > 0;
    }
    if (!any) break;
    CK(cudaMemsetAsync(c->d_stat_hist.p, 0, h_hist.size() * sizeof(unsigned), c->stream));
    chain_select_hist_kernel<<<grid, 256, 0, c->stream>>>(dx, total, sc, ps, c->d_stat_hist.p);
    c->launches += 1;
    CK(cudaMemcpyAsync(h_hist.data(), c->d_stat_hist.p, h_hist.size() * sizeof(unsigned), cudaMemcpyDeviceToHost,
                       c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int col = 0; col < ncols; ++col)
      for (int s = 0; s < nsel[col]; ++s) {
        const unsigned* h = h_hist.data() + ((size_t)col * kStatMaxSel + group_of[col][s]) * 256;
        unsigned long long rem = sel[col][s].rem;
        int b = 0;
        for (; b < 255; ++b) {
          if (rem < h[b]) break;
          rem -= h[b];
        }
        if (rem >= h[b]) return fail("mbb_chain_stats: rank outside the histogram (internal error)");
        sel[col][s].rem = rem;
        sel[col][s].prefix = (sel[col][s].prefix << 8) | (unsigned long long)b;
      }
  }
  end_timing(c);
  CK(cudaGetLastError());
  for (int col = 0; col < ncols; ++col)
    for (int s = 0; s < nsel[col]; ++s) {
      const double v = stat_unkey(sel[col][s].prefix);
      (s & 1 ? q_hi : q_lo)[col * nq + s / 2] = v;
    }
  return 0;
}

// see include/mbb_b200.h
int mbb_ensemble_fit(mbb_ctx* c, int64_t nsrc, int nwalkers, int64_t nburn, int64_t nsteps, double a,
                     uint64_t seed, uint64_t step0, int64_t src0, double* pos, double* lnprob, int have_lnprob,
                     int32_t* naccept, int32_t* status, double* stats, double* chain, double* chain_lnprob,
                     int64_t chain_nsrc, int thin, int mem) {
  if (!c) return fail("null context");
  if (chain_nsrc <= 0) chain_nsrc = nsrc;
  if (chain_nsrc < nsrc) return fail("chain_nsrc is smaller than nsrc");
  if (nsrc <= 0 || nsteps < 0 || nburn < 0) return fail("nsrc must be positive, nburn / nsteps non-negative");
  if (nwalkers < 2 || (nwalkers & 1)) return fail("the number of walkers must be even");
  if (nwalkers <= 10) return fail("need more than 2*dim = 10 walkers");
  if (!(a > 1.0)) return fail("stretch scale a must be > 1");
  if (!pos || !lnprob) return fail("null pos/lnprob pointer");
  if (src0 < 0) return fail("src0 must be non-negative");
  if (!c->bands_set || (!c->has_ivar && !c->has_cinv)) return fail("bands/data not set");
  if (c->data_nb != c->nb) return fail("data has a different number of bands than the band table");
  if (nsrc > c->nsrc) return fail("more sources than mbb_set_data provided");
  if (nburn + nsteps >= ((int64_t)1 << 31)) return fail("too many iterations for one call");
  if (thin < 1) thin = 1;
  Use u(c);
  const int h = nwalkers / 2;
  const long long nwk = nsrc * nwalkers, nh = nsrc * h;
  if (nwk >= (1LL << 31)) return fail("too many walkers for one call (>= 2^31); shard the sources");
  const bool host = mem != MBB_DEVICE;
  const int64_t nrec_total = nsteps / thin;
  CK(ensure_tables(c));      // node tables and the cold-path block match the current model / priors
  // delta-band FAST configurations whose ensembles fit in shared memory take the source-resident kernel
  // (the switch is read per call so that tests can compare the paths in one process)
  const bool no_fuse = getenv("MBB_B200_NO_FUSED_SAMPLER") != nullptr;
  const EnsResidentPlan plan = ens_resident_plan(c, nwalkers, stats != nullptr);
  const bool resident_ok = !no_fuse && c->math_mode != MBB_MATH_FAITHFUL && c->nn == c->nb &&
                           c->nb <= kMaxDeltaNB && plan.ok;
  const StretchScale sc = stretch_scale(a);
  DataRef dref;
  dref.flux = c->d_flux.p;
  dref.ivar = c->has_ivar ? c->d_ivar.p : nullptr;
  dref.cinv = c->has_cinv ? c->d_cinv.p : nullptr;
  dref.chol = c->has_cinv && c->cinv_is_chol ? 1 : 0;
  dref.nsrc = c->nsrc;
  dref.nb = c->nb;

  // ---- host buffers, no chain: chunks of sources flow through three slots, each with its own stream, so
  // that the upload of one chunk's starting ensembles and the download of another's results run behind
  // the sampler working on a third (the resident kernel treats sources independently)
  {
    static const bool no_pipe = getenv("MBB_B200_NO_FIT_PIPELINE") != nullptr;
    const int64_t min_chunk = 2048;
    if (host && resident_ok && !chain && !chain_lnprob && !no_pipe && nsrc >= 2 * min_chunk && nburn + nsteps > 0) {
      if (ensure_slots(c)) return 1;
      int64_t nchunks = 8;
      if (nsrc / nchunks < min_chunk) nchunks = nsrc / min_chunk;
      const int64_t CH = (nsrc + nchunks - 1) / nchunks;
      begin_timing(c);
      for (int64_t k = 0; k * CH < nsrc; ++k) {
        mbb_ctx::Slot& sl = c->slots[k % 3];
        CK(cudaStreamSynchronize(sl.s));
        const int64_t s0 = k * CH, ns = (nsrc - s0) < CH ? (nsrc - s0) : CH;
        const size_t nw_c = (size_t)ns * nwalkers;
        CK(sl.epos.reserve((size_t)CH * nwalkers * 5));
        CK(sl.elnp.reserve((size_t)CH * nwalkers));
        CK(sl.enacc.reserve((size_t)CH * nwalkers));
        CK(sl.est.reserve((size_t)CH * nwalkers));
        if (stats) CK(sl.estats.reserve((size_t)CH * kFitStats));
        if (k == 0) CK(cudaStreamWaitEvent(sl.s, c->ev0, 0));
        CK(cudaMemcpyAsync(sl.epos.p, pos + (size_t)s0 * nwalkers * 5, nw_c * 5 * sizeof(double),
                           cudaMemcpyHostToDevice, sl.s));
        if (have_lnprob)
          CK(cudaMemcpyAsync(sl.elnp.p, lnprob + (size_t)s0 * nwalkers, nw_c * sizeof(double),
                             cudaMemcpyHostToDevice, sl.s));
        CK(cudaMemsetAsync(sl.enacc.p, 0, nw_c * sizeof(int), sl.s));
        CK(cudaMemsetAsync(sl.est.p, 0, nw_c * sizeof(int), sl.s));
        if (!have_lnprob) {
          EvalArgs e{};
          e.n = (long long)nw_c; e.e0 = s0 * nwalkers; e.wps = nwalkers; e.layout = MBB_AOS;
          e.pars = sl.epos.p; e.src_index = nullptr; e.out = sl.elnp.p; e.status = sl.est.p;
          if (launch_loglike(c, sl.s, e)) return 1;
        }
        EnsFit g{};
        g.pos = sl.epos.p; g.lnp = sl.elnp.p; g.nacc = sl.enacc.p; g.status = sl.est.p;
        g.stats = stats ? sl.estats.p : nullptr;
        g.nsrc = ns; g.src0 = src0 + s0; g.dsrc0 = s0; g.chain_nsrc = ns;
        g.nw = nwalkers; g.h = h; g.G = plan.G;
        g.niter = (int)(nburn + nsteps);
        g.main_from = (int)nburn;
        g.main_done = 0;
        g.thin = thin;
        g.nrec = (int)nrec_total;
        g.merge = 0;
        g.keys = philox_keys(seed);
        g.step0 = step0;
        g.sc = sc;
        cudaError_t err = cudaSuccess;
        dispatch3<LaunchEnsResident>(c->opthin != 0, c->noalpha == 0, true, c, g, dref, plan.smem, sl.s,
                                     &sl.escratch, &err);
        CK(err);
        c->launches += 1;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(pos + (size_t)s0 * nwalkers * 5, sl.epos.p, nw_c * 5 * sizeof(double),
                           cudaMemcpyDeviceToHost, sl.s));
        CK(cudaMemcpyAsync(lnprob + (size_t)s0 * nwalkers, sl.elnp.p, nw_c * sizeof(double),
                           cudaMemcpyDeviceToHost, sl.s));
        if (naccept)
          CK(cudaMemcpyAsync(naccept + (size_t)s0 * nwalkers, sl.enacc.p, nw_c * sizeof(int),
                             cudaMemcpyDeviceToHost, sl.s));
        if (status)
          CK(cudaMemcpyAsync(status + (size_t)s0 * nwalkers, sl.est.p, nw_c * sizeof(int),
                             cudaMemcpyDeviceToHost, sl.s));
        if (stats)
          CK(cudaMemcpyAsync(stats + (size_t)s0 * kFitStats, sl.estats.p, (size_t)ns * kFitStats * sizeof(double),
                             cudaMemcpyDeviceToHost, sl.s));
      }
      for (auto& sl : c->slots) CK(cudaStreamSynchronize(sl.s));
      end_timing(c);
      return 0;
    }
  }

  double *dpos = pos, *dlnp = lnprob, *dstats = stats;
  int *dnacc = naccept, *dst = status;
  if (host) {
    CK(c->d_epos.reserve((size_t)nwk * 5));
    CK(c->d_elnp.reserve((size_t)nwk));
    CK(cudaMemcpyAsync(c->d_epos.p, pos, (size_t)nwk * 5 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (have_lnprob)
      CK(cudaMemcpyAsync(c->d_elnp.p, lnprob, (size_t)nwk * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    dpos = c->d_epos.p;
    dlnp = c->d_elnp.p;
    dnacc = nullptr;
    dst = nullptr;
    if (stats) { CK(c->d_estats.reserve((size_t)nsrc * kFitStats)); dstats = c->d_estats.p; }
  }
  if (!dnacc) { CK(c->d_enacc.reserve((size_t)nwk)); dnacc = c->d_enacc.p; }
  if (!dst) { CK(c->d_est.reserve((size_t)nwk)); dst = c->d_est.p; }
  CK(cudaMemsetAsync(dnacc, 0, (size_t)nwk * sizeof(int), c->stream));
  CK(cudaMemsetAsync(dst, 0, (size_t)nwk * sizeof(int), c->stream));
  begin_timing(c);
  if (!have_lnprob) {
    // log-probability of the starting ensemble (emcee computes it once up front)
    EvalArgs e{};
    e.n = nwk; e.e0 = 0; e.wps = nwalkers; e.layout = MBB_AOS;
    e.pars = dpos; e.src_index = nullptr; e.out = dlnp; e.status = dst;
    if (launch_loglike(c, c->stream, e)) return 1;
  }
  const bool aligned = ((uintptr_t)dpos % 16 == 0) && ((uintptr_t)dlnp % 16 == 0) &&
                       (host || (((uintptr_t)chain % 16 == 0) && ((uintptr_t)chain_lnprob % 16 == 0)));
  const bool resident = resident_ok && aligned;
  if (!resident) {
    CK(c->d_eq.reserve((size_t)nh * 5));
    CK(c->d_eqlnp.reserve((size_t)nh));
    CK(c->d_eqst.reserve((size_t)nh));
  }
  // One segment = iterations [i0, i1) of the call (burn-in first); its records go to cdst / ldst.
  auto run_segment = [&](int64_t i0, int64_t i1, double* cdst, double* ldst, int64_t cns) -> int {
    const int64_t main_done = i0 > nburn ? i0 - nburn : 0;
    const int64_t main_here = i1 > nburn ? i1 - (i0 > nburn ? i0 : nburn) : 0;
    const int64_t nrec = (main_done + main_here) / thin - main_done / thin;
    if (resident) {
      EnsFit g{};
      g.pos = dpos; g.lnp = dlnp; g.nacc = dnacc; g.status = dst; g.stats = dstats;
      g.chain = cdst; g.chain_lnp = ldst;
      g.nsrc = nsrc; g.src0 = src0; g.dsrc0 = 0; g.chain_nsrc = cns; g.nw = nwalkers; g.h = h; g.G = plan.G;
      g.niter = (int)(i1 - i0);
      g.main_from = (int)(nburn > i0 ? (nburn - i0 < i1 - i0 ? nburn - i0 : i1 - i0) : 0);
      g.main_done = main_done;
      g.thin = thin;
      g.nrec = (int)nrec;
      g.merge = main_done > 0;
      g.keys = philox_keys(seed);
      g.step0 = step0 + (uint64_t)i0;
      g.sc = sc;
      cudaError_t err = cudaSuccess;
      dispatch3<LaunchEnsResident>(c->opthin != 0, c->noalpha == 0, true, c, g, dref, plan.smem, c->stream,
                                   &c->d_escratch, &err);
      CK(err);
      c->launches += 1;
      CK(cudaGetLastError());
      return 0;
    }
    EnsArgs g{};
    g.pos = dpos; g.lnp = dlnp; g.nacc = dnacc; g.status = dst;
    g.q = c->d_eq.p; g.qlnp = c->d_eqlnp.p; g.qst = c->d_eqst.p;
    g.nsrc = nsrc; g.src0 = src0; g.nw = nwalkers; g.h = h; g.keys = philox_keys(seed); g.sc = sc;
    const unsigned grid = (unsigned)((nh + 255) / 256);
    int64_t kept = 0;
    // small ensembles over tabulated passbands (FAST, full tables): the half-step in two launches
    const bool fused_half = getenv("MBB_B200_NO_FUSED_HALFSTEP") == nullptr &&
                            getenv("MBB_B200_NO_SMALL_NODES") == nullptr && c->math_mode == MBB_MATH_FAST &&
                            !(c->nn == c->nb && c->nb <= kMaxDeltaNB) && c->nn > kSmallMaxNodes &&
                            nh <= kSmallNodesMaxN;
    for (int64_t it = i0; it < i1; ++it) {
      const bool is_main = it >= nburn;
      g.count = is_main ? 1 : 0;
      for (int half = 0; half < 2; ++half) {
        g.half = half;
        g.hstep = 2 * (step0 + (uint64_t)it) + (uint64_t)half;
        EvalArgs e{};
        e.n = nh; e.e0 = 0; e.wps = h; e.layout = MBB_AOS;
        e.pars = g.q; e.src_index = nullptr; e.out = g.qlnp; e.status = g.qst;
        if (fused_half) {
          set_wps_division(e);
          cudaError_t err = cudaSuccess;
          dispatch3<LaunchEnsHalfStepFused>(c->opthin != 0, c->noalpha == 0, true, c, c->stream, g, e, dref, &err);
          CK(err);
          c->launches += 2;
          continue;
        }
        ens_propose_kernel<<<grid, 256, 0, c->stream>>>(g);
        if (launch_loglike(c, c->stream, e)) return 1;
        ens_accept_kernel<<<grid, 256, 0, c->stream>>>(g);
        c->launches += 2;
      }
      const int64_t jm = it - nburn;
      if (is_main && (jm + 1) % thin == 0) {
        if (dstats) {
          ens_stats_kernel<<<(unsigned)nsrc, kStatsThreads, 0, c->stream>>>(dpos, dlnp, dnacc, dstats, nwalkers,
                                                                          jm + 1 > thin ? 1 : 0, (double)(jm + 1));
          c->launches += 1;
        }
        if (cdst)
          CK(cudaMemcpyAsync(cdst + (size_t)kept * cns * nwalkers * 5, dpos, (size_t)nwk * 5 * sizeof(double),
                             cudaMemcpyDeviceToDevice, c->stream));
        if (ldst)
          CK(cudaMemcpyAsync(ldst + (size_t)kept * cns * nwalkers, dlnp, (size_t)nwk * sizeof(double),
                             cudaMemcpyDeviceToDevice, c->stream));
        ++kept;
      }
    }
    CK(cudaGetLastError());
    return 0;
  };

  const int64_t total = nburn + nsteps;
  if (!host || (!chain && !chain_lnprob) || nrec_total == 0) {
    if (total > 0 && run_segment(0, total, host ? nullptr : chain, host ? nullptr : chain_lnprob, chain_nsrc))
      return 1;
    if (total == 0 && dstats) CK(cudaMemsetAsync(dstats, 0, (size_t)nsrc * kFitStats * sizeof(double), c->stream));
  } else {
    // chain to host memory: segments of R records through two device buffers; the D2H copy of
    // one segment runs on its own stream behind the sampler working on the next
    if (!c->copy_stream) {
      CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&c->seg_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->seg_copied[i], cudaEventDisableTiming));
      }
    }
    const size_t rec_bytes = (size_t)nwk * 48;
    // (MBB_B200_CHAIN_SEGMENT_MB: segment budget in MiB, default 256; read per call so that tests can shrink it)
    const char* seg_env = getenv("MBB_B200_CHAIN_SEGMENT_MB");
    const size_t seg_mb = seg_env && atoll(seg_env) > 0 ? (size_t)atoll(seg_env) : 256;
    int64_t R = (int64_t)((seg_mb << 20) / rec_bytes);
    if (R < 1) R = 1;
    if (R > nrec_total) R = nrec_total;
    for (int i = 0; i < 2; ++i) {
      if (chain) CK(c->d_echain[i].reserve((size_t)R * nwk * 5));
      if (chain_lnprob) CK(c->d_echainl[i].reserve((size_t)R * nwk));
    }
    int64_t i0 = 0, rec0 = 0;
    for (int seg = 0; rec0 < nrec_total || i0 < total; ++seg) {
      const int b = seg & 1;
      const int64_t recs = (nrec_total - rec0) < R ? (nrec_total - rec0) : R;
      // the segment ends with its last record; the final one also takes the unrecorded tail
      int64_t i1 = nburn + (rec0 + recs) * thin;
      if (rec0 + recs >= nrec_total) i1 = total;
      if (seg >= 2) CK(cudaStreamWaitEvent(c->stream, c->seg_copied[b], 0));
      if (run_segment(i0, i1, chain ? c->d_echain[b].p : nullptr, chain_lnprob ? c->d_echainl[b].p : nullptr, nsrc))
        return 1;
      CK(cudaEventRecord(c->seg_done[b], c->stream));
      CK(cudaStreamWaitEvent(c->copy_stream, c->seg_done[b], 0));
      // one record = nsrc contiguous ensembles; the host arrays hold chain_nsrc sources per record
      if (chain && recs > 0)
        CK(cudaMemcpy2DAsync(chain + (size_t)rec0 * chain_nsrc * nwalkers * 5,
                             (size_t)chain_nsrc * nwalkers * 5 * sizeof(double), c->d_echain[b].p,
                             (size_t)nwk * 5 * sizeof(double), (size_t)nwk * 5 * sizeof(double), (size_t)recs,
                             cudaMemcpyDeviceToHost, c->copy_stream));
      if (chain_lnprob && recs > 0)
        CK(cudaMemcpy2DAsync(chain_lnprob + (size_t)rec0 * chain_nsrc * nwalkers,
                             (size_t)chain_nsrc * nwalkers * sizeof(double), c->d_echainl[b].p,
                             (size_t)nwk * sizeof(double), (size_t)nwk * sizeof(double), (size_t)recs,
                             cudaMemcpyDeviceToHost, c->copy_stream));
      CK(cudaEventRecord(c->seg_copied[b], c->copy_stream));
      i0 = i1;
      rec0 += recs;
      if (i0 >= total) break;
    }
  }
  end_timing(c);
  CK(cudaGetLastError());
  if (host) {
    CK(cudaMemcpyAsync(pos, dpos, (size_t)nwk * 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(lnprob, dlnp, (size_t)nwk * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (naccept)
      CK(cudaMemcpyAsync(naccept, dnacc, (size_t)nwk * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (status)
      CK(cudaMemcpyAsync(status, dst, (size_t)nwk * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (stats)
      CK(cudaMemcpyAsync(stats, dstats, (size_t)nsrc * kFitStats * sizeof(double), cudaMemcpyDeviceToHost,
                         c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->copy_stream) CK(cudaStreamSynchronize(c->copy_stream));
  }
  return 0;
}

int mbb_ensemble_run(mbb_ctx* c, int64_t nsrc, int nwalkers, int64_t nsteps, double a, uint64_t seed,
                     uint64_t step0, double* pos, double* lnprob, int have_lnprob, int32_t* naccept,
                     int32_t* status, double* chain, double* chain_lnprob, int thin, int mem) {
  return mbb_ensemble_fit(c, nsrc, nwalkers, 0, nsteps, a, seed, step0, 0, pos, lnprob, have_lnprob, naccept,
                          status, nullptr, chain, chain_lnprob, 0, thin, mem);
}

int mbb_fp64_peak(mbb_ctx* c, int iters, double* tflops) {
  if (!c || !tflops) return fail("null argument");
  if (iters <= 0) iters = 20000;
  Use u(c);
  CK(c->d_out.reserve(16));
  const int blocks = c->sm_count * 8, threads = 256;
  fp64_peak_kernel<<<blocks, threads, 0, c->stream>>>(c->d_out.p, 1000, 1.0);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    begin_timing(c);
    fp64_peak_kernel<<<blocks, threads, 0, c->stream>>>(c->d_out.p, iters, 1.0 + rep);
    end_timing(c);
    c->launches += 1;
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
  *tflops = flops / (best * 1e-3) / 1e12;
  return 0;
}

}  // extern "C"
