// Lean FP64 exp / expm1 / division for the FAST arithmetic mode.
//
// Why not libdevice: on sm_100a its exp/expm1/division spend more issue slots on
// integer fix-ups, special-case branches and constant materialisation than on
// FP64 work (ncu, profiles/r01_*: 27% of the executed instructions of the first
// likelihood kernel were FP64; the issue stage, not the FP64 pipe, was the
// limiter).  The family below works in units of 1/64 octave:
//
//     exp(a*b) = 2^(y/64),  y = a*b*64/ln2 = 64 m + (j - 32) + f,  |f| <= 1/2
//              = 2^m * T_j * 2^(f/64),        T_j = 2^((j-32)/64)
//
//   * the argument is handed over as a PRODUCT with the second factor already
//     scaled by 64/ln2 (node tables hold log(lambda/lambda_norm)*64/ln2,
//     per-walker constants are scaled once in the setup), so the reduction is
//     3-4 FP64 instructions for any exp in the node loop -- no separate
//     product, no Cody-Waite constants;
//   * 2^(f/64) - 1 = f g(f) with g of degree 4 (4 DFMA + 1 DMUL);
//   * no branches and, on the hot path, no clamps: the caller proves per walker
//     that every exponent stays inside the double range (`safe` flag of
//     FastSed); CLAMP=true instantiations saturate the binary exponent for the
//     rest;
//   * expm1(x) = sT p + (sT - 1) with sT = 2^m T_j: j = 32 (T = 1 exactly)
//     covers |x| < ln2/128 and sT - 1 is exact (Sterbenz) for the neighbouring
//     entries; the rounding of T_j (<= 2^-53) bounds the relative error of
//     expm1 by 2^-53 sT/|sT - 1| <= 2e-14 (at |x| ~ 0.0054), <= 2e-16 for |x| > 1;
//   * the 64-entry table is stored rotated and with pre-adjusted high words so
//     that index and scale come straight from the low word of the magic-number
//     sum: shift, mask, LDS.64, integer multiply-add.  In shared memory it is
//     replicated 16x (8 KB, copy c in bank pair c) and lane l reads copy l & 15,
//     so a lookup with 32 different indices is bank-conflict-free (2 wavefronts;
//     ncu showed the unreplicated table at ~9 wavefronts per lookup and the
//     shared-memory pipe, not the FP64 pipe, saturated);
//   * reciprocal = MUFU.RCP64H seed + one cubic correction (3 FP64), arranged
//     by the callers so that it runs beside the exp chains, not after them.
// 9-10 FP64 instructions per exp, 10-11 per expm1, against 29/35 + ~30 fix-up
// instructions in libdevice.  Accuracy (tests/test_device_logic_cpu.py, vs
// mpmath): exp <= 1.5 ulp, expm1 <= 3 ulp away from the |x| ~ 0.01 band
// described above; constants from tools/gen_poly.py.
#pragma once
#include <cmath>
#include <cstring>

#include "mbb_model_defs.cuh"

namespace mbb {

constexpr double kMagic = 6755399441055744.0;            // 1.5*2^52: low word of x+kMagic = round(x)

// Two table sizes, chosen per kernel through the TS template parameter of the
// functions below (bit 4 of TS set = 256 entries):
//   64 entries, g of degree 4   -- per-walker setup code and the delta-band kernels;
//   256 entries, g of degree 3  -- one DFMA less per exponential for a 32 KB instead of
//                                  8 KB replicated table: the node loops of the passband kernels.
// Measured on B200 (whole-library builds of either size, all GPU tests green):
// loglike_nodes_kernel cfg2 6.00 -> 5.75 ms and the Gauss-rule thread kernel 1.44 -> 1.39 ms
// with 256 entries, but loglike_delta_kernel cfg5 1.196 -> 1.223 ms (66 instead of 41 KB of
// shared memory per CTA).  "64" in the comments of this file and of mbb_model.cuh stands
// for the table size N in use; quantities "in 1/64-octave units" of the 256 flavour are the
// 64 ones times 4, exactly (fast_sed_rescale256, node tables built with that factor).
// Accuracy with 256 entries: exp unchanged (<= 1.5 ulp); expm1 near 0 <= 20 ulp (g is
// good to 4.8e-18 absolute = 3.5e-15 of f g(f)) and <= 8.4e-14 in the |x| ~ 0.0014-0.09 band.
constexpr int kTab256 = 16;                              // TS flag: the 256-entry table
template <int TS>
struct TabCfg {
  static constexpr int bits = (TS & kTab256) ? 8 : 6;
  static constexpr int n = 1 << bits;
  static constexpr int mask = n - 1;
  static constexpr int half = n / 2;
  static constexpr int hishift = 20 - bits;              // entry index -> exponent-field units
  static constexpr int stride = TS & 15;                 // log2 of the index stride (replication)
  static constexpr int deg = bits == 6 ? 4 : 3;          // degree of g, 2^(f/N) - 1 = f g(f)
  static constexpr double cscale = bits == 6 ? 1.0 : 4.0;   // N/ln2 relative to 64/ln2
};
constexpr double kC64Hi = 92.332482616893657;            // 64/ln2 as a double-double
constexpr double kC64Lo = 1.3027375194195861e-15;

// Table entry i serves the exponent residue i = k & (N-1): T_j with j = (i + N/2) & (N-1),
// as a bit pattern whose high word is pre-adjusted by 0x80000 - (j << hishift), so that
// hi + (k << hishift) is the high word of 2^m T_j, m = (k + N/2) >> bits.
// f*g(f) = 2^(f/N) - 1 on |f| <= 0.5005: max abs error 2.4e-18 (N = 64), 4.8e-18 (N = 256).
#define MBB_LEAN_G64 {0.010830424696249145, 5.8649049550517742e-05, 2.1173137155457974e-07, \
                      5.7328587073751599e-10, 1.2417854561126839e-12}
#define MBB_LEAN_G256 {0.0027076061740622767, 3.665565596910102e-06, 3.3083029843180382e-09, \
                       2.2393953279597525e-12}
#if defined(__CUDACC__)
__device__ const unsigned long long kExp2Tab_dev[64] = {
#include "mbb_exptab.inc"
};
__device__ const unsigned long long kExp2Tab256_dev[256] = {
#include "mbb_exptab256.inc"
};
__device__ __align__(16) const unsigned long long kLogTab_dev[256] = {
#include "mbb_logtab.inc"
};
__device__ __constant__ double kLeanG_dev[5] = MBB_LEAN_G64;
__device__ __constant__ double kLeanG256_dev[4] = MBB_LEAN_G256;
#endif

inline const double* exp2_tab_host() {
  static const unsigned long long tab[64] = {
#include "mbb_exptab.inc"
  };
  return reinterpret_cast<const double*>(tab);
}
inline const double* exp2_tab256_host() {
  static const unsigned long long tab[256] = {
#include "mbb_exptab256.inc"
  };
  return reinterpret_cast<const double*>(tab);
}

// The plain (unreplicated) tables: global memory (L1-resident) on the device, a
// static array on the host.  Index stride 1 (TS = 0, or kTab256 for the 256-entry one).
MBB_HD const double* exp2_tab_default() {
#if defined(__CUDA_ARCH__)
  return reinterpret_cast<const double*>(kExp2Tab_dev);
#else
  return exp2_tab_host();
#endif
}
MBB_HD const double* exp2_tab256_default() {
#if defined(__CUDA_ARCH__)
  return reinterpret_cast<const double*>(kExp2Tab256_dev);
#else
  return exp2_tab256_host();
#endif
}
// shared-memory copies: entry i, copy c at double index i*16 + c
constexpr int kTabRepShift = 4;                          // TS of the replicated 64-entry table
constexpr int kTabRepDoubles = 64 << kTabRepShift;
constexpr int kTabRep256 = kTabRepShift | kTab256;       // TS of the replicated 256-entry table
constexpr int kTabRep256Doubles = 256 << kTabRepShift;

template <int TS>
MBB_HD double lean_g_coef(int i) {
#if defined(__CUDA_ARCH__)
  return TabCfg<TS>::bits == 6 ? kLeanG_dev[i] : kLeanG256_dev[i];
#else
  const double c64[5] = MBB_LEAN_G64;
  const double c256[4] = MBB_LEAN_G256;
  return TabCfg<TS>::bits == 6 ? c64[i] : c256[i];
#endif
}

// ---- bit-level helpers (identical results on host and device) ----
MBB_HD int hi32_of(double t) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(t);
#else
  unsigned long long u;
  memcpy(&u, &t, 8);
  return (int)(unsigned)(u >> 32);
#endif
}
MBB_HD int lo32_of(double t) {
#if defined(__CUDA_ARCH__)
  return __double2loint(t);
#else
  unsigned long long u;
  memcpy(&u, &t, 8);
  return (int)(unsigned)(u & 0xffffffffu);
#endif
}
MBB_HD double from_hilo(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  const unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo;
  double r;
  memcpy(&r, &u, 8);
  return r;
#endif
}

// min(x, ~cap) for x >= 0 (also +inf and NaN -> ~cap): one integer instruction on
// the high word; when clamped the result lies in [cap, cap + one high-word step).
template <int CAP_HI_WORD>
MBB_HD double clamp_pos(double x) {
  const int h = hi32_of(x);
  return from_hilo(h < CAP_HI_WORD ? h : CAP_HI_WORD, lo32_of(x));
}
constexpr int kHi700 = 0x4085e000;      // high word of 700.0
// high word of 64624.0 (x4 for the 256-entry flavour) ~ 700*N/ln2
template <int TS>
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int hi700c() { return TabCfg<TS>::bits == 6 ? 0x40ef8e00 : 0x410f8e00; }

// Reduced exponent: y = k + f in units of 1/64 octave, k = round(y), |f| <= 1/2.
// Valid while |y| < 2^31 (callers gate their parameters).
struct Red {
  double f;
  int k;
};

// y = a*(b_hi + b_lo)
MBB_HD Red red_prod(double a, double b_hi, double b_lo) {
  const double t = fma(a, b_hi, kMagic);
  const double kf = t - kMagic;
  Red r;
  r.f = fma(a, b_lo, fma(a, b_hi, -kf));
  r.k = lo32_of(t);
  return r;
}

// y = a*b, b a single double (node tables: the rounding of b costs |y| 2^-53 in y)
MBB_HD Red red_prod1(double a, double b) {
  const double t = fma(a, b, kMagic);
  const double kf = t - kMagic;
  Red r;
  r.f = fma(a, b, -kf);
  r.k = lo32_of(t);
  return r;
}

// plain argument in natural units
MBB_HD Red red_x(double x) { return red_prod(x, kC64Hi, kC64Lo); }

// y = -yc for an argument already in 1/64-octave units (3 instructions)
MBB_HD Red red_neg_scaled(double yc) {
  const double t = kMagic - yc;
  const double kf = t - kMagic;
  Red r;
  r.f = -(yc + kf);
  r.k = lo32_of(t);
  return r;
}

// y = (a_hi + a_lo) + b*c, all in 1/64-octave units (slow path of the optically
// thick node: the two exponents may individually overflow)
MBB_HD Red red_sum_prod(double a_hi, double a_lo, double b, double c) {
  const double y = b * c;
  const double ylo = fma(b, c, -y);
  const double s = a_hi + y;
  const double bb = s - a_hi;
  const double err = (a_hi - (s - bb)) + (y - bb);
  const double t = s + kMagic;
  const double kf = t - kMagic;
  Red r;
  r.f = (s - kf) + (err + (a_lo + ylo));
  r.k = lo32_of(t);
  return r;
}

// 2^m T_j: table lookup + scaling.  TS = log2 of the index stride (0: plain
// table, kTabRepShift: the replicated shared-memory copy, `tab` then points at
// the calling lane's copy).  CLAMP=false: one integer multiply-add (the caller
// guarantees m in [-1022, 1023]); CLAMP=true saturates m.
template <int TS, bool CLAMP>
MBB_HD double scaled_T(const double* tab, int k) {
  using C = TabCfg<TS>;
#if defined(__CUDA_ARCH__) && !defined(MBB_NO_ASM_LOOKUP)
  double tb;
  if constexpr (C::stride != 0) {
    // the replicated shared-memory copy: mask, one scaled add, LDS.64 -- written out because
    // ptxas otherwise re-derives the lane's copy offset for every lookup (shift, mask, or, LEA)
    unsigned addr;
    asm("{\n\t.reg .u32 t;\n\tand.b32 t, %1, %2;\n\tmad.lo.u32 %0, t, %3, %4;\n\t}"
        : "=r"(addr)
        : "r"(k), "n"(C::mask), "n"(8 << C::stride), "r"((unsigned)__cvta_generic_to_shared(tab)));
    asm("ld.shared.f64 %0, [%1];" : "=d"(tb) : "r"(addr));
  } else {
    tb = tab[k & C::mask];
  }
#else
  const double tb = tab[(k & C::mask) << C::stride];
#endif
  if (!CLAMP) return from_hilo(hi32_of(tb) + (int)((unsigned)k << C::hishift), lo32_of(tb));
  const int j = (k + C::half) & C::mask;
  int m = (k + C::half) >> C::bits;
  m = m > 1023 ? 1023 : m;
  m = m < -1022 ? -1022 : m;
  return from_hilo(hi32_of(tb) - 0x80000 + (j << C::hishift) + (int)((unsigned)m << 20), lo32_of(tb));
}

// p = 2^(f/N) - 1
template <int TS>
MBB_HD double lean_p(double f) {
  double g = lean_g_coef<TS>(TabCfg<TS>::deg);
#pragma unroll
  for (int i = TabCfg<TS>::deg - 1; i >= 0; --i) g = fma(g, f, lean_g_coef<TS>(i));
  return g * f;
}

template <int TS, bool CLAMP>
MBB_HD double exp_red(const Red r, const double* tab) {
  const double sT = scaled_T<TS, CLAMP>(tab, r.k);
  return fma(sT, lean_p<TS>(r.f), sT);
}

// scale * exp(y): the product scale*sT runs beside the polynomial, not after it
template <int TS, bool CLAMP>
MBB_HD double exp_red_times(const Red r, const double* tab, double scale) {
  const double sT = scaled_T<TS, CLAMP>(tab, r.k) * scale;
  return fma(sT, lean_p<TS>(r.f), sT);
}

template <int TS, bool CLAMP>
MBB_HD double expm1_red(const Red r, const double* tab) {
  const double sT = scaled_T<TS, CLAMP>(tab, r.k);
  return fma(sT, lean_p<TS>(r.f), sT - 1.0);
}

// 1 - exp(y) for the reduced exponent of y (y <= 0 in every use)
template <int TS, bool CLAMP>
MBB_HD double one_minus_exp_red(const Red r, const double* tab) {
  const double sT = scaled_T<TS, CLAMP>(tab, r.k);
  return fma(-sT, lean_p<TS>(r.f), 1.0 - sT);
}

// 1/b for normal b (no subnormal / zero handling), <= 1 ulp
MBB_HD double rcp_fast(double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  e = fma(-b, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / b;
#endif
}

// a / b with one residual correction (<= 1 ulp): per-walker quotients
MBB_HD double div_fast(double a, double b) {
  const double r = rcp_fast(b);
  const double y = a * r;
  return fma(fma(-b, y, a), r, y);
}

// 1/b in 3 FP64 instructions (node loop): seed r0 = MUFU.RCP64H(b) with
// |e| = |1 - b r0| <= 2^-19.9 (measured on B200, tools/rcp_probe.cu), then
// r0 (1 + e + e^2): relative error e^3 ~ 2^-59.7 plus two roundings.
MBB_HD double rcp_cubic(double b) {
#if defined(__CUDA_ARCH__)
  double r0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));
  const double e = fma(-b, r0, 1.0);
  return fma(fma(r0, e, r0), e, r0);
#else
  return 1.0 / b;
#endif
}

// generic saturating forms (per-walker setup code)
MBB_HD double exp_l(double x) { return exp_red<0, true>(red_x(x), exp2_tab_default()); }
MBB_HD double expm1_l(double x) { return expm1_red<0, true>(red_x(x), exp2_tab_default()); }
// saturating near 700: for optical depths t that only feed 1 - exp(-t)
MBB_HD double exp_tau(double x) { return clamp_pos<kHi700>(exp_l(x)); }

// log(x) for positive normal x (no zero / subnormal / inf / NaN handling: fast_setup gates its
// parameters first).  x = 2^e m, m in [1, 2), j = top 7 mantissa bits:
//     log x = e ln2 - log c_j + log1p(r),   r = m c_j - 1 (one FMA, |r| <= 2^-8),
// log1p by its series through r^6 (truncation r^7/7 <= 2e-18).  11 FP64 instructions, one
// 16-byte table load (2 KB table, L1-resident), ~8 integer ones -- libdevice log: 29 FP64 + ~50
// others on sm_100a.  Absolute error <= 2.2e-16 max(1, |log x|) (1 ulp) against mpmath
// (tests/test_device_logic_cpu.py).
constexpr double kLn2Hi = 0.69314718055994529;
constexpr double kLn2Lo = 2.3190468138462996e-17;
MBB_HD double log_l(double x) {
#if defined(__CUDA_ARCH__)
  const double2* lt = reinterpret_cast<const double2*>(kLogTab_dev);
#else
  static const unsigned long long lt_bits[256] = {
#include "mbb_logtab.inc"
  };
  struct D2 { double x, y; };
  const D2* lt = reinterpret_cast<const D2*>(lt_bits);
#endif
  const int hi = hi32_of(x);
  const int j = (hi >> 13) & 127;
  const double ed = (double)((hi >> 20) - 1023);
  const double m = from_hilo((hi & 0x000fffff) | 0x3ff00000, lo32_of(x));
#if defined(__CUDA_ARCH__)
  const double2 cl = __ldg(lt + j);
#else
  const D2 cl = lt[j];
#endif
  const double r = fma(m, cl.x, -1.0);
  double p = fma(r, -1.0 / 6.0, 0.2);
  p = fma(r, p, -0.25);
  p = fma(r, p, 1.0 / 3.0);
  p = fma(r, p, -0.5);
  const double s = fma(ed, kLn2Hi, cl.y);
  const double t = fma(ed, kLn2Lo, r);
  return s + fma(r * r, p, t);
}

// The lean exp family extracts the binary exponent from the low 32 bits of
// a*b + 1.5*2^52: valid for |x| < 2^31 ln2/64 ~ 2.3e7.  fast_setup gates the
// parameters (T >= 1e-3 K, beta, alpha <= 300) so every argument formed from
// them stays far inside that range.
constexpr double kFastMaxIndex = 300.0;
constexpr double kFastMinT = 1e-3;
// |exponent| (natural units) below which no CLAMP is needed
constexpr double kSafeExp = 690.0;

}  // namespace mbb
