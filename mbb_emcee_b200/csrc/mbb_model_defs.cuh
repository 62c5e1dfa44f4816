// Shared definitions of the device/host model code: qualifiers, the physical
// constants the reference hardwires, and the per-evaluation status codes.
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define MBB_HD __host__ __device__ __forceinline__
#define MBB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define MBB_HD inline
#define MBB_HD_NOINLINE inline
#endif

namespace mbb {

// Physical constants exactly as the reference hardwires them
// (modified_blackbody.py:15-18, fnu.pyx:14-15).
constexpr double kH = 6.6260693e-34;      // J s
constexpr double kK = 1.3806505e-23;      // J / K
constexpr double kCum = 299792458e6;      // um / s
constexpr double kUmToGHz = 299792458e-3; // um * GHz
constexpr double kInf = __builtin_huge_val();

// Per-evaluation status.  0 and 1 are normal outcomes; the rest map to the
// exceptions the reference raises (SURVEY.md 8b "Errors").
enum Status : int {
  ST_OK = 0,
  ST_BELOW_LOWLIM = 1,   // likelihood.py:806-807 -> -inf
  ST_BAD_ALPHA = 2,      // modified_blackbody.py:219-221 ValueError
  ST_BAD_BETA = 3,       // :222-224 ValueError
  ST_BRACKET_LOW = 4,    // :294-300 ValueError
  ST_BRACKET_HIGH = 5,   // :310-316 ValueError
  ST_NO_CONVERGE = 6,    // brentq RuntimeError
  ST_OVERFLOW = 7,       // :326-328 OverflowError
  ST_PEAK_BRACKET = 8,   // :612-630 Exception
  ST_NONFINITE = 9       // NaN/inf parameters or result
};


MBB_HD bool finite_d(double x) { return x - x == 0.0; }

}  // namespace mbb
