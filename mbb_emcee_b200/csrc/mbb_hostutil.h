// Host-side helpers of the launcher that the CPU test-suite checks directly.
#pragma once

namespace mbb {

// floor(g / d) for every 32-bit g without a division (Granlund & Montgomery 1994, "Division by
// invariant integers using multiplication"):  t = umulhi(mul, g);  q = (t + ((g - t) >> sh1)) >> sh2
// with l = ceil(log2 d), mul = floor(2^32 (2^l - d) / d) + 1, sh1 = min(l, 1), sh2 = max(l - 1, 0).
// mul == 0 marks "d does not fit 32 bits": the device falls back to 64-bit division.
struct WpsDivision {
  unsigned mul;
  int sh1, sh2;
};

inline WpsDivision wps_division(unsigned long long d) {
  WpsDivision r = {0u, 0, 0};
  if (d == 0) d = 1;
  if (d >> 32) return r;
  int l = 0;
  while ((1ull << l) < d) ++l;
  r.mul = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  r.sh1 = l < 1 ? l : 1;
  r.sh2 = l > 1 ? l - 1 : 0;
  return r;
}

inline unsigned wps_divide(const WpsDivision& d, unsigned g) {
  const unsigned t = (unsigned)(((unsigned long long)d.mul * g) >> 32);
  return (t + ((g - t) >> d.sh1)) >> d.sh2;
}

}  // namespace mbb
