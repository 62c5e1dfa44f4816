// Philox4x32-10 counter-based generator (Salmon, Moraes, Dror, Shaw 2011) and
// the uniform conversion used by the device ensemble sampler.  Host+device so
// the CPU suite can check it against the published known-answer vectors.
#pragma once

#if defined(__CUDACC__)
#define MBB_PHILOX_HD __host__ __device__ inline
#else
#define MBB_PHILOX_HD inline
#endif

namespace mbb {

struct Philox {
  unsigned c[4];
};

MBB_PHILOX_HD void philox_round(unsigned c[4], unsigned k0, unsigned k1) {
  const unsigned long long p0 = 0xD2511F53ull * c[0];
  const unsigned long long p1 = 0xCD9E8D57ull * c[2];
  const unsigned n0 = (unsigned)(p1 >> 32) ^ c[1] ^ k0;
  const unsigned n1 = (unsigned)p1;
  const unsigned n2 = (unsigned)(p0 >> 32) ^ c[3] ^ k1;
  const unsigned n3 = (unsigned)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

MBB_PHILOX_HD Philox philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                                unsigned k0, unsigned k1) {
  Philox r;
  r.c[0] = c0; r.c[1] = c1; r.c[2] = c2; r.c[3] = c3;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(r.c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return r;
}

// The same generator with the ten round keys (k0 + i W0, k1 + i W1) formed beforehand: the
// sampler kernels carry them in their parameter block (constant bank), where the key
// additions of every call would otherwise take two uniform-datapath issue slots per round.
struct PhiloxKeys {
  unsigned k[20];
};
MBB_PHILOX_HD PhiloxKeys philox_keys(unsigned long long seed) {
  PhiloxKeys K;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    K.k[2 * i] = k0;
    K.k[2 * i + 1] = k1;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return K;
}
MBB_PHILOX_HD void philox_mulhilo(unsigned m, unsigned c, unsigned& hi, unsigned& lo) {
#if defined(__CUDA_ARCH__)
  hi = __umulhi(m, c);        // ptxas pairs the two into one IMAD.WIDE.U32
  lo = m * c;
#else
  const unsigned long long p = (unsigned long long)m * c;
  hi = (unsigned)(p >> 32);
  lo = (unsigned)p;
#endif
}
MBB_PHILOX_HD Philox philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, const PhiloxKeys& K) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    unsigned hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
    philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    c0 = hi1 ^ c1 ^ K.k[2 * i];
    c2 = hi0 ^ c3 ^ K.k[2 * i + 1];
    c1 = lo1;
    c3 = lo0;
  }
  Philox r;
  r.c[0] = c0; r.c[1] = c1; r.c[2] = c2; r.c[3] = c3;
  return r;
}

// 53-bit uniform in (0, 1) from two 32-bit words
MBB_PHILOX_HD double u53(unsigned hi, unsigned lo) {
  const unsigned long long m = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
  return ((double)m + 0.5) * 1.1102230246251565e-16;   // 2^-53
}

// the three fields the sampler cuts out of ONE Philox block (mbb_ensemble.cuh): 52 bits from
// word 0 and the top 20 bits of word 1 (m + 0.5 is exact below 2^52: strictly inside (0, 1)) ...
// (m + 0.5) 2^-52 without an integer conversion: m goes into the mantissa of d = 1 + m 2^-52 and
// d - (1 - 2^-53) = (2m + 1) 2^-53 is exact (an odd integer below 2^53 times a power of two).
MBB_PHILOX_HD double philox_bits_to_double(unsigned hi, unsigned lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)hi, (int)lo);
#else
  const unsigned long long b = ((unsigned long long)hi << 32) | lo;
  double d;
  __builtin_memcpy(&d, &b, 8);
  return d;
#endif
}
MBB_PHILOX_HD double u52w(unsigned w0, unsigned w1_top20) {
  const double d = philox_bits_to_double(0x3ff00000u | (w0 >> 12), (w0 << 20) | w1_top20);
  return d - (1.0 - 1.1102230246251565e-16);       // 1 - 2^-53
}
// ... and 43 bits from word 3 and the low 11 bits of word 1
// (m + 0.5) 2^-43, m < 2^43: d = 1 + m 2^-43 (m shifted up by 9 mantissa bits), minus 1 - 2^-44
MBB_PHILOX_HD double u43(unsigned w3, unsigned w1_low11) {
  const double d = philox_bits_to_double(0x3ff00000u | (w3 >> 12), (w3 << 20) | (w1_low11 << 9));
  return d - (1.0 - 5.6843418860808015e-14);       // 1 - 2^-44
}

}  // namespace mbb
