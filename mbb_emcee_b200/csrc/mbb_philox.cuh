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

// 53-bit uniform in (0, 1) from two 32-bit words
MBB_PHILOX_HD double u53(unsigned hi, unsigned lo) {
  const unsigned long long m = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
  return ((double)m + 0.5) * 1.1102230246251565e-16;   // 2^-53
}

// the three fields the sampler cuts out of ONE Philox block (mbb_ensemble.cuh): 52 bits from
// word 0 and the top 20 bits of word 1 (m + 0.5 is exact below 2^52: strictly inside (0, 1)) ...
MBB_PHILOX_HD double u52w(unsigned w0, unsigned w1_top20) {
  const unsigned long long m = ((unsigned long long)w0 << 20) | (unsigned long long)w1_top20;
  return ((double)m + 0.5) * 2.2204460492503131e-16;   // 2^-52
}
// ... and 43 bits from word 3 and the low 11 bits of word 1
MBB_PHILOX_HD double u43(unsigned w3, unsigned w1_low11) {
  const unsigned long long m = ((unsigned long long)w3 << 11) | (unsigned long long)w1_low11;
  return ((double)m + 0.5) * 1.1368683772161603e-13;   // 2^-43
}

}  // namespace mbb
