// zg_philox.cuh -- Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11), host and device.
// GPT.sample (main.zig:198-207) draws from std.rand.DefaultPrng re-seeded with the wall clock on EVERY call, which
// nothing can reproduce.  Here the uniform draw of sampling step `step` of sequence `sequence` is a pure function of
// (seed, step, sequence): counter = (step lo, step hi, sequence lo, sequence hi), key = (seed lo, seed hi).  No state, so
// generate() can sample on the device without a host round trip, the result does not depend on how sequences are
// sharded over GPUs, and `--seed N` reproduces a run.
#pragma once
#include <stdint.h>

namespace zg {

#ifdef __CUDACC__
#define ZG_HD __host__ __device__ __forceinline__
#else
#define ZG_HD inline
#endif

struct Philox4 { uint32_t v[4]; };

ZG_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// uniform in [0, 1) with 24 random bits: exactly representable in fp32, the type weightedIndex scans with
ZG_HD float philox_uniform(uint64_t seed, uint64_t step, uint64_t sequence) {
  const Philox4 r = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)sequence, (uint32_t)(sequence >> 32),
                                  (uint32_t)seed, (uint32_t)(seed >> 32));
  return (float)(r.v[0] >> 8) * (1.0f / 16777216.0f);
}

}  // namespace zg
