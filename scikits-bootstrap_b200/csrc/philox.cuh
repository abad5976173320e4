// Philox4x32 counter-based generator (Salmon, Moraes, Dror, Shaw; SC'11; 7 rounds, see below) and
// the counter layout every kernel shares.  Host+device so the CPU test-suite can
// exercise the very same code (b200boot_host_* probes).
//
// Replaces numpy's PCG64 stream behind rng.integers at
// /root/reference/src/scikits/bootstrap/bootstrap.py:677-680: the index of
// (resample r, cell t, position p) is a pure function of (seed, r, t, p), so any
// shard of the resample range reproduces the same values on any GPU count.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

namespace b200boot {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

// counter word 3 carries a purpose tag so the streams never overlap
constexpr uint32_t kPurposeIndex = 0u;     // index words
constexpr uint32_t kPurposeRedraw = 1u;    // Lemire rejection redraws
constexpr uint32_t kPurposeBinomial = 2u;  // tile-occupancy uniforms
constexpr uint32_t kPurposeBlock = 3u;     // moving-block start draws
constexpr uint32_t kPurposeShuffle = 4u;   // subsample sort keys

struct Words4 {
  uint32_t w[4];
};

// Rounds used by every stream of the library: Philox4x32-7, the smallest round count the
// SC'11 paper reports as Crush-resistant (it passes SmallCrush, Crush and BigCrush; -10 is
// Random123's default and adds a safety margin).  Three rounds fewer are worth 20 % of the
// whole resample kernel on B200 (profiles/r02_k1_variants.md), the
// multiplies of the rounds being its busiest pipe.  -DB2_PHILOX_ROUNDS=10 builds the library
// with the margin; b200boot_philox_rounds() reports what a build uses.
#ifndef B2_PHILOX_ROUNDS
#define B2_PHILOX_ROUNDS 7
#endif
constexpr int kPhiloxRounds = B2_PHILOX_ROUNDS;

template <int ROUNDS>
B2_HD Words4 philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                          uint32_t k1) {
#pragma unroll
  for (int round = 0; round < ROUNDS; ++round) {
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  Words4 r;
  r.w[0] = c0;
  r.w[1] = c1;
  r.w[2] = c2;
  r.w[3] = c3;
  return r;
}
B2_HD Words4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                           uint32_t k1) {
  return philox4x32_r<10>(c0, c1, c2, c3, k0, k1);
}

// Round keys hoisted out of the rounds: they depend on the seed only, so a kernel
// computes them once and every xor takes them as warp-uniform operands.
struct RoundKeys {
  uint32_t k0[10], k1[10];
};

B2_HD RoundKeys make_round_keys(uint64_t seed) {
  RoundKeys rk;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    rk.k0[i] = a;
    rk.k1[i] = b;
    a += kPhiloxW0;
    b += kPhiloxW1;
  }
  return rk;
}

B2_HD Words4 philox4x32_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const RoundKeys& rk) {
#pragma unroll
  for (int round = 0; round < kPhiloxRounds; ++round) {
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0;
    const uint64_t p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ rk.k0[round];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ rk.k1[round];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  Words4 r;
  r.w[0] = c0;
  r.w[1] = c1;
  r.w[2] = c2;
  r.w[3] = c3;
  return r;
}

// Per-(resample, cell) stream descriptor.
//   key = (seed lo, seed hi)
//   c0  = block index inside the stream        c1 = cell id
//   c2  = resample id (low 32)                 c3 = resample id bits 32..47 | purpose << 16
struct Stream {
  uint32_t k0, k1, c1, c2, c3;
};

B2_HD Stream make_stream(uint64_t seed, int64_t resample, uint32_t cell, uint32_t purpose) {
  Stream s;
  s.k0 = (uint32_t)seed;
  s.k1 = (uint32_t)(seed >> 32);
  s.c1 = cell;
  s.c2 = (uint32_t)resample;
  s.c3 = ((uint32_t)((uint64_t)resample >> 32) & 0xFFFFu) | (purpose << 16);
  return s;
}

B2_HD Words4 stream_block(const Stream& s, uint32_t block) {
  return philox4x32_r<kPhiloxRounds>(block, s.c1, s.c2, s.c3, s.k0, s.k1);
}
// same stream with pre-expanded round keys (rk must come from the stream's seed)
B2_HD Words4 stream_block(const Stream& s, uint32_t block, const RoundKeys& rk) {
  return philox4x32_keys(block, s.c1, s.c2, s.c3, rk);
}

// 53-bit uniform in the open interval (0,1) from two words.
B2_HD double uniform53(uint32_t hi, uint32_t lo) {
  const uint64_t m = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
  return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

}  // namespace b200boot
