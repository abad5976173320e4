// Tile plan and the per-cell index streams shared by K1 (fused gather+reduce),
// K2 (index materialiser) and the host probes.  One definition, so what
// `bootstrap_indices` yields is by construction what the fused kernel draws.
#pragma once
#include <stdint.h>

#include "binomial.cuh"
#include "philox.cuh"

namespace b200boot {

// bytes of shared memory a CTA may spend on resident rows
constexpr int64_t kSmemDataBytes = 224 * 1024;

struct Plan {
  int32_t log2_tile;
  int32_t n_arrays;
  int64_t n_full;   // full power-of-two tiles
  int64_t rem;      // rows in the trailing cell (== n_rows when n_full == 0)
  int64_t n_tiles;  // n_full + (rem > 0)
  int64_t n_rows;
  B2_HD int64_t tile() const { return (int64_t)1 << log2_tile; }
  B2_HD int64_t cell_size(int64_t t) const { return t < n_full ? tile() : rem; }
};

inline Plan make_plan(int64_t n_rows, int n_arrays) {
  Plan p;
  p.n_arrays = n_arrays;
  p.n_rows = n_rows;
  // largest power of two with n_arrays * 8 B * 2^k <= 200 KB
  p.log2_tile = n_arrays <= 1 ? 14 : (n_arrays <= 3 ? 13 : 12);
  const int64_t single_cap = kSmemDataBytes / (8 * (int64_t)n_arrays);
  if (n_rows <= single_cap) {
    p.n_full = 0;
    p.rem = n_rows;
    p.n_tiles = 1;
  } else {
    p.n_full = n_rows >> p.log2_tile;
    p.rem = n_rows - (p.n_full << p.log2_tile);
    p.n_tiles = p.n_full + (p.rem > 0 ? 1 : 0);
  }
  return p;
}

// ---- draws inside one cell --------------------------------------------------
// Power-of-two cell: a Philox block yields 8 indices, word w -> (w & mask) and
// ((w >> 16) & mask).  Exactly uniform, no rejection.  log2 size <= 16.
//
// General cell of `size` rows: one word per index, Lemire multiply-shift
// idx = (w * size) >> 32; words whose low product is < 2^32 mod size are
// rejected and redrawn from the kPurposeRedraw stream keyed by the position, so
// the result is exactly uniform and still a pure function of the counters.

#if defined(__CUDACC__)
#define B2_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define B2_HD_NOINLINE static
#endif

// Rare path (probability size / 2^32 per draw): kept out of line.
B2_HD_NOINLINE uint32_t lemire_redraw(Stream rs, uint32_t position, uint32_t size,
                                      uint32_t thresh) {
  const uint32_t base_c3 = (rs.c3 & 0xFFFFu) | (kPurposeRedraw << 16);
  for (uint32_t attempt = 0;; ++attempt) {
    rs.c3 = base_c3 | ((attempt >> 2) << 20);
    const Words4 w = stream_block(rs, position);
    const uint64_t prod = (uint64_t)w.w[attempt & 3u] * size;
    if ((uint32_t)prod >= thresh) return (uint32_t)(prod >> 32);
  }
}

// Calls f(j, position, index) for draw j = 0..7 of block `q` of a pow2 cell.
// CHECK = false promises that the whole block lies below n_draws.
template <bool CHECK, typename F>
B2_HD void pow2_block(const Stream& s, uint32_t q, uint32_t mask, int64_t n_draws, F&& f) {
  const Words4 w = stream_block(s, q);
  const int64_t base = (int64_t)q * 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!CHECK || base + 2 * j < n_draws) f(2 * j, base + 2 * j, w.w[j] & mask);
    if (!CHECK || base + 2 * j + 1 < n_draws) f(2 * j + 1, base + 2 * j + 1, (w.w[j] >> 16) & mask);
  }
}

// Same for a general cell: draws j = 0..3 of block `q`.
template <bool CHECK, typename F>
B2_HD void lemire_block(const Stream& s, uint32_t q, uint32_t size, uint32_t thresh,
                        int64_t n_draws, F&& f) {
  const Words4 w = stream_block(s, q);
  const int64_t base = (int64_t)q * 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!CHECK || base + j < n_draws) {
      const uint64_t prod = (uint64_t)w.w[j] * size;
      uint32_t idx = (uint32_t)(prod >> 32);
      if ((uint32_t)prod < thresh) idx = lemire_redraw(s, (uint32_t)(base + j), size, thresh);
      f(j, base + j, idx);
    }
  }
}

B2_HD uint32_t lemire_threshold(uint32_t size) { return (uint32_t)(0u - size) % size; }

// ---- tile occupancies -------------------------------------------------------
// counts[t * stride] for t in [0, n_tiles): conditional-binomial chain over the
// cells in order; the last cell takes what is left.
template <typename I>
B2_HD void tile_count_chain(uint64_t seed, int64_t resample, const Plan& plan, I* counts,
                            int64_t stride) {
  int64_t n_left = plan.n_rows;   // draws still to place
  int64_t rows_left = plan.n_rows;
  for (int64_t t = 0; t + 1 < plan.n_tiles; ++t) {
    const int64_t sz = plan.cell_size(t);
    const Stream s = make_stream(seed, resample, (uint32_t)t, kPurposeBinomial);
    const int64_t k = binomial_ratio(n_left, sz, rows_left, s);
    counts[t * stride] = (I)k;
    n_left -= k;
    rows_left -= sz;
  }
  counts[(plan.n_tiles - 1) * stride] = (I)n_left;
}

}  // namespace b200boot
