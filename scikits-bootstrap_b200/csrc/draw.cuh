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
// Full power-of-two tiles are drawn per bank class (see class_block below).
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

// Calls f(j, position, index) for draw j = 0..3 of block `q` of a general cell.
// CHECK = false promises that the whole block lies below n_draws.  `rk`, when
// given, must hold the round keys of the stream's seed (hoisted by the caller).
template <bool CHECK, typename F>
B2_HD void lemire_block(const Stream& s, uint32_t q, uint32_t size, uint32_t thresh,
                        int64_t n_draws, F&& f, const RoundKeys* rk = nullptr) {
  const Words4 w = rk ? stream_block(s, q, *rk) : stream_block(s, q);
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

// ---- bank classes of a full tile -------------------------------------------------
// A full tile of 2^L rows is cut into 16 classes by (row mod 16): class c owns the
// rows whose float64 sits in shared-memory bank pair (2c, 2c+1).  A lane that only
// ever reads its own class never conflicts with the other 15 lanes of its half-warp.
// The draws of a tile are therefore split over the classes (multinomial, 16 equal
// cells; class_count_tree) and each class has its own stream, cell id
// 0x80000000 | t << 4 | c.  A Philox word yields three (L-4)-bit row numbers j:
//   j_A = (w >> 7) & m,  j_B = (w >> 17) & m,  j_C = rotl(w, 5) & m      (m = 2^(L-4) - 1)
// i.e. 12 draws per block; the row inside the tile is j * 16 + c.
#ifndef B2_K1_SHF
#define B2_K1_SHF 1     // field shifts on the ALU pipe, not as multiply-high (measured: -0.8 %)
#endif
constexpr int kClasses = 16;
constexpr int kClassDrawsPerBlock = 12;

B2_HD uint32_t class_cell_id(int64_t t, int c) { return 0x80000000u | ((uint32_t)t << 4) | (uint32_t)c; }

B2_HD uint32_t rotl32(uint32_t x, int r) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(x, x, r);
#else
  return (x << r) | (x >> (32 - r));
#endif
}

// the three row numbers of word w, pre-shifted by 7 (j << 7 = byte offset of row j of
// class 0 for 8-byte elements); m7 = m << 7
B2_HD void class_fields_shifted(uint32_t w, uint32_t m7, uint32_t* f) {
  f[0] = w & m7;
#if defined(__CUDA_ARCH__) && !B2_K1_SHF
  // w >> 10 as a multiply-high: off the ALU pipe, onto the one the Philox multiplies use
  uint32_t hi;
  asm("mul.hi.u32 %0, %1, 4194304;" : "=r"(hi) : "r"(w));
  f[1] = hi & m7;
#else
  f[1] = (w >> 10) & m7;
#endif
  f[2] = rotl32(w, 12) & m7;
}

// Calls f(j, position, row_in_tile) for draw j = 0..11 of block q of class c.
template <bool CHECK, typename F>
B2_HD void class_block(const Stream& s, uint32_t q, uint32_t m7, int c, int64_t n_draws, F&& f) {
  const Words4 w = stream_block(s, q);
  const int64_t base = (int64_t)q * kClassDrawsPerBlock;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t fld[3];
    class_fields_shifted(w.w[i], m7, fld);
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (!CHECK || base + 3 * i + k < n_draws)
        f(3 * i + k, base + 3 * i + k, (fld[k] >> 7) * kClasses + (uint32_t)c);
  }
}

// Splits the n_t draws of full tile t over its 16 equal classes by a binary tree of even
// splits: node (level l, index j) holds the draws of the classes [j * 16 >> l, (j+1) * 16 >> l)
// and hands Bin(n, 1/2) of them to its lower half.  15 nodes, heap numbering
// node = 2^l + j, each with its own stream (cell id 0x80000000 | t << 4 | node, purpose
// binomial).  All nodes of a level are independent, so a kernel can run a level densely
// (k0_class_tree_kernel); this sequential form is what K2, the host probes and single threads
// use, and gives the same values.  out[c * stride].
B2_HD uint32_t class_node_cell_id(int64_t t, int node) { return 0x80000000u | ((uint32_t)t << 4) | (uint32_t)node; }

template <typename I>
B2_HD void class_count_tree(uint64_t seed, int64_t resample, int64_t t, int64_t n_t, I* out,
                            int64_t stride) {
  int64_t cnt[kClasses];
  cnt[0] = n_t;
  for (int node = 1; node < kClasses; ++node) {          // heap order: parents before children
    int level = 0;
    while ((2 << level) <= node) ++level;
    const int span = kClasses >> level;                  // classes under a node of this level
    const int first = (node - (1 << level)) * span;
    const int64_t n = cnt[first];
    const Stream s = make_stream(seed, resample, class_node_cell_id(t, node), kPurposeBinomial);
    const int64_t k = binomial_half(n, s);
    cnt[first] = k;
    cnt[first + span / 2] = n - k;
  }
  for (int c = 0; c < kClasses; ++c) out[c * stride] = (I)cnt[c];
}

// ---- tile occupancies -------------------------------------------------------
// counts[t * stride] for t in [0, n_tiles): the multinomial over the cells is sampled
// in two levels so that no chain is longer than ~2 x 32 steps: first the occupancies
// of groups of kTileGroup consecutive cells (conditional binomials over the groups,
// cell id 0x40000000 | g), then inside each group a chain over its cells (cell id t);
// the last group / cell takes what is left.  Plans with at most kTileGroup cells have
// a single group and only the inner chain.
constexpr int kTileGroup = 32;

B2_HD uint32_t group_cell_id(int64_t g) { return 0x40000000u | (uint32_t)g; }
B2_HD int64_t n_tile_groups(const Plan& plan) { return (plan.n_tiles + kTileGroup - 1) / kTileGroup; }

// rows covered by cells [t0, t1)
B2_HD int64_t rows_of_cells(const Plan& plan, int64_t t0, int64_t t1) {
  const int64_t full = (t1 < plan.n_full ? t1 : plan.n_full) - (t0 < plan.n_full ? t0 : plan.n_full);
  return full * plan.tile() + ((t1 > plan.n_full && t0 <= plan.n_full) ? plan.rem : 0);
}

// level 1: draws per group; gcounts[g * stride]
template <typename I>
B2_HD void group_count_chain(uint64_t seed, int64_t resample, const Plan& plan, I* gcounts, int64_t stride) {
  const int64_t ng = n_tile_groups(plan);
  int64_t n_left = plan.n_rows, rows_left = plan.n_rows;
  for (int64_t g = 0; g + 1 < ng; ++g) {
    const int64_t rows = rows_of_cells(plan, g * kTileGroup, (g + 1) * kTileGroup);
    const Stream s = make_stream(seed, resample, group_cell_id(g), kPurposeBinomial);
    const int64_t k = binomial_ratio(n_left, rows, rows_left, s);
    gcounts[g * stride] = (I)k;
    n_left -= k;
    rows_left -= rows;
  }
  gcounts[(ng - 1) * stride] = (I)n_left;
}

// level 2: the n_group draws of group g over its cells; counts[t * stride]
template <typename I>
B2_HD void tile_chain_in_group(uint64_t seed, int64_t resample, const Plan& plan, int64_t g,
                               int64_t n_group, I* counts, int64_t stride) {
  const int64_t t0 = g * kTileGroup;
  const int64_t t1 = (t0 + kTileGroup < plan.n_tiles) ? t0 + kTileGroup : plan.n_tiles;
  int64_t n_left = n_group, rows_left = rows_of_cells(plan, t0, t1);
  for (int64_t t = t0; t + 1 < t1; ++t) {
    const int64_t sz = plan.cell_size(t);
    const Stream s = make_stream(seed, resample, (uint32_t)t, kPurposeBinomial);
    const int64_t k = binomial_ratio(n_left, sz, rows_left, s);
    counts[t * stride] = (I)k;
    n_left -= k;
    rows_left -= sz;
  }
  counts[(t1 - 1) * stride] = (I)n_left;
}

// both levels, sequentially (host probes, single-thread uses)
template <typename I>
B2_HD void tile_count_chain(uint64_t seed, int64_t resample, const Plan& plan, I* counts,
                            int64_t stride) {
  const int64_t ng = n_tile_groups(plan);
  int64_t n_left = plan.n_rows, rows_left = plan.n_rows;
  for (int64_t g = 0; g < ng; ++g) {
    int64_t n_g = n_left;
    if (g + 1 < ng) {
      const int64_t rows = rows_of_cells(plan, g * kTileGroup, (g + 1) * kTileGroup);
      const Stream s = make_stream(seed, resample, group_cell_id(g), kPurposeBinomial);
      n_g = binomial_ratio(n_left, rows, rows_left, s);
      n_left -= n_g;
      rows_left -= rows;
    }
    tile_chain_in_group<I>(seed, resample, plan, g, n_g, counts, stride);
  }
}

}  // namespace b200boot
