// K1mr — rank statistics (median, quantiles) of LONG (N, D) data in ROW space.
//
// The reference gathers whole rows, x[indices] (bootstrap.py:431), so every column of a
// resample sees the same rows.  For columns longer than K1m's resident limit (28 672 rows) the
// 1-D path K1mt works in the rank space of its single column; with D > 1 columns that would
// give every column its own rows.  Here the resample stays in row space:
//   * once per job, every column is sorted as (key, row) pairs; rank[i][d] = position of row i
//     in column d's ascending order (ties in any fixed order: equal values give equal results);
//   * a resample's k-th order statistic of column d is sorted[d][p_k], p_k = the k-th smallest
//     of the drawn rows' ranks rank[idx_j][d], j = 0..N-1 — found exactly by a two-level
//     counting select over the rank bits (shared-memory histograms of the high, then the low
//     half of the bits), without sorting the resample;
//   * the index sets are the materialised draws of the tiled plan (K2: exactly what
//     `bootstrap_indices` yields and what K1 consumes) or the caller's own (parity mode).
// One CTA per (resample, group of four columns): the index row is read once per group, the
// rank table (int32 [N][D], L2-resident) is gathered per drawn row.
#pragma once
#include "common.cuh"
#include "k1m_median.cuh"
#include "k45_interval.cuh"

namespace b200boot {

constexpr int kRowsThreads = 1024;
constexpr int kRowsGroup = 4;                   // columns per CTA
constexpr int kRowsMaxBins = 4096;              // per level: N <= 2^24 rows
constexpr int64_t kRowsMaxRows = (int64_t)1 << 24;

// rank[idx[d][p] * n_cols + d] = p and sorted[d][p] = value of the key, p < n_rows
__global__ void __launch_bounds__(256) rows_rank_kernel(const uint64_t* __restrict__ keys,
                                                        const uint32_t* __restrict__ idx, int64_t n_rows,
                                                        int64_t n_pad, int64_t n_cols, int64_t d0, int64_t nd,
                                                        int32_t* __restrict__ rank, double* __restrict__ sorted) {
  const int64_t total = nd * n_rows;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t dl = e / n_rows, p = e - dl * n_rows;
    const int64_t d = d0 + dl;
    rank[(int64_t)idx[dl * n_pad + p] * n_cols + d] = (int32_t)p;
    sorted[d * n_rows + p] = value_of(keys[dl * n_pad + p]);
  }
}

struct RowsParams {
  const int32_t* rank;       // [N][D]
  const double* sorted;      // [D][N]
  const int64_t* indices;    // [n_sets][N]
  int64_t n_rows, n_cols, n_sets;   // n_cols: row stride of the rank table
  int32_t shift;             // low bits per rank: bins of level 2 = 1 << shift
  int32_t hi_bins;           // bins of level 1 = ceil(N / 2^shift)
  int32_t bins;              // histogram stride in shared memory: max(hi_bins, 2^shift) rounded up to 32
  RankRule rule;
  double* out;               // [D][out_ld]
  int64_t out_ld;
};

// first bin whose cumulative count exceeds k (k below the total): warp-cooperative over n_bins
// (a multiple of 32 is not required); returns the bin and writes the count below it
__device__ __forceinline__ int rows_find_bin(const unsigned int* hist, int n_bins, long long k, long long* below_out) {
  const int lane = threadIdx.x & 31;
  const int per = (n_bins + 31) / 32;
  const int b0 = lane * per, b1 = min(n_bins, b0 + per);
  unsigned int mine = 0;
  for (int b = b0; b < b1; ++b) mine += hist[b];
  unsigned int inc = mine;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned int o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  const unsigned int hit = __ballot_sync(0xffffffffu, (long long)inc > k);
  const int owner = hit ? __ffs(hit) - 1 : 31;
  int bin = n_bins - 1;
  long long below = 0;
  if (lane == owner) {
    below = (long long)(inc - mine);
    for (int b = b0; b < b1; ++b) {
      const long long c = hist[b];
      if (k < below + c) {
        bin = b;
        break;
      }
      below += c;
    }
  }
  bin = __shfl_sync(0xffffffffu, bin, owner);
  below = __shfl_sync(0xffffffffu, below, owner);
  *below_out = below;
  return bin;
}

// the group's ranks of one drawn row.  The rows are random: every lane of a load touches its own
// sector, so a load costs the L1 32 tag cycles whatever its width — four 32-bit loads per row made
// the kernel L1-tag-bound (6e10 rows/s at N = 1e6: 4 cycles per row and SM).  When the row stride
// is a multiple of four columns the group's 16 bytes come with one 128-bit load.
__device__ __forceinline__ void rows_load_ranks(const int32_t* r, bool ok, int nc, bool vec, unsigned int* p) {
  static_assert(kRowsGroup == 4, "one uint4 per row group");
  if (vec) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(r));
    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
#pragma unroll
    for (int c = 0; c < kRowsGroup; ++c)
      if (!ok || c >= nc) p[c] = 0xFFFFFFFFu;
  } else {
#pragma unroll
    for (int c = 0; c < kRowsGroup; ++c) p[c] = (ok && c < nc) ? (unsigned int)__ldg(r + c) : 0xFFFFFFFFu;
  }
}

__global__ void __launch_bounds__(kRowsThreads) k1mr_select_kernel(const RowsParams P, const int64_t cols_left) {
  // level 1: one histogram per column; level 2: one per (column, target rank)
  extern __shared__ unsigned int rows_hist[];            // [kRowsGroup][2][P.bins]
  __shared__ int s_bin[kRowsGroup][2];
  __shared__ long long s_krem[kRowsGroup][2];
  __shared__ int s_pos[kRowsGroup][2];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t set = blockIdx.x;
  const int64_t d0 = (int64_t)blockIdx.y * kRowsGroup;
  const int nc = (int)min((int64_t)kRowsGroup, cols_left - d0);
  const int64_t* idx = P.indices + set * P.n_rows;
  const int lo_bins = 1 << P.shift;
  const unsigned int lo_mask = (unsigned int)lo_bins - 1u;
  const int two = P.rule.k_hi != P.rule.k_lo ? 1 : 0;
  const bool vec = (P.n_cols & 3) == 0 && (reinterpret_cast<uintptr_t>(P.rank + d0) & 15) == 0;
  const int bins = P.bins;
  auto h1 = [&](int c) { return rows_hist + (size_t)c * 2 * bins; };
  auto h2 = [&](int c, int t) { return rows_hist + ((size_t)c * 2 + t) * bins; };

  for (int i = tid; i < nc * 2 * bins; i += kRowsThreads) rows_hist[i] = 0u;
  __syncthreads();
  // four index entries and their rank rows in flight per thread: the rank gathers are random L2
  // reads and their latency is hidden by loads in flight only (N = 1e6: 153 -> see DESIGN 5)
  constexpr int kRowsBatch = 4;
  for (int64_t j0 = tid; j0 < P.n_rows; j0 += (int64_t)kRowsThreads * kRowsBatch) {
    unsigned int p[kRowsBatch][kRowsGroup];
#pragma unroll
    for (int b = 0; b < kRowsBatch; ++b) {
      const int64_t j = j0 + (int64_t)b * kRowsThreads;
      const int64_t i = j < P.n_rows ? idx[j] : -1;
      const bool ok = (uint64_t)i < (uint64_t)P.n_rows;          // out-of-range entries are ignored
      const int32_t* r = P.rank + (ok ? i : 0) * P.n_cols + d0;
      rows_load_ranks(r, ok, nc, vec, p[b]);
    }
#pragma unroll
    for (int b = 0; b < kRowsBatch; ++b)
#pragma unroll
      for (int c = 0; c < kRowsGroup; ++c)
        if (p[b][c] != 0xFFFFFFFFu) atomicAdd(h1(c) + (p[b][c] >> P.shift), 1u);
  }
  __syncthreads();
  if (warp < nc * 2) {
    const int c = warp >> 1, t = warp & 1;
    long long below;
    const long long k = t ? P.rule.k_hi : P.rule.k_lo;
    const int bin = rows_find_bin(h1(c), P.hi_bins, k, &below);
    if ((tid & 31) == 0) {
      s_bin[c][t] = bin;
      s_krem[c][t] = k - below;
    }
  }
  __syncthreads();
  for (int i = tid; i < nc * 2 * bins; i += kRowsThreads) rows_hist[i] = 0u;
  __syncthreads();
  for (int64_t j0 = tid; j0 < P.n_rows; j0 += (int64_t)kRowsThreads * kRowsBatch) {
    unsigned int p[kRowsBatch][kRowsGroup];
#pragma unroll
    for (int b = 0; b < kRowsBatch; ++b) {
      const int64_t j = j0 + (int64_t)b * kRowsThreads;
      const int64_t i = j < P.n_rows ? idx[j] : -1;
      const bool ok = (uint64_t)i < (uint64_t)P.n_rows;
      const int32_t* r = P.rank + (ok ? i : 0) * P.n_cols + d0;
      rows_load_ranks(r, ok, nc, vec, p[b]);
    }
#pragma unroll
    for (int b = 0; b < kRowsBatch; ++b)
#pragma unroll
      for (int c = 0; c < kRowsGroup; ++c) {
        if (p[b][c] == 0xFFFFFFFFu) continue;
        const int hb = (int)(p[b][c] >> P.shift);
        if (hb == s_bin[c][0]) atomicAdd(h2(c, 0) + (p[b][c] & lo_mask), 1u);
        if (two && hb == s_bin[c][1]) atomicAdd(h2(c, 1) + (p[b][c] & lo_mask), 1u);
      }
  }
  __syncthreads();
  if (warp < nc * 2) {
    const int c = warp >> 1, t = warp & 1;
    if (t == 0 || two) {
      long long below;
      const int bin = rows_find_bin(h2(c, t), lo_bins, s_krem[c][t], &below);
      if ((tid & 31) == 0) s_pos[c][t] = (s_bin[c][t] << P.shift) | bin;
    }
  }
  __syncthreads();
  if (tid < nc) {
    const double* col = P.sorted + (d0 + tid) * P.n_rows;
    const double a = col[s_pos[tid][0]];
    const double b = two ? col[s_pos[tid][1]] : a;
    P.out[(d0 + tid) * P.out_ld + set] = rank_combine(P.rule, a, b);
  }
}

// ---- K1m1: rank statistics of ONE resident column ---------------------------------------------
// np.median / np.quantile of a 1-D array that fits one cell (N <= 28 672) — the most common call.
// K1m spreads columns over warps and has a single column's searches done by one warp (N = 1e4,
// n = 1e4: 4.4 ms).  Here a group of threads owns a resample: the drawn rows' ranks (u16 table in
// shared memory, shared by the CTA) are counted into the group's own u16 count table (packed
// 32-bit atomics; a count cannot exceed N < 2^16), and a prefix scan over the N counts finds the
// positions of the wanted order statistics.  The draws are the single-cell Lemire stream, or the
// caller's index sets: row space, exactly what K1m computes.
// GROUP = 32 (a warp per resample, N <= 2048: eight resamples per CTA) or 256 (the CTA).
struct M1Params {
  const int32_t* rank;       // [N] position of row i in ascending order
  const double* sorted;      // [N]
  const int64_t* indices;    // [n_sets][N] or null: draw from (seed, resample_lo + set)
  int64_t n_rows, n_sets;
  uint64_t seed;
  int64_t resample_lo;
  RankRule rule;
  double* out;               // [n_sets]
  RoundKeys rk;
};

template <int GROUP>
__global__ void __launch_bounds__(256) k1m1_kernel(const M1Params P) {
  extern __shared__ __align__(16) unsigned char m1_smem[];
  constexpr int kGroups = 256 / GROUP;
  const int n = (int)P.n_rows;
  const int words = (n + 1) / 2;                                  // u16 pairs
  uint16_t* s_rank = reinterpret_cast<uint16_t*>(m1_smem);        // [n], padded to words * 2
  uint32_t* s_cnt_all = reinterpret_cast<uint32_t*>(m1_smem + (size_t)words * 4);
  __shared__ uint32_t s_part[8];
  __shared__ int s_pos[kGroups][2];
  const int tid = threadIdx.x, g = tid / GROUP, t = tid % GROUP;
  uint32_t* s_cnt = s_cnt_all + (size_t)g * words;
  auto group_sync = [&]() {
    if (GROUP == 32)
      __syncwarp();
    else
      __syncthreads();
  };
  for (int i = tid; i < n; i += 256) s_rank[i] = (uint16_t)__ldg(P.rank + i);
  __syncthreads();
  const uint32_t thresh = P.indices ? 0u : lemire_threshold((uint32_t)n);
  const uint32_t n_blocks = (uint32_t)((n + 3) / 4);
  const int two = P.rule.k_hi != P.rule.k_lo ? 1 : 0;
  // every group runs the same number of rounds (the CTA-wide barrier of GROUP = 256 needs it)
  const int64_t n_rounds = (P.n_sets + (int64_t)gridDim.x * kGroups - 1) / ((int64_t)gridDim.x * kGroups);
  for (int64_t round = 0; round < n_rounds; ++round) {
    const int64_t set = (round * gridDim.x + blockIdx.x) * kGroups + g;
    const bool live = set < P.n_sets;
    for (int i = t; i < words; i += GROUP) s_cnt[i] = 0u;
    group_sync();
    if (live) {
      auto count = [&](uint32_t row) {
        const uint32_t p = s_rank[row];
        atomicAdd(s_cnt + (p >> 1), 1u << (16u * (p & 1u)));
      };
      if (P.indices) {
        const int64_t* idx = P.indices + set * P.n_rows;
        for (int j = t; j < n; j += GROUP) {
          const int64_t i = idx[j];
          if ((uint64_t)i < (uint64_t)n) count((uint32_t)i);
        }
      } else {
        const Stream st = make_stream(P.seed, P.resample_lo + set, 0u, kPurposeIndex);
        for (uint32_t q = t; q < n_blocks; q += GROUP)
          lemire_block<true>(st, q, (uint32_t)n, thresh, (int64_t)n, [&](int, int64_t, uint32_t idx) { count(idx); }, &P.rk);
      }
    }
    group_sync();
    // prefix scan over the counts: thread t owns the words [w0, w1)
    const int per = (words + GROUP - 1) / GROUP;
    const int w0 = min(words, t * per), w1 = min(words, w0 + per);
    uint32_t mine = 0;
    for (int i = w0; i < w1; ++i) {
      const uint32_t c = s_cnt[i];
      mine += (c & 0xFFFFu) + (c >> 16);
    }
    uint32_t inc = mine;                                          // inclusive scan inside the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
      if ((tid & 31) >= off) inc += o;
    }
    uint32_t below = inc - mine;                                  // draws at positions before w0
    if (GROUP != 32) {
      if ((tid & 31) == 31) s_part[tid >> 5] = inc;
      __syncthreads();
      for (int w = 0; w < (tid >> 5); ++w) below += s_part[w];
    }
    if (live) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        if (which == 1 && !two) break;
        const long long k = which ? P.rule.k_hi : P.rule.k_lo;
        if (k >= (long long)below && k < (long long)below + (long long)mine) {
          long long acc = below;
          for (int i = w0; i < w1; ++i) {
            const uint32_t c = s_cnt[i];
            const long long c0 = c & 0xFFFFu, c1 = c >> 16;
            if (k < acc + c0) {
              s_pos[g][which] = 2 * i;
              break;
            }
            if (k < acc + c0 + c1) {
              s_pos[g][which] = 2 * i + 1;
              break;
            }
            acc += c0 + c1;
          }
        }
      }
    }
    group_sync();
    if (live && t == 0) {
      const double a = P.sorted[s_pos[g][0]];
      const double b = two ? P.sorted[s_pos[g][1]] : a;
      P.out[set] = rank_combine(P.rule, a, b);
    }
    group_sync();
  }
}

inline size_t m1_smem_bytes(int64_t n_rows, int group) {
  const size_t words = (size_t)(n_rows + 1) / 2;
  return words * 4 + (size_t)(256 / group) * words * 4;
}

// ---- NaN propagation for rank statistics ----------------------------------------------------
// np.median / np.quantile return NaN as soon as the sample holds one (the reference's
// statfunction(x[indices]) therefore gives NaN for every resample that draws a NaN row), whereas
// the rank kernels order NaNs last and would pick a finite order statistic.  Data with NaNs are
// rare, so the rank kernels stay as they are and the affected columns are patched afterwards:
// nan_columns_kernel flags them, nan_fixup_kernel sets the statistic of every index set that
// contains a NaN row of a flagged column to NaN, nan_poison_rows_kernel does the same for the
// jackknife summary (leave-one-out samples of a column with a NaN are NaN, bar at most one).
__global__ void __launch_bounds__(256) nan_columns_kernel(const double* __restrict__ data, int64_t n_rows,
                                                          int64_t n_cols, int64_t ld, int32_t* __restrict__ flags) {
  // 32 columns per CTA column (coalesced rows), 8 row groups; blockIdx.y strides over the rows so
  // that a single long column is still scanned by many CTAs; flags[n_cols] = any column flagged
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t d = (int64_t)blockIdx.x * 32 + cx;
  bool bad = false;
  if (d < n_cols)
    for (int64_t i = (int64_t)blockIdx.y * 8 + ry; i < n_rows; i += (int64_t)gridDim.y * 8) {
      const double v = data[i * ld + d];
      bad = bad || v != v;
    }
  if (bad) {
    atomicExch(flags + d, 1);
    atomicExch(flags + n_cols, 1);
  }
}

// out[set] = NaN when any of src[idx[set][j] * stride] is NaN; one warp per index set
__global__ void __launch_bounds__(256) nan_fixup_kernel(const double* __restrict__ src, int64_t stride, int64_t n_rows,
                                                        const int64_t* __restrict__ idx, int64_t n_sets,
                                                        double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t set = warp_global; set < n_sets; set += n_warps) {
    bool bad = false;
    for (int64_t j = lane; j < n_rows; j += 32) {
      const int64_t i = idx[set * n_rows + j];
      if ((uint64_t)i < (uint64_t)n_rows) {
        const double v = src[i * stride];
        bad = bad || v != v;
      }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) out[set] = nan("");
  }
}

// rows of a [n_cols][width] table whose column is flagged become NaN
__global__ void __launch_bounds__(256) nan_poison_rows_kernel(double* __restrict__ table, int64_t n_cols, int width,
                                                              const int32_t* __restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_cols * width) return;
  if (flags[e / width]) table[e] = nan("");
}

// shared memory of one CTA: the histograms are as long as the job's rank bits need (N = 1e6: 32 KB,
// four CTAs of 512 threads per SM; the 128 KB of N = 2^24 allow one:
// 25 % occupancy against random L2 gathers)
inline size_t rows_smem_bytes(int bins = kRowsMaxBins) { return (size_t)kRowsGroup * 2 * bins * sizeof(unsigned int); }

}  // namespace b200boot
