// K1m / K3m — column-batched resample median in rank space.
//
// Replaces `statfunction(x[indices])` with np.median(axis=0) over (N, D) data
// (/root/reference/src/scikits/bootstrap/bootstrap.py:431: all columns share one
// index set per resample) and the leave-one-out loop (:595-599) for the median.
//
// Each column is sorted once (bitonic on (key, row) pairs).  A resample is then
// described by the draw-count vector c[row] (shared by all columns); the median
// of the resampled column is the sorted value at the first sorted position whose
// cumulative count exceeds the middle rank (mean of the two middle ranks for
// even N, as np.median).  A CTA keeps the sorted-order row permutation of a group
// of columns in shared memory, rebuilds c[] for a few resamples at a time from
// the same Philox/Lemire stream K1 and K2 use for single-cell plans, and each
// warp scans one (resample, column) pair.
#pragma once
#include "common.cuh"
#include "draw.cuh"
#include "k1_resample.cuh"
#include "k45_interval.cuh"

namespace b200boot {

constexpr int kMedThreads = 1024;
constexpr int kMedWarps = kMedThreads / 32;
constexpr int64_t kMedMaxRows = kSmemDataBytes / 8;   // single-cell plans only (28 672 rows)

// ---- per-column (key, row) bitonic sort -------------------------------------------
__device__ __forceinline__ bool pair_gt(uint64_t ka, uint32_t ia, uint64_t kb, uint32_t ib) {
  return ka > kb || (ka == kb && ia > ib);
}

// keys[d][i] = key(data[i][d]), idx[d][i] = i; padding sorts last
__global__ void __launch_bounds__(256) psort_load_kernel(const double* __restrict__ data, int64_t n,
                                                         int64_t n_cols, int64_t ld, int64_t n_pad,
                                                         uint64_t* __restrict__ keys,
                                                         uint32_t* __restrict__ idx) {
  const int64_t total = n_pad * n_cols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t d = e % n_cols, i = e / n_cols;      // d fastest: coalesced reads of a row
    keys[d * n_pad + i] = i < n ? key_of(data[i * ld + d]) : ~0ull;
    idx[d * n_pad + i] = i < n ? (uint32_t)i : 0xFFFFFFFFu;
  }
}

__global__ void __launch_bounds__(256) psort_step_kernel(uint64_t* __restrict__ keys,
                                                         uint32_t* __restrict__ idx, int64_t n_pad,
                                                         int64_t j, int64_t k) {
  const int64_t d = blockIdx.y;
  uint64_t* kc = keys + d * n_pad;
  uint32_t* ic = idx + d * n_pad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
    const int64_t p = i ^ j;
    if (p > i) {
      const uint64_t a = kc[i], b = kc[p];
      const uint32_t ia = ic[i], ib = ic[p];
      const bool ascending = (i & k) == 0;
      if (pair_gt(a, ia, b, ib) == ascending) {
        kc[i] = b; kc[p] = a;
        ic[i] = ib; ic[p] = ia;
      }
    }
  }
}

__global__ void __launch_bounds__(1024) psort_local_kernel(uint64_t* __restrict__ keys,
                                                           uint32_t* __restrict__ idx, int64_t n_pad,
                                                           int64_t k_first, int64_t k_last,
                                                           int64_t j_first) {
  __shared__ uint64_t sk[2048];
  __shared__ uint32_t si[2048];
  const int64_t d = blockIdx.y;
  const int64_t seg0 = (int64_t)blockIdx.x * 2048;
  uint64_t* kc = keys + d * n_pad + seg0;
  uint32_t* ic = idx + d * n_pad + seg0;
  const int seg_n = (int)min((int64_t)2048, n_pad);
  for (int i = threadIdx.x; i < seg_n; i += 1024) {
    sk[i] = kc[i];
    si[i] = ic[i];
  }
  __syncthreads();
  for (int64_t k = k_first; k <= k_last; k <<= 1) {
    const int64_t j0 = (k == k_first) ? j_first : (k >> 1);
    for (int64_t j = j0; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < seg_n; i += 1024) {
        const int p = i ^ (int)j;
        if (p > i) {
          const uint64_t a = sk[i], b = sk[p];
          const uint32_t ia = si[i], ib = si[p];
          const bool ascending = ((seg0 + i) & k) == 0;
          if (pair_gt(a, ia, b, ib) == ascending) {
            sk[i] = b; sk[p] = a;
            si[i] = ib; si[p] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < seg_n; i += 1024) {
    kc[i] = sk[i];
    ic[i] = si[i];
  }
}

// Sorted values and the row permutation of every column in two shared-memory friendly
// layouts.  Sorted positions are cut into 32 lane segments of `seg` (even) positions:
// position s = lane * seg + i.
//   pairs[d][k * 32 + lane] (u32): two rows of lane's segment (low / high half), in the
//                               bank-friendly visiting order of median_schedule_kernel
//   lin  [d][s]                 (u16): row at position s
// Slots past N point at the always-zero count slot N.
__global__ void __launch_bounds__(256) median_pack_kernel(const uint64_t* __restrict__ keys,
                                                          const uint32_t* __restrict__ idx, int64_t n,
                                                          int64_t n_pad, int seg,
                                                          double* __restrict__ sorted_vals,
                                                          uint16_t* __restrict__ lin) {
  const int64_t d = blockIdx.y;
  const int64_t slots = (int64_t)seg * 32;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += stride) {
    lin[d * slots + s] = s < n ? (uint16_t)idx[d * n_pad + s] : (uint16_t)n;
    if (s < n) sorted_vals[d * n + s] = value_of(keys[d * n_pad + s]);
  }
}

// The segment sums do not care in which order a lane visits the positions of its own
// segment, so the visiting order is chosen to keep the count-table gathers of K1m off each
// other's banks.  With B interleaved tables a row's counts are one 2B-byte load; the lanes
// that share a shared-memory wavefront (8 for 16-byte loads, 16 for 8-byte, 32 for 2-byte)
// are conflict-free when their rows differ in `bank_key`.  Lane l therefore wants key
// (step + l) mod M for its first row of a step and (step + l + M/2) mod M for the second;
// rows are handed out per key in sorted order while they last, what is left fills the gaps.
__device__ __forceinline__ int bank_key(uint32_t row, int batch) {
  return batch == 8 ? (int)(row & 7u) : (batch == 4 ? (int)(row & 15u) : (int)((row >> 1) & 31u));
}

__global__ void __launch_bounds__(128) median_schedule_kernel(const uint32_t* __restrict__ idx, int64_t n,
                                                              int64_t n_pad, int seg, int batch,
                                                              int64_t n_cols, uint32_t* __restrict__ pairs) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_cols * 32) return;
  const int64_t d = g >> 5;
  const int lane = (int)(g & 31);
  const int M = batch == 8 ? 8 : (batch == 4 ? 16 : 32);
  const int64_t base = (int64_t)lane * seg;
  const uint32_t* col = idx + d * n_pad;
  auto row_at = [&](int p) -> uint32_t { return base + p < n ? col[base + p] : (uint32_t)n; };
  uint16_t* pr = reinterpret_cast<uint16_t*>(pairs + d * ((int64_t)seg * 16));
  auto slot_ptr = [&](int s) -> uint16_t* { return pr + ((int64_t)(s >> 1) * 32 + lane) * 2 + (s & 1); };

  int cursor[32];   // next unused position of each key inside the segment (seg: none left)
  for (int k = 0; k < M; ++k) cursor[k] = seg;
  for (int p = seg - 1; p >= 0; --p) cursor[bank_key(row_at(p), batch)] = p;
  for (int s = 0; s < seg; ++s) {
    const int want = ((s >> 1) + lane + (s & 1) * (M / 2)) & (M - 1);
    const int p = cursor[want];
    if (p < seg) {
      *slot_ptr(s) = (uint16_t)row_at(p);
      int q = p + 1;
      while (q < seg && bank_key(row_at(q), batch) != want) ++q;
      cursor[want] = q;
    } else {
      *slot_ptr(s) = 0xFFFFu;   // gap, filled below
    }
  }
  int s = 0;
  for (int k = 0; k < M; ++k) {
    for (int p = cursor[k]; p < seg; ++p) {
      const uint32_t r = row_at(p);
      if (bank_key(r, batch) != k) continue;
      while (*slot_ptr(s) != 0xFFFFu) ++s;
      *slot_ptr(s) = (uint16_t)r;
    }
  }
}

// ---- which order statistics make the statistic ----------------------------------------
// np.median is the mean of the two middle order statistics (one for odd N); np.quantile /
// np.percentile (method 'linear') interpolate between order statistics floor(h) and
// floor(h) + 1, h = (N - 1) q.  The host fixes the ranks and the weight (it knows numpy's
// rounding of h); the kernels fetch the two order statistics and combine them here.
struct RankRule {
  int64_t k_lo, k_hi;   // 0-based ranks, k_lo <= k_hi <= k_lo + 1
  double gamma;         // interpolation weight h - floor(h)   (mode 1)
  int32_t mode;         // 0: mean of the two (np.median), 1: numpy's _lerp
};
inline RankRule median_rule(int64_t n) { return RankRule{(n - 1) / 2, n / 2, 0.5, 0}; }

__device__ __forceinline__ double rank_combine(const RankRule& r, double a, double b) {
  if (r.mode == 0) return (a + b) / 2.0;
  // numpy/lib/_function_base_impl.py _lerp: a + (b - a) t, or b - (b - a)(1 - t) for t >= 0.5;
  // explicit roundings so that no multiply-add is contracted
  const double diff = __dsub_rn(b, a);
  return r.gamma >= 0.5 ? __dsub_rn(b, __dmul_rn(diff, __dsub_rn(1.0, r.gamma)))
                        : __dadd_rn(a, __dmul_rn(diff, r.gamma));
}

// ---- K1m ---------------------------------------------------------------------------
struct K1mParams {
  const double* sorted_vals;   // [D][N]
  const uint32_t* pairs;       // [D][seg * 16]
  const uint16_t* lin;         // [D][seg * 32]
  int64_t n_rows, n_cols;
  int32_t seg;                 // positions per lane segment (even)
  int32_t cols_per_item;       // columns whose permutation is resident per item
  int32_t n_colchunks;
  int32_t rchunk, n_rchunks;   // resamples per item
  int64_t n_items;
  uint64_t seed;
  int64_t resample_lo, n_local;
  double* out;                 // [D][n_local]
  int32_t* work_counter;
  int32_t count_words;         // uint32 words of one count table: ceil((N + 1) * B / 2), padded
  const int64_t* indices;      // parity mode: caller-supplied index sets [n_local][N], else null
  RankRule rule;
  int32_t tables;              // short-column kernel: groups of 8 count tables resident at once
  int32_t col_block;           // short-column kernel: columns per work item
  int32_t n_colblocks;
};

// First sorted position whose cumulative draw count (resample b of the batch) exceeds k.
// incl / excl: the lane's inclusive / exclusive segment prefix for that resample.
// With k2 >= k (the second rank of an even-N median or of an interpolated quantile) the
// same pass also looks for the first position whose cumulative count exceeds k2 and
// returns it through *pos2 — or -1 when that position lies beyond the chunk that holds
// the first one (the caller then searches for it separately; rare).
template <int B>
__device__ __forceinline__ int median_find(const uint16_t* __restrict__ cnt,
                                           const uint16_t* __restrict__ lin, int seg_len, int lane,
                                           int b, uint32_t incl, uint32_t excl, uint32_t k,
                                           uint32_t k2 = 0, int* pos2 = nullptr) {
  const uint32_t hit = __ballot_sync(0xffffffffu, incl > k);
  const int seg = __ffs(hit) - 1;                       // lane segment holding the crossing
  uint32_t running = __shfl_sync(0xffffffffu, excl, seg);
  const uint16_t* ls = lin + seg * seg_len;
  for (int c0 = 0; c0 < seg_len; c0 += 32) {
    const int i = c0 + lane;
    uint32_t inc = i < seg_len ? (uint32_t)cnt[(int)ls[i] * B + b] : 0u;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += o;
    }
    inc += running;
    const uint32_t h = __ballot_sync(0xffffffffu, inc > k);
    if (h) {
      if (pos2) {
        const uint32_t h2 = __ballot_sync(0xffffffffu, inc > k2);
        *pos2 = h2 ? seg * seg_len + c0 + __ffs(h2) - 1 : -1;
      }
      return seg * seg_len + c0 + __ffs(h) - 1;
    }
    running = __shfl_sync(0xffffffffu, inc, 31);
  }
  if (pos2) *pos2 = -1;
  return seg * seg_len + seg_len - 1;   // unreachable when counts sum to N
}

// Four resamples at a time for short segments (seg_len <= 64): eight lanes per resample,
// sub-lane u of group g scanning positions u*P .. u*P+P-1 (P = ceil(seg_len / 8)) of the
// segment that holds resample (b0 + g)'s crossing.  incl4 / excl4: this lane's inclusive /
// exclusive segment prefixes for resamples b0 .. b0+3.  Every lane returns the positions
// found for ITS group's resample; *pos_hi = -1 when the second rank lies beyond that
// segment (the caller falls back to median_find for it).
template <int B, int PMAX>
__device__ __forceinline__ int median_find4(const uint16_t* __restrict__ cnt,
                                            const uint16_t* __restrict__ lin, int seg_len, int lane, int b0,
                                            const uint32_t* incl4, const uint32_t* excl4, uint32_t k_lo,
                                            uint32_t k_hi, bool two, int* pos_hi) {
  const int g = lane >> 3, u = lane & 7;
  int my_seg = 0;
  uint32_t my_run = 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const uint32_t hit = __ballot_sync(0xffffffffu, incl4[r] > k_lo);
    const int sg = hit ? __ffs(hit) - 1 : 0;
    const uint32_t run = __shfl_sync(0xffffffffu, excl4[r], sg);
    if (g == r) {
      my_seg = sg;
      my_run = run;
    }
  }
  const int P = (seg_len + 7) >> 3;              // <= PMAX
  const uint16_t* ls = lin + my_seg * seg_len + u * P;
  const uint16_t* cb = cnt + (b0 + g);
  const int left = seg_len - u * P;              // positions of the segment from this sub-lane on
  uint32_t c[PMAX];
  uint32_t local = 0;
#pragma unroll
  for (int i = 0; i < PMAX; ++i) {
    c[i] = (i < P && i < left) ? (uint32_t)cb[(int)ls[i] * B] : 0u;
    local += c[i];
  }
  uint32_t inc = local;
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off, 8);
    if (u >= off) inc += o;
  }
  const uint32_t cum_incl = my_run + inc;
  // every sub-lane walks its own positions once for both ranks; the owners' results are taken
  uint32_t running = cum_incl - local;
  int w_lo = P - 1, w_hi = P - 1;
  bool f_lo = false, f_hi = false;
#pragma unroll
  for (int i = 0; i < PMAX; ++i) {
    running += c[i];
    if (!f_lo && running > k_lo) {
      w_lo = i;
      f_lo = true;
    }
    if (!f_hi && running > k_hi) {
      w_hi = i;
      f_hi = true;
    }
  }
  const int base = my_seg * seg_len + u * P;
  const uint32_t m_lo = (__ballot_sync(0xffffffffu, cum_incl > k_lo) >> (8 * g)) & 0xFFu;
  const int own_lo = m_lo ? __ffs(m_lo) - 1 : 7;
  const int pl = __shfl_sync(0xffffffffu, base + w_lo, 8 * g + own_lo);
  int ph = -1;
  if (two) {
    const uint32_t m_hi = (__ballot_sync(0xffffffffu, cum_incl > k_hi) >> (8 * g)) & 0xFFu;
    const int own_hi = m_hi ? __ffs(m_hi) - 1 : 7;
    const int w = __shfl_sync(0xffffffffu, base + w_hi, 8 * g + own_hi);
    ph = m_hi ? w : -1;
  }
  *pos_hi = ph;
  return pl;
}

// B = resamples whose draw-count tables are interleaved in shared memory
// (cnt[row * B + b], u16): one wide load per sorted position serves B resamples and the
// B segment sums are carried packed (two u16 per 32-bit add).
template <int B>
__global__ void __launch_bounds__(kMedThreads, 1) k1m_median_kernel(const K1mParams P) {
  static_assert(B == 8 || B == 4 || B == 1, "batch of 8 / 4 (interleaved) or 1");
  constexpr int NP = B == 1 ? 1 : B / 2;           // packed u16-pair accumulators
  extern __shared__ __align__(128) unsigned char smem_raw[];
  volatile int* s_item = reinterpret_cast<volatile int*>(smem_raw);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_raw + 128);                 // count table
  uint32_t* s_pairs = s_cnt + P.count_words;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg_len = P.seg;
  const int pair_words = seg_len * 16;             // per column
  const int n = (int)P.n_rows;
  const uint32_t thresh = lemire_threshold((uint32_t)n);
  const int n_blocks = (n + 3) / 4;
  const uint32_t k_hi = (uint32_t)P.rule.k_hi, k_lo = (uint32_t)P.rule.k_lo;
  const bool two = k_lo != k_hi;                   // e.g. the two middle ranks of an even N
  const uint16_t* cnt16 = reinterpret_cast<const uint16_t*>(s_cnt);

  int resident = -1;
  for (;;) {
    if (tid == 0) *s_item = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int64_t item = *s_item;
    __syncthreads();
    if (item >= P.n_items) break;
    const int cc = (int)(item / P.n_rchunks);
    const int rc = (int)(item - (int64_t)cc * P.n_rchunks);
    const int64_t d0 = (int64_t)cc * P.cols_per_item;
    const int dc = (int)min((int64_t)P.cols_per_item, P.n_cols - d0);
    uint16_t* s_lin = reinterpret_cast<uint16_t*>(s_pairs + (size_t)P.cols_per_item * pair_words);
    if (cc != resident) {
      const uint32_t* src = P.pairs + d0 * pair_words;
      for (int64_t e = tid; e < (int64_t)dc * pair_words; e += kMedThreads) s_pairs[e] = src[e];
      const uint32_t* lsrc = reinterpret_cast<const uint32_t*>(P.lin + d0 * (int64_t)seg_len * 32);
      uint32_t* ldst = reinterpret_cast<uint32_t*>(s_lin);
      for (int64_t e = tid; e < (int64_t)dc * pair_words; e += kMedThreads) ldst[e] = lsrc[e];
      resident = cc;
      __syncthreads();
    }
    const int64_t r_begin = (int64_t)rc * P.rchunk;
    const int64_t r_end = min(r_begin + (int64_t)P.rchunk, P.n_local);
    for (int64_t rb = r_begin; rb < r_end; rb += B) {
      const int nb = (int)min((int64_t)B, r_end - rb);
      for (int e = tid; e < P.count_words; e += kMedThreads) s_cnt[e] = 0u;
      __syncthreads();
      auto bump = [&](int b, uint32_t row) {
        const uint32_t slot = row * B + b;
        atomicAdd(s_cnt + (slot >> 1), 1u << ((slot & 1u) * 16));
      };
      if (P.indices) {   // parity mode: count the caller's indices
        for (int e = tid; e < nb * n; e += kMedThreads) {
          const int b = e / n, pos = e - b * n;
          bump(b, (uint32_t)P.indices[(rb + b) * (int64_t)n + pos]);
        }
      } else {           // the single-cell Lemire stream of K1 / K2 (cell 0)
        for (int e = tid; e < nb * n_blocks; e += kMedThreads) {
          const int b = e / n_blocks, q = e - b * n_blocks;
          const Stream st = make_stream(P.seed, P.resample_lo + rb + b, 0u, kPurposeIndex);
          lemire_block<true>(st, (uint32_t)q, (uint32_t)n, thresh, (int64_t)n,
                             [&](int, int64_t, uint32_t idx) { bump(b, idx); });
        }
      }
      __syncthreads();
      for (int j = warp; j < dc; j += kMedWarps) {
        const uint32_t* pp = s_pairs + (size_t)j * pair_words;
        const uint16_t* lin = s_lin + (size_t)j * seg_len * 32;
        // packed segment sums: resamples (2p, 2p+1) in acc[p]
        uint32_t acc[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p] = 0u;
        for (int i2 = 0; i2 < seg_len / 2; ++i2) {
          const uint32_t two = pp[i2 * 32 + lane];
          const uint32_t r0 = two & 0xFFFFu, r1 = two >> 16;
          if (B == 8) {
            const uint4 a = *reinterpret_cast<const uint4*>(cnt16 + r0 * 8);
            const uint4 c = *reinterpret_cast<const uint4*>(cnt16 + r1 * 8);
            acc[0] += a.x + c.x;
            acc[1 % NP] += a.y + c.y;
            acc[2 % NP] += a.z + c.z;
            acc[3 % NP] += a.w + c.w;
          } else if (B == 4) {
            const uint2 a = *reinterpret_cast<const uint2*>(cnt16 + r0 * 4);
            const uint2 c = *reinterpret_cast<const uint2*>(cnt16 + r1 * 4);
            acc[0] += a.x + c.x;
            acc[1 % NP] += a.y + c.y;
          } else {
            acc[0] += (uint32_t)cnt16[r0] + (uint32_t)cnt16[r1];
          }
        }
        // packed inclusive scans over the lanes (totals are N < 2^16: no carry between halves)
        uint32_t inc[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) inc[p] = acc[p];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc[p], off);
            if (lane >= off) inc[p] += o;
          }
        }
        const double* sv = P.sorted_vals + (d0 + j) * P.n_rows;
        auto prefix_of = [&](int b, uint32_t* incl, uint32_t* excl) {
          const uint32_t pi = inc[(b / 2) % NP], pa = acc[(b / 2) % NP];
          *incl = (b & 1) ? (pi >> 16) : (pi & 0xFFFFu);
          *excl = *incl - ((b & 1) ? (pa >> 16) : (pa & 0xFFFFu));
        };
        if (B >= 4 && seg_len <= 64) {
          // short segments: four resamples per pass, eight lanes each
#pragma unroll
          for (int b0 = 0; b0 < B; b0 += 4) {
            if (b0 < nb) {
              uint32_t incl4[4], excl4[4];
#pragma unroll
              for (int r = 0; r < 4; ++r) prefix_of(b0 + r, &incl4[r], &excl4[r]);
              int s_hi;
              const int s_lo = seg_len <= 32
                                   ? median_find4<B, 4>(cnt16, lin, seg_len, lane, b0, incl4, excl4, k_lo, k_hi, two, &s_hi)
                                   : median_find4<B, 8>(cnt16, lin, seg_len, lane, b0, incl4, excl4, k_lo, k_hi, two, &s_hi);
              const int g = lane >> 3;
              const bool valid = b0 + g < nb;
              if (two) {
                // second rank beyond the crossing segment (rare): full-warp search for that resample
                const uint32_t miss = __ballot_sync(0xffffffffu, valid && s_hi < 0);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                  if ((miss >> (8 * r)) & 1u) {
                    const int f = median_find<B>(cnt16, lin, seg_len, lane, b0 + r, incl4[r], excl4[r], k_hi);
                    if (g == r) s_hi = f;
                  }
                }
              }
              if (valid && (lane & 7) == 0) {
                const double med = two ? rank_combine(P.rule, sv[s_lo], sv[s_hi]) : sv[s_lo];
                P.out[(d0 + j) * P.n_local + rb + b0 + g] = med;
              }
            }
          }
        } else {
#pragma unroll
          for (int b = 0; b < B; ++b) {
            if (b < nb) {
              uint32_t incl, excl;
              prefix_of(b, &incl, &excl);
              double med;
              if (two) {
                // one pass finds the lower rank and, almost always, the upper one next to it
                int s_hi;
                const int s_lo = median_find<B>(cnt16, lin, seg_len, lane, b, incl, excl, k_lo, k_hi, &s_hi);
                if (s_hi < 0) s_hi = median_find<B>(cnt16, lin, seg_len, lane, b, incl, excl, k_hi);
                med = rank_combine(P.rule, sv[s_lo], sv[s_hi]);
              } else {
                med = sv[median_find<B>(cnt16, lin, seg_len, lane, b, incl, excl, k_hi)];
              }
              if (lane == 0) P.out[(d0 + j) * P.n_local + rb + b] = med;
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- K1m for short columns (segments of at most 64 positions, eight interleaved tables) ----
// The draw-count tables depend on the resample only, so rebuilding them for every group of
// resident columns (k1m_median_kernel) repeats the index draws D/32 times.  Here a work item
// is (block of many columns) x (T*8 resamples): the CTA builds T groups of eight interleaved
// tables once, then every warp walks its columns of the block: the column's sorted-order
// permutation is read straight from global memory (coalesced; `lin` is staged in a private
// 64*seg-byte slice for the searches) and used against all T table groups.
__global__ void __launch_bounds__(kMedThreads, 1) k1m_short_kernel(const K1mParams P) {
  constexpr int B = 8, NP = 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  volatile int* s_item = reinterpret_cast<volatile int*>(smem_raw);
  uint32_t* s_cnt = reinterpret_cast<uint32_t*>(smem_raw + 128);                 // T table groups
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg_len = P.seg;
  const int pair_words = seg_len * 16;             // per column (pairs, and lin as u32 words)
  uint32_t* s_lin32 = s_cnt + (size_t)P.tables * P.count_words + (size_t)warp * pair_words;
  const uint16_t* s_lin = reinterpret_cast<const uint16_t*>(s_lin32);
  const int n = (int)P.n_rows;
  const uint32_t thresh = lemire_threshold((uint32_t)n);
  const int n_blocks = (n + 3) / 4;
  const uint32_t k_hi = (uint32_t)P.rule.k_hi, k_lo = (uint32_t)P.rule.k_lo;
  const bool two = k_lo != k_hi;
  const int res_per_item = P.tables * B;

  for (;;) {
    if (tid == 0) *s_item = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int64_t item = *s_item;
    __syncthreads();
    if (item >= P.n_items) break;
    const int rc = (int)(item / P.n_colblocks);
    const int cb = (int)(item - (int64_t)rc * P.n_colblocks);
    const int64_t r_begin = (int64_t)rc * res_per_item;
    const int nres = (int)min((int64_t)res_per_item, P.n_local - r_begin);
    const int64_t c_begin = (int64_t)cb * P.col_block;
    const int64_t c_end = min(c_begin + (int64_t)P.col_block, P.n_cols);

    for (int e = tid; e < P.tables * P.count_words; e += kMedThreads) s_cnt[e] = 0u;
    __syncthreads();
    auto bump = [&](int r, uint32_t row) {       // resample r of the item: table group r / 8, slot r % 8
      const uint32_t slot = row * B + (uint32_t)(r & 7);
      atomicAdd(s_cnt + (size_t)(r >> 3) * P.count_words + (slot >> 1), 1u << ((slot & 1u) * 16));
    };
    if (P.indices) {
      for (int e = tid; e < nres * n; e += kMedThreads) {
        const int r = e / n, pos = e - r * n;
        bump(r, (uint32_t)P.indices[(r_begin + r) * (int64_t)n + pos]);
      }
    } else {
      for (int e = tid; e < nres * n_blocks; e += kMedThreads) {
        const int r = e / n_blocks, q = e - r * n_blocks;
        const Stream st = make_stream(P.seed, P.resample_lo + r_begin + r, 0u, kPurposeIndex);
        lemire_block<true>(st, (uint32_t)q, (uint32_t)n, thresh, (int64_t)n,
                           [&](int, int64_t, uint32_t idx) { bump(r, idx); });
      }
    }
    __syncthreads();

    const int n_groups = (nres + B - 1) / B;
    for (int64_t j = c_begin + warp; j < c_end; j += kMedWarps) {
      // stage this column's linear permutation in the warp's slice
      {
        const uint4* src = reinterpret_cast<const uint4*>(P.lin + j * (int64_t)seg_len * 32);
        uint4* dst = reinterpret_cast<uint4*>(s_lin32);
        for (int w = lane; w < pair_words / 4; w += 32) dst[w] = __ldg(src + w);
      }
      __syncwarp();
      const uint32_t* pp = P.pairs + j * (int64_t)pair_words;
      const double* sv = P.sorted_vals + j * P.n_rows;
      for (int t = 0; t < n_groups; ++t) {
        const int nb = min(B, nres - t * B);
        const uint16_t* cnt16 = reinterpret_cast<const uint16_t*>(s_cnt + (size_t)t * P.count_words);
        uint32_t acc[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) acc[p] = 0u;
#pragma unroll 4
        for (int i2 = 0; i2 < seg_len / 2; ++i2) {
          const uint32_t pr = __ldg(pp + i2 * 32 + lane);
          const uint32_t r0 = pr & 0xFFFFu, r1 = pr >> 16;
          const uint4 a = *reinterpret_cast<const uint4*>(cnt16 + r0 * 8);
          const uint4 c = *reinterpret_cast<const uint4*>(cnt16 + r1 * 8);
          acc[0] += a.x + c.x;
          acc[1] += a.y + c.y;
          acc[2] += a.z + c.z;
          acc[3] += a.w + c.w;
        }
        uint32_t inc[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) inc[p] = acc[p];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc[p], off);
            if (lane >= off) inc[p] += o;
          }
        }
#pragma unroll
        for (int b0 = 0; b0 < B; b0 += 4) {
          if (b0 < nb) {
            uint32_t incl4[4], excl4[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int b = b0 + r;
              const uint32_t pi = inc[b / 2], pa = acc[b / 2];
              incl4[r] = (b & 1) ? (pi >> 16) : (pi & 0xFFFFu);
              excl4[r] = incl4[r] - ((b & 1) ? (pa >> 16) : (pa & 0xFFFFu));
            }
            int s_hi;
            const int s_lo = seg_len <= 32
                                 ? median_find4<B, 4>(cnt16, s_lin, seg_len, lane, b0, incl4, excl4, k_lo, k_hi, two, &s_hi)
                                 : median_find4<B, 8>(cnt16, s_lin, seg_len, lane, b0, incl4, excl4, k_lo, k_hi, two, &s_hi);
            const int g = lane >> 3;
            const bool valid = b0 + g < nb;
            if (two) {
              const uint32_t miss = __ballot_sync(0xffffffffu, valid && s_hi < 0);
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                if ((miss >> (8 * r)) & 1u) {
                  const int f = median_find<B>(cnt16, s_lin, seg_len, lane, b0 + r, incl4[r], excl4[r], k_hi);
                  if (g == r) s_hi = f;
                }
              }
            }
            if (valid && (lane & 7) == 0) {
              const double med = two ? rank_combine(P.rule, sv[s_lo], sv[s_hi]) : sv[s_lo];
              P.out[j * P.n_local + r_begin + t * B + b0 + g] = med;
            }
          }
        }
      }
      __syncwarp();   // the slice is overwritten by the next column
    }
    __syncthreads();
  }
}

// ---- K3m: leave-one-out medians by rank of the deleted row ---------------------------
__device__ __forceinline__ double loo_median(const double* __restrict__ sv, int64_t n, int64_t r,
                                             const RankRule& loo) {
  const int64_t m = n - 1;
  if (m <= 0) return NAN;
  auto kth = [&](int64_t k) { return sv[k < r ? k : k + 1]; };
  const double hi = kth(loo.k_hi);
  return loo.k_lo == loo.k_hi ? hi : rank_combine(loo, kth(loo.k_lo), hi);
}

__device__ __forceinline__ double block_sum_256(double v, double* s_red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < 8; ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(256) k3m_jackknife_median_kernel(const double* __restrict__ sorted_vals,
                                                                   int64_t n, const RankRule full,
                                                                   const RankRule loo,
                                                                   double* __restrict__ out) {
  __shared__ double s_red[8];
  const int64_t d = blockIdx.x;
  const double* sv = sorted_vals + d * n;
  // accumulate relative to the first leave-one-out value: identical values give
  // d == 0 exactly (undefined acceleration) rather than rounding noise
  const double u = loo_median(sv, n, 0, loo);
  double acc = 0.0;
  for (int64_t r = threadIdx.x; r < n; r += 256) acc += loo_median(sv, n, r, loo) - u;
  const double jmean = u + block_sum_256(acc, s_red) / (double)n;
  double a2 = 0.0, a3 = 0.0;
  for (int64_t r = threadIdx.x; r < n; r += 256) {
    const double dv = jmean - loo_median(sv, n, r, loo);
    a2 += dv * dv;
    a3 += dv * dv * dv;
  }
  const double d2 = block_sum_256(a2, s_red);
  const double d3 = block_sum_256(a3, s_red);
  if (threadIdx.x == 0) {
    double* o = out + d * 8;
    o[0] = full.k_lo == full.k_hi ? sv[full.k_hi] : rank_combine(full, sv[full.k_lo], sv[full.k_hi]);
    o[1] = jmean;
    o[2] = d2;
    o[3] = d3;
    o[4] = d3 / (6.0 * pow(d2, 1.5));
    o[5] = o[6] = o[7] = 0.0;
  }
}

// ---- K1mt: median of a long 1-D array in rank space ----------------------------------
// For N beyond the shared-memory-resident limit the data are sorted once and the tile
// plan of K1 is laid over the SORTED array.  The median's tile follows from the tile
// occupancies alone (prefix sum over the K0 counts); only that one tile's draws are
// generated (same class / Lemire streams as K1 and K2) and histogrammed over the
// tile's positions, and a scan finds the sorted position holding the middle rank.
// Cost per resample is O(T + tile) instead of O(N).  The resample is, by
// construction, sorted(x)[bootstrap_indices(...)]: see DESIGN.md.
constexpr int kMtThreads = 256;

struct K1mtParams {
  const double* sorted_vals;   // [N] ascending
  Plan plan;
  uint64_t seed;
  int64_t resample_lo, n_local;
  const int32_t* counts;       // [n_tiles][n_local]
  double* out;                 // [n_local]
  RankRule rule;
};

// CTA-wide exclusive scan of one value per thread; returns the exclusive prefix and the
// total through *total (s_scan holds kMtThreads / 32 + 1 words)
__device__ __forceinline__ uint32_t cta_exclusive_scan(uint32_t v, uint32_t* s_scan, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  __syncthreads();
  if (lane == 31) s_scan[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
  for (int w = 0; w < kMtThreads / 32; ++w) {
    const uint32_t x = s_scan[w];
    if (w < warp) base += x;
    tot += x;
  }
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(kMtThreads) k1mt_median_tiled_kernel(const K1mtParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* s_hist = reinterpret_cast<uint32_t*>(smem_raw);              // packed u16 pairs, tile/2 words
  __shared__ uint32_t s_scan[kMtThreads / 32 + 1];
  __shared__ int32_t s_cls[kClasses];
  __shared__ int32_t s_blk[kClasses + 1];
  __shared__ long long s_found[2];                                       // tile, draws before it
  __shared__ int s_pos;
  const int tid = threadIdx.x;
  const Plan& plan = P.plan;
  const int64_t n = plan.n_rows;
  const int64_t tile = plan.tile();
  const uint32_t m7 = ((1u << (plan.log2_tile - 4)) - 1u) << 7;
  const int n_ranks = P.rule.k_lo == P.rule.k_hi ? 1 : 2;
  const int64_t rank0 = P.rule.k_lo;                                      // 0-based rank(s) wanted

  for (int64_t rl = blockIdx.x; rl < P.n_local; rl += gridDim.x) {
    const int64_t r = P.resample_lo + rl;
    int64_t hist_tile = -1;
    double vals[2] = {0.0, 0.0};
    for (int which = 0; which < n_ranks; ++which) {
      const int64_t rank = rank0 + which;
      // (a) tile holding the rank: first t with cumulative occupancy > rank
      if (tid == 0) s_found[0] = -1;
      __syncthreads();
      int64_t carried = 0;
      for (int64_t t0 = 0; t0 < plan.n_tiles; t0 += kMtThreads) {
        const int64_t t = t0 + tid;
        const uint32_t c = t < plan.n_tiles ? (uint32_t)__ldg(P.counts + t * P.n_local + rl) : 0u;
        uint32_t total;
        const uint32_t excl = cta_exclusive_scan(c, s_scan, &total);
        const int64_t before = carried + excl;
        if (t < plan.n_tiles && before <= rank && rank < before + (int64_t)c) {
          s_found[0] = t;
          s_found[1] = before;
        }
        carried += total;
        __syncthreads();
        if (s_found[0] >= 0) break;
      }
      const int64_t t_star = s_found[0];
      const uint32_t j = (uint32_t)(rank - s_found[1]);                    // rank inside the tile
      const int64_t sz = plan.cell_size(t_star);
      // (b) histogram of the tile's draws over its positions
      if (t_star != hist_tile) {
        const int words = (int)((sz + 2) / 2);
        for (int i = tid; i < words; i += kMtThreads) s_hist[i] = 0u;
        const int64_t n_t = (int64_t)__ldg(P.counts + t_star * P.n_local + rl);
        if (t_star < plan.n_full) {
          if (tid == 0) {
            class_count_tree<int32_t>(P.seed, r, t_star, n_t, s_cls, 1);
            int acc = 0;
            for (int c = 0; c < kClasses; ++c) {
              s_blk[c] = acc;
              acc += (s_cls[c] + kClassDrawsPerBlock - 1) / kClassDrawsPerBlock;
            }
            s_blk[kClasses] = acc;
          }
          __syncthreads();
          const int n_blocks = s_blk[kClasses];
          for (int b = tid; b < n_blocks; b += kMtThreads) {
            int c = 0;
            while (b >= s_blk[c + 1]) ++c;
            const Stream st = make_stream(P.seed, r, class_cell_id(t_star, c), kPurposeIndex);
            class_block<true>(st, (uint32_t)(b - s_blk[c]), m7, c, (int64_t)s_cls[c],
                              [&](int, int64_t, uint32_t idx) { atomicAdd(s_hist + (idx >> 1), 1u << ((idx & 1u) * 16)); });
          }
        } else {
          __syncthreads();
          const Stream st = make_stream(P.seed, r, (uint32_t)t_star, kPurposeIndex);
          const uint32_t thresh = lemire_threshold((uint32_t)sz);
          const int n_blocks = (int)((n_t + 3) / 4);
          for (int b = tid; b < n_blocks; b += kMtThreads)
            lemire_block<true>(st, (uint32_t)b, (uint32_t)sz, thresh, n_t,
                               [&](int, int64_t, uint32_t idx) { atomicAdd(s_hist + (idx >> 1), 1u << ((idx & 1u) * 16)); });
        }
        __syncthreads();
        hist_tile = t_star;
      }
      // (c) first position whose cumulative count exceeds j
      const uint16_t* h16 = reinterpret_cast<const uint16_t*>(s_hist);
      const int per = (int)((sz + kMtThreads - 1) / kMtThreads);
      const int p0 = tid * per;
      uint32_t mine = 0;
      for (int i = 0; i < per; ++i)
        if (p0 + i < sz) mine += h16[p0 + i];
      uint32_t total;
      const uint32_t excl = cta_exclusive_scan(mine, s_scan, &total);
      if (excl <= j && j < excl + mine) {
        uint32_t run = excl;
        for (int i = 0; i < per; ++i) {
          run += h16[p0 + i];
          if (run > j) {
            s_pos = p0 + i;
            break;
          }
        }
      }
      __syncthreads();
      const int64_t row0 = t_star < plan.n_full ? (t_star << plan.log2_tile) : (plan.n_full << plan.log2_tile);
      vals[which] = P.sorted_vals[row0 + s_pos];
      __syncthreads();
    }
    if (tid == 0) P.out[rl] = n_ranks == 1 ? vals[0] : rank_combine(P.rule, vals[0], vals[1]);
    (void)tile;
    (void)n;
  }
}

// ---- host side ---------------------------------------------------------------------
struct MedianWs {
  int32_t* work_counter;
  uint64_t* keys;
  uint32_t* idx;
  double* sorted_vals;
  uint32_t* pairs;
  uint16_t* lin;
  int64_t n_pad;
  int seg;          // positions per lane segment, even
};

inline MedianWs median_carve(Carver& cv, int64_t n_rows, int64_t n_cols) {
  MedianWs w{};
  w.n_pad = 1;
  while (w.n_pad < n_rows) w.n_pad <<= 1;
  w.seg = (int)(((n_rows + 31) / 32 + 1) & ~(int64_t)1);
  w.work_counter = cv.take<int32_t>(64);
  w.keys = cv.take<uint64_t>((size_t)w.n_pad * n_cols);
  w.idx = cv.take<uint32_t>((size_t)w.n_pad * n_cols);
  w.sorted_vals = cv.take<double>((size_t)n_rows * n_cols);
  w.pairs = cv.take<uint32_t>((size_t)w.seg * 16 * n_cols);
  w.lin = cv.take<uint16_t>((size_t)w.seg * 32 * n_cols);
  return w;
}

inline size_t median_ws_bytes(int64_t n_rows, int64_t n_cols, int64_t n_local) {
  if (n_rows > kMedMaxRows) {   // sorted 1-D path: tile occupancies only
    const Plan plan = make_plan(n_rows, 1);
    return pad256((size_t)plan.n_tiles * (size_t)(n_local > 0 ? n_local : 1) * 4);
  }
  Carver cv(reinterpret_cast<void*>(256), ~(size_t)0 >> 1);
  median_carve(cv, n_rows, n_cols);
  return cv.used;
}

inline int median_check(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, void* ws) {
  if (!data || n_rows < 1 || n_cols < 1 || ld < n_cols) return fail(B200BOOT_EINVAL, "median: bad arguments");
  if (n_rows > kMedMaxRows)
    return fail(B200BOOT_EUNSUPPORTED,
                "median: n_rows=%lld exceeds the shared-memory-resident limit of %lld rows",
                (long long)n_rows, (long long)kMedMaxRows);
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  return 0;
}

// bitonic sort of (key, idx) pairs: n_seq sequences of n_pad (power of two) pairs each
inline int psort_pairs(uint64_t* keys, uint32_t* idx, int64_t n_pad, unsigned n_seq, cudaStream_t s) {
  if (n_pad <= 1) return 0;
  const int64_t seg = n_pad < 2048 ? n_pad : 2048;
  const dim3 glocal((unsigned)(n_pad / seg), n_seq);
  int64_t fb = (n_pad + 1023) / 1024;
  if (fb > 1024) fb = 1024;
  const dim3 gflat((unsigned)fb, n_seq);
  psort_local_kernel<<<glocal, 1024, 0, s>>>(keys, idx, n_pad, 2, seg, 1);
  B2_LAUNCH_CHECK();
  for (int64_t k = seg * 2; k <= n_pad; k <<= 1) {
    for (int64_t j = k >> 1; j >= 2048; j >>= 1) {
      psort_step_kernel<<<gflat, 256, 0, s>>>(keys, idx, n_pad, j, k);
      B2_LAUNCH_CHECK();
    }
    psort_local_kernel<<<glocal, 1024, 0, s>>>(keys, idx, n_pad, k, k, 1024);
    B2_LAUNCH_CHECK();
  }
  return 0;
}

// sorts every column and packs sorted values + permutation into the workspace
inline int median_batch(int64_t n_rows) {
  // eight / four interleaved count tables when they fit in 64 KB, else one resample at a time
  return (n_rows + 1) * 16 <= 64 * 1024 ? 8 : ((n_rows + 1) * 8 <= 64 * 1024 ? 4 : 1);
}

inline int median_prepare(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, MedianWs& w,
                          cudaStream_t s) {
  const int64_t n_pad = w.n_pad;
  const int64_t total = n_pad * n_cols;
  int64_t lb = (total + 255) / 256;
  if (lb > 148 * 32) lb = 148 * 32;
  psort_load_kernel<<<(int)lb, 256, 0, s>>>(data, n_rows, n_cols, ld, n_pad, w.keys, w.idx);
  B2_LAUNCH_CHECK();
  for (int64_t d0 = 0; d0 < n_cols; d0 += 65535) {
    const unsigned nd = (unsigned)(n_cols - d0 < 65535 ? n_cols - d0 : 65535);
    uint64_t* kc = w.keys + d0 * n_pad;
    uint32_t* ic = w.idx + d0 * n_pad;
    int rc = psort_pairs(kc, ic, n_pad, nd, s);
    if (rc) return rc;
    const dim3 gpack((unsigned)((w.seg * 32 + 255) / 256), nd);
    median_pack_kernel<<<gpack, 256, 0, s>>>(kc, ic, n_rows, n_pad, w.seg, w.sorted_vals + d0 * n_rows,
                                             w.lin + d0 * w.seg * 32);
    B2_LAUNCH_CHECK();
    median_schedule_kernel<<<(unsigned)(((int64_t)nd * 32 + 127) / 128), 128, 0, s>>>(
        ic, n_rows, n_pad, w.seg, median_batch(n_rows), (int64_t)nd, w.pairs + d0 * w.seg * 16);
    B2_LAUNCH_CHECK();
  }
  return 0;
}

inline int median_resample(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, uint64_t seed,
                           int64_t resample_lo, int64_t resample_hi, double* stat_out, void* ws,
                           size_t ws_bytes, cudaStream_t s, const RankRule& rule,
                           const int64_t* indices = nullptr, double* jack_out = nullptr,
                           const RankRule* loo = nullptr) {
  int rc = median_check(data, n_rows, n_cols, ld, ws);
  if (rc) return rc;
  const int64_t n_local = resample_hi - resample_lo;
  if (n_local < 0 || resample_lo < 0) return fail(B200BOOT_EINVAL, "bad resample range");
  if (n_local == 0) return 0;
  if (!stat_out) return fail(B200BOOT_EINVAL, "null output");
  Carver cv(ws, ws_bytes);
  MedianWs w = median_carve(cv, n_rows, n_cols);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  rc = median_prepare(data, n_rows, n_cols, ld, w, s);
  if (rc) return rc;
  if (jack_out && loo) {   // the sorted columns are at hand: K3m costs one small launch
    for (int64_t d0 = 0; d0 < n_cols; d0 += 1 << 30) {
      const int64_t nd = std::min<int64_t>((int64_t)1 << 30, n_cols - d0);
      k3m_jackknife_median_kernel<<<(unsigned)nd, 256, 0, s>>>(w.sorted_vals + d0 * n_rows, n_rows, rule, *loo,
                                                               jack_out + d0 * 8);
      B2_LAUNCH_CHECK();
    }
  }

  K1mParams P{};
  P.sorted_vals = w.sorted_vals;
  P.pairs = w.pairs;
  P.lin = w.lin;
  P.n_rows = n_rows;
  P.n_cols = n_cols;
  P.seg = w.seg;
  const int batch = median_batch(n_rows);
  P.count_words = (int32_t)((((n_rows + 1) * batch + 1) / 2 + 31) & ~(int64_t)31);
  const int64_t cnt_bytes = (int64_t)P.count_words * 4;
  const int64_t perm_bytes = (int64_t)w.seg * 32 * 4;      // pairs + linear copy
  const int sms = sm_count();
  P.seed = seed;
  P.resample_lo = resample_lo;
  P.n_local = n_local;
  P.out = stat_out;
  P.work_counter = w.work_counter;
  P.indices = indices;
  P.rule = rule;
  B2_CUDA(cudaMemsetAsync(w.work_counter, 0, 256, s));
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    const int bytes = (int)(kSmemHeader + kSmemDataBytes);
    cudaError_t e = cudaFuncSetAttribute(k1m_median_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k1m_median_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k1m_median_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k1m_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    return e;
  }));

  // short columns: tables for many resamples stay resident, columns stream through
  const int64_t slice_bytes = (int64_t)w.seg * 64 * kMedWarps;            // one lin slice per warp
  const int64_t tables = std::min<int64_t>(8, (kSmemDataBytes - 128 - slice_bytes) / cnt_bytes);
  if (batch == 8 && w.seg <= 64 && tables >= 2) {
    P.tables = (int32_t)tables;
    const int64_t res_per_item = tables * 8;
    P.n_rchunks = (int32_t)((n_local + res_per_item - 1) / res_per_item);
    int64_t ncb = ((int64_t)sms * 4 + P.n_rchunks - 1) / P.n_rchunks;
    ncb = std::max<int64_t>(1, std::min<int64_t>(ncb, (n_cols + kMedWarps - 1) / kMedWarps));
    int64_t cbk = (n_cols + ncb - 1) / ncb;
    cbk = ((cbk + kMedWarps - 1) / kMedWarps) * kMedWarps;
    P.col_block = (int32_t)cbk;
    P.n_colblocks = (int32_t)((n_cols + cbk - 1) / cbk);
    P.n_items = (int64_t)P.n_rchunks * P.n_colblocks;
    if (P.n_items >= ((int64_t)1 << 31)) return fail(B200BOOT_EINVAL, "too many work items");
    const size_t smem = 128 + (size_t)tables * cnt_bytes + (size_t)slice_bytes;
    const int grid = (int)std::min<int64_t>(sms, P.n_items);
    k1m_short_kernel<<<grid, kMedThreads, smem, s>>>(P);
    B2_LAUNCH_CHECK();
    return 0;
  }

  int64_t cpi = (kSmemDataBytes - 128 - cnt_bytes) / perm_bytes;
  if (cpi < 1) return fail(B200BOOT_EUNSUPPORTED, "median: column does not fit in shared memory");
  cpi = std::min<int64_t>(std::min<int64_t>(cpi, 32), n_cols);
  P.cols_per_item = (int32_t)cpi;
  P.n_colchunks = (int32_t)((n_cols + cpi - 1) / cpi);
  int64_t want_r = ((int64_t)sms * 8 + P.n_colchunks - 1) / P.n_colchunks;
  int64_t rchunk = (n_local + want_r - 1) / want_r;
  rchunk = ((rchunk + batch - 1) / batch) * batch;
  P.rchunk = (int32_t)std::min<int64_t>(std::max<int64_t>(rchunk, batch), 1 << 20);
  P.n_rchunks = (int32_t)((n_local + P.rchunk - 1) / P.rchunk);
  P.n_items = (int64_t)P.n_colchunks * P.n_rchunks;
  if (P.n_items >= ((int64_t)1 << 31)) return fail(B200BOOT_EINVAL, "too many work items");
  const size_t smem = 128 + (size_t)cnt_bytes + (size_t)cpi * perm_bytes;
  const int grid = (int)std::min<int64_t>(sms, P.n_items);
  if (batch == 8)
    k1m_median_kernel<8><<<grid, kMedThreads, smem, s>>>(P);
  else if (batch == 4)
    k1m_median_kernel<4><<<grid, kMedThreads, smem, s>>>(P);
  else
    k1m_median_kernel<1><<<grid, kMedThreads, smem, s>>>(P);
  B2_LAUNCH_CHECK();
  return 0;
}

// Median of a long sorted 1-D array (n_rows beyond the resident limit).
inline int median_resample_sorted(const double* sorted_vals, int64_t n_rows, uint64_t seed,
                                  int64_t resample_lo, int64_t resample_hi, double* stat_out, void* ws,
                                  size_t ws_bytes, cudaStream_t s, const RankRule& rule) {
  if (!sorted_vals || !stat_out || n_rows < 1 || n_rows >= ((int64_t)1 << 31))
    return fail(B200BOOT_EINVAL, "median_sorted: bad arguments");
  const int64_t n_local = resample_hi - resample_lo;
  if (n_local < 0 || resample_lo < 0) return fail(B200BOOT_EINVAL, "bad resample range");
  if (n_local == 0) return 0;
  const Plan plan = make_plan(n_rows, 1);
  if (plan.n_tiles <= 1)
    return fail(B200BOOT_EINVAL, "median_sorted is for multi-tile plans; use resample_median_columns");
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  int32_t* counts = cv.take<int32_t>((size_t)plan.n_tiles * n_local);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  k0_tile_counts_kernel<<<(unsigned)((n_local + 127) / 128), 128, 0, s>>>(seed, resample_lo, n_local, plan,
                                                                        counts);
  B2_LAUNCH_CHECK();
  K1mtParams P{};
  P.sorted_vals = sorted_vals;
  P.plan = plan;
  P.seed = seed;
  P.resample_lo = resample_lo;
  P.n_local = n_local;
  P.counts = counts;
  P.out = stat_out;
  P.rule = rule;
  const size_t smem = (size_t)(plan.tile() + 2) * 2 + 16;
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(k1mt_median_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  }));
  const int grid = (int)std::min<int64_t>((int64_t)sm_count() * 6, n_local);
  k1mt_median_tiled_kernel<<<grid, kMtThreads, smem, s>>>(P);
  B2_LAUNCH_CHECK();
  return 0;
}

inline int median_jackknife_sorted(const double* sorted_vals, int64_t n_rows, double* out, cudaStream_t s,
                                   const RankRule& full, const RankRule& loo) {
  if (!sorted_vals || !out || n_rows < 1) return fail(B200BOOT_EINVAL, "bad arguments");
  k3m_jackknife_median_kernel<<<1, 256, 0, s>>>(sorted_vals, n_rows, full, loo, out);
  B2_LAUNCH_CHECK();
  return 0;
}

inline int median_jackknife(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, double* out,
                            void* ws, size_t ws_bytes, cudaStream_t s, const RankRule& full,
                            const RankRule& loo) {
  int rc = median_check(data, n_rows, n_cols, ld, ws);
  if (rc) return rc;
  if (!out) return fail(B200BOOT_EINVAL, "null output");
  Carver cv(ws, ws_bytes);
  MedianWs w = median_carve(cv, n_rows, n_cols);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  rc = median_prepare(data, n_rows, n_cols, ld, w, s);
  if (rc) return rc;
  for (int64_t d0 = 0; d0 < n_cols; d0 += 1 << 30) {
    const int64_t nd = std::min<int64_t>((int64_t)1 << 30, n_cols - d0);
    k3m_jackknife_median_kernel<<<(unsigned)nd, 256, 0, s>>>(w.sorted_vals + d0 * n_rows, n_rows, full, loo,
                                                             out + d0 * 8);
    B2_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace b200boot
