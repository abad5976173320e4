// K4 (count reduction), K5 (radix-descent order-statistic select) and the
// column sort used only for return_dist=True.
//
// Replace, in /root/reference/src/scikits/bootstrap/bootstrap.py:
//   :583       np.sum(stat < ostat, axis=0)              -> K4
//   :861-866   np.mean([compfunction(s) ...], axis=0)     -> K4
//   :445 + :485-490  stat.sort(axis=0); stat[nvals]       -> K5 (no sort)
//   :445 with return_dist                                  -> sort_columns
// Statistic vectors are column-major: stat[d * n + r].
#pragma once
#include "common.cuh"

namespace b200boot {

// Order-preserving map double -> uint64 (ascending; every NaN maps above +inf,
// matching numpy's "NaNs sort last").
__device__ __forceinline__ uint64_t key_of(double x) {
  uint64_t b = (uint64_t)__double_as_longlong(x);
  if (x != x) b = 0x7FF8000000000000ull;   // canonical +NaN
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(uint64_t k) {
  const uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
  return __longlong_as_double((long long)b);
}

// ------------------------------------------------------------------ K4 --------
__global__ void __launch_bounds__(256) k4_count_kernel(const double* __restrict__ stat, int64_t n,
                                                       int op, const double* __restrict__ t0,
                                                       const double* __restrict__ t1,
                                                       unsigned long long* __restrict__ counts) {
  const int64_t d = blockIdx.y;
  const double* col = stat + d * n;
  const double a = t0[d];
  const double b = t1 ? t1[d] : 0.0;
  unsigned int local = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const double s = col[r];
    bool hit;
    switch (op) {
      case B200BOOT_CMP_LT: hit = s < a; break;
      case B200BOOT_CMP_LE: hit = s <= a; break;
      case B200BOOT_CMP_GT: hit = s > a; break;
      case B200BOOT_CMP_GE: hit = s >= a; break;
      default: hit = (a <= s) && (s <= b); break;
    }
    local += hit ? 1u : 0u;
  }
  local = __reduce_add_sync(0xffffffffu, local);
  __shared__ unsigned int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(&s_cnt, local);
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) atomicAdd(counts + d, (unsigned long long)s_cnt);
}

// ------------------------------------------------------------------ K5 --------
// State per select (alpha a, column d), index a * n_cols + d:
//   prefix : key bits fixed so far (MSB first), krem : rank inside that bucket.
// Pass p (0..7) histograms byte (7 - p) of the keys whose higher bytes equal
// the prefix; pick() walks the 256 bins and descends.
constexpr int kSelMaxAlphas = 8;   // per launch; the host loops for more

struct SelectState {
  uint64_t* prefix;   // [A * D]
  int64_t* krem;      // [A * D]
  int64_t* hist;      // [A * D * 256]
};

inline size_t select_ws_bytes(int n_alphas, int64_t n_cols) {
  const size_t ad = (size_t)n_alphas * (size_t)n_cols;
  return pad256(ad * 8) + pad256(ad * 8) + pad256(ad * 256 * 8);
}
inline SelectState select_carve(Carver& cv, int n_alphas, int64_t n_cols) {
  const size_t ad = (size_t)n_alphas * (size_t)n_cols;
  SelectState s;
  s.prefix = cv.take<uint64_t>(ad);
  s.krem = cv.take<int64_t>(ad);
  s.hist = cv.take<int64_t>(ad * 256);
  return s;
}

__global__ void __launch_bounds__(256) k5_begin_kernel(const int64_t* __restrict__ ranks, int64_t ad,
                                                       SelectState st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ad) {
    st.prefix[i] = 0;
    st.krem[i] = ranks[i];
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = i; j < ad * 256; j += stride) st.hist[j] = 0;
}

__global__ void __launch_bounds__(256) k5_hist_kernel(const double* __restrict__ stat, int64_t n,
                                                      int64_t n_cols, int a_lo, int a_cnt, int pass,
                                                      SelectState st) {
  __shared__ unsigned int s_hist[kSelMaxAlphas][256];
  __shared__ uint64_t s_prefix[kSelMaxAlphas];
  const int64_t d = blockIdx.y;
  for (int i = threadIdx.x; i < a_cnt * 256; i += blockDim.x) (&s_hist[0][0])[i] = 0;
  if (threadIdx.x < a_cnt) s_prefix[threadIdx.x] = st.prefix[(int64_t)(a_lo + threadIdx.x) * n_cols + d];
  __syncthreads();
  const int shift = 56 - 8 * pass;
  const double* col = stat + d * n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride) {
    const uint64_t k = key_of(col[r]);
    const unsigned int digit = (unsigned int)(k >> shift) & 0xFFu;
    for (int a = 0; a < a_cnt; ++a) {
      const bool match = pass == 0 || ((k ^ s_prefix[a]) >> (shift + 8)) == 0;
      if (match) atomicAdd(&s_hist[a][digit], 1u);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < a_cnt * 256; i += blockDim.x) {
    const unsigned int v = (&s_hist[0][0])[i];
    if (v) {
      const int a = i >> 8, bin = i & 255;
      atomicAdd(reinterpret_cast<unsigned long long*>(
                    st.hist + ((int64_t)(a_lo + a) * n_cols + d) * 256 + bin),
                (unsigned long long)v);
    }
  }
}

__global__ void __launch_bounds__(256) k5_pick_kernel(int64_t ad, int pass, SelectState st) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ad) return;
  int64_t* h = st.hist + i * 256;
  int64_t k = st.krem[i];
  int digit = 255;
  int64_t below = 0;
  for (int b = 0; b < 256; ++b) {
    const int64_t c = h[b];
    if (k < below + c) {
      digit = b;
      break;
    }
    below += c;
  }
  for (int b = 0; b < 256; ++b) h[b] = 0;   // ready for the next pass
  st.krem[i] = k - below;
  st.prefix[i] |= (uint64_t)digit << (56 - 8 * pass);
}

__global__ void __launch_bounds__(256) k5_end_kernel(int64_t ad, SelectState st,
                                                     double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ad) out[i] = value_of(st.prefix[i]);
}

// Whole radix descent of one short column inside one CTA (one launch instead of
// begin + 8 x (hist + pick) + end): the small-problem path, where launch count is
// what the call costs.  Warp a walks the bins of select a; same results as the
// multi-launch path (both are exact).
constexpr int kSelSmallThreads = 1024;
constexpr int64_t kSelSmallMaxRows = 32768;
constexpr int64_t kSelSmallMaxCols = (int64_t)1 << 30;

// The eight passes over one column for up to kSelMaxAlphas selects; s_prefix = 0 and s_krem = the
// ranks on entry (set before a barrier or by the calling threads tid < a_cnt), s_prefix = the
// keys of the order statistics on exit (after the final barrier).  All threads of the CTA call it.
__device__ __forceinline__ void select_small_descend(const double* __restrict__ col, int64_t n, int a_cnt,
                                                     unsigned int (*s_hist)[256], uint64_t* s_prefix,
                                                     long long* s_krem) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nthreads = blockDim.x;                 // 256 .. 1024 by column length, >= 32 * a_cnt
  for (int pass = 0; pass < 8; ++pass) {
    for (int i = tid; i < a_cnt * 256; i += nthreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const int shift = 56 - 8 * pass;
    // Statistic values crowd into a few bins (pass 0 sees little more than sign and exponent), and
    // increments on one shared-memory address serialise.  Per warp and value: the lanes whose bin
    // equals that of the first counted lane are added by that lane in one atomic (two ballots and
    // a shuffle); only lanes with another bin issue their own.
    const int64_t n_round = ((n + nthreads - 1) / nthreads) * nthreads;      // uniform trip count for the ballots
    for (int64_t r = tid; r < n_round; r += nthreads) {
      const bool live = r < n;
      const uint64_t k = live ? key_of(col[r]) : 0ull;
      const unsigned int digit = (unsigned int)(k >> shift) & 0xFFu;
      for (int a = 0; a < a_cnt; ++a) {
        const bool hit = live && (pass == 0 || ((k ^ s_prefix[a]) >> (shift + 8)) == 0);
        const unsigned int any_hit = __ballot_sync(0xffffffffu, hit);
        if (any_hit == 0u) continue;                                           // warp-uniform
        const int src = __ffs(any_hit) - 1;
        const unsigned int d0 = __shfl_sync(0xffffffffu, digit, src);
        const unsigned int same = __ballot_sync(0xffffffffu, hit && digit == d0);
        if (lane == src)
          atomicAdd(&s_hist[a][d0], (unsigned int)__popc(same));
        else if (hit && digit != d0)
          atomicAdd(&s_hist[a][digit], 1u);
      }
    }
    __syncthreads();
    if (warp < a_cnt) {
      // lane l owns bins 8l .. 8l+7; first bin whose cumulative count exceeds krem
      const long long k = s_krem[warp];
      unsigned int mine = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) mine += s_hist[warp][lane * 8 + b];
      unsigned int inc = mine;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned int o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
      }
      const unsigned int hit = __ballot_sync(0xffffffffu, (long long)inc > k);
      const int owner = hit ? __ffs(hit) - 1 : 31;
      if (lane == owner) {
        long long below = (long long)inc;      // no bin found (rank beyond the column): as k5_pick
        int digit = 255;
        if (hit) {
          below = (long long)(inc - mine);
          for (int b = 0; b < 8; ++b) {
            const long long c = s_hist[warp][lane * 8 + b];
            if (k < below + c) {
              digit = lane * 8 + b;
              break;
            }
            below += c;
          }
        }
        s_krem[warp] = k - below;
        s_prefix[warp] |= (uint64_t)digit << shift;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSelSmallThreads) k5_select_small_kernel(
    const double* __restrict__ stat, int64_t n, int64_t n_cols, const int64_t* __restrict__ ranks,
    int a_lo, int a_cnt, double* __restrict__ out) {
  __shared__ unsigned int s_hist[kSelMaxAlphas][256];
  __shared__ uint64_t s_prefix[kSelMaxAlphas];
  __shared__ long long s_krem[kSelMaxAlphas];
  const int tid = threadIdx.x;
  const int64_t d = blockIdx.x;
  if (tid < a_cnt) {
    s_prefix[tid] = 0;
    s_krem[tid] = ranks[(int64_t)(a_lo + tid) * n_cols + d];
  }
  select_small_descend(stat + d * n, n, a_cnt, s_hist, s_prefix, s_krem);
  if (tid < a_cnt) out[(int64_t)(a_lo + tid) * n_cols + d] = value_of(s_prefix[tid]);
}

// ------------------------------------------------------------------ sort ------
// Bitonic network over keys padded to a power of two with the maximum key.
__global__ void __launch_bounds__(256) sort_load_kernel(const double* __restrict__ stat, int64_t n,
                                                        int64_t n_pad, uint64_t* __restrict__ keys) {
  const int64_t d = blockIdx.y;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride)
    keys[d * n_pad + i] = i < n ? key_of(stat[d * n + i]) : ~0ull;
}
__global__ void __launch_bounds__(256) sort_step_kernel(uint64_t* __restrict__ keys, int64_t n_pad,
                                                        int64_t j, int64_t k) {
  const int64_t d = blockIdx.y;
  uint64_t* col = keys + d * n_pad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += stride) {
    const int64_t partner = i ^ j;
    if (partner > i) {
      const uint64_t a = col[i], b = col[partner];
      const bool ascending = (i & k) == 0;
      if ((a > b) == ascending) {
        col[i] = b;
        col[partner] = a;
      }
    }
  }
}
// All steps with j < 1024 of one k-stage (and whole stages with k <= 1024) inside
// shared memory: one 2048-key segment per CTA.
__global__ void __launch_bounds__(1024) sort_local_kernel(uint64_t* __restrict__ keys, int64_t n_pad,
                                                          int64_t k_first, int64_t k_last,
                                                          int64_t j_first) {
  __shared__ uint64_t s[2048];
  const int64_t d = blockIdx.y;
  uint64_t* col = keys + d * n_pad + (int64_t)blockIdx.x * 2048;
  const int64_t seg0 = (int64_t)blockIdx.x * 2048;
  const int64_t seg_n = min((int64_t)2048, n_pad);
  for (int i = threadIdx.x; i < seg_n; i += 1024) s[i] = col[i];
  __syncthreads();
  for (int64_t k = k_first; k <= k_last; k <<= 1) {
    const int64_t j0 = (k == k_first) ? j_first : (k >> 1);
    for (int64_t j = j0; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < seg_n; i += 1024) {
        const int partner = i ^ (int)j;
        if (partner > i) {
          const uint64_t a = s[i], b = s[partner];
          const bool ascending = ((seg0 + i) & k) == 0;
          if ((a > b) == ascending) {
            s[i] = b;
            s[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < seg_n; i += 1024) col[i] = s[i];
}
__global__ void __launch_bounds__(256) sort_store_kernel(const uint64_t* __restrict__ keys, int64_t n,
                                                         int64_t n_pad, double* __restrict__ stat) {
  const int64_t d = blockIdx.y;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    stat[d * n + i] = value_of(keys[d * n_pad + i]);
}

}  // namespace b200boot
