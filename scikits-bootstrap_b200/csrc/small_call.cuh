// Small calls: when the data are one resident cell and the statistic vector is short, a whole
// ci() is launch- and round-trip-bound (cfg1: N = 1e3, n_samples = 1e4).  This file holds
//   center_small_kernel    means + centred rows of short arrays (var / std / Pearson / fit);
//   k3_small_kernel        the closed-form jackknife (bootstrap.py:595-615) in one CTA — what
//                          b200boot_jackknife runs for short arrays;
//   ci_small_fused_kernel  a whole short ci() in ONE launch: resampling (K1's single-cell path),
//                          the jackknife, the z0 count (bootstrap.py:583), the BCa levels
//                          (:609-629), the order-statistic ranks (:456, :480) and their exact
//                          select (:485-490).
// The tail's levels are SPECULATIVE: nppf / ncdf are evaluated on the device with the
// reference's formulas, but the host recomputes the ranks with its own (verbatim) math from
// the returned count and jackknife summary and only accepts the device's order statistics
// when the ranks agree — otherwise it selects again with its own ranks.  The result is
// therefore always the host tail's, as in the general path.
#pragma once
#include "common.cuh"
#include "k3_jackknife.cuh"
#include "k45_interval.cuh"
#include "stats.cuh"

namespace b200boot {

constexpr int kSmallThreads = 1024;
constexpr int64_t kSmallMaxRows = 32768;       // single-CTA jackknife / centring up to here
constexpr int kSmallResultDoubles = 32;        // [0..7] order statistics, [8..15] ranks, [16] less, [17..24] jackknife summary

// Sum over the CTA's 1024 threads in a fixed order; every thread gets the total.
template <int NV>
__device__ __forceinline__ void block_sum_small(double* acc, double (*s_part)[NV]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                      // s_part may still be read from a previous call
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double w = warp_sum(acc[k]);
    if (lane == 0) s_part[warp][k] = w;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0.0;
    for (int w = 0; w < kSmallThreads / 32; ++w) t += s_part[w][k];
    acc[k] = t;
  }
}

// Means of up to two short arrays and the rows centred on them.
__global__ void __launch_bounds__(kSmallThreads) center_small_kernel(const double* __restrict__ d0,
                                                                     const double* __restrict__ d1, int64_t n,
                                                                     JackState* st, double* __restrict__ c0,
                                                                     double* __restrict__ c1) {
  __shared__ double s_part[kSmallThreads / 32][2];
  double acc[2] = {0.0, 0.0};
  for (int64_t i = threadIdx.x; i < n; i += kSmallThreads) {
    acc[0] += d0[i];
    if (d1) acc[1] += d1[i];
  }
  block_sum_small<2>(acc, s_part);
  const double m0 = acc[0] / (double)n, m1 = acc[1] / (double)n;
  if (threadIdx.x == 0) {
    st->center[0] = m0;
    st->center[1] = m1;
  }
  for (int64_t i = threadIdx.x; i < n; i += kSmallThreads) {
    c0[i] = d0[i] - m0;
    if (d1) c1[i] = d1[i] - m1;
  }
}

// ---- the reference's normal cdf / inverse cdf on the device (speculative ranks only) ---------
// Same operations in the same order as _interval.py (bootstrap.py:75-119).
__device__ __forceinline__ double horner6(const double* c, double t) {
  double r = c[0];
#pragma unroll
  for (int i = 1; i < 6; ++i) r = __dadd_rn(__dmul_rn(r, t), c[i]);
  return r;
}
__device__ __forceinline__ double acklam_tail(double t) {
  const double a[6] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e00,
                       -2.549732539343734e00, 4.374664141464968e00, 2.938163982698783e00};
  const double b[4] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e00,
                       3.754408661907416e00};
  double den = b[0];
#pragma unroll
  for (int i = 1; i < 4; ++i) den = __dadd_rn(__dmul_rn(den, t), b[i]);
  return horner6(a, t) / __dadd_rn(__dmul_rn(den, t), 1.0);
}
__device__ inline double dev_nppf(double p) {
  if (p != p) return p;
  if (p <= 0.0) return p == 0.0 ? -INFINITY : NAN;
  if (p >= 1.0) return p == 1.0 ? INFINITY : NAN;
  if (p < 0.02425) return acklam_tail(sqrt(-2.0 * log(p)));
  if (p <= 0.97575) {
    const double a[6] = {-3.969683028665376e01, 2.209460984245205e02, -2.759285104469687e02,
                         1.383577518672690e02, -3.066479806614716e01, 2.506628277459239e00};
    const double b[5] = {-5.447609879822406e01, 1.615858368580409e02, -1.556989798598866e02,
                         6.680131188771972e01, -1.328068155288572e01};
    const double q = p - 0.5;
    const double s = __dmul_rn(q, q);
    double den = b[0];
#pragma unroll
    for (int i = 1; i < 5; ++i) den = __dadd_rn(__dmul_rn(den, s), b[i]);
    return __dmul_rn(horner6(a, s), q) / __dadd_rn(__dmul_rn(den, s), 1.0);
  }
  return -acklam_tail(sqrt(-2.0 * log(1.0 - p)));
}
__device__ inline double dev_ncdf(double x) { return 0.5 * (1.0 + erf(x / 1.4142135623730951)); }

// Tail of a short scalar-statistic call in one CTA.  `jack` = k3 output (ostat at [0], accel
// at [4]) or null for method 'pi' without it.  res: kSmallResultDoubles doubles.
struct SmallTailParams {
  const double* stat;     // [n_samples]
  int64_t n_samples;
  const double* jack;     // [8] or null
  double alphas[kSelMaxAlphas];
  int n_alphas;
  int bca;                // 1: BCa levels, 0: the alphas themselves
  int staged;             // the statistic vector is copied to shared memory first
  int n_bins;             // buckets of the select: kBktBins (staged) or kBktBinsWide
  const int64_t* ranks;   // select-only mode: the ranks are given (stride rank_stride), no count / levels
  int64_t rank_stride;
  int64_t res_stride;     // res[a * res_stride] (select-only mode; the tail writes its block densely)
  int cmp_op;             // pval mode (n_alphas == 0): res[16] = #{stat op t0 [, t1]}, a b200boot_cmp; else -1
  double cmp_t0, cmp_t1;
  double* res;
};

// ---- the fused small call: ONE launch for K1 + jackknife + tail -----------------------------
// A cfg1-sized ci() is bound by launch latencies and by whatever runs on a single SM.  The fused
// kernel removes the launch boundaries and shortens the single-SM part:
//   * every CTA loads the (short) rows into shared memory and, for centred statistics, centres
//     them itself — the same thread-strided sums in the same order in every CTA, so all CTAs hold
//     the bits center_small_kernel would have written;
//   * the last CTA of the grid runs the closed-form jackknife (k3_small's passes) while the others
//     resample (one warp per resample over the single-cell Lemire stream, exactly K1's
//     lemire_accumulate + finish_stat: bit-equal to the general path);
//   * the CTA that finishes last (ticket counter) runs the tail: z0 count, speculative BCa levels
//     and ranks as before, and a BUCKET select instead of the 8-pass radix descent: min / max of
//     the statistic vector, a monotone linear map into 4096 buckets, one histogram pass, one scan,
//     then the few values sharing the bucket of each rank are ranked exactly by their keys.  The
//     map is monotone non-decreasing in the value, so the k-th smallest value lies in the bucket
//     whose cumulative count first exceeds k: the result is sort(stat)[k] bit for bit.  Degenerate
//     vectors (all equal, non-finite range, NaN ranks, a crowded bucket) take the radix descent.
//   * the 32-double result block is also stored straight into the caller's pinned host mirror
//     when that memory is device-accessible (no copy operation behind the kernel).
// Measured on B200 (cfg1, BCa): three launches + copy (K1 23 us, jackknife 10.5 us, single-CTA
// tail 54 us, of which the ballot-aggregated 8-pass radix descent was ~45 us — itself the best
// of five selects tried: __match_any aggregation 176 us, shared-memory bitonic sort 165 us,
// 64-step bit descent 36-48 us, quickselect 100 us) = 98 us of device time per call; this
// kernel: 42 us (launch + row load ~24 us with an empty statistic vector, ~5 us per round of
// 4704 resamples, the rest the tail).
constexpr int kBktBins = 4096;                  // buckets when the vector is staged next to them
constexpr int kBktBinsWide = 32768;             // long (unstaged) vectors: 128 KB of buckets
constexpr int64_t kTailMaxValues = (int64_t)1 << 21;   // single-CTA tail up to here
constexpr int kBktCand = 256;                   // candidates kept per rank; more -> radix descent
constexpr int64_t kFusedStageMax = 16384;       // statistic vectors up to here are staged (128 KB)
constexpr int64_t kFusedMaxCellDoubles = 28672; // rows x arrays resident (224 KB), K1's single-cell limit

struct SmallFusedParams {
  const double* data[2];
  int64_t n_rows;
  int64_t smem_stride;      // doubles between the two arrays in shared memory
  int stat;
  int centre;               // rows are centred on their means first
  uint64_t seed;
  RoundKeys rk;
  double* stat_out;         // [n_samples]
  JackState* jack_state;
  double* jack_out;         // [8] device, or null: no jackknife
  int32_t* done;            // ticket counter: zero on entry, reset on exit
  double* res_host;         // device-accessible pinned mirror of tail.res, or null
  SmallTailParams tail;
};

// dynamic shared memory of the tail, behind the staged vector
struct TailScratch {
  unsigned int* hist;                       // [kBktBins]
  uint64_t* cand;                           // [kSelMaxAlphas][kBktCand]
  unsigned int (*hist2)[256];               // [kSelMaxAlphas][256]: radix descent
};
inline __host__ __device__ size_t tail_stage_bytes(int64_t n) { return ((size_t)n * 8 + 15) & ~(size_t)15; }
inline __host__ __device__ size_t tail_scratch_bytes(int n_bins = kBktBins) {
  return (size_t)n_bins * 4 + (size_t)kSelMaxAlphas * kBktCand * 8 + (size_t)kSelMaxAlphas * 256 * 4;
}

__device__ __forceinline__ int bucket_of(double v, double lo, double scale, int n_bins) {
  const int b = (int)((v - lo) * scale);    // v >= lo: non-negative, monotone in v
  return b < n_bins - 1 ? b : n_bins - 1;
}

// f(r, v) for every element of the vector, thread-strided, four loads in flight per thread (an
// unstaged vector comes from L2).  The single-CTA tail is instruction-bound, not latency-bound:
// 1e5 values cost ~97 us with or without the batching (~47 thread-instructions per value visit,
// three visits; ncu source view).
template <typename F>
__device__ __forceinline__ void for_each_value(const double* __restrict__ src, bool from_global, int64_t n, F&& f) {
  constexpr int kBatch = 4;
  const int64_t step = (int64_t)kSmallThreads * kBatch;
  for (int64_t r0 = threadIdx.x; r0 < n; r0 += step) {
    double v[kBatch];
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int64_t r = r0 + (int64_t)k * kSmallThreads;
      v[k] = r < n ? (from_global ? __ldcg(src + r) : src[r]) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kBatch; ++k) {
      const int64_t r = r0 + (int64_t)k * kSmallThreads;
      if (r < n) f(r, v[k]);
    }
  }
}

// All 1024 threads of one CTA.  `smem`: (staged ? n doubles : 0) + tail_scratch_bytes().
__device__ __noinline__ void small_tail_bucket(const SmallTailParams& P, unsigned char* smem, double* res_host) {
  __shared__ double s_red[kSmallThreads / 32][2];
  __shared__ unsigned int s_cnt_red[kSmallThreads / 32][2];
  __shared__ unsigned int s_scan[kSmallThreads / 32];
  __shared__ uint64_t s_prefix[kSelMaxAlphas];
  __shared__ long long s_krem[kSelMaxAlphas], s_rank[kSelMaxAlphas];
  __shared__ int s_bucket[kSelMaxAlphas];
  __shared__ unsigned int s_ccnt[kSelMaxAlphas];
  __shared__ double s_val[kSelMaxAlphas];
  __shared__ int s_slow;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the parameter block lives in local memory (the function is not inlined): what the per-value
  // loops need is copied to registers once, or every value would reload it
  const int64_t n = P.n_samples;
  const int A = P.n_alphas;
  const bool staged = P.staged != 0, bca = P.bca != 0;
  const int cmp_op = P.cmp_op;
  const double cmp_t0 = P.cmp_t0, cmp_t1 = P.cmp_t1;
  const double* const gstat = P.stat;
  double* s_stat = reinterpret_cast<double*>(smem);
  unsigned char* scratch = smem + (staged ? tail_stage_bytes(n) : 0);
  TailScratch T;
  T.hist = reinterpret_cast<unsigned int*>(scratch);
  const int n_bins = P.n_bins;                 // a multiple of 4 * kSmallThreads
  T.cand = reinterpret_cast<uint64_t*>(scratch + (size_t)n_bins * 4);
  T.hist2 = reinterpret_cast<unsigned int(*)[256]>(scratch + (size_t)n_bins * 4 + (size_t)kSelMaxAlphas * kBktCand * 8);

  // pass A: stage, min / max of the non-NaN values, NaN count, #{stat < ostat}
  const double ostat = bca ? __ldcg(P.jack) : 0.0;
  double lo = INFINITY, hi = -INFINITY;
  unsigned int less = 0, nans = 0;
  for_each_value(gstat, true, n, [&](int64_t r, double v) {
    if (staged) s_stat[r] = v;
    if (v != v) {
      ++nans;
    } else {
      lo = fmin(lo, v);
      hi = fmax(hi, v);
    }
    less += (bca && v < ostat) ? 1u : 0u;                          // strict '<'  bootstrap.py:583
    if (cmp_op >= 0) {                                               // pval criterion  bootstrap.py:861-866
      bool hit;
      switch (cmp_op) {
        case B200BOOT_CMP_LT: hit = v < cmp_t0; break;
        case B200BOOT_CMP_LE: hit = v <= cmp_t0; break;
        case B200BOOT_CMP_GT: hit = v > cmp_t0; break;
        case B200BOOT_CMP_GE: hit = v >= cmp_t0; break;
        default: hit = (cmp_t0 <= v) && (v <= cmp_t1); break;
      }
      less += hit ? 1u : 0u;
    }
  });
  for (int i = tid; i < n_bins; i += kSmallThreads) T.hist[i] = 0u;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  less = __reduce_add_sync(0xffffffffu, less);
  nans = __reduce_add_sync(0xffffffffu, nans);
  if (lane == 0) {
    s_red[warp][0] = lo;
    s_red[warp][1] = hi;
    s_cnt_red[warp][0] = less;
    s_cnt_red[warp][1] = nans;
  }
  if (tid == 0) s_slow = 0;
  if (tid < kSelMaxAlphas) s_ccnt[tid] = 0u;
  __syncthreads();
  lo = INFINITY, hi = -INFINITY, less = 0, nans = 0;
  for (int w = 0; w < kSmallThreads / 32; ++w) {
    lo = fmin(lo, s_red[w][0]);
    hi = fmax(hi, s_red[w][1]);
    less += s_cnt_red[w][0];
    nans += s_cnt_red[w][1];
  }
  const double* stat = staged ? s_stat : gstat;
  const double* jack = P.jack;
  if (A == 0) {                                      // pval mode: the count is the result
    if (tid == 0) {
      P.res[16] = (double)less;
      if (res_host) {
        res_host[16] = (double)less;
        __threadfence_system();
      }
    }
    return;
  }

  if (P.ranks != nullptr) {
    if (tid < A) {
      const long long r = (long long)__ldg(P.ranks + tid * P.rank_stride);
      s_rank[tid] = r < 0 ? 0 : (r > (long long)(n - 1) ? (long long)(n - 1) : r);
    }
  } else if (tid < A) {
    // speculative levels and ranks (the host verifies them)
    double aval = P.alphas[tid];
    if (P.bca) {
      const double z0 = dev_nppf((1.0 * (double)less) / (double)n);
      const double zs = z0 + dev_nppf(P.alphas[tid]);
      aval = dev_ncdf(z0 + zs / (1.0 - __ldcg(jack + 4) * zs));
    }
    double v = rint((double)(n - 1) * aval);          // half to even, as np.round
    if (v != v) v = 0.0;                               // nan_to_num
    v = fmin(fmax(v, 0.0), (double)(n - 1));           // clip, as the host does before selecting
    s_rank[tid] = (long long)v;
    P.res[8 + tid] = v;
    if (res_host) res_host[8 + tid] = v;
  }
  if (tid == 0 && P.ranks == nullptr) {
    P.res[16] = (double)less;
    if (res_host) res_host[16] = (double)less;
  }
  if (tid < 8 && jack && P.ranks == nullptr) {
    const double j = __ldcg(jack + tid);
    P.res[17 + tid] = j;
    if (res_host) res_host[17 + tid] = j;
  }

  // bucket select
  const double range = hi - lo;
  const double scale = (double)n_bins / range;
  const bool fast = hi > lo && isfinite(range) && isfinite(scale) && isfinite(lo);
  if (fast) {
    for_each_value(stat, !staged, n, [&](int64_t, double v) {
      if (v == v) atomicAdd(&T.hist[bucket_of(v, lo, scale, n_bins)], 1u);
    });
    __syncthreads();
    // exclusive scan over the buckets: thread t owns the `per` consecutive buckets from t * per
    const int per = n_bins / kSmallThreads;
    const uint4* mine4 = reinterpret_cast<const uint4*>(T.hist + (size_t)tid * per);
    unsigned int mine = 0;
    for (int g = 0; g < per / 4; ++g) {
      const uint4 h4 = mine4[g];
      mine += h4.x + h4.y + h4.z + h4.w;
    }
    unsigned int inc = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int o = __shfl_up_sync(0xffffffffu, inc, off);
      if (lane >= off) inc += o;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    unsigned int base = 0;
    for (int w = 0; w < warp; ++w) base += s_scan[w];
    const long long below0 = (long long)base + (long long)(inc - mine);
    for (int a = 0; a < A; ++a) {
      const long long k = s_rank[a];
      if (k >= n - (long long)nans) {                  // a NaN is asked for: radix descent
        if (tid == 0) s_slow = 1;
      } else if (k >= below0 && k < below0 + (long long)mine) {
        long long below = below0;
        for (int b = 0; b < per; ++b) {
          const long long c = T.hist[(size_t)tid * per + b];
          if (k < below + c) {
            s_bucket[a] = tid * per + b;
            s_krem[a] = k - below;
            break;
          }
          below += c;
        }
      }
    }
    __syncthreads();
    if (!s_slow) {
      // the values of each rank's bucket, as keys
      int my_bucket[kSelMaxAlphas];
#pragma unroll
      for (int a = 0; a < kSelMaxAlphas; ++a) my_bucket[a] = a < A ? s_bucket[a] : -1;
      for_each_value(stat, !staged, n, [&](int64_t, double v) {
        if (v != v) return;
        const int b = bucket_of(v, lo, scale, n_bins);
#pragma unroll
        for (int a = 0; a < kSelMaxAlphas; ++a)
          if (b == my_bucket[a]) {
            const unsigned int pos = atomicAdd(&s_ccnt[a], 1u);
            if (pos < (unsigned int)kBktCand) T.cand[a * kBktCand + pos] = key_of(v);
          }
      });
      __syncthreads();
      for (int a = 0; a < A; ++a)
        if (s_ccnt[a] > (unsigned int)kBktCand && tid == 0) s_slow = 1;
      __syncthreads();
      if (!s_slow) {
        // exact rank inside the bucket: four ranks at a time, 256 threads each
        for (int a0 = 0; a0 < A; a0 += kSmallThreads / kBktCand) {
          const int a = a0 + tid / kBktCand, t = tid % kBktCand;
          if (a < A && t < (int)s_ccnt[a]) {
            const uint64_t* c = T.cand + a * kBktCand;
            const uint64_t key = c[t];
            const int cnt = (int)s_ccnt[a];
            int before = 0;
            for (int j = 0; j < cnt; ++j) before += (c[j] < key || (c[j] == key && j < t)) ? 1 : 0;
            if (before == (int)s_krem[a]) s_val[a] = value_of(key);
          }
        }
        __syncthreads();
      }
    }
  }
  if (!fast || s_slow) {
    __syncthreads();
    if (tid < A) {
      s_prefix[tid] = 0;
      s_krem[tid] = s_rank[tid];
    }
    __syncthreads();
    select_small_descend(stat, n, A, T.hist2, s_prefix, s_krem);
    if (tid < A) s_val[tid] = value_of(s_prefix[tid]);
    __syncthreads();
  }
  if (tid < A) {
    P.res[P.ranks != nullptr ? tid * P.res_stride : (int64_t)tid] = s_val[tid];
    if (res_host) res_host[tid] = s_val[tid];
  }
  if (res_host) __threadfence_system();
}

// The tail as its own launch, for statistic vectors of any origin (the general path's K1, a
// gathered sharded vector) up to kTailMaxValues values: one CTA, one result block, one host wait
// instead of count -> host -> 8 x (hist + pick) -> host.
// nan_probe (may be null): the largest value of a sorted column; res[25] = 1 when it is NaN, i.e. the
// column holds NaNs and numpy's rule for rank statistics applies (the caller then takes the general path).
__global__ void __launch_bounds__(kSmallThreads, 1) ci_tail_kernel(const SmallTailParams P, double* res_host,
                                                                 const double* nan_probe) {
  extern __shared__ __align__(128) unsigned char tail_smem[];
  if (threadIdx.x == 0) {
    const double flag = (nan_probe != nullptr && __ldcg(nan_probe) != __ldcg(nan_probe)) ? 1.0 : 0.0;
    P.res[25] = flag;
    if (res_host) res_host[25] = flag;
  }
  small_tail_bucket(P, tail_smem, res_host);
}

// K5 for tables of short columns (cfg4: 10^4 columns of 10^4 values): one CTA per column runs
// the bucket select with the caller's ranks — three passes over the column instead of the eight
// of the radix descent.  out[(a_lo + a) * n_cols + d].
__global__ void __launch_bounds__(kSmallThreads, 1) k5_select_bucket_kernel(const double* __restrict__ stat, int64_t n,
                                                                          int64_t n_cols,
                                                                          const int64_t* __restrict__ ranks, int a_lo,
                                                                          int a_cnt, int staged,
                                                                          double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char sel_smem[];
  for (int64_t d = blockIdx.x; d < n_cols; d += gridDim.x) {
    SmallTailParams P;
    P.stat = stat + d * n;
    P.n_samples = n;
    P.jack = nullptr;
    P.n_alphas = a_cnt;
    P.bca = 0;
    P.staged = staged;
    P.n_bins = kBktBins;
    P.ranks = ranks + (int64_t)a_lo * n_cols + d;
    P.rank_stride = n_cols;
    P.res = out + (int64_t)a_lo * n_cols + d;
    P.res_stride = n_cols;
    P.cmp_op = -1;
    small_tail_bucket(P, sel_smem, nullptr);
    __syncthreads();
  }
}

// Closed-form jackknife of the resident rows by the whole CTA (the passes of k3_small_kernel).
template <int MS>
__device__ __noinline__ void k3_small_body(const double* d0, const double* d1, int64_t n_rows, int stat, JackState* st,
                              double* out, double (*s_part)[kMaxMoments]) {
  constexpr int NM = MomTraits<MS>::kMoments;
  const double n = (double)n_rows;
  double acc[kMaxMoments];
#pragma unroll
  for (int k = 0; k < kMaxMoments; ++k) acc[k] = 0.0;
  const ContributionF<MS> contrib{d0, d1};
  for (int64_t i = threadIdx.x; i < n_rows; i += kSmallThreads) {
    double c[kMaxMoments];
    contrib(i, c);
#pragma unroll
    for (int k = 0; k < NM; ++k) acc[k] += c[k];
  }
  block_sum_small<kMaxMoments>(acc, s_part);
  if (threadIdx.x == 0) StoreTotalsG<MS>{st, stat, n, d0, d1}(acc);
  __syncthreads();
  __threadfence_block();
  double s1[kMaxMoments] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t i = threadIdx.x; i < n_rows; i += kSmallThreads)
    s1[0] += leave_one_out<MS>(st, stat, n, d0, d1, i) - st->j0;
  block_sum_small<kMaxMoments>(s1, s_part);
  if (threadIdx.x == 0) StoreJmeanG{st, n}(s1);
  __syncthreads();
  __threadfence_block();
  double s2[kMaxMoments] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (int64_t i = threadIdx.x; i < n_rows; i += kSmallThreads) {
    const double d = st->jmean - leave_one_out<MS>(st, stat, n, d0, d1, i);
    s2[0] += d * d;
    s2[1] += d * d * d;
  }
  block_sum_small<kMaxMoments>(s2, s_part);
  if (threadIdx.x == 0) StoreAccelG{st, out}(s2);
}

// Closed-form jackknife of a short array (pair) as its own launch; out[8] as StoreAccelG writes it.
template <int MS>
__global__ void __launch_bounds__(kSmallThreads) k3_small_kernel(const double* __restrict__ d0,
                                                                 const double* __restrict__ d1, int64_t n_rows,
                                                                 int stat, JackState* st, double* __restrict__ out) {
  __shared__ double s_part[kSmallThreads / 32][kMaxMoments];
  k3_small_body<MS>(d0, d1, n_rows, stat, st, out, s_part);
}

template <int MS>
__global__ void __launch_bounds__(kSmallThreads, 1) ci_small_fused_kernel(const SmallFusedParams P) {
  constexpr int NA = MomTraits<MS>::kArrays;
  constexpr int NM = MomTraits<MS>::kMoments;
  extern __shared__ __align__(128) unsigned char sf_smem[];
  __shared__ double s_part[kSmallThreads / 32][kMaxMoments];
  __shared__ double s_centre[2];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n_rows = P.n_rows;
  double* s_data = reinterpret_cast<double*>(sf_smem);
  double* s_d1 = s_data + P.smem_stride;

  // rows -> shared memory (centred on their means where the statistic asks for it, with the
  // arithmetic of center_small_kernel)
  if (P.centre) {
    double acc[kMaxMoments] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int64_t i = tid; i < n_rows; i += kSmallThreads) {
      acc[0] += P.data[0][i];
      if (NA == 2) acc[1] += P.data[1][i];
    }
    block_sum_small<kMaxMoments>(acc, s_part);
    const double m0 = acc[0] / (double)n_rows, m1 = acc[1] / (double)n_rows;
    if (tid == 0) {
      s_centre[0] = m0;
      s_centre[1] = m1;
    }
    for (int64_t i = tid; i < n_rows; i += kSmallThreads) {
      s_data[i] = P.data[0][i] - m0;
      if (NA == 2) s_d1[i] = P.data[1][i] - m1;
    }
  } else {
    if (tid == 0) s_centre[0] = s_centre[1] = 0.0;
    for (int64_t i = tid; i < n_rows; i += kSmallThreads) {
      s_data[i] = P.data[0][i];
      if (NA == 2) s_d1[i] = P.data[1][i];
    }
  }
  __syncthreads();

  const bool with_jack = P.jack_out != nullptr;
  const bool jack_cta = with_jack && gridDim.x > 1 && blockIdx.x == gridDim.x - 1;
  const int n_res_ctas = (with_jack && gridDim.x > 1) ? gridDim.x - 1 : gridDim.x;
  if (!jack_cta) {
    const unsigned char* s_bytes = sf_smem;
    const uint32_t stride_bytes = (uint32_t)(P.smem_stride * 8);
    const int64_t step = (int64_t)n_res_ctas * (kSmallThreads / 32);
    for (int64_t rl = (int64_t)blockIdx.x * (kSmallThreads / 32) + warp; rl < P.tail.n_samples; rl += step) {
      double tot[NM];
      const Stream st = make_stream(P.seed, rl, 0u, kPurposeIndex);
      lemire_accumulate<MS>(s_bytes, stride_bytes, P.rk, st, n_rows, (uint32_t)n_rows, lane, tot);
      if (lane == 0) {
        double m[kMaxMoments];
#pragma unroll
        for (int k = 0; k < NM; ++k) m[k] = 0.0 + tot[k];      // as the finisher's sum over one cell
        P.stat_out[rl] = finish_stat(P.stat, m, (double)n_rows, s_centre);
      }
    }
  }
  if (with_jack && (jack_cta || gridDim.x == 1)) {
    if (tid == 0) {
      P.jack_state->center[0] = s_centre[0];
      P.jack_state->center[1] = s_centre[1];
    }
    __syncthreads();
    __threadfence_block();
    k3_small_body<MS>(s_data, NA == 2 ? s_d1 : nullptr, n_rows, P.stat, P.jack_state, P.jack_out, s_part);
  }

  // the CTA that finishes last runs the tail
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(P.done, 1) == (int)gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  small_tail_bucket(P.tail, sf_smem, P.res_host);
  if (tid == 0) *P.done = 0;          // the ticket counter is left ready for the next call
}

}  // namespace b200boot
