// Small calls: when the data are one resident cell and the statistic vector is short, a whole
// ci() is launch- and round-trip-bound (cfg1: N = 1e3, n_samples = 1e4).  This file holds the
// single-CTA kernels that shorten such a call to three launches and one device-to-host copy:
//   center_small_kernel   means + centred rows of short arrays (var / std / Pearson / fit);
//   k3_small_kernel       the closed-form jackknife (bootstrap.py:595-615) in one CTA — also
//                         what b200boot_jackknife runs for short arrays, so the short and the
//                         general call agree bit for bit;
//   ci_tail_small_kernel  z0 count (bootstrap.py:583), BCa levels (:609-629), order-statistic
//                         ranks (:456, :480) and their exact radix select (:485-490) in one CTA.
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

// Closed-form jackknife of a short array (pair): the three passes of jack_passes in one CTA.
// out[8] as StoreAccelG writes it.
template <int MS>
__global__ void __launch_bounds__(kSmallThreads) k3_small_kernel(const double* __restrict__ d0,
                                                                 const double* __restrict__ d1, int64_t n_rows,
                                                                 int stat, JackState* st, double* __restrict__ out) {
  constexpr int NM = MomTraits<MS>::kMoments;
  __shared__ double s_part[kSmallThreads / 32][kMaxMoments];
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
  // jackknife mean, relative to the leave-one-out value of row 0
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
  double* res;
};

// The tail runs in ONE CTA, i.e. on one SM: whatever it does to 1e4 values costs about a
// microsecond per 1e5 thread-instructions.  Measured alternatives for the select of cfg1
// (n = 1e4, two ranks; whole tail kernel, B200): byte-wise radix descent with ballot-aggregated
// histogram atomics over the vector staged in shared memory — this version — 48 us; the same
// with __match_any_sync aggregation 176 us; a bitonic sort of the keys in shared memory 165 us
// (105 stages x 128 KB through a 128 B/clk pipe); a 64-step bit descent with register-resident
// keys 36-48 us; quickselect with register-resident keys 100 us.  The select is a third of the
// device time of a cfg1 call and a tenth of its wall time (the rest is host Python).
constexpr int64_t kSmallStageMax = 24576;     // 192 KB of float64

__global__ void __launch_bounds__(kSmallThreads) ci_tail_small_kernel(const SmallTailParams P) {
  extern __shared__ double s_stat[];
  __shared__ unsigned int s_hist[kSelMaxAlphas][256];
  __shared__ uint64_t s_prefix[kSelMaxAlphas];
  __shared__ long long s_krem[kSelMaxAlphas];
  __shared__ unsigned int s_less;
  const int tid = threadIdx.x;
  const int64_t n = P.n_samples;
  const double* stat = P.stat;
  if (P.staged) {
    for (int64_t r = tid; r < n; r += kSmallThreads) s_stat[r] = P.stat[r];
    stat = s_stat;
  }
  if (tid == 0) s_less = 0;
  __syncthreads();
  if (P.bca) {
    const double ostat = P.jack[0];
    unsigned int local = 0;
    for (int64_t r = tid; r < n; r += kSmallThreads) local += stat[r] < ostat ? 1u : 0u;   // strict '<'
    local = __reduce_add_sync(0xffffffffu, local);
    if ((tid & 31) == 0 && local) atomicAdd(&s_less, local);
  }
  __syncthreads();
  if (tid < P.n_alphas) {
    double aval = P.alphas[tid];
    if (P.bca) {
      const double z0 = dev_nppf((1.0 * (double)s_less) / (double)n);
      const double zs = z0 + dev_nppf(P.alphas[tid]);
      aval = dev_ncdf(z0 + zs / (1.0 - P.jack[4] * zs));
    }
    double v = rint((double)(n - 1) * aval);          // half to even, as np.round
    if (v != v) v = 0.0;                               // nan_to_num
    v = fmin(fmax(v, 0.0), (double)(n - 1));           // clip, as the host does before selecting
    s_krem[tid] = (long long)v;
    s_prefix[tid] = 0;
    P.res[8 + tid] = v;
  }
  if (tid == 0) P.res[16] = (double)s_less;
  if (tid < 8 && P.jack) P.res[17 + tid] = P.jack[tid];
  __syncthreads();
  select_small_descend(stat, n, P.n_alphas, s_hist, s_prefix, s_krem);
  if (tid < P.n_alphas) P.res[tid] = value_of(s_prefix[tid]);
}

}  // namespace b200boot
