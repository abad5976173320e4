// method='abc' (bootstrap.py:507-567) for the weighted mean, np.average(x, weights=w).
//
// The reference evaluates 2N + 3 + A weighted means whose weight vectors differ from 1/N by
// O(epsilon / N), and takes first and second differences of them; at the default epsilon the
// second differences are dominated by the rounding of those means, so the reference's answer
// depends on numpy's arithmetic bit for bit.  These kernels therefore reproduce numpy's
// evaluation exactly: products and weights formed by single IEEE operations (no FMA
// contraction), sums in numpy's pairwise order (umath/loops_utils.h.src, pairwise_sum: below 8
// elements sequential, up to 128 eight strided accumulators combined as
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential tail, above that split at
// n/2 rounded down to a multiple of 8), np.average = sum(x * w) / sum(w).  The O(N^2) part —
// the 2N perturbed means — runs here; the O(N) arithmetic after it stays numpy on the host,
// the same expressions as the reference's.
#pragma once
#include "common.cuh"

namespace b200boot {

// numpy's pairwise sum of f(lo) .. f(lo + n - 1)
template <typename F>
__device__ double np_pairwise_sum(const F& f, int64_t lo, int64_t n) {
  if (n < 8) {
    double res = 0.0;
    for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, f(lo + i));
    return res;
  }
  if (n <= 128) {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(lo + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, f(lo + i));
    return res;
  }
  int64_t n2 = n / 2;
  n2 -= n2 % 8;
  const double a = np_pairwise_sum(f, lo, n2);
  const double b = np_pairwise_sum(f, lo + n2, n - n2);
  return __dadd_rn(a, b);
}

// np.average(x, weights=w) for a weight functor
template <typename W>
__device__ double np_average(const double* __restrict__ x, int64_t n, const W& w) {
  const double s = np_pairwise_sum([&](int64_t j) { return __dmul_rn(x[j], w(j)); }, 0, n);
  const double scl = np_pairwise_sum(w, 0, n);
  return __ddiv_rn(s, scl);
}

// tp[i] = average(x, p0 + ep * (e_i - p0)), tm[i] = average(x, p0 - ep * (e_i - p0)), p0 = 1/N
// (bootstrap.py:531-537): thread per i; all threads read the same x[j] together (broadcast).
__global__ void __launch_bounds__(128) abc_identity_averages_kernel(const double* __restrict__ x, int64_t n_rows,
                                                                    double ep, double* __restrict__ tp,
                                                                    double* __restrict__ tm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  const double p0 = __ddiv_rn(1.0, (double)n_rows);
  const double d_on = __dsub_rn(1.0, p0), d_off = __dsub_rn(0.0, p0);      // rows of I - p0
  const double w_on_p = __dadd_rn(p0, __dmul_rn(ep, d_on)), w_off_p = __dadd_rn(p0, __dmul_rn(ep, d_off));
  const double w_on_m = __dsub_rn(p0, __dmul_rn(ep, d_on)), w_off_m = __dsub_rn(p0, __dmul_rn(ep, d_off));
  tp[i] = np_average(x, n_rows, [&](int64_t j) { return j == i ? w_on_p : w_off_p; });
  tm[i] = np_average(x, n_rows, [&](int64_t j) { return j == i ? w_on_m : w_off_m; });
}

// out[v] = average(x, weights[v][:]) for explicit weight vectors (t0, the cq pair, the A endpoints)
__global__ void __launch_bounds__(32) abc_weighted_averages_kernel(const double* __restrict__ x, int64_t n_rows,
                                                                   const double* __restrict__ weights,
                                                                   int64_t n_vectors, double* __restrict__ out) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_vectors) return;
  if (weights == nullptr) {          // np.average(x) without weights = np.mean(x): pairwise sum / N
    out[v] = __ddiv_rn(np_pairwise_sum([&](int64_t j) { return x[j]; }, 0, n_rows), (double)n_rows);
    return;
  }
  const double* w = weights + v * n_rows;
  out[v] = np_average(x, n_rows, [&](int64_t j) { return w[j]; });
}

}  // namespace b200boot
