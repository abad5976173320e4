// K3 — closed-form leave-one-out pass for the moment statistics.
//
// Replaces the O(N^2) Python loop at
// /root/reference/src/scikits/bootstrap/bootstrap.py:595-599 (and its numba
// O(N) special case :639-650) plus the moments at :609-615.  Each pass is a
// grid-stride reduction with a fixed grid and a fixed combine order, so the
// result is bit-reproducible.
#pragma once
#include "common.cuh"
#include "stats.cuh"

namespace b200boot {

constexpr int kRedBlocks = 592;   // 4 CTAs per SM on 148 SMs
constexpr int kRedThreads = 256;

// Scalars the passes hand to one another; lives in the workspace.
struct JackState {
  double center[2];          // per-array mean used for pre-centring
  double tot[kMaxMoments];   // contribution totals over all rows
  double ostat;              // statistic on the data itself
  double j0;                 // leave-one-out statistic of row 0 (shift of the jackknife mean)
  double jmean;              // mean of the leave-one-out statistics
  double d2, d3;             // sum (jmean - j_i)^2, ^3
  double accel;              // d3 / (6 d2^1.5)
};

// Generic pass: every row i yields NV values via f(i, v); block partials go to
// block_out[block][NV].
template <int NV, typename F>
__global__ void __launch_bounds__(kRedThreads) reduce_rows_kernel(int64_t n, F f,
                                                                  double* __restrict__ block_out) {
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double v[NV];
    f(i, v);
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] += v[k];
  }
  __shared__ double s_part[kRedThreads / 32][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double w = warp_sum(acc[k]);
    if (lane == 0) s_part[warp][k] = w;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
    for (int w = 0; w < kRedThreads / 32; ++w) t += s_part[w][threadIdx.x];
    block_out[(int64_t)blockIdx.x * NV + threadIdx.x] = t;
  }
}

// Single-block combine of the block partials, then g(total) on thread 0.
template <int NV, typename G>
__global__ void __launch_bounds__(kRedThreads) reduce_final_kernel(const double* __restrict__ block_out,
                                                                   int n_blocks, G g) {
  __shared__ double s[kRedThreads][NV];
  double acc[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) acc[k] = 0.0;
  for (int b = threadIdx.x; b < n_blocks; b += kRedThreads)
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] += block_out[(int64_t)b * NV + k];
#pragma unroll
  for (int k = 0; k < NV; ++k) s[threadIdx.x][k] = acc[k];
  __syncthreads();
  for (int off = kRedThreads / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off)
#pragma unroll
      for (int k = 0; k < NV; ++k) s[threadIdx.x][k] += s[threadIdx.x + off][k];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double tot[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) tot[k] = s[0][k];
    g(tot);
  }
}

// ---- pass functors -------------------------------------------------------------
struct RawSumsF {   // per-array plain sums (for the centring means)
  const double* d0;
  const double* d1;
  __device__ void operator()(int64_t i, double* v) const {
    v[0] = d0[i];
    v[1] = d1 ? d1[i] : 0.0;
  }
};
struct StoreCentersG {
  JackState* st;
  double n;
  __device__ void operator()(const double* tot) const {
    st->center[0] = tot[0] / n;
    st->center[1] = tot[1] / n;
  }
};

__global__ void __launch_bounds__(256) center_rows_kernel(const double* __restrict__ src,
                                                          double* __restrict__ dst, int64_t n,
                                                          const JackState* st, int which) {
  const double c = st->center[which];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = src[i] - c;
}

template <int MS>
struct ContributionF {
  const double* d0;
  const double* d1;
  __device__ void operator()(int64_t i, double* c) const {
    double v[2];
    v[0] = d0[i];
    v[1] = (MomTraits<MS>::kArrays == 2) ? d1[i] : 0.0;
    contribution<MS>(v, c);
  }
};
template <int MS>
struct StoreTotalsG {
  JackState* st;
  int stat;
  double n;
  const double* d0;
  const double* d1;
  __device__ void operator()(const double* tot) const {
    for (int k = 0; k < MomTraits<MS>::kMoments; ++k) st->tot[k] = tot[k];
    st->ostat = finish_stat(stat, tot, n, st->center);
    // leave-one-out value of row 0: the jackknife mean is accumulated relative to it,
    // so identical leave-one-out values give d == 0 exactly (acceleration undefined,
    // as the reference intends: bootstrap.py:616-625) instead of rounding noise
    double v[2], c[kMaxMoments], m[kMaxMoments];
    v[0] = d0[0];
    v[1] = (MomTraits<MS>::kArrays == 2) ? d1[0] : 0.0;
    contribution<MS>(v, c);
    for (int k = 0; k < MomTraits<MS>::kMoments; ++k) m[k] = tot[k] - c[k];
    st->j0 = finish_stat(stat, m, n - 1.0, st->center);
  }
};

template <int MS>
__device__ __forceinline__ double leave_one_out(const JackState* st, int stat, double n,
                                                const double* d0, const double* d1, int64_t i) {
  double v[2], c[kMaxMoments], m[kMaxMoments];
  v[0] = d0[i];
  v[1] = (MomTraits<MS>::kArrays == 2) ? d1[i] : 0.0;
  contribution<MS>(v, c);
#pragma unroll
  for (int k = 0; k < MomTraits<MS>::kMoments; ++k) m[k] = st->tot[k] - c[k];
  return finish_stat(stat, m, n - 1.0, st->center);
}

template <int MS>
struct JackSumF {
  const JackState* st;
  int stat;
  double n;
  const double* d0;
  const double* d1;
  __device__ void operator()(int64_t i, double* v) const {
    v[0] = leave_one_out<MS>(st, stat, n, d0, d1, i) - st->j0;
  }
};
struct StoreJmeanG {
  JackState* st;
  double n;
  __device__ void operator()(const double* tot) const { st->jmean = st->j0 + tot[0] / n; }
};

template <int MS>
struct JackDevF {
  const JackState* st;
  int stat;
  double n;
  const double* d0;
  const double* d1;
  __device__ void operator()(int64_t i, double* v) const {
    const double d = st->jmean - leave_one_out<MS>(st, stat, n, d0, d1, i);
    v[0] = d * d;
    v[1] = d * d * d;
  }
};
struct StoreAccelG {
  JackState* st;
  double* out;
  __device__ void operator()(const double* tot) const {
    st->d2 = tot[0];
    st->d3 = tot[1];
    st->accel = tot[1] / (6.0 * pow(tot[0], 1.5));
    out[0] = st->ostat;
    out[1] = st->jmean;
    out[2] = st->d2;
    out[3] = st->d3;
    out[4] = st->accel;
    out[5] = st->center[0];
    out[6] = st->center[1];
    out[7] = 0.0;
  }
};

}  // namespace b200boot
