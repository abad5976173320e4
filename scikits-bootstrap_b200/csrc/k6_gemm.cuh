// K6 — multi-column moment statistics as (multinomial count matrix) x (data) GEMM.
//
// For (N, D) data every column shares the index set of a resample
// (/root/reference/src/scikits/bootstrap/bootstrap.py:431), so the column sums of resample r
// are  sum_i c[r][i] * x[i][:]  with c[r][i] the number of times row i was drawn: one FP64 GEMM
// C[n x N] * X[N x D] for all resamples and columns.  The GEMM itself is a plain library
// DGEMM (cuBLAS, issued by the host mirror through torch.mm); what is hand-written here is the
// count matrix (the single-cell Lemire stream of K1 / K2, so the rows are the multiplicities of
// exactly the indices bootstrap_indices yields) and the finisher that turns the moment sums
// into the statistic.  On B200 the DGEMM runs at 35 TFLOP/s = 1.7e13 resampled points/s,
// against 3.2e12 for the shared-memory gather of K1c.
#pragma once
#include "common.cuh"
#include "draw.cuh"
#include "stats.cuh"

namespace b200boot {

constexpr int kK6Threads = 512;

// out[r][i] (float64, row-major [n_local][N]) = multiplicity of row i in resample resample_lo + r.
// One CTA per resample at a time; counts are built with shared-memory atomics (N <= 28 672).
__global__ void __launch_bounds__(kK6Threads) k6_count_matrix_kernel(uint64_t seed, int64_t n_rows,
                                                                     int64_t resample_lo, int64_t n_local,
                                                                     double* __restrict__ out) {
  extern __shared__ uint32_t s_count[];
  const int n = (int)n_rows;
  const uint32_t thresh = lemire_threshold((uint32_t)n);
  const int n_blocks = (n + 3) / 4;
  for (int64_t rl = blockIdx.x; rl < n_local; rl += gridDim.x) {
    for (int i = threadIdx.x; i < n; i += kK6Threads) s_count[i] = 0u;
    __syncthreads();
    const Stream st = make_stream(seed, resample_lo + rl, 0u, kPurposeIndex);
    for (int q = threadIdx.x; q < n_blocks; q += kK6Threads)
      lemire_block<true>(st, (uint32_t)q, (uint32_t)n, thresh, (int64_t)n,
                         [&](int, int64_t, uint32_t idx) { atomicAdd(s_count + idx, 1u); });
    __syncthreads();
    double* row = out + rl * n_rows;
    for (int i = threadIdx.x; i < n; i += kK6Threads) row[i] = (double)s_count[i];
    __syncthreads();
  }
}

// Count matrix from materialised index sets (any N: the tiled plans of K2): out[r][i] += 1 for
// every index i of set r.  `out` must be zeroed by the caller; FP64 atomics land in L2.
__global__ void __launch_bounds__(256) k6_counts_from_indices_kernel(const int64_t* __restrict__ idx,
                                                                     int64_t n_sets, int64_t n_rows,
                                                                     double* __restrict__ out) {
  const int64_t total = n_sets * n_rows;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t r = e / n_rows;
    const int64_t i = idx[e];
    if ((uint64_t)i < (uint64_t)n_rows) atomicAdd(out + r * n_rows + i, 1.0);   // out-of-range entries are ignored
  }
}

// s1 [D][n_local] holds sum_i c x (in place -> the statistic); s2 (var / std) holds
// sum_i c x^2 of the column-centred data.
__global__ void __launch_bounds__(256) k6_finish_kernel(int stat, double count, int64_t total,
                                                        double* __restrict__ s1, const double* __restrict__ s2) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    double m[2];
    m[0] = s1[e];
    m[1] = s2 ? s2[e] : 0.0;
    s1[e] = finish_stat(stat, m, count);
  }
}

inline bool gemm_supported(int stat, int64_t n_rows) {
  const int ms = moment_set_of(stat);
  return (ms == kMomSum || ms == kMomSum2) && n_rows >= 1 && n_rows <= kSmemDataBytes / 8;
}

}  // namespace b200boot
