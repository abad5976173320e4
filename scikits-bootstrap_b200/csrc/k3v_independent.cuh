// multi='independent' with a general two-sample statistic f(s_a(a), s_b(b)).
//
// Reference: every array is resampled on its own (bootstrap.py:438-444, 683-691) and the BCa
// jackknife pools sum_k N_k leave-one-out values, array-major: delete row j of array k, keep the
// others whole (bootstrap.py:601-607, 708-714).  With s_a, s_b scalar registry statistics and f
// one of the binary operations below, a leave-one-out value is f(loo_a[j], s_b) or
// f(s_a, loo_b[j]); the kernels here emit the per-array leave-one-out VECTORS (closed forms, as
// K3 / K3m use them) and reduce the pooled moments the acceleration needs.
#pragma once
#include "common.cuh"
#include "k1m_median.cuh"
#include "k3_jackknife.cuh"

namespace b200boot {

__device__ __forceinline__ double combine_op(int op, double a, double b) {
  switch (op) {
    case B200BOOT_OP_SUB: return a - b;
    case B200BOOT_OP_ADD: return a + b;
    case B200BOOT_OP_DIV: return a / b;
    case B200BOOT_OP_MUL: return a * b;
    default: return log(a / b);            // B200BOOT_OP_LOG_RATIO
  }
}

__global__ void __launch_bounds__(256) combine_kernel(int op, const double* __restrict__ a,
                                                      const double* __restrict__ b, double* __restrict__ out,
                                                      int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = combine_op(op, a[i], b[i]);
}

// leave-one-out statistic of every row of a moment statistic (totals already in *st)
template <int MS>
__global__ void __launch_bounds__(256) loo_values_kernel(const JackState* st, int stat, int64_t n_rows,
                                                         const double* __restrict__ d0,
                                                         const double* __restrict__ d1, double* __restrict__ out) {
  const double n = (double)n_rows;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rows; i += stride)
    out[i] = leave_one_out<MS>(st, stat, n, d0, d1, i);
}

// leave-one-out rank statistic for every SORTED position (the pooled moments do not depend on
// the order of the values)
__global__ void __launch_bounds__(256) loo_values_sorted_kernel(const double* __restrict__ sv, int64_t n,
                                                                const RankRule loo, double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += stride)
    out[r] = loo_median(sv, n, r, loo);
}

// pooled leave-one-out value i of [0, n_a + n_b): array-major
struct PooledLoo {
  int op;
  const double* loo_a;
  int64_t n_a;
  const double* s_a;      // device scalars: the statistics on the whole arrays
  const double* loo_b;
  int64_t n_b;
  const double* s_b;
  __device__ __forceinline__ double at(int64_t i) const {
    return i < n_a ? combine_op(op, loo_a[i], *s_b) : combine_op(op, *s_a, loo_b[i - n_a]);
  }
};
struct PooledSumF {       // relative to the first value: identical values give d == 0 exactly
  PooledLoo p;
  __device__ void operator()(int64_t i, double* v) const { v[0] = p.at(i) - p.at(0); }
};
struct StorePooledMeanG {
  PooledLoo p;
  JackState* st;
  __device__ void operator()(const double* tot) const {
    st->j0 = p.at(0);
    st->jmean = st->j0 + tot[0] / (double)(p.n_a + p.n_b);
  }
};
struct PooledDevF {
  PooledLoo p;
  const JackState* st;
  __device__ void operator()(int64_t i, double* v) const {
    const double d = st->jmean - p.at(i);
    v[0] = d * d;
    v[1] = d * d * d;
  }
};
struct StorePooledG {
  PooledLoo p;
  JackState* st;
  double* out;
  __device__ void operator()(const double* tot) const {
    out[0] = combine_op(p.op, *p.s_a, *p.s_b);
    out[1] = st->jmean;
    out[2] = tot[0];
    out[3] = tot[1];
    out[4] = tot[1] / (6.0 * pow(tot[0], 1.5));
    out[5] = out[6] = out[7] = 0.0;
  }
};

}  // namespace b200boot
