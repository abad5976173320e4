// K1c / K3c — column-batched moment statistics over (N, D) data.
//
// `statfunction(x[indices])` with an axis-0 statistic over 2-D data gathers whole ROWS:
// every column sees the same index set (/root/reference/src/scikits/bootstrap/
// bootstrap.py:431).  For data sets that fit in shared memory as a row-major block of
// W columns (W = 32, 16, 8 or 4, the widest that fits), one set of draws serves W
// columns at once: the lanes of a group hold one column each, the drawn row number is
// broadcast by a shuffle and the row is read as one contiguous, conflict-free
// shared-memory access.  Index generation is amortised over the W columns and no
// warp reduction is needed (each lane finishes its own column).
//
// The draws are the single-cell Lemire stream (cell 0) that K1, K2 and K1m use for
// plans with one cell, so column d of the result equals the 1-D run on column d and
// `bootstrap_indices` replays the resamples.
#pragma once
#include "common.cuh"
#include "draw.cuh"
#include "k1_resample.cuh"
#include "stats.cuh"

namespace b200boot {

constexpr int kColThreads = 1024;
constexpr int kColWarps = kColThreads / 32;

// widest column block whose n_rows x W float64 tile fits in shared memory (0: none)
inline int column_block_width(int64_t n_rows) {
  for (int w = 32; w >= 4; w >>= 1)
    if (n_rows * w * 8 <= kSmemDataBytes) return w;
  return 0;
}

struct K1cParams {
  const double* data;      // [N][ld]
  const double* centers;   // [D] per-column shift (var/std) or null
  int64_t n_rows, n_cols, ld;
  int stat;
  uint64_t seed;
  int64_t resample_lo, n_local;
  double* out;             // [D][n_local]
  int32_t* work_counter;
  int32_t n_colblocks, rchunk, n_rchunks;
  int64_t n_items;
  RoundKeys rk;
};

template <int MS, int W>
__global__ void __launch_bounds__(kColThreads, 1) k1c_columns_kernel(const K1cParams P) {
  constexpr int NM = MomTraits<MS>::kMoments;
  constexpr int G = 32 / W;                      // resamples handled side by side in a warp
  extern __shared__ __align__(128) unsigned char smem_raw[];
  volatile int* s_item = reinterpret_cast<volatile int*>(smem_raw);
  double* s_tile = reinterpret_cast<double*>(smem_raw + kSmemHeader);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int col = lane % W, grp = lane / W;
  const uint32_t n = (uint32_t)P.n_rows;
  const uint32_t thresh = lemire_threshold(n);
  const uint32_t n_blocks = (n + 3) / 4;         // Philox blocks per resample

  int resident = -1;
  for (;;) {
    if (tid == 0) *s_item = atomicAdd(P.work_counter, 1);
    __syncthreads();
    const int64_t item = *s_item;
    __syncthreads();
    if (item >= P.n_items) break;
    const int cb = (int)(item / P.n_rchunks);
    const int rc = (int)(item - (int64_t)cb * P.n_rchunks);
    const int64_t d0 = (int64_t)cb * W;
    if (cb != resident) {
      // row-major block: s_tile[i * W + c] = data[i][d0 + c] - center[d0 + c]
      for (int64_t e = tid; e < (int64_t)n * W; e += kColThreads) {
        const int64_t i = e / W;
        const int c = (int)(e - i * W);
        double v = 0.0;
        if (d0 + c < P.n_cols) {
          v = __ldg(P.data + i * P.ld + d0 + c);
          if (P.centers) v -= __ldg(P.centers + d0 + c);
        }
        s_tile[e] = v;
      }
      resident = cb;
      __syncthreads();
    }
    const int64_t r_begin = (int64_t)rc * P.rchunk;
    const int64_t r_end = min(r_begin + (int64_t)P.rchunk, P.n_local);
    for (int64_t rl0 = r_begin + (int64_t)warp * G; rl0 < r_end; rl0 += (int64_t)kColWarps * G) {
      const int64_t rl = rl0 + grp;
      const bool live = rl < r_end;
      const Stream st = make_stream(P.seed, P.resample_lo + (live ? rl : r_begin), 0u, kPurposeIndex);
      double acc[NM][4];
#pragma unroll
      for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[m][j] = 0.0;
      // this lane's column inside a row; rows are W * 8 bytes apart
      const unsigned char* col_base = reinterpret_cast<const unsigned char*>(s_tile) + col * 8;
      constexpr uint32_t kRowBytes = W * 8;
      const int src0 = grp * W;
      // groups of W whole blocks (4 W draws): no bound checks, fully unrolled
      const uint32_t n_whole = n / 4;                    // blocks with 4 valid draws
      uint32_t q0 = 0;
      for (; q0 + W <= n_whole; q0 += W) {
        // row numbers are < 2^16: two per 32-bit word halve the shuffle traffic (shuffles
        // share the LSU data pipe with the row reads, which is what binds this kernel)
        uint32_t pair01 = 0u, pair23 = 0u;
        lemire_block<false>(st, q0 + (uint32_t)col, n, thresh, (int64_t)n,
                            [&](int j, int64_t, uint32_t idx) {
                              if (j == 0) pair01 = idx;
                              if (j == 1) pair01 |= idx << 16;
                              if (j == 2) pair23 = idx;
                              if (j == 3) pair23 |= idx << 16;
                            }, &P.rk);
#pragma unroll
        for (int src = 0; src < W; ++src) {
          const uint32_t p01 = __shfl_sync(0xffffffffu, pair01, src0 + src);
          const uint32_t p23 = __shfl_sync(0xffffffffu, pair23, src0 + src);
          const uint32_t off[4] = {(p01 & 0xFFFFu) * kRowBytes, (p01 >> 16) * kRowBytes,
                                   (p23 & 0xFFFFu) * kRowBytes, (p23 >> 16) * kRowBytes};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double v = *reinterpret_cast<const double*>(col_base + off[j]);
            acc[0][j] += v;
            if (NM == 2) acc[NM - 1][j] += v * v;
          }
        }
      }
      // what is left: fewer than W whole blocks and possibly one partial block
      if (q0 < n_blocks) {
        uint32_t mine[4] = {0u, 0u, 0u, 0u};
        const uint32_t q = q0 + (uint32_t)col;
        if (q < n_blocks)
          lemire_block<true>(st, q, n, thresh, (int64_t)n,
                             [&](int j, int64_t, uint32_t idx) { mine[j] = idx * kRowBytes; }, &P.rk);
        const int n_src = (int)(n_blocks - q0);
        for (int src = 0; src < n_src; ++src) {
          const uint32_t base = (q0 + (uint32_t)src) * 4u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t off = __shfl_sync(0xffffffffu, mine[j], src0 + src);
            if (base + j < n) {
              const double v = *reinterpret_cast<const double*>(col_base + off);
              acc[0][j] += v;
              if (NM == 2) acc[NM - 1][j] += v * v;
            }
          }
        }
      }
      if (live && d0 + col < P.n_cols) {
        double m[kMaxMoments];
#pragma unroll
        for (int k = 0; k < NM; ++k) m[k] = (acc[k][0] + acc[k][1]) + (acc[k][2] + acc[k][3]);
        P.out[(d0 + col) * P.n_local + rl] = finish_stat(P.stat, m, (double)n);
      }
    }
  }
}

// ---- K3c: per-column means and closed-form jackknife, one CTA per 32 columns ----------
// out[d][8] as b200boot_jackknife.  blockDim = (32 columns, 8 row groups).
__device__ __forceinline__ double colblock_sum(double v, double (*s_red)[33]) {
  __syncthreads();
  s_red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  double t = 0.0;
  for (int y = 0; y < 8; ++y) t += s_red[y][threadIdx.x];
  return t;
}

// per-column means (the shift of var / std), same summation order as K3c's first pass
__global__ void __launch_bounds__(256) k1c_column_means_kernel(const double* __restrict__ data,
                                                               int64_t n_rows, int64_t n_cols, int64_t ld,
                                                               double* __restrict__ centers) {
  __shared__ double s_red[8][33];
  const int64_t d = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const bool live = d < n_cols;
  double a = 0.0;
  for (int64_t i = threadIdx.y; i < n_rows; i += 8) a += live ? __ldg(data + i * ld + d) : 0.0;
  const double mean = colblock_sum(a, s_red) / (double)n_rows;
  if (live && threadIdx.y == 0) centers[d] = mean;
}

template <int MS>
__global__ void __launch_bounds__(256) k3c_jackknife_columns_kernel(const double* __restrict__ data,
                                                                    int64_t n_rows, int64_t n_cols,
                                                                    int64_t ld, int stat, int center,
                                                                    double* __restrict__ centers_out,
                                                                    double* __restrict__ out) {
  constexpr int NM = MomTraits<MS>::kMoments;
  __shared__ double s_red[8][33];
  const int64_t d = (int64_t)blockIdx.x * 32 + threadIdx.x;
  const bool live = d < n_cols;
  const double n = (double)n_rows;
  auto value = [&](int64_t i) { return live ? __ldg(data + i * ld + d) : 0.0; };
  // pass 0: column mean (the shift of var / std)
  double shift = 0.0;
  if (center) {
    double a = 0.0;
    for (int64_t i = threadIdx.y; i < n_rows; i += 8) a += value(i);
    shift = colblock_sum(a, s_red) / n;
    if (live && threadIdx.y == 0 && centers_out) centers_out[d] = shift;
  }
  // pass 1: contribution totals
  double tot[NM];
  {
    double a[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) a[k] = 0.0;
    for (int64_t i = threadIdx.y; i < n_rows; i += 8) {
      double v[1] = {value(i) - shift}, c[NM];
      contribution<MS>(v, c);
#pragma unroll
      for (int k = 0; k < NM; ++k) a[k] += c[k];
    }
#pragma unroll
    for (int k = 0; k < NM; ++k) tot[k] = colblock_sum(a[k], s_red);
  }
  auto loo = [&](int64_t i) {
    double v[1] = {value(i) - shift}, c[NM], m[NM];
    contribution<MS>(v, c);
#pragma unroll
    for (int k = 0; k < NM; ++k) m[k] = tot[k] - c[k];
    return finish_stat(stat, m, n - 1.0);
  };
  // pass 2: jackknife mean, accumulated relative to the first leave-one-out value
  const double j0 = loo(0);
  double a1 = 0.0;
  for (int64_t i = threadIdx.y; i < n_rows; i += 8) a1 += loo(i) - j0;
  const double jmean = j0 + colblock_sum(a1, s_red) / n;
  // pass 3: second and third moments of the deviations
  double a2 = 0.0, a3 = 0.0;
  for (int64_t i = threadIdx.y; i < n_rows; i += 8) {
    const double dv = jmean - loo(i);
    a2 += dv * dv;
    a3 += dv * dv * dv;
  }
  const double d2 = colblock_sum(a2, s_red);
  const double d3 = colblock_sum(a3, s_red);
  if (live && threadIdx.y == 0) {
    double* o = out + d * 8;
    o[0] = finish_stat(stat, tot, n);
    o[1] = jmean;
    o[2] = d2;
    o[3] = d3;
    o[4] = d3 / (6.0 * pow(d2, 1.5));
    o[5] = shift;
    o[6] = o[7] = 0.0;
  }
}

// ---- host side --------------------------------------------------------------------------
inline bool columns_supported(int stat, int64_t n_rows) {
  const int ms = moment_set_of(stat);
  return (ms == kMomSum || ms == kMomSum2) && column_block_width(n_rows) > 0 && n_rows >= 1;
}

template <int MS>
int launch_k3c(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, int stat, double* centers,
               double* out, cudaStream_t s) {
  const dim3 block(32, 8);
  const unsigned grid = (unsigned)((n_cols + 31) / 32);
  k3c_jackknife_columns_kernel<MS><<<grid, block, 0, s>>>(data, n_rows, n_cols, ld, stat,
                                                          stat_needs_centering(stat) ? 1 : 0, centers, out);
  B2_LAUNCH_CHECK();
  return 0;
}

inline int columns_jackknife(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, int stat,
                             double* centers, double* out, cudaStream_t s) {
  if (moment_set_of(stat) == kMomSum) return launch_k3c<kMomSum>(data, n_rows, n_cols, ld, stat, centers, out, s);
  return launch_k3c<kMomSum2>(data, n_rows, n_cols, ld, stat, centers, out, s);
}

template <int MS, int W>
int launch_k1c(const K1cParams& P0, cudaStream_t s) {
  K1cParams P = P0;
  const int sms = sm_count();
  P.n_colblocks = (int32_t)((P.n_cols + W - 1) / W);
  constexpr int G = 32 / W;
  const int64_t per_pass = (int64_t)kColWarps * G;          // resamples per CTA sweep
  int64_t want_r = ((int64_t)sms * 8 + P.n_colblocks - 1) / P.n_colblocks;
  int64_t rchunk = (P.n_local + want_r - 1) / want_r;
  rchunk = ((rchunk + per_pass - 1) / per_pass) * per_pass;
  P.rchunk = (int32_t)std::min<int64_t>(std::max<int64_t>(rchunk, per_pass), 1 << 20);
  P.n_rchunks = (int32_t)((P.n_local + P.rchunk - 1) / P.rchunk);
  P.n_items = (int64_t)P.n_colblocks * P.n_rchunks;
  if (P.n_items >= ((int64_t)1 << 31)) return fail(B200BOOT_EINVAL, "too many work items");
  static PerDeviceOnce configured;     // one per <MS, W> instantiation
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(k1c_columns_kernel<MS, W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kSmemHeader + kSmemDataBytes));
  }));
  const size_t smem = kSmemHeader + (size_t)P.n_rows * W * 8;
  const int grid = (int)std::min<int64_t>(sms, P.n_items);
  k1c_columns_kernel<MS, W><<<grid, kColThreads, smem, s>>>(P);
  B2_LAUNCH_CHECK();
  return 0;
}

template <int MS>
int dispatch_k1c(const K1cParams& P, int w, cudaStream_t s) {
  switch (w) {
    case 32: return launch_k1c<MS, 32>(P, s);
    case 16: return launch_k1c<MS, 16>(P, s);
    case 8: return launch_k1c<MS, 8>(P, s);
    default: return launch_k1c<MS, 4>(P, s);
  }
}

// ws: work counter + per-column centres
inline size_t columns_ws_bytes(int64_t n_cols) { return 256 + pad256((size_t)n_cols * 8) + 256; }

inline int columns_resample(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, int stat,
                            uint64_t seed, int64_t resample_lo, int64_t resample_hi, double* stat_out,
                            void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!data || !stat_out || n_cols < 1 || ld < n_cols) return fail(B200BOOT_EINVAL, "columns: bad arguments");
  if (!columns_supported(stat, n_rows))
    return fail(B200BOOT_EUNSUPPORTED, "columns: statistic %d with %lld rows is not column-batched", stat,
                (long long)n_rows);
  const int64_t n_local = resample_hi - resample_lo;
  if (n_local < 0 || resample_lo < 0) return fail(B200BOOT_EINVAL, "bad resample range");
  if (n_local == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  int32_t* counter = cv.take<int32_t>(64);
  double* centers = cv.take<double>((size_t)n_cols);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  K1cParams P{};
  P.data = data;
  P.centers = nullptr;
  if (stat_needs_centering(stat)) {
    const dim3 block(32, 8);
    k1c_column_means_kernel<<<(unsigned)((n_cols + 31) / 32), block, 0, s>>>(data, n_rows, n_cols, ld, centers);
    B2_LAUNCH_CHECK();
    P.centers = centers;
  }
  P.n_rows = n_rows;
  P.n_cols = n_cols;
  P.ld = ld;
  P.stat = stat;
  P.seed = seed;
  P.resample_lo = resample_lo;
  P.n_local = n_local;
  P.out = stat_out;
  P.work_counter = counter;
  P.rk = make_round_keys(seed);
  B2_CUDA(cudaMemsetAsync(counter, 0, 256, s));
  const int w = column_block_width(n_rows);
  if (moment_set_of(stat) == kMomSum) return dispatch_k1c<kMomSum>(P, w, s);
  return dispatch_k1c<kMomSum2>(P, w, s);
}

}  // namespace b200boot
