// libb200boot.so — C ABI (include/b200boot.h) over the sm_100a kernels.
// Host side only enqueues work on the caller's stream; all scratch comes from
// the caller's workspace.
#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "draw.cuh"
#include "k1_resample.cuh"
#include "k1c_columns.cuh"
#include "k1m_median.cuh"
#include "k1mr_rows.cuh"
#include "k6_gemm.cuh"
#include "k6i_mma.cuh"
#include "k6i_pipe.cuh"
#include "k1mg_gemm.cuh"
#include "k3_jackknife.cuh"
#include "k45_interval.cuh"
#include "small_call.cuh"
#include "k3v_independent.cuh"
#include "abc_method.cuh"
#include "stats.cuh"

using namespace b200boot;

namespace {

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// optional event timing of the occupancy kernels (K0) and of the K1 launch on the caller's
// stream (bench.py); per host thread, like the calls it times
struct K1Profile {
  bool enabled = false;
  bool valid = false;
  cudaEvent_t begin = nullptr, start = nullptr, stop = nullptr;   // K0 from begin to start, K1 from start to stop
};
thread_local K1Profile g_prof;

inline int blocks_for(int64_t n, int threads, int64_t cap) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// ---- workspace layouts ---------------------------------------------------------
struct ResampleWs {
  int32_t* counts;
  uint16_t* cls;        // [n_full][n_local][16]
  int32_t* gcounts;     // [n_groups][n_local] level-1 occupancies
  HalfTables half_tables;   // BTRS constants of Bin(n, 1/2) and log(n!), n < kHalfTableSize
  int32_t* overflow;    // set when a class count did not fit its uint16 slot
  double* partial;
  int32_t* work_counter;
  JackState* jack;
  double* red;          // block partials of the reductions
  double* centered[2];
};

HalfTables take_half_tables(Carver& cv) {
  HalfTables t;
  t.consts = cv.take<HalfConsts>((size_t)kHalfTableSize);
  t.log_fact = cv.take<double>((size_t)kHalfTableSize);
  return t;
}

ResampleWs carve_resample(Carver& cv, const Plan& plan, int stat, int64_t n_local) {
  ResampleWs w{};
  const int ms = moment_set_of(stat);
  const int nm = ms == kMomSum ? 1 : (ms == kMomPair5 ? 5 : 2);
  w.work_counter = cv.take<int32_t>(64);
  w.overflow = cv.take<int32_t>(64);          // directly behind the work counter: one memset clears both
  w.jack = cv.take<JackState>(1);
  w.red = cv.take<double>((size_t)kRedBlocks * kMaxMoments);
  w.counts = plan.n_tiles > 1 ? cv.take<int32_t>((size_t)plan.n_tiles * n_local) : nullptr;
  w.gcounts = plan.n_tiles > 1 ? cv.take<int32_t>((size_t)n_tile_groups(plan) * std::max<int64_t>(n_local, 1)) : nullptr;
  w.cls = plan.n_full > 0 ? cv.take<uint16_t>((size_t)plan.n_full * n_local * kClasses) : nullptr;
  if (plan.n_full > 0) w.half_tables = take_half_tables(cv);
  w.partial = cv.take<double>((size_t)plan.n_tiles * nm * std::max<int64_t>(n_local, 1));
  if (stat_needs_centering(stat))
    for (int a = 0; a < plan.n_arrays; ++a) w.centered[a] = cv.take<double>((size_t)plan.n_rows);
  return w;
}

template <int NV, typename F, typename G>
int run_reduction(int64_t n, F f, G g, double* red, cudaStream_t s) {
  reduce_rows_kernel<NV, F><<<kRedBlocks, kRedThreads, 0, s>>>(n, f, red);
  B2_LAUNCH_CHECK();
  reduce_final_kernel<NV, G><<<1, kRedThreads, 0, s>>>(red, kRedBlocks, g);
  B2_LAUNCH_CHECK();
  return 0;
}

// Centres the arrays on their means when the statistic asks for it; returns the
// pointers the moment kernels should read.
int prepare_rows(const double* const* data, int n_arrays, int64_t n_rows, int stat, ResampleWs& w,
                 const double** rows, cudaStream_t s) {
  rows[0] = data[0];
  rows[1] = n_arrays > 1 ? data[1] : nullptr;
  B2_CUDA(cudaMemsetAsync(w.jack, 0, sizeof(JackState), s));
  if (!stat_needs_centering(stat)) return 0;
  if (n_rows <= kSmallMaxRows) {         // short arrays: means and centred rows in one single-CTA launch
    center_small_kernel<<<1, kSmallThreads, 0, s>>>(rows[0], rows[1], n_rows, w.jack, w.centered[0],
                                                    n_arrays > 1 ? w.centered[1] : nullptr);
    B2_LAUNCH_CHECK();
    for (int a = 0; a < n_arrays; ++a) rows[a] = w.centered[a];
    return 0;
  }
  int rc = run_reduction<2>(n_rows, RawSumsF{rows[0], rows[1]}, StoreCentersG{w.jack, (double)n_rows},
                            w.red, s);
  if (rc) return rc;
  for (int a = 0; a < n_arrays; ++a) {
    center_rows_kernel<<<blocks_for(n_rows, 256, 2368), 256, 0, s>>>(rows[a], w.centered[a], n_rows,
                                                                    w.jack, a);
    B2_LAUNCH_CHECK();
    rows[a] = w.centered[a];
  }
  return 0;
}

// K0, tile level: occupancies of the cells (two-level chains when `gcounts` is given).
int launch_tile_occupancies(uint64_t seed, int64_t resample_lo, int64_t n_local, const Plan& plan,
                            int32_t* counts, int32_t* gcounts, cudaStream_t s) {
  if (plan.n_tiles <= 1) return 0;
  if (gcounts) {
    const int64_t ng = n_tile_groups(plan);
    if (ng > 1) {
      k0_group_counts_kernel<<<blocks_for(n_local, 128, 1 << 30), 128, 0, s>>>(seed, resample_lo, n_local, plan,
                                                                             gcounts);
      B2_LAUNCH_CHECK();
    }
    k0_tiles_in_groups_kernel<<<blocks_for(ng * n_local, 128, 1 << 30), 128, 0, s>>>(seed, resample_lo, n_local,
                                                                                    plan, gcounts, counts);
    B2_LAUNCH_CHECK();
  } else {
    k0_tile_counts_kernel<<<blocks_for(n_local, 128, 1 << 30), 128, 0, s>>>(seed, resample_lo, n_local, plan,
                                                                          counts);
    B2_LAUNCH_CHECK();
  }
  return 0;
}

// K0 as the inspection API and K2 see it: tile occupancies, then the 16-way class split of
// every full tile as int32 records (one thread per pair walks the split tree).
int launch_occupancies(uint64_t seed, int64_t resample_lo, int64_t n_local, const Plan& plan,
                       int32_t* counts, int32_t* cls, cudaStream_t s) {
  int rc = launch_tile_occupancies(seed, resample_lo, n_local, plan, counts, nullptr, s);
  if (rc) return rc;
  if (plan.n_full > 0 && cls) {
    const int64_t total = plan.n_full * n_local;
    const int64_t blocks = std::min<int64_t>((total + 127) / 128, (int64_t)sm_count() * 64);
    k0_class_counts_kernel<<<(unsigned)blocks, 128, 0, s>>>(seed, resample_lo, n_local, plan.n_full, counts, cls);
    B2_LAUNCH_CHECK();
  }
  return 0;
}

// K0 of the resample path: two-level tile chains, then the class split tree level by level
// (k0_class_tree_kernel) into uint16 records.  Same values as launch_occupancies.
int launch_occupancies_fast(uint64_t seed, int64_t resample_lo, int64_t n_local, const Plan& plan,
                            int32_t* counts, int32_t* gcounts, uint16_t* cls, HalfTables half_tables,
                            int32_t* overflow, cudaStream_t s) {
  int rc = launch_tile_occupancies(seed, resample_lo, n_local, plan, counts, gcounts, s);
  if (rc) return rc;
  if (plan.n_full > 0 && cls) {
    const int64_t total = plan.n_full * n_local;
    if (total >= ((int64_t)1 << 31)) return fail(B200BOOT_EINVAL, "too many (tile, resample) pairs");
    k0_half_table_kernel<<<kHalfTableSize / 256, 256, 0, s>>>(half_tables);
    B2_LAUNCH_CHECK();
    const int64_t n_batches = (total + kTreeBatch - 1) / kTreeBatch;
    const int grid = (int)std::min<int64_t>(n_batches, (int64_t)sm_count() * 32);
    k0_class_tree_kernel<<<grid, kTreeThreads, 0, s>>>(seed, resample_lo, n_local, plan.n_full, counts, half_tables,
                                                       cls, overflow, make_round_keys(seed));
    B2_LAUNCH_CHECK();
  }
  return 0;
}

int check_moment_stat(int stat, int n_arrays, int64_t n_rows) {
  if (moment_set_of(stat) < 0)
    return fail(B200BOOT_EUNSUPPORTED, "statistic id %d is not a moment statistic of the device registry",
                stat);
  if (n_arrays != stat_arrays(stat))
    return fail(B200BOOT_EINVAL, "statistic id %d takes %d array(s), got %d", stat, stat_arrays(stat),
                n_arrays);
  if (n_rows < 1 || n_rows >= ((int64_t)1 << 31))
    return fail(B200BOOT_EINVAL, "n_rows=%lld outside [1, 2^31)", (long long)n_rows);
  return 0;
}

// FAST: the cell's shared-window address is a compile-time constant (see kFastSmemBase)
template <int MS, bool FAST>
int launch_k1_variant(const K1Params& P, size_t smem, int grid, cudaStream_t s) {
  static PerDeviceOnce configured;     // one per <MS, FAST> instantiation
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(k1_resample_kernel<MS, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kSmemHeader + kSmemDataBytes));
  }));
  k1_resample_kernel<MS, FAST><<<grid, kK1Threads, smem, s>>>(P);
  B2_LAUNCH_CHECK();
  return 0;
}

// One-time probe per device, under std::call_once: is the dynamic shared memory of a kernel
// without static shared memory at 0x400, as the FAST variants assume?  The probe runs on its
// own non-blocking stream with a pinned result word, so it neither touches the legacy
// stream nor allocates device memory; any failure selects the generic-address variants.
inline bool fast_smem_addressing() {
  static PerDeviceOnce probed;
  static std::atomic<int> verdict[kMaxDevices];   // 0 unknown / no, 1 yes
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return false;
  probed.run([dev] {
    int reserved = 0;
    if (cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev) != cudaSuccess ||
        reserved != 1024)
      return cudaSuccess;
    uint32_t* h_out = nullptr;
    cudaStream_t ps = nullptr;
    if (cudaHostAlloc(&h_out, sizeof(uint32_t), cudaHostAllocMapped) == cudaSuccess) {
      uint32_t* d_out = nullptr;
      *h_out = 0;
      if (cudaHostGetDevicePointer(&d_out, h_out, 0) == cudaSuccess &&
          cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking) == cudaSuccess) {
        probe_dynamic_smem_base_kernel<<<1, 32, 1024, ps>>>(d_out);
        if (cudaStreamSynchronize(ps) == cudaSuccess && *h_out == 1024u) verdict[dev].store(1);
        cudaStreamDestroy(ps);
      }
      cudaFreeHost(h_out);
    }
    cudaGetLastError();   // a failed probe must not leave a sticky-looking error for the caller
    return cudaSuccess;
  });
  return verdict[dev].load() == 1;
}

template <int MS>
int launch_k1(const K1Params& P, size_t smem, int grid, cudaStream_t s) {
  // the 2-array fast variant also assumes 2^13-row tiles (kFastStride2)
  const bool fast = fast_smem_addressing() && P.plan.n_full > 0 &&
                    (MomTraits<MS>::kArrays == 1 || P.plan.log2_tile == 13);
  return fast ? launch_k1_variant<MS, true>(P, smem, grid, s) : launch_k1_variant<MS, false>(P, smem, grid, s);
}

int dispatch_k1(int ms, const K1Params& P, size_t smem, int grid, cudaStream_t s) {
  switch (ms) {
    case kMomSum: return launch_k1<kMomSum>(P, smem, grid, s);
    case kMomSum2: return launch_k1<kMomSum2>(P, smem, grid, s);
    case kMomAB: return launch_k1<kMomAB>(P, smem, grid, s);
    case kMomXW: return launch_k1<kMomXW>(P, smem, grid, s);
    default: return launch_k1<kMomPair5>(P, smem, grid, s);
  }
}

template <int MS>
int launch_k1i(const double* d0, const double* d1, int64_t n_rows, const int64_t* idx, int64_t n_sets,
               double* partial, cudaStream_t s) {
  const int grid = blocks_for(n_sets * 32, 256, 148 * 8);
  k1_indexed_kernel<MS><<<grid, 256, 0, s>>>(d0, d1, n_rows, idx, n_sets, partial);
  B2_LAUNCH_CHECK();
  return 0;
}

template <int MS>
int jack_passes(const double** rows, int64_t n_rows, int stat, ResampleWs& w, double* out,
                cudaStream_t s) {
  const double n = (double)n_rows;
  if (n_rows <= kSmallMaxRows) {         // short arrays: the three passes in one single-CTA launch
    k3_small_kernel<MS><<<1, kSmallThreads, 0, s>>>(rows[0], rows[1], n_rows, stat, w.jack, out);
    B2_LAUNCH_CHECK();
    return 0;
  }
  int rc = run_reduction<MomTraits<MS>::kMoments>(n_rows, ContributionF<MS>{rows[0], rows[1]},
                                                  StoreTotalsG<MS>{w.jack, stat, n, rows[0], rows[1]}, w.red, s);
  if (rc) return rc;
  rc = run_reduction<1>(n_rows, JackSumF<MS>{w.jack, stat, n, rows[0], rows[1]}, StoreJmeanG{w.jack, n},
                        w.red, s);
  if (rc) return rc;
  return run_reduction<2>(n_rows, JackDevF<MS>{w.jack, stat, n, rows[0], rows[1]},
                          StoreAccelG{w.jack, out}, w.red, s);
}

// leave-one-out statistic of every row (multi='independent'): totals by the jackknife passes,
// then one elementwise kernel
template <int MS>
int loo_values(const double** rows, int64_t n_rows, int stat, ResampleWs& w, double* values, double* summary,
               cudaStream_t s) {
  int rc = jack_passes<MS>(rows, n_rows, stat, w, summary, s);       // leaves the totals in w.jack
  if (rc) return rc;
  loo_values_kernel<MS><<<blocks_for(n_rows, 256, 148 * 16), 256, 0, s>>>(w.jack, stat, n_rows, rows[0], rows[1],
                                                                         values);
  B2_LAUNCH_CHECK();
  return 0;
}

}  // namespace

namespace {

// ---- K1mg: medians of many short columns through the tensor cores (csrc/k1mg_gemm.cuh) --------
// measured against K1m (tools/mg_crossover.py, n = 1e4): N = 1000: D = 2 0.12 vs 0.25 ms, D = 32 0.14 vs 0.33,
// D = 64 0.18 vs 0.41, D = 1e4 (cfg4) 8.1 vs 24 ms; N = 200, D = 8: 0.09 vs 0.12; N = 64, D = 16: 0.09 vs 0.11
constexpr int64_t kMgMinCols = 2;
constexpr int64_t kMgMaxRows = 1024;      // resident count tile (128 KB) + two digit stages + two permutation blocks
constexpr size_t kMgStreamMaxWs = (size_t)3 << 30;   // longer columns: only while the images stay moderate

inline int64_t mg_k_pad(int64_t n_rows) { return ((n_rows + kI8KChunk - 1) / kI8KChunk) * kI8KChunk; }
inline int64_t mg_col_tiles(int64_t n_cols) { return (n_cols + kMgColsPerTile - 1) / kMgColsPerTile; }
size_t median_gemm_ws_bytes(int64_t n_rows, int64_t n_cols, int64_t n_local);
inline bool median_gemm_applies(int64_t n_rows, int64_t n_cols, int64_t n_local) {
  if (n_cols < kMgMinCols || n_rows < 32) return false;
  if (n_rows <= kMgMaxRows) return true;
  // streamed form (1024 < N <= 28 672): measured against K1m / the K1m1 loop in tools/mg_crossover.py
  return n_rows <= kMedMaxRows && median_gemm_ws_bytes(n_rows, n_cols, n_local) <= kMgStreamMaxWs;
}

struct MgWs {
  int32_t* flags;
  uint64_t* keys;
  uint32_t* idx;
  double* sorted;
  int32_t* rank;
  uint8_t* counts;
  unsigned char* a_img;
  unsigned char* b_img;
  uint16_t* perm_img;
  int64_t n_pad, k_pad;
};

MgWs mg_carve(Carver& cv, int64_t n_rows, int64_t n_cols, int64_t n_local) {
  MgWs w{};
  w.n_pad = 1;
  while (w.n_pad < n_rows) w.n_pad <<= 1;
  w.k_pad = mg_k_pad(n_rows);
  const int64_t m_tiles = (std::max<int64_t>(n_local, 1) + kI8M - 1) / kI8M;
  w.flags = cv.take<int32_t>(64);
  w.keys = cv.take<uint64_t>((size_t)w.n_pad * n_cols);
  w.idx = cv.take<uint32_t>((size_t)w.n_pad * n_cols);
  w.sorted = cv.take<double>((size_t)n_rows * n_cols);
  w.rank = cv.take<int32_t>((size_t)n_rows * n_cols);
  w.counts = cv.take<uint8_t>((size_t)std::max<int64_t>(n_local, 1) * w.k_pad);
  w.a_img = cv.take<unsigned char>((size_t)m_tiles * (size_t)(w.k_pad / kI8KChunk) * kI8AChunkBytes);
  w.b_img = cv.take<unsigned char>((size_t)mg_col_tiles(n_cols) * (size_t)(w.k_pad / kI8KChunk) * kI8BChunkBytes);
  w.perm_img = n_rows <= kMgMaxRows ? cv.take<uint16_t>((size_t)mg_col_tiles(n_cols) * kMgColsPerTile * (size_t)w.k_pad)
                                    : nullptr;
  return w;
}

size_t median_gemm_ws_bytes(int64_t n_rows, int64_t n_cols, int64_t n_local) {
  if (n_cols < kMgMinCols || n_rows < 32 || n_rows > kMedMaxRows) return 0;
  Carver cv(reinterpret_cast<void*>(256), ~(size_t)0 >> 1);
  mg_carve(cv, n_rows, n_cols, n_local);
  return cv.used;
}

int median_gemm(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, uint64_t seed, int64_t resample_lo,
                int64_t n_local, double* stat_out, void* ws, size_t ws_bytes, cudaStream_t s, const RankRule& full,
                const RankRule& loo, double* jack_out) {
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  MgWs w = mg_carve(cv, n_rows, n_cols, n_local);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  // every column sorted as (key, row) pairs: sorted values, permutation (idx) and rank table
  psort_load_kernel<<<blocks_for(w.n_pad * n_cols, 256, 148 * 32), 256, 0, s>>>(data, n_rows, n_cols, ld, w.n_pad,
                                                                                w.keys, w.idx);
  B2_LAUNCH_CHECK();
  for (int64_t d0 = 0; d0 < n_cols; d0 += 65535) {
    const unsigned nd = (unsigned)std::min<int64_t>(65535, n_cols - d0);
    int rc = psort_pairs(w.keys + d0 * w.n_pad, w.idx + d0 * w.n_pad, w.n_pad, nd, s);
    if (rc) return rc;
  }
  rows_rank_kernel<<<blocks_for(n_rows * n_cols, 256, 148 * 32), 256, 0, s>>>(w.keys, w.idx, n_rows, w.n_pad, n_cols, 0,
                                                                              n_cols, w.rank, w.sorted);
  B2_LAUNCH_CHECK();
  if (jack_out) {
    for (int64_t d0 = 0; d0 < n_cols; d0 += 1 << 30) {
      const int64_t nd = std::min<int64_t>((int64_t)1 << 30, n_cols - d0);
      k3m_jackknife_median_kernel<<<(unsigned)nd, 256, 0, s>>>(w.sorted + d0 * n_rows, n_rows, full, loo,
                                                               jack_out + d0 * 8);
      B2_LAUNCH_CHECK();
    }
  }
  // the resamples' count matrix (the multiplicities of the single-cell Lemire stream) and both images
  B2_CUDA(cudaMemsetAsync(w.flags, 0, 256, s));
  {
    static PerDeviceOnce configured;
    B2_CUDA(configured.run([] {
      return cudaFuncSetAttribute(k6i_count_bytes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    }));
    const int grid = (int)std::min<int64_t>((int64_t)sm_count() * 4, n_local);
    k6i_count_bytes_kernel<<<grid, 256, (size_t)w.k_pad, s>>>(seed, n_rows, w.k_pad, resample_lo, n_local, w.counts,
                                                             w.flags);
    B2_LAUNCH_CHECK();
  }
  const int64_t m_tiles = (n_local + kI8M - 1) / kI8M;
  const int64_t n_col_tiles = mg_col_tiles(n_cols);
  const int n_kchunks = (int)(w.k_pad / kI8KChunk);
  k6i_image_kernel<kI8M><<<blocks_for(m_tiles * kI8M * (w.k_pad / 16), 256, 148 * 16), 256, 0, s>>>(
      w.counts, n_local, w.k_pad, m_tiles, w.a_img);
  B2_LAUNCH_CHECK();
  const int seg = (int)((n_rows + kMgT - 1) / kMgT);
  k1mg_indicator_image_kernel<<<blocks_for(n_col_tiles * kMgColsPerTile * (w.k_pad / 16), 256, 148 * 16), 256, 0, s>>>(
      w.rank, n_rows, n_cols, w.k_pad, n_col_tiles, seg, w.b_img);
  B2_LAUNCH_CHECK();
  // ~32 work items per CTA: fine enough to balance, coarse enough that reloading the count tile
  // per item (128 KB against seg_tiles x 256 KB of indicator chunks) stays a few per cent
  const int64_t want_segs = std::max<int64_t>(1, (32 * (int64_t)sm_count() + m_tiles - 1) / m_tiles);
  const int64_t seg_tiles = std::max<int64_t>(1, (n_col_tiles + want_segs - 1) / want_segs);
  const int64_t n_items = ((n_col_tiles + seg_tiles - 1) / seg_tiles) * m_tiles;
  const int grid = (int)std::min<int64_t>(sm_count(), n_items);
  const size_t limit = 227 * 1024;
  if (n_rows > kMgMaxRows) {
    // longer columns: count chunks streamed next to the indicator chunks, walks over L2
    static PerDeviceOnce configured_stream;
    B2_CUDA(configured_stream.run([] {
      return cudaFuncSetAttribute(k1mg_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }));
    int n_stages = kPipeMaxStages;
    while (n_stages > 2 && k1mg_stream_smem_bytes(n_stages) > limit) --n_stages;
    K1mgStreamParams Q{};
    Q.a_img = w.a_img;
    Q.b_img = w.b_img;
    Q.counts = w.counts;
    Q.perm = w.idx;
    Q.perm_ld = w.n_pad;
    Q.sorted = w.sorted;
    Q.m_total = n_local;
    Q.m_tiles = m_tiles;
    Q.n_kchunks = n_kchunks;
    Q.n_stages = n_stages;
    Q.n_rows = n_rows;
    Q.n_cols = n_cols;
    Q.n_col_tiles = n_col_tiles;
    Q.seg_tiles = seg_tiles;
    Q.seg = seg;
    Q.rule = full;
    Q.out = stat_out;
    Q.out_ld = n_local;
    Q.overflow = w.flags;
    k1mg_stream_kernel<<<grid, kPipeThreads, k1mg_stream_smem_bytes(n_stages), s>>>(Q);
    B2_LAUNCH_CHECK();
    return 0;
  }
  k1mg_perm_image_kernel<<<blocks_for(n_col_tiles * kMgColsPerTile * w.k_pad, 256, 148 * 16), 256, 0, s>>>(
      w.idx, n_rows, w.n_pad, n_cols, w.k_pad, n_col_tiles, w.perm_img);
  B2_LAUNCH_CHECK();
  int n_stages = kPipeMaxStages;
  while (n_stages > 2 && k1mg_smem_bytes(n_kchunks, n_stages) > limit) --n_stages;
  const size_t smem = k1mg_smem_bytes(n_kchunks, n_stages);
  if (smem > limit) return fail(B200BOOT_EUNSUPPORTED, "K1mg: shared memory");
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(k1mg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }));
  K1mgParams P{};
  P.a_img = w.a_img;
  P.b_img = w.b_img;
  P.perm_img = w.perm_img;
  P.sorted = w.sorted;
  P.m_total = n_local;
  P.m_tiles = m_tiles;
  P.n_kchunks = n_kchunks;
  P.n_stages = n_stages;
  P.n_rows = n_rows;
  P.n_cols = n_cols;
  P.n_col_tiles = n_col_tiles;
  P.seg = seg;
  P.rule = full;
  P.out = stat_out;
  P.out_ld = n_local;
  P.overflow = w.flags;
  P.seg_tiles = seg_tiles;
  k1mg_kernel<<<grid, kPipeThreads, smem, s>>>(P);
  B2_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int b200boot_version(void) { return B200BOOT_ABI_VERSION; }
int b200boot_philox_rounds(void) { return kPhiloxRounds; }
const char* b200boot_last_error(void) { return err_buf(); }

int b200boot_make_plan(int64_t n_rows, int n_arrays, b200boot_plan* out) {
  if (!out || n_rows < 1 || n_arrays < 1 || n_arrays > 4)
    return fail(B200BOOT_EINVAL, "make_plan: bad arguments");
  const Plan p = make_plan(n_rows, n_arrays);
  out->log2_tile = p.log2_tile;
  out->n_arrays = p.n_arrays;
  out->n_full = p.n_full;
  out->rem = p.rem;
  out->n_tiles = p.n_tiles;
  return 0;
}

size_t b200boot_workspace_bytes(int stat_id, int64_t n_rows, int n_arrays, int64_t n_cols,
                                int64_t n_local, int n_alphas) {
  if (n_rows < 1 || n_arrays < 1 || n_local < 0 || n_cols < 1) return 0;
  size_t total = 4096;
  if (stat_id == kStatMedian) {
    total += std::max(median_ws_bytes(n_rows, n_cols, n_local),
                      median_gemm_applies(n_rows, n_cols, n_local) ? median_gemm_ws_bytes(n_rows, n_cols, n_local) : (size_t)0);
  } else if (n_cols > 1) {
    total += columns_ws_bytes(n_cols);
  } else {
    const Plan plan = make_plan(n_rows, n_arrays);
    Carver cv(reinterpret_cast<void*>(256), ~(size_t)0 >> 1);
    carve_resample(cv, plan, stat_id, n_local);
    total += cv.used;
  }
  // select state, or sort keys (padded to a power of two), whichever is larger
  const size_t sel = select_ws_bytes(std::max(n_alphas, 1), n_cols);
  int64_t n_pad = 1;
  while (n_pad < std::max<int64_t>(n_local, 1)) n_pad <<= 1;
  const size_t srt = pad256((size_t)n_pad * (size_t)n_cols * 8);
  total += std::max(sel, srt);
  return total;
}

int b200boot_resample_stat(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                           uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                           double* stat_out, void* ws, size_t ws_bytes, void* stream) {
  return b200boot_resample_stat_after(data, n_arrays, n_rows, stat_id, seed, resample_lo, resample_hi, stat_out,
                                      ws, ws_bytes, stream, nullptr);
}

}  // extern "C"

namespace {

// K0 + K1 + finish.  The occupancy chains (K0) do not read the data, so the data may still be on
// their way while K0 runs: either the caller copies them on another stream and hands over the
// event recorded behind the copy (`data_ready_event`), or the rows are still in HOST memory
// (`host_data`, pageable or pinned) and are copied here, on `copy_stream`, after K0 has been
// enqueued — the host-side staging of a pageable array (8 ms for 80 MB) and the transfer then run
// under K0 instead of in front of it.
int resample_stat_impl(const double* const* data, int n_arrays, int64_t n_rows, int stat_id, uint64_t seed,
                       int64_t resample_lo, int64_t resample_hi, double* stat_out, void* ws, size_t ws_bytes,
                       void* stream, void* data_ready_event, const double* const* host_data, void* copy_stream) {
  int rc = check_moment_stat(stat_id, n_arrays, n_rows);
  if (rc) return rc;
  const int64_t n_local = resample_hi - resample_lo;
  if (n_local < 0 || resample_lo < 0) return fail(B200BOOT_EINVAL, "bad resample range");
  if (n_local == 0) return 0;
  if (!data || !stat_out) return fail(B200BOOT_EINVAL, "null pointer");
  cudaStream_t s = as_stream(stream);
  const Plan plan = make_plan(n_rows, n_arrays);
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  ResampleWs w = carve_resample(cv, plan, stat_id, n_local);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);

  // The occupancy chains do not read the data: when the caller is still copying it in on
  // another stream, only what follows them waits for the copy.
  cudaEvent_t ready = reinterpret_cast<cudaEvent_t>(data_ready_event);
  cudaEvent_t ev_entry = nullptr, ev_copied = nullptr;
  if (host_data) {
    // the device rows may have earlier users on `s` (a recycled allocation): the copy stream
    // starts behind everything enqueued on `s` so far, but not behind K0
    B2_CUDA(cudaEventCreateWithFlags(&ev_entry, cudaEventDisableTiming));
    B2_CUDA(cudaEventCreateWithFlags(&ev_copied, cudaEventDisableTiming));
    B2_CUDA(cudaEventRecord(ev_entry, s));
  }
  B2_CUDA(cudaMemsetAsync(w.work_counter, 0, 512, s));   // work counter and (next 256 B) the overflow flag
  if (g_prof.enabled) B2_CUDA(cudaEventRecord(g_prof.begin, s));
  rc = launch_occupancies_fast(seed, resample_lo, n_local, plan, w.counts, w.gcounts, w.cls, w.half_tables,
                               w.overflow, s);
  if (rc) return rc;
  if (host_data) {
    cudaStream_t cs = as_stream(copy_stream);
    B2_CUDA(cudaStreamWaitEvent(cs, ev_entry, 0));
    for (int a = 0; a < n_arrays; ++a)
      B2_CUDA(cudaMemcpyAsync(const_cast<double*>(data[a]), host_data[a], (size_t)n_rows * sizeof(double),
                              cudaMemcpyHostToDevice, cs));
    B2_CUDA(cudaEventRecord(ev_copied, cs));
    B2_CUDA(cudaStreamWaitEvent(s, ev_copied, 0));
    B2_CUDA(cudaEventDestroy(ev_entry));          // released once the work that uses them has run
    B2_CUDA(cudaEventDestroy(ev_copied));
  } else if (ready) {
    B2_CUDA(cudaStreamWaitEvent(s, ready, 0));
  }
  // centring (var / std / Pearson / fit) reads the data: behind the copy, and behind K0
  const double* rows[2];
  rc = prepare_rows(data, n_arrays, n_rows, stat_id, w, rows, s);
  if (rc) return rc;

  const int ms = moment_set_of(stat_id);
  const int nm = ms == kMomSum ? 1 : (ms == kMomPair5 ? 5 : 2);
  const int sms = sm_count();
  // resamples per work item: aim at >= 32 items per CTA (tail balance), whole warps,
  // <= 512 (an item then lasts ~0.5 ms; reloading the cell costs ~3 us).  Measured at the 8-GPU
  // shard of cfg5 (12 500 resamples, K1 ms): 512 -> 35.16, 384 -> 35.20, 256 -> 35.32, 128 -> 35.75
  int64_t want_chunks = ((int64_t)sms * 32 + plan.n_tiles - 1) / plan.n_tiles;
  int64_t chunk = (n_local + want_chunks - 1) / want_chunks;
  const int64_t per_pass = kK1Warps * (B2_K1_PAIR ? 2 : 1);   // resamples a CTA works on at a time
  chunk = ((chunk + per_pass - 1) / per_pass) * per_pass;
  chunk = std::min<int64_t>(std::max<int64_t>(chunk, per_pass), 512);
  const int64_t n_chunks = (n_local + chunk - 1) / chunk;

  K1Params P{};
  P.data[0] = rows[0];
  P.data[1] = rows[1];
  P.plan = plan;
  P.seed = seed;
  P.resample_lo = resample_lo;
  P.n_local = n_local;
  P.counts = w.counts;
  P.cls = w.cls;
  P.rk = make_round_keys(seed);
  P.partial = w.partial;
  P.work_counter = w.work_counter;
  P.chunk = (int32_t)chunk;
  P.n_chunks = (int32_t)n_chunks;
  P.n_items = plan.n_tiles * n_chunks;
  if (P.n_items >= ((int64_t)1 << 31)) return fail(B200BOOT_EINVAL, "too many work items");
  P.smem_stride = plan.n_full > 0 ? plan.tile() : ((plan.rem + 1) & ~(int64_t)1);
  P.stat = stat_id;
  P.center = w.jack->center;
  P.direct_out = plan.n_tiles == 1 ? stat_out : nullptr;     // one cell: K1 finishes the statistic itself
  const size_t smem = kSmemHeader + (size_t)P.smem_stride * 8 * n_arrays;
  const int grid = (int)std::min<int64_t>(sms, P.n_items);
  if (g_prof.enabled) B2_CUDA(cudaEventRecord(g_prof.start, s));
  rc = dispatch_k1(ms, P, smem, grid, s);
  if (rc) return rc;
  if (g_prof.enabled) {
    B2_CUDA(cudaEventRecord(g_prof.stop, s));
    g_prof.valid = true;
  }
  if (P.direct_out) return 0;

  k1_finish_kernel<<<blocks_for(n_local, 256, 1 << 30), 256, 0, s>>>(w.partial, plan.n_tiles, nm, n_local,
                                                                     stat_id, (double)n_rows, w.jack->center,
                                                                     stat_out, w.overflow);
  B2_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

int b200boot_resample_stat_after(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                                 uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                 double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                 void* data_ready_event) {
  return resample_stat_impl(data, n_arrays, n_rows, stat_id, seed, resample_lo, resample_hi, stat_out, ws, ws_bytes,
                            stream, data_ready_event, nullptr, nullptr);
}

int b200boot_resample_stat_host(const double* const* host_data, double* const* data_dev, int n_arrays,
                                int64_t n_rows, int stat_id, uint64_t seed, int64_t resample_lo,
                                int64_t resample_hi, double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                void* copy_stream) {
  if (!host_data || !data_dev || !copy_stream || copy_stream == stream)
    return fail(B200BOOT_EINVAL, "resample_stat_host: host rows, device rows and a separate copy stream are required");
  for (int a = 0; a < n_arrays && a < 2; ++a)
    if (!host_data[a] || !data_dev[a]) return fail(B200BOOT_EINVAL, "resample_stat_host: null pointer");
  if (resample_hi <= resample_lo)
    return fail(B200BOOT_EINVAL, "resample_stat_host: empty resample range (the rows would not be copied)");
  return resample_stat_impl(data_dev, n_arrays, n_rows, stat_id, seed, resample_lo, resample_hi, stat_out, ws,
                            ws_bytes, stream, nullptr, host_data, copy_stream);
}

int b200boot_resample_stat_indexed(const double* const* data, int n_arrays, int64_t n_rows,
                                   int stat_id, const int64_t* indices, int64_t n_sets,
                                   double* stat_out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_moment_stat(stat_id, n_arrays, n_rows);
  if (rc) return rc;
  if (n_sets < 0) return fail(B200BOOT_EINVAL, "n_sets < 0");
  if (n_sets == 0) return 0;
  if (!data || !indices || !stat_out) return fail(B200BOOT_EINVAL, "null pointer");
  cudaStream_t s = as_stream(stream);
  Plan plan = make_plan(n_rows, n_arrays);
  plan.n_tiles = 1;   // partial layout [m][n_sets]
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  ResampleWs w = carve_resample(cv, plan, stat_id, n_sets);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  const double* rows[2];
  rc = prepare_rows(data, n_arrays, n_rows, stat_id, w, rows, s);
  if (rc) return rc;
  const int ms = moment_set_of(stat_id);
  const int nm = ms == kMomSum ? 1 : (ms == kMomPair5 ? 5 : 2);
  switch (ms) {
    case kMomSum: rc = launch_k1i<kMomSum>(rows[0], rows[1], n_rows, indices, n_sets, w.partial, s); break;
    case kMomSum2: rc = launch_k1i<kMomSum2>(rows[0], rows[1], n_rows, indices, n_sets, w.partial, s); break;
    case kMomAB: rc = launch_k1i<kMomAB>(rows[0], rows[1], n_rows, indices, n_sets, w.partial, s); break;
    case kMomXW: rc = launch_k1i<kMomXW>(rows[0], rows[1], n_rows, indices, n_sets, w.partial, s); break;
    default: rc = launch_k1i<kMomPair5>(rows[0], rows[1], n_rows, indices, n_sets, w.partial, s); break;
  }
  if (rc) return rc;
  k1_finish_kernel<<<blocks_for(n_sets, 256, 1 << 30), 256, 0, s>>>(w.partial, 1, nm, n_sets, stat_id,
                                                                    (double)n_rows, w.jack->center, stat_out, nullptr);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_stream_synchronize(void* stream) {
  B2_CUDA(cudaStreamSynchronize(as_stream(stream)));
  return 0;
}

int b200boot_ci_small_supported(int stat_id, int n_arrays, int64_t n_rows, int64_t n_samples, int n_alphas) {
  if (moment_set_of(stat_id) < 0 || stat_outputs(stat_id) != 1 || n_arrays != stat_arrays(stat_id)) return 0;
  if (n_rows < 1 || n_rows * n_arrays > kFusedMaxCellDoubles || make_plan(n_rows, n_arrays).n_tiles != 1) return 0;
  if (n_samples < 1 || n_samples > kSelSmallMaxRows || n_alphas < 1 || n_alphas > kSelMaxAlphas) return 0;
  return 1;
}

}  // extern "C"

namespace {

template <int MS>
int launch_small_fused(const SmallFusedParams& P, int grid, size_t smem, cudaStream_t s) {
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(ci_small_fused_kernel<MS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kFusedMaxCellDoubles * 8));
  }));
  ci_small_fused_kernel<MS><<<grid, kSmallThreads, smem, s>>>(P);
  B2_LAUNCH_CHECK();
  return 0;
}

// the device address of a pinned host block the kernel may store to, or null (then the result
// block comes back by a copy operation); the last answer is remembered per host thread
double* device_view_of_pinned(double* host) {
  static thread_local double* last_host = nullptr;
  static thread_local double* last_dev = nullptr;
  if (host == nullptr) return nullptr;
  if (host == last_host) return last_dev;
  cudaPointerAttributes attr{};
  double* dev = nullptr;
  if (cudaPointerGetAttributes(&attr, host) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
      attr.devicePointer != nullptr)
    dev = static_cast<double*>(attr.devicePointer);
  else
    (void)cudaGetLastError();
  last_host = host;
  last_dev = dev;
  return dev;
}

int enqueue_ci_small(const double* const* data, int n_arrays, int64_t n_rows, int stat_id, uint64_t seed,
                     int64_t n_samples, const double* alphas, int n_alphas, int bca, int need_jack,
                     double* stat_out, double* result_dev, double* result_host, void* ws, size_t ws_bytes,
                     cudaStream_t s, int cmp_op = -1, double cmp_t0 = 0.0, double cmp_t1 = 0.0) {
  int rc = check_moment_stat(stat_id, n_arrays, n_rows);
  if (rc) return rc;
  if (!b200boot_ci_small_supported(stat_id, n_arrays, n_rows, n_samples, cmp_op >= 0 ? 1 : n_alphas))
    return fail(B200BOOT_EUNSUPPORTED, "ci_small: shape outside the small-call path");
  if (!data || (!alphas && n_alphas > 0) || !stat_out || !result_dev) return fail(B200BOOT_EINVAL, "null pointer");
  if (cmp_op >= 0) n_alphas = 0, bca = 0, need_jack = 0;      // pval: the tail only counts
  if (bca) need_jack = 1;
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  JackState* jack_state = cv.take<JackState>(1);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  SmallFusedParams P{};
  P.data[0] = data[0];
  P.data[1] = n_arrays > 1 ? data[1] : nullptr;
  P.n_rows = n_rows;
  P.smem_stride = (n_rows + 1) & ~(int64_t)1;
  P.stat = stat_id;
  P.centre = stat_needs_centering(stat_id) ? 1 : 0;
  P.seed = seed;
  P.rk = make_round_keys(seed);
  P.stat_out = stat_out;
  P.jack_state = jack_state;
  P.jack_out = need_jack ? result_dev + kSmallResultDoubles : nullptr;     // 8 doubles behind the result block
  P.done = reinterpret_cast<int32_t*>(result_dev + kSmallResultDoubles + 8);
  P.res_host = device_view_of_pinned(result_host);
  P.tail.stat = stat_out;
  P.tail.n_samples = n_samples;
  P.tail.jack = P.jack_out;
  for (int a = 0; a < n_alphas; ++a) P.tail.alphas[a] = alphas[a];
  P.tail.n_alphas = n_alphas;
  P.tail.bca = bca ? 1 : 0;
  P.tail.res = result_dev;
  P.tail.staged = n_samples <= kFusedStageMax ? 1 : 0;
  P.tail.n_bins = kBktBins;
  P.tail.ranks = nullptr;
  P.tail.cmp_op = cmp_op;
  P.tail.cmp_t0 = cmp_t0;
  P.tail.cmp_t1 = cmp_t1;
  const size_t cell = (size_t)P.smem_stride * 8 * n_arrays;
  const size_t tail = (P.tail.staged ? tail_stage_bytes(n_samples) : 0) + tail_scratch_bytes();
  const size_t smem = std::max(cell, tail);
  // one warp per resample; the last CTA is the jackknife's when there is more than one
  const int64_t res_ctas = (n_samples + kSmallThreads / 32 - 1) / (kSmallThreads / 32);
  const int grid = (int)std::min<int64_t>(sm_count(), res_ctas + (need_jack ? 1 : 0));
  switch (moment_set_of(stat_id)) {
    case kMomSum: rc = launch_small_fused<kMomSum>(P, grid, smem, s); break;
    case kMomSum2: rc = launch_small_fused<kMomSum2>(P, grid, smem, s); break;
    case kMomAB: rc = launch_small_fused<kMomAB>(P, grid, smem, s); break;
    case kMomXW: rc = launch_small_fused<kMomXW>(P, grid, smem, s); break;
    default: rc = launch_small_fused<kMomPair5>(P, grid, smem, s); break;
  }
  if (rc) return rc;
  if (result_host && P.res_host == nullptr)
    B2_CUDA(cudaMemcpyAsync(result_host, result_dev, kSmallResultDoubles * sizeof(double), cudaMemcpyDeviceToHost, s));
  return 0;
}

}  // namespace

extern "C" {

int b200boot_ci_small(const double* const* data, int n_arrays, int64_t n_rows, int stat_id, uint64_t seed,
                      int64_t n_samples, const double* alphas, int n_alphas, int bca, int need_jack,
                      double* stat_out, double* result_dev, double* result_host, void* ws, size_t ws_bytes,
                      void* stream) {
  return enqueue_ci_small(data, n_arrays, n_rows, stat_id, seed, n_samples, alphas, n_alphas, bca, need_jack,
                          stat_out, result_dev, result_host, ws, ws_bytes, as_stream(stream));
}

int b200boot_ci_tail_supported(int64_t n_samples, int n_alphas) {
  return n_samples >= 1 && n_samples <= kTailMaxValues && n_alphas >= 1 && n_alphas <= kSelMaxAlphas ? 1 : 0;
}

}  // extern "C"

namespace {
int enqueue_ci_tail(const double* stat, int64_t n_samples, const double* jack, const double* alphas, int n_alphas,
                    int bca, double* result_dev, double* result_host, cudaStream_t s, const double* nan_probe) {
  if (!stat || !alphas || !result_dev || (bca && !jack)) return fail(B200BOOT_EINVAL, "ci_tail: null pointer");
  if (!b200boot_ci_tail_supported(n_samples, n_alphas))
    return fail(B200BOOT_EUNSUPPORTED, "ci_tail: 1 <= n_samples <= 2^21, 1 <= n_alphas <= 8");
  SmallTailParams P{};
  P.stat = stat;
  P.n_samples = n_samples;
  P.jack = jack;
  for (int a = 0; a < n_alphas; ++a) P.alphas[a] = alphas[a];
  P.n_alphas = n_alphas;
  P.bca = bca ? 1 : 0;
  P.res = result_dev;
  P.staged = n_samples <= kFusedStageMax ? 1 : 0;
  P.n_bins = P.staged ? kBktBins : kBktBinsWide;
  P.ranks = nullptr;
  P.cmp_op = -1;
  const size_t smem = (P.staged ? tail_stage_bytes(n_samples) : 0) + tail_scratch_bytes(P.n_bins);
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(ci_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)std::max(tail_stage_bytes(kFusedStageMax) + tail_scratch_bytes(kBktBins),
                                              tail_scratch_bytes(kBktBinsWide)));
  }));
  double* host_view = device_view_of_pinned(result_host);
  ci_tail_kernel<<<1, kSmallThreads, smem, s>>>(P, host_view, nan_probe);
  B2_LAUNCH_CHECK();
  if (result_host && host_view == nullptr)
    B2_CUDA(cudaMemcpyAsync(result_host, result_dev, kSmallResultDoubles * sizeof(double), cudaMemcpyDeviceToHost, s));
  return 0;
}
}  // namespace

extern "C" {

int b200boot_ci_tail(const double* stat, int64_t n_samples, const double* jack, const double* alphas, int n_alphas,
                     int bca, double* result_dev, double* result_host, void* stream) {
  return enqueue_ci_tail(stat, n_samples, jack, alphas, n_alphas, bca, result_dev, result_host, as_stream(stream),
                         nullptr);
}

int b200boot_ci_small_host(const double* const* host_data, double* const* staging, double* const* data_dev,
                           int n_arrays, int64_t n_rows, int stat_id, uint64_t seed, int64_t n_samples,
                           const double* alphas, int n_alphas, int bca, int need_jack, double* stat_out,
                           double* result_dev, double* result_host, void* ws, size_t ws_bytes, void* stream) {
  if (!host_data || !staging || !data_dev || n_arrays < 1 || n_arrays > 2 || n_rows < 1)
    return fail(B200BOOT_EINVAL, "ci_small_host: bad arguments");
  if (!result_host) return fail(B200BOOT_EINVAL, "ci_small_host: result_host is required");
  cudaStream_t s = as_stream(stream);
  const size_t bytes = (size_t)n_rows * sizeof(double);
  for (int a = 0; a < n_arrays; ++a) {
    if (!host_data[a] || !staging[a] || !data_dev[a]) return fail(B200BOOT_EINVAL, "ci_small_host: null pointer");
    memcpy(staging[a], host_data[a], bytes);           // pageable -> pinned: the copy below is a plain DMA
    B2_CUDA(cudaMemcpyAsync(data_dev[a], staging[a], bytes, cudaMemcpyHostToDevice, s));
  }
  const int rc = enqueue_ci_small(data_dev, n_arrays, n_rows, stat_id, seed, n_samples, alphas, n_alphas, bca,
                                  need_jack, stat_out, result_dev, result_host, ws, ws_bytes, s);
  if (rc) return rc;
  B2_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int b200boot_pval_small(const double* const* data, const double* const* host_data, double* const* staging,
                        int n_arrays, int64_t n_rows, int stat_id, uint64_t seed, int64_t n_samples, int cmp_op,
                        double t0, double t1, double* stat_out, double* result_dev, double* result_host, void* ws,
                        size_t ws_bytes, void* stream) {
  if (cmp_op < B200BOOT_CMP_LT || cmp_op > B200BOOT_CMP_BETWEEN) return fail(B200BOOT_EINVAL, "pval_small: bad criterion");
  if (!data || !result_host || n_arrays < 1 || n_arrays > 2 || n_rows < 1)
    return fail(B200BOOT_EINVAL, "pval_small: bad arguments");
  cudaStream_t s = as_stream(stream);
  if (host_data) {                                  // rows still on the host: staged like b200boot_ci_small_host
    if (!staging) return fail(B200BOOT_EINVAL, "pval_small: staging buffers are required for host rows");
    const size_t bytes = (size_t)n_rows * sizeof(double);
    for (int a = 0; a < n_arrays; ++a) {
      if (!host_data[a] || !staging[a] || !data[a]) return fail(B200BOOT_EINVAL, "pval_small: null pointer");
      memcpy(staging[a], host_data[a], bytes);
      B2_CUDA(cudaMemcpyAsync(const_cast<double*>(data[a]), staging[a], bytes, cudaMemcpyHostToDevice, s));
    }
  }
  const int rc = enqueue_ci_small(data, n_arrays, n_rows, stat_id, seed, n_samples, nullptr, 0, 0, 0, stat_out,
                                  result_dev, result_host, ws, ws_bytes, s, cmp_op, t0, t1);
  if (rc) return rc;
  B2_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int b200boot_tile_counts(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample_lo,
                         int64_t resample_hi, int32_t* counts, int32_t* class_counts, void* stream) {
  const int64_t n_local = resample_hi - resample_lo;
  if (n_rows < 1 || n_arrays < 1 || n_local < 0 || !counts) return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_local == 0) return 0;
  const Plan plan = make_plan(n_rows, n_arrays);
  return launch_occupancies(seed, resample_lo, n_local, plan, counts, class_counts, as_stream(stream));
}

size_t b200boot_tile_counts_packed_ws_bytes(int64_t n_rows, int n_arrays, int64_t n_local) {
  if (n_rows < 1 || n_arrays < 1 || n_local < 0) return 0;
  const Plan plan = make_plan(n_rows, n_arrays);
  return 4096 + pad256((size_t)n_tile_groups(plan) * std::max<int64_t>(n_local, 1) * 4) +
         pad256((size_t)kHalfTableSize * sizeof(HalfConsts)) + pad256((size_t)kHalfTableSize * sizeof(double));
}

int b200boot_tile_counts_packed(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample_lo,
                                int64_t resample_hi, int32_t* counts, uint16_t* class_counts,
                                void* ws, size_t ws_bytes, void* stream) {
  const int64_t n_local = resample_hi - resample_lo;
  if (n_rows < 1 || n_arrays < 1 || n_local < 0 || !counts) return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_local == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  const Plan plan = make_plan(n_rows, n_arrays);
  Carver cv(ws, ws_bytes);
  int32_t* overflow = cv.take<int32_t>(64);
  int32_t* gcounts = cv.take<int32_t>((size_t)n_tile_groups(plan) * n_local);
  const HalfTables table = take_half_tables(cv);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  cudaStream_t s = as_stream(stream);
  B2_CUDA(cudaMemsetAsync(overflow, 0, 256, s));
  return launch_occupancies_fast(seed, resample_lo, n_local, plan, counts, gcounts, class_counts, table, overflow, s);
}

int b200boot_fill_indices(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample_lo,
                          int64_t resample_hi, int64_t* out, void* ws, size_t ws_bytes,
                          void* stream) {
  const int64_t n_local = resample_hi - resample_lo;
  if (n_rows < 1 || n_rows >= ((int64_t)1 << 31) || n_arrays < 1 || n_local < 0 || !out)
    return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_local == 0) return 0;
  cudaStream_t s = as_stream(stream);
  const Plan plan = make_plan(n_rows, n_arrays);
  int32_t* counts = nullptr;
  int32_t* cls = nullptr;
  if (plan.n_tiles > 1) {
    if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
      return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
    Carver cv(ws, ws_bytes);
    counts = cv.take<int32_t>((size_t)plan.n_tiles * n_local);
    cls = cv.take<int32_t>((size_t)plan.n_full * n_local * kClasses);
    if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
    int rc = launch_occupancies(seed, resample_lo, n_local, plan, counts, cls, s);
    if (rc) return rc;
  }
  k2_fill_indices_kernel<<<blocks_for(n_local * plan.n_tiles * 32, 256, 148 * 8), 256, 0, s>>>(plan, seed, resample_lo,
                                                                               n_local, counts, cls, out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_fill_moving_block(uint64_t seed, int64_t n_rows, int64_t block_length, int wrap,
                               int64_t resample_lo, int64_t resample_hi, int64_t* out, void* stream) {
  const int64_t n_local = resample_hi - resample_lo;
  if (n_rows < 1 || n_rows >= ((int64_t)1 << 31) || block_length < 1 || n_local < 0 || !out)
    return fail(B200BOOT_EINVAL, "bad arguments");
  // numpy draws starts from [0, N) when wrapping, else from [0, N - L)  (bootstrap.py:777-783)
  const int64_t start_bound = wrap ? n_rows : n_rows - block_length;
  if (start_bound < 1) return fail(B200BOOT_EINVAL, "low >= high: no admissible block start");
  if (n_local == 0) return 0;
  k2_moving_block_kernel<<<blocks_for(n_local * 32, 256, 148 * 8), 256, 0, as_stream(stream)>>>(
      seed, n_rows, block_length, start_bound, wrap ? 1 : 0, resample_lo, n_local, out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_resample_stat_moving_block(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                                        uint64_t seed, int64_t block_length, int wrap, int64_t resample_lo,
                                        int64_t resample_hi, double* stat_out, void* ws, size_t ws_bytes,
                                        void* stream) {
  int rc = check_moment_stat(stat_id, n_arrays, n_rows);
  if (rc) return rc;
  const int64_t n_local = resample_hi - resample_lo;
  if (n_local < 0 || resample_lo < 0 || block_length < 1) return fail(B200BOOT_EINVAL, "bad arguments");
  const int64_t start_bound = wrap ? n_rows : n_rows - block_length;
  if (start_bound < 1) return fail(B200BOOT_EINVAL, "low >= high: no admissible block start");
  if (n_local == 0) return 0;
  if (!data || !stat_out) return fail(B200BOOT_EINVAL, "null pointer");
  cudaStream_t s = as_stream(stream);
  const Plan plan = make_plan(n_rows, n_arrays);
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  ResampleWs w = carve_resample(cv, plan, stat_id, 0);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  const double* rows[2];
  rc = prepare_rows(data, n_arrays, n_rows, stat_id, w, rows, s);
  if (rc) return rc;
  const int grid = blocks_for(n_local * 32, 256, (int64_t)sm_count() * 8);
#define B2_K1B(MS)                                                                                              \
  k1b_moving_block_kernel<MS><<<grid, 256, 0, s>>>(rows[0], rows[1], n_rows, block_length, start_bound, wrap ? 1 : 0, \
                                                   seed, resample_lo, n_local, stat_id, w.jack->center, stat_out)
  switch (moment_set_of(stat_id)) {
    case kMomSum: B2_K1B(kMomSum); break;
    case kMomSum2: B2_K1B(kMomSum2); break;
    case kMomAB: B2_K1B(kMomAB); break;
    case kMomXW: B2_K1B(kMomXW); break;
    default: B2_K1B(kMomPair5); break;
  }
#undef B2_K1B
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_fill_subsample(uint64_t seed, int64_t n_rows, int64_t size, int64_t sample_lo,
                            int64_t sample_hi, int64_t* out, void* ws, size_t ws_bytes, void* stream) {
  const int64_t n_local = sample_hi - sample_lo;
  if (n_rows < 1 || n_rows >= ((int64_t)1 << 31) || size < 1 || size > n_rows || n_local < 0 || !out)
    return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_local == 0) return 0;
  if (n_local > 65535) return fail(B200BOOT_EINVAL, "at most 65535 samples per call");
  cudaStream_t s = as_stream(stream);
  int64_t n_pad = 1;
  while (n_pad < n_rows) n_pad <<= 1;
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  uint64_t* keys = cv.take<uint64_t>((size_t)n_pad * n_local);
  uint32_t* idx = cv.take<uint32_t>((size_t)n_pad * n_local);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  k2_shuffle_keys_kernel<<<blocks_for(((n_pad + 1) / 2) * n_local, 256, 148 * 16), 256, 0, s>>>(
      seed, n_rows, n_pad, sample_lo, n_local, keys, idx);
  B2_LAUNCH_CHECK();
  int rc = psort_pairs(keys, idx, n_pad, (unsigned)n_local, s);
  if (rc) return rc;
  k2_subsample_store_kernel<<<blocks_for(size * n_local, 256, 148 * 16), 256, 0, s>>>(idx, n_pad, size,
                                                                                      n_local, out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_gather(const double* data, int64_t n_rows, const int64_t* idx, int64_t n_idx,
                    double* out, void* stream) {
  if (!data || !idx || !out || n_rows < 1 || n_idx < 0) return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_idx == 0) return 0;
  gather_kernel<<<blocks_for(n_idx, 256, 148 * 16), 256, 0, as_stream(stream)>>>(data, idx, n_idx, out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_jackknife(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                       double* out, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_moment_stat(stat_id, n_arrays, n_rows);
  if (rc) return rc;
  if (!data || !out) return fail(B200BOOT_EINVAL, "null pointer");
  cudaStream_t s = as_stream(stream);
  const Plan plan = make_plan(n_rows, n_arrays);
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  ResampleWs w = carve_resample(cv, plan, stat_id, 0);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  const double* rows[2];
  rc = prepare_rows(data, n_arrays, n_rows, stat_id, w, rows, s);
  if (rc) return rc;
  switch (moment_set_of(stat_id)) {
    case kMomSum: return jack_passes<kMomSum>(rows, n_rows, stat_id, w, out, s);
    case kMomSum2: return jack_passes<kMomSum2>(rows, n_rows, stat_id, w, out, s);
    case kMomAB: return jack_passes<kMomAB>(rows, n_rows, stat_id, w, out, s);
    case kMomXW: return jack_passes<kMomXW>(rows, n_rows, stat_id, w, out, s);
    default:
      for (int o = 0; o < stat_outputs(stat_id); ++o) {   // out[o][8]
        rc = jack_passes<kMomPair5>(rows, n_rows, stat_component(stat_id, o), w, out + 8 * o, s);
        if (rc) return rc;
      }
      return 0;
  }
}

int b200boot_count(const double* stat, int64_t n_local, int64_t n_cols, int cmp_op,
                   const double* t0, const double* t1, int64_t* counts, void* stream) {
  if (!stat || !t0 || !counts || n_local < 0 || n_cols < 1 || n_cols > 65535 * 64LL)
    return fail(B200BOOT_EINVAL, "bad arguments");
  if (cmp_op < 0 || cmp_op > B200BOOT_CMP_BETWEEN) return fail(B200BOOT_EINVAL, "bad comparison op");
  if (cmp_op == B200BOOT_CMP_BETWEEN && !t1) return fail(B200BOOT_EINVAL, "BETWEEN needs t1");
  cudaStream_t s = as_stream(stream);
  B2_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_cols * 8, s));
  if (n_local == 0) return 0;
  for (int64_t d0 = 0; d0 < n_cols; d0 += 65535) {
    const int64_t nd = std::min<int64_t>(65535, n_cols - d0);
    dim3 grid(blocks_for(n_local, 256 * 8, 64), (unsigned)nd);
    k4_count_kernel<<<grid, 256, 0, s>>>(stat + d0 * n_local, n_local, cmp_op, t0 + d0, t1 ? t1 + d0 : nullptr,
                                         reinterpret_cast<unsigned long long*>(counts + d0));
    B2_LAUNCH_CHECK();
  }
  return 0;
}

// ---- K5 -----------------------------------------------------------------------
static int select_state(int n_alphas, int64_t n_cols, void* ws, size_t ws_bytes, SelectState* st) {
  if (n_alphas < 1 || n_cols < 1) return fail(B200BOOT_EINVAL, "bad select shape");
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  *st = select_carve(cv, n_alphas, n_cols);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  return 0;
}

int b200boot_select_begin(const int64_t* ranks, int n_alphas, int64_t n_cols, void* ws,
                          size_t ws_bytes, void* stream) {
  SelectState st;
  int rc = select_state(n_alphas, n_cols, ws, ws_bytes, &st);
  if (rc) return rc;
  if (!ranks) return fail(B200BOOT_EINVAL, "null ranks");
  const int64_t ad = (int64_t)n_alphas * n_cols;
  k5_begin_kernel<<<blocks_for(ad * 256, 256, 148 * 8), 256, 0, as_stream(stream)>>>(ranks, ad, st);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_select_hist(const double* stat, int64_t n_local, int64_t n_cols, int n_alphas,
                         int pass, void* ws, size_t ws_bytes, void* stream) {
  SelectState st;
  int rc = select_state(n_alphas, n_cols, ws, ws_bytes, &st);
  if (rc) return rc;
  if (pass < 0 || pass > 7 || n_local < 0) return fail(B200BOOT_EINVAL, "bad pass / n_local");
  if (n_local == 0) return 0;
  if (!stat) return fail(B200BOOT_EINVAL, "null stat");
  for (int a_lo = 0; a_lo < n_alphas; a_lo += kSelMaxAlphas) {
    const int a_cnt = std::min(kSelMaxAlphas, n_alphas - a_lo);
    for (int64_t d0 = 0; d0 < n_cols; d0 += 65535) {
      const int64_t nd = std::min<int64_t>(65535, n_cols - d0);
      dim3 grid(blocks_for(n_local, 256 * 8, 64), (unsigned)nd);
      SelectState sub = st;
      sub.prefix += d0;
      sub.krem += d0;
      sub.hist += d0 * 256;
      k5_hist_kernel<<<grid, 256, 0, as_stream(stream)>>>(stat + d0 * n_local, n_local, n_cols, a_lo, a_cnt,
                                                          pass, sub);
      B2_LAUNCH_CHECK();
    }
  }
  return 0;
}

int b200boot_select_pick(int n_alphas, int64_t n_cols, int pass, void* ws, size_t ws_bytes,
                         void* stream) {
  SelectState st;
  int rc = select_state(n_alphas, n_cols, ws, ws_bytes, &st);
  if (rc) return rc;
  if (pass < 0 || pass > 7) return fail(B200BOOT_EINVAL, "bad pass");
  const int64_t ad = (int64_t)n_alphas * n_cols;
  k5_pick_kernel<<<blocks_for(ad, 256, 1 << 30), 256, 0, as_stream(stream)>>>(ad, pass, st);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_select_end(int n_alphas, int64_t n_cols, double* out, void* ws, size_t ws_bytes,
                        void* stream) {
  SelectState st;
  int rc = select_state(n_alphas, n_cols, ws, ws_bytes, &st);
  if (rc) return rc;
  if (!out) return fail(B200BOOT_EINVAL, "null out");
  const int64_t ad = (int64_t)n_alphas * n_cols;
  k5_end_kernel<<<blocks_for(ad, 256, 1 << 30), 256, 0, as_stream(stream)>>>(ad, st, out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_select_hist_buffer(int n_alphas, int64_t n_cols, void* ws, size_t ws_bytes,
                                int64_t** hist, int64_t* n_elems) {
  SelectState st;
  int rc = select_state(n_alphas, n_cols, ws, ws_bytes, &st);
  if (rc) return rc;
  if (!hist || !n_elems) return fail(B200BOOT_EINVAL, "null out pointer");
  *hist = st.hist;
  *n_elems = (int64_t)n_alphas * n_cols * 256;
  return 0;
}

int b200boot_select(const double* stat, int64_t n_local, int64_t n_cols, const int64_t* ranks,
                    int n_alphas, double* out, void* ws, size_t ws_bytes, void* stream) {
  if (stat && ranks && out && n_alphas >= 1 && n_cols >= 1 && n_cols <= kSelSmallMaxCols && n_local >= 1 &&
      n_local <= kSelSmallMaxRows) {
    // short columns: the whole select in one launch per group of ranks
    static PerDeviceOnce configured;
    B2_CUDA(configured.run([] {
      return cudaFuncSetAttribute(k5_select_bucket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(tail_stage_bytes(kFusedStageMax) + tail_scratch_bytes(kBktBins)));
    }));
    for (int a0 = 0; a0 < n_alphas; a0 += kSelMaxAlphas) {
      const int ac = std::min(kSelMaxAlphas, n_alphas - a0);
      if (n_local >= 2048) {
        // bucket select: one CTA per column at a time, the column staged when it fits (cfg4, 10^4
        // columns of 10^4 values: 1.07 ms staged, 1.2 ms unstaged, 2.86 ms for the radix descent)
        const int staged = n_local <= kFusedStageMax ? 1 : 0;
        const size_t smem = (staged ? tail_stage_bytes(n_local) : 0) + tail_scratch_bytes(kBktBins);
        const int per_sm = smem <= 100 * 1024 ? 2 : 1;
        k5_select_bucket_kernel<<<(unsigned)std::min<int64_t>(n_cols, (int64_t)sm_count() * per_sm), kSmallThreads, smem,
                                  as_stream(stream)>>>(stat, n_local, n_cols, ranks, a0, ac, staged, out);
        B2_LAUNCH_CHECK();
        continue;
      }
      // 8 warps walk the bins of up to 8 selects; more threads only pay off for longer columns
      const int threads = n_local <= 2048 ? 256 : (n_local <= 8192 ? 512 : kSelSmallThreads);
      k5_select_small_kernel<<<(unsigned)n_cols, threads, 0, as_stream(stream)>>>(
          stat, n_local, n_cols, ranks, a0, ac, out);
      B2_LAUNCH_CHECK();
    }
    return 0;
  }
  int rc = b200boot_select_begin(ranks, n_alphas, n_cols, ws, ws_bytes, stream);
  for (int pass = 0; pass < 8 && !rc; ++pass) {
    rc = b200boot_select_hist(stat, n_local, n_cols, n_alphas, pass, ws, ws_bytes, stream);
    if (!rc) rc = b200boot_select_pick(n_alphas, n_cols, pass, ws, ws_bytes, stream);
  }
  if (!rc) rc = b200boot_select_end(n_alphas, n_cols, out, ws, ws_bytes, stream);
  return rc;
}

int b200boot_sort_columns(double* stat, int64_t n, int64_t n_cols, void* ws, size_t ws_bytes,
                          void* stream) {
  if (!stat || n < 0 || n_cols < 1 || n_cols > 65535) return fail(B200BOOT_EINVAL, "bad arguments");
  if (n <= 1) return 0;
  cudaStream_t s = as_stream(stream);
  int64_t n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  uint64_t* keys = cv.take<uint64_t>((size_t)n_pad * n_cols);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  const dim3 gflat(blocks_for(n_pad, 256 * 4, 4096), (unsigned)n_cols);
  sort_load_kernel<<<gflat, 256, 0, s>>>(stat, n, n_pad, keys);
  B2_LAUNCH_CHECK();
  const int64_t seg = std::min<int64_t>(2048, n_pad);
  const dim3 glocal((unsigned)(n_pad / seg), (unsigned)n_cols);
  sort_local_kernel<<<glocal, 1024, 0, s>>>(keys, n_pad, 2, seg, 1);
  B2_LAUNCH_CHECK();
  for (int64_t k = seg * 2; k <= n_pad; k <<= 1) {
    for (int64_t j = k >> 1; j >= 2048; j >>= 1) {
      sort_step_kernel<<<gflat, 256, 0, s>>>(keys, n_pad, j, k);
      B2_LAUNCH_CHECK();
    }
    sort_local_kernel<<<glocal, 1024, 0, s>>>(keys, n_pad, k, k, 1024);
    B2_LAUNCH_CHECK();
  }
  sort_store_kernel<<<gflat, 256, 0, s>>>(keys, n, n_pad, stat);
  B2_LAUNCH_CHECK();
  return 0;
}

// ---- median and quantiles (K1m / K3m) ------------------------------------------------
namespace {
// full-sample and leave-one-out rules from the ABI struct (NULL: np.median)
int rank_rules(const b200boot_rank_rule* rule, int64_t n_rows, RankRule* full, RankRule* loo) {
  if (!rule) {
    *full = median_rule(n_rows);
    *loo = median_rule(n_rows > 1 ? n_rows - 1 : 1);
    return 0;
  }
  *full = RankRule{rule->k_lo, rule->k_hi, rule->gamma, rule->mode};
  *loo = RankRule{rule->loo_k_lo, rule->loo_k_hi, rule->loo_gamma, rule->mode};
  const int64_t m = n_rows > 1 ? n_rows - 1 : 1;
  const bool ok = (rule->mode == 0 || rule->mode == 1) && full->k_lo >= 0 && full->k_lo <= full->k_hi &&
                  full->k_hi <= full->k_lo + 1 && full->k_hi < n_rows && loo->k_lo >= 0 &&
                  loo->k_lo <= loo->k_hi && loo->k_hi <= loo->k_lo + 1 && loo->k_hi < m &&
                  full->gamma >= 0.0 && full->gamma <= 1.0 && loo->gamma >= 0.0 && loo->gamma <= 1.0;
  return ok ? 0 : fail(B200BOOT_EINVAL, "rank rule: ranks outside the sample or weight outside [0, 1]");
}

// K1m1: one resident column (csrc/k1mr_rows.cuh).  Sorts the column once ((key, row) pairs ->
// sorted values + rank table), then a warp (N <= 2048) or a CTA per resample.
int median_single_column(const double* data, int64_t n_rows, int64_t ld, uint64_t seed, int64_t resample_lo,
                         int64_t n_sets, const int64_t* indices, double* stat_out, void* ws, size_t ws_bytes,
                         cudaStream_t s, const RankRule& full, const RankRule& loo, double* jack_out,
                         const double** sorted_out = nullptr) {
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  int64_t n_pad = 1;
  while (n_pad < n_rows) n_pad <<= 1;
  Carver cv(ws, ws_bytes);
  uint64_t* keys = cv.take<uint64_t>((size_t)n_pad);
  uint32_t* idx = cv.take<uint32_t>((size_t)n_pad);
  double* sorted = cv.take<double>((size_t)n_rows);
  int32_t* rank = cv.take<int32_t>((size_t)n_rows);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  if (sorted_out) *sorted_out = sorted;
  psort_load_kernel<<<blocks_for(n_pad, 256, 148 * 32), 256, 0, s>>>(data, n_rows, 1, ld, n_pad, keys, idx);
  B2_LAUNCH_CHECK();
  int rc = psort_pairs(keys, idx, n_pad, 1u, s);
  if (rc) return rc;
  rows_rank_kernel<<<blocks_for(n_rows, 256, 148 * 32), 256, 0, s>>>(keys, idx, n_rows, n_pad, 1, 0, 1, rank, sorted);
  B2_LAUNCH_CHECK();
  if (jack_out) {
    k3m_jackknife_median_kernel<<<1, 256, 0, s>>>(sorted, n_rows, full, loo, jack_out);
    B2_LAUNCH_CHECK();
  }
  M1Params P{};
  P.rank = rank;
  P.sorted = sorted;
  P.indices = indices;
  P.n_rows = n_rows;
  P.n_sets = n_sets;
  P.seed = seed;
  P.resample_lo = resample_lo;
  P.rule = full;
  P.out = stat_out;
  P.rk = make_round_keys(seed);
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    cudaError_t e = cudaFuncSetAttribute(k1m1_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)m1_smem_bytes(2048, 32));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k1m1_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)m1_smem_bytes(kMedMaxRows, 256));
    return e;
  }));
  if (n_rows <= 2048) {
    const size_t smem = m1_smem_bytes(n_rows, 32);
    const int grid = (int)std::min<int64_t>((n_sets + 7) / 8, (int64_t)sm_count() * 8);
    k1m1_kernel<32><<<grid, 256, smem, s>>>(P);
  } else {
    const size_t smem = m1_smem_bytes(n_rows, 256);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / smem));
    const int grid = (int)std::min<int64_t>(n_sets, (int64_t)sm_count() * per_sm);
    k1m1_kernel<256><<<grid, 256, smem, s>>>(P);
  }
  B2_LAUNCH_CHECK();
  return 0;
}
}  // namespace

size_t b200boot_rank_columns_ws_bytes(int64_t n_rows, int64_t n_cols) {
  if (n_rows < 1 || n_cols < 1) return 0;
  int64_t n_pad = 1;
  while (n_pad < n_rows) n_pad <<= 1;
  const int64_t nd = std::min<int64_t>(n_cols, 8);             // columns sorted per pass
  return 4096 + pad256((size_t)n_pad * nd * 8) + pad256((size_t)n_pad * nd * 4);
}

int b200boot_rank_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, double* sorted_out,
                          int32_t* rank_out, void* ws, size_t ws_bytes, void* stream) {
  return b200boot_rank_columns_ld(data, n_rows, n_cols, ld, sorted_out, rank_out, n_cols, ws, ws_bytes, stream);
}

int b200boot_rank_columns_ld(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, double* sorted_out,
                             int32_t* rank_out, int64_t rank_ld, void* ws, size_t ws_bytes, void* stream) {
  if (!data || !sorted_out || !rank_out || n_rows < 1 || n_cols < 1 || ld < n_cols || rank_ld < n_cols ||
      n_rows >= ((int64_t)1 << 31))
    return fail(B200BOOT_EINVAL, "rank_columns: bad arguments");
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  cudaStream_t s = as_stream(stream);
  int64_t n_pad = 1;
  while (n_pad < n_rows) n_pad <<= 1;
  // as many columns per pass as the workspace holds (12 bytes per padded entry)
  const size_t per_col = pad256((size_t)n_pad * 8) + pad256((size_t)n_pad * 4);
  const int64_t fit = ws_bytes > 4096 ? (int64_t)((ws_bytes - 4096) / per_col) : 0;
  const int64_t step = std::min<int64_t>(std::min<int64_t>(fit, n_cols), 65535);
  if (step < 1) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", (size_t)4096 + per_col);
  for (int64_t d0 = 0; d0 < n_cols; d0 += step) {
    const int64_t nd = std::min<int64_t>(step, n_cols - d0);
    Carver cv(ws, ws_bytes);
    uint64_t* keys = cv.take<uint64_t>((size_t)n_pad * nd);
    uint32_t* idx = cv.take<uint32_t>((size_t)n_pad * nd);
    if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
    psort_load_kernel<<<blocks_for(n_pad * nd, 256, 148 * 32), 256, 0, s>>>(data + d0, n_rows, nd, ld, n_pad, keys, idx);
    B2_LAUNCH_CHECK();
    int rc = psort_pairs(keys, idx, n_pad, (unsigned)nd, s);
    if (rc) return rc;
    rows_rank_kernel<<<blocks_for(n_rows * nd, 256, 148 * 32), 256, 0, s>>>(keys, idx, n_rows, n_pad, rank_ld, d0, nd,
                                                                            rank_out, sorted_out);
    B2_LAUNCH_CHECK();
  }
  return 0;
}

int b200boot_nan_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t* flags,
                         void* stream) {
  if (!data || !flags || n_rows < 1 || n_cols < 1 || ld < n_cols) return fail(B200BOOT_EINVAL, "nan_columns: bad arguments");
  cudaStream_t s = as_stream(stream);
  B2_CUDA(cudaMemsetAsync(flags, 0, (size_t)(n_cols + 1) * sizeof(int32_t), s));
  // enough CTAs over the rows to fill the device when there are few columns
  const int64_t col_ctas = (n_cols + 31) / 32;
  const int64_t row_ctas = std::max<int64_t>(1, std::min<int64_t>((n_rows + 255) / 256, (4 * (int64_t)sm_count()) / col_ctas));
  nan_columns_kernel<<<dim3((unsigned)col_ctas, (unsigned)std::min<int64_t>(row_ctas, 65535)), 256, 0, s>>>(
      data, n_rows, n_cols, ld, flags);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_nan_fixup(const double* src, int64_t stride, int64_t n_rows, const int64_t* indices, int64_t n_sets,
                       double* out, void* stream) {
  if (!src || !indices || !out || stride < 1 || n_rows < 1 || n_sets < 0)
    return fail(B200BOOT_EINVAL, "nan_fixup: bad arguments");
  if (n_sets == 0) return 0;
  nan_fixup_kernel<<<blocks_for(n_sets * 32, 256, 148 * 8), 256, 0, as_stream(stream)>>>(src, stride, n_rows, indices,
                                                                                        n_sets, out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_nan_poison_rows(double* table, int64_t n_cols, int width, const int32_t* flags, void* stream) {
  if (!table || !flags || n_cols < 1 || width < 1) return fail(B200BOOT_EINVAL, "nan_poison_rows: bad arguments");
  nan_poison_rows_kernel<<<blocks_for(n_cols * width, 256, 1 << 20), 256, 0, as_stream(stream)>>>(table, n_cols, width,
                                                                                                 flags);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_median_rows_supported(int64_t n_rows) { return n_rows >= 1 && n_rows <= kRowsMaxRows ? 1 : 0; }

int b200boot_median_rows_indexed(const double* sorted_vals, const int32_t* rank, int64_t n_rows, int64_t n_cols,
                                 const int64_t* indices, int64_t n_sets, double* stat_out, int64_t out_ld,
                                 void* stream, const b200boot_rank_rule* rule) {
  return b200boot_median_rows_indexed_ld(sorted_vals, rank, n_cols, n_rows, n_cols, indices, n_sets, stat_out, out_ld,
                                         stream, rule);
}

int b200boot_median_rows_indexed_ld(const double* sorted_vals, const int32_t* rank, int64_t rank_ld, int64_t n_rows,
                                    int64_t n_cols, const int64_t* indices, int64_t n_sets, double* stat_out,
                                    int64_t out_ld, void* stream, const b200boot_rank_rule* rule) {
  if (!sorted_vals || !rank || !indices || !stat_out || n_cols < 1 || rank_ld < n_cols || n_sets < 0 || out_ld < n_sets)
    return fail(B200BOOT_EINVAL, "median_rows_indexed: bad arguments");
  if (!b200boot_median_rows_supported(n_rows))
    return fail(B200BOOT_EUNSUPPORTED, "median_rows_indexed: 1 <= n_rows <= 2^24");
  if (n_sets == 0) return 0;
  RowsParams P{};
  RankRule loo;
  int rc = rank_rules(rule, n_rows, &P.rule, &loo);
  if (rc) return rc;
  int bits = 0;
  while (((int64_t)1 << bits) < n_rows) ++bits;
  P.shift = (bits + 1) / 2;
  P.hi_bins = (int32_t)((n_rows + ((int64_t)1 << P.shift) - 1) >> P.shift);
  P.bins = (std::max<int32_t>(P.hi_bins, (int32_t)1 << P.shift) + 31) & ~31;
  P.rank = rank;
  P.sorted = sorted_vals;
  P.indices = indices;
  P.n_rows = n_rows;
  P.n_cols = n_cols;
  P.n_sets = n_sets;
  P.out = stat_out;
  P.out_ld = out_ld;
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(k1mr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)rows_smem_bytes());
  }));
  cudaStream_t s = as_stream(stream);
  const int64_t groups = (n_cols + kRowsGroup - 1) / kRowsGroup;
  for (int64_t g0 = 0; g0 < groups; g0 += 65535) {
    RowsParams Q = P;
    const int64_t ng = std::min<int64_t>(65535, groups - g0);
    Q.rank = P.rank + g0 * kRowsGroup;
    Q.sorted = P.sorted + g0 * kRowsGroup * n_rows;
    Q.out = P.out + g0 * kRowsGroup * out_ld;
    Q.n_cols = rank_ld;                 // row stride of the rank table
    k1mr_select_kernel<<<dim3((unsigned)n_sets, (unsigned)ng), kRowsThreads, rows_smem_bytes(P.bins), s>>>(Q, n_cols - g0 * kRowsGroup);
    B2_LAUNCH_CHECK();
  }
  return 0;
}

int b200boot_resample_median_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                     uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                     double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                     const b200boot_rank_rule* rule, double* jack_out) {
  RankRule full, loo;
  if (int rc = rank_rules(rule, n_rows, &full, &loo)) return rc;
  // two or more resident columns: prefix counts on the tensor cores + short walks (K1mg; the
  // streamed form for 1024 < N <= 28 672 while its images fit 3 GB: tools/mg_crossover.py, n = 1e4:
  // N = 1e4: D = 2 0.47 vs 0.77 ms, D = 32 1.06 vs 8.3, D = 256 5.4 vs 36.9; N = 28 000, D = 32: 3.4 vs 45.6)
  static const bool use_mg = getenv("B200BOOT_K1MG") == nullptr || atoi(getenv("B200BOOT_K1MG")) != 0;
  if (use_mg && data && stat_out && resample_hi > resample_lo &&
      median_gemm_applies(n_rows, n_cols, resample_hi - resample_lo))
    return median_gemm(data, n_rows, n_cols, ld, seed, resample_lo, resample_hi - resample_lo, stat_out, ws, ws_bytes,
                       as_stream(stream), full, loo, jack_out);
  // one column — or a few long ones, where K1m would leave all but a few warps idle — go column
  // by column through K1m1 (same seed: same index sets, the rows stay aligned across columns)
  // (measured, n = 1e4, K1m vs the loop: N = 3000: D = 2 0.76 vs 0.53 ms, D = 4 0.77 vs 0.80;
  // N = 1e4: D = 8 5.45 vs 3.18, D = 16 6.51 vs 6.23; N = 28 000: D = 16 31.6 vs 23.0, D = 32 49.5 vs 45.8)
  const int few = n_rows <= 2048 ? 1 : (n_rows <= 4096 ? 2 : (n_rows <= 8192 ? 4 : (n_rows <= 16384 ? 16 : 32)));
  if (data && stat_out && n_rows >= 1 && n_rows <= kMedMaxRows && resample_hi > resample_lo && n_cols <= few) {
    const int64_t n_local = resample_hi - resample_lo;
    for (int64_t d = 0; d < n_cols; ++d) {
      int rc = median_single_column(data + d, n_rows, ld, seed, resample_lo, n_local, nullptr, stat_out + d * n_local,
                                    ws, ws_bytes, as_stream(stream), full, loo, jack_out ? jack_out + 8 * d : nullptr);
      if (rc) return rc;
    }
    return 0;
  }
  return median_resample(data, n_rows, n_cols, ld, seed, resample_lo, resample_hi, stat_out, ws,
                         ws_bytes, as_stream(stream), full, nullptr, jack_out, &loo);
}

int b200boot_resample_median_columns_indexed(const double* data, int64_t n_rows, int64_t n_cols,
                                             int64_t ld, const int64_t* indices, int64_t n_sets,
                                             double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                             const b200boot_rank_rule* rule) {
  if (!indices) return fail(B200BOOT_EINVAL, "null indices");
  RankRule full, loo;
  if (int rc = rank_rules(rule, n_rows, &full, &loo)) return rc;
  if (n_cols == 1 && data && stat_out && n_rows >= 1 && n_rows <= kMedMaxRows && n_sets > 0)
    return median_single_column(data, n_rows, ld, 0, 0, n_sets, indices, stat_out, ws, ws_bytes, as_stream(stream),
                                full, loo, nullptr);
  return median_resample(data, n_rows, n_cols, ld, 0, 0, n_sets, stat_out, ws, ws_bytes,
                         as_stream(stream), full, indices);
}

int b200boot_ci_median_small_supported(int64_t n_rows, int64_t n_samples, int n_alphas) {
  return n_rows >= 1 && n_rows <= kMedMaxRows && b200boot_ci_tail_supported(n_samples, n_alphas) ? 1 : 0;
}

int b200boot_ci_median_small(const double* data, const double* host_data, double* staging, int64_t n_rows,
                             uint64_t seed, int64_t n_samples, const double* alphas, int n_alphas, int bca,
                             int need_jack, const b200boot_rank_rule* rule, double* stat_out, double* result_dev,
                             double* result_host, void* ws, size_t ws_bytes, void* stream) {
  if (!data || !alphas || !stat_out || !result_dev || !result_host)
    return fail(B200BOOT_EINVAL, "ci_median_small: null pointer");
  if (!b200boot_ci_median_small_supported(n_rows, n_samples, n_alphas))
    return fail(B200BOOT_EUNSUPPORTED, "ci_median_small: one resident column, n_samples <= 2^21, <= 8 levels");
  RankRule full, loo;
  if (int rc = rank_rules(rule, n_rows, &full, &loo)) return rc;
  cudaStream_t s = as_stream(stream);
  if (host_data) {
    if (!staging) return fail(B200BOOT_EINVAL, "ci_median_small: a staging buffer is required for host rows");
    memcpy(staging, host_data, (size_t)n_rows * sizeof(double));
    B2_CUDA(cudaMemcpyAsync(const_cast<double*>(data), staging, (size_t)n_rows * sizeof(double),
                            cudaMemcpyHostToDevice, s));
  }
  if (bca) need_jack = 1;
  double* jack = need_jack ? result_dev + kSmallResultDoubles : nullptr;
  const double* sorted = nullptr;
  int rc = median_single_column(data, n_rows, 1, seed, 0, n_samples, nullptr, stat_out, ws, ws_bytes, s, full, loo,
                                jack, &sorted);
  if (rc) return rc;
  rc = enqueue_ci_tail(stat_out, n_samples, jack, alphas, n_alphas, bca, result_dev, result_host, s,
                       sorted + (n_rows - 1));
  if (rc) return rc;
  B2_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int b200boot_jackknife_median_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                      double* out, void* ws, size_t ws_bytes, void* stream,
                                      const b200boot_rank_rule* rule) {
  RankRule full, loo;
  if (int rc = rank_rules(rule, n_rows, &full, &loo)) return rc;
  return median_jackknife(data, n_rows, n_cols, ld, out, ws, ws_bytes, as_stream(stream), full, loo);
}

int b200boot_columns_block_width(int stat_id, int64_t n_rows) {
  return columns_supported(stat_id, n_rows) ? column_block_width(n_rows) : 0;
}

int b200boot_resample_stat_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                   int stat_id, uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                   double* stat_out, void* ws, size_t ws_bytes, void* stream) {
  return columns_resample(data, n_rows, n_cols, ld, stat_id, seed, resample_lo, resample_hi, stat_out, ws,
                          ws_bytes, as_stream(stream));
}

int b200boot_jackknife_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, int stat_id,
                               double* out, void* stream) {
  if (!data || !out || n_rows < 1 || n_cols < 1 || ld < n_cols) return fail(B200BOOT_EINVAL, "bad arguments");
  const int ms = moment_set_of(stat_id);
  if (ms != kMomSum && ms != kMomSum2)
    return fail(B200BOOT_EUNSUPPORTED, "statistic %d is not a single-array moment statistic", stat_id);
  return columns_jackknife(data, n_rows, n_cols, ld, stat_id, nullptr, out, as_stream(stream));
}

int b200boot_resample_median_sorted(const double* sorted_vals, int64_t n_rows, uint64_t seed,
                                    int64_t resample_lo, int64_t resample_hi, double* stat_out, void* ws,
                                    size_t ws_bytes, void* stream, const b200boot_rank_rule* rule) {
  RankRule full, loo;
  if (int rc = rank_rules(rule, n_rows, &full, &loo)) return rc;
  return median_resample_sorted(sorted_vals, n_rows, seed, resample_lo, resample_hi, stat_out, ws, ws_bytes,
                                as_stream(stream), full);
}

int b200boot_jackknife_median_sorted(const double* sorted_vals, int64_t n_rows, double* out, void* stream,
                                     const b200boot_rank_rule* rule) {
  RankRule full, loo;
  if (int rc = rank_rules(rule, n_rows, &full, &loo)) return rc;
  return median_jackknife_sorted(sorted_vals, n_rows, out, as_stream(stream), full, loo);
}

// ---- multi='independent': combine, leave-one-out vectors, pooled moments ------------------
int b200boot_combine(int op, const double* a, const double* b, double* out, int64_t n, void* stream) {
  if (!a || !b || !out || n < 0 || op < 0 || op > B200BOOT_OP_LOG_RATIO) return fail(B200BOOT_EINVAL, "bad arguments");
  if (n == 0) return 0;
  combine_kernel<<<blocks_for(n, 256, 148 * 16), 256, 0, as_stream(stream)>>>(op, a, b, out, n);
  B2_LAUNCH_CHECK();
  return 0;
}


int b200boot_jackknife_values(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                              double* values, double* summary, void* ws, size_t ws_bytes, void* stream) {
  int rc = check_moment_stat(stat_id, n_arrays, n_rows);
  if (rc) return rc;
  if (!data || !values || !summary) return fail(B200BOOT_EINVAL, "null pointer");
  if (stat_outputs(stat_id) != 1) return fail(B200BOOT_EUNSUPPORTED, "leave-one-out vectors need a scalar statistic");
  cudaStream_t s = as_stream(stream);
  const Plan plan = make_plan(n_rows, n_arrays);
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  ResampleWs w = carve_resample(cv, plan, stat_id, 0);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  const double* rows[2];
  rc = prepare_rows(data, n_arrays, n_rows, stat_id, w, rows, s);
  if (rc) return rc;
  switch (moment_set_of(stat_id)) {
    case kMomSum: return loo_values<kMomSum>(rows, n_rows, stat_id, w, values, summary, s);
    case kMomSum2: return loo_values<kMomSum2>(rows, n_rows, stat_id, w, values, summary, s);
    case kMomAB: return loo_values<kMomAB>(rows, n_rows, stat_id, w, values, summary, s);
    case kMomXW: return loo_values<kMomXW>(rows, n_rows, stat_id, w, values, summary, s);
    default: return loo_values<kMomPair5>(rows, n_rows, stat_id, w, values, summary, s);
  }
}

int b200boot_jackknife_values_sorted(const double* sorted_vals, int64_t n_rows, double* values,
                                     double* summary, void* stream, const b200boot_rank_rule* rule) {
  if (!sorted_vals || !values || !summary || n_rows < 1) return fail(B200BOOT_EINVAL, "bad arguments");
  RankRule full, loo;
  if (int rc = rank_rules(rule, n_rows, &full, &loo)) return rc;
  cudaStream_t s = as_stream(stream);
  int rc = median_jackknife_sorted(sorted_vals, n_rows, summary, s, full, loo);
  if (rc) return rc;
  loo_values_sorted_kernel<<<blocks_for(n_rows, 256, 148 * 16), 256, 0, s>>>(sorted_vals, n_rows, loo, values);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_jackknife_pooled(int op, const double* loo_a, int64_t n_a, const double* s_a,
                              const double* loo_b, int64_t n_b, const double* s_b, double* out,
                              void* ws, size_t ws_bytes, void* stream) {
  if (!loo_a || !loo_b || !s_a || !s_b || !out || n_a < 1 || n_b < 1 || op < 0 || op > B200BOOT_OP_LOG_RATIO)
    return fail(B200BOOT_EINVAL, "bad arguments");
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  Carver cv(ws, ws_bytes);
  JackState* st = cv.take<JackState>(1);
  double* red = cv.take<double>((size_t)kRedBlocks * kMaxMoments);
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  cudaStream_t s = as_stream(stream);
  const PooledLoo p{op, loo_a, n_a, s_a, loo_b, n_b, s_b};
  int rc = run_reduction<1>(n_a + n_b, PooledSumF{p}, StorePooledMeanG{p, st}, red, s);
  if (rc) return rc;
  return run_reduction<2>(n_a + n_b, PooledDevF{p, st}, StorePooledG{p, st, out}, red, s);
}

// ---- method='abc': numpy-faithful weighted means --------------------------------------------
int b200boot_abc_identity_averages(const double* x, int64_t n_rows, double ep, double* tp, double* tm,
                                   void* stream) {
  if (!x || !tp || !tm || n_rows < 1) return fail(B200BOOT_EINVAL, "bad arguments");
  abc_identity_averages_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, as_stream(stream)>>>(x, n_rows, ep, tp, tm);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_weighted_averages(const double* x, int64_t n_rows, const double* weights, int64_t n_vectors,
                               double* out, void* stream) {
  if (!x || !out || n_rows < 1 || n_vectors < 0 || (!weights && n_vectors != 1))
    return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_vectors == 0) return 0;
  abc_weighted_averages_kernel<<<(unsigned)((n_vectors + 31) / 32), 32, 0, as_stream(stream)>>>(x, n_rows, weights,
                                                                                               n_vectors, out);
  B2_LAUNCH_CHECK();
  return 0;
}

// ---- K6: count matrix and finisher around a library DGEMM ------------------------------
int b200boot_gemm_supported(int stat_id, int64_t n_rows) { return gemm_supported(stat_id, n_rows) ? 1 : 0; }

int b200boot_count_matrix(uint64_t seed, int64_t n_rows, int64_t resample_lo, int64_t resample_hi,
                          double* out, void* stream) {
  const int64_t n_local = resample_hi - resample_lo;
  if (!out || n_rows < 1 || n_local < 0 || resample_lo < 0) return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_rows > kSmemDataBytes / 8)
    return fail(B200BOOT_EUNSUPPORTED, "count matrix: n_rows=%lld beyond the single-cell limit", (long long)n_rows);
  if (n_local == 0) return 0;
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    return cudaFuncSetAttribute(k6_count_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(kSmemHeader + kSmemDataBytes));
  }));
  const size_t smem = (size_t)n_rows * 4;
  const int per_sm = (int)std::max<int64_t>(1, std::min<int64_t>(4, (int64_t)(200 * 1024) / (int64_t)std::max<size_t>(smem, 1)));
  const int grid = (int)std::min<int64_t>((int64_t)sm_count() * per_sm, n_local);
  k6_count_matrix_kernel<<<grid, kK6Threads, smem, as_stream(stream)>>>(seed, n_rows, resample_lo, n_local, out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_counts_from_indices(const int64_t* indices, int64_t n_sets, int64_t n_rows, double* out,
                                 void* stream) {
  if (!indices || !out || n_sets < 0 || n_rows < 1) return fail(B200BOOT_EINVAL, "bad arguments");
  if (n_sets == 0) return 0;
  cudaStream_t s = as_stream(stream);
  B2_CUDA(cudaMemsetAsync(out, 0, (size_t)n_sets * (size_t)n_rows * 8, s));
  k6_counts_from_indices_kernel<<<blocks_for(n_sets * n_rows, 256, 148 * 16), 256, 0, s>>>(indices, n_sets, n_rows,
                                                                                         out);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_finish_columns(int stat_id, int64_t n_rows, int64_t n_cols, int64_t n_local, double* s1,
                            const double* s2, void* stream) {
  if (!s1 || n_rows < 1 || n_cols < 1 || n_local < 0) return fail(B200BOOT_EINVAL, "bad arguments");
  const int ms = moment_set_of(stat_id);
  if (ms != kMomSum && ms != kMomSum2)
    return fail(B200BOOT_EUNSUPPORTED, "statistic %d is not a single-array moment statistic", stat_id);
  if (ms == kMomSum2 && !s2) return fail(B200BOOT_EINVAL, "var / std need the second-moment sums");
  const int64_t total = n_cols * n_local;
  if (total == 0) return 0;
  k6_finish_kernel<<<blocks_for(total, 256, 148 * 16), 256, 0, as_stream(stream)>>>(stat_id, (double)n_rows, total,
                                                                                  s1, s2);
  B2_LAUNCH_CHECK();
  return 0;
}

// ---- K6i: the count x data contraction on tcgen05 (exact integer split) -----------------
namespace {
int64_t k6i_k_pad(int64_t n_rows) { return ((n_rows + kI8KChunk - 1) / kI8KChunk) * kI8KChunk; }
int64_t k6i_cols_pad(int64_t n_cols) { return ((n_cols + kI8ColsPerTile - 1) / kI8ColsPerTile) * kI8ColsPerTile; }

int launch_k6i(const K6iParams& P, cudaStream_t s) {
  const bool resident = P.k_pad / kI8KChunk <= kI8MaxResidentChunks;
  const size_t smem = k6i_smem_bytes(P.k_pad, resident);
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    cudaError_t e = cudaFuncSetAttribute(k6i_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)k6i_smem_bytes(kI8MaxResidentChunks * kI8KChunk, true));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k6i_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)k6i_smem_bytes(kI8KChunk, false));
    return e;
  }));
  const int64_t m_tiles = (P.m_total + kI8M - 1) / kI8M;
  // column splits: fill the SMs when there are few resample tiles
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(P.n_col_tiles, (2 * (int64_t)sm_count()) / std::max<int64_t>(m_tiles, 1)));
  splits = std::min<int64_t>(splits, 65535);
  const dim3 grid((unsigned)m_tiles, (unsigned)splits);
  if (resident)
    k6i_mma_kernel<true><<<grid, kI8Threads, smem, s>>>(P);
  else
    k6i_mma_kernel<false><<<grid, kI8Threads, smem, s>>>(P);
  B2_LAUNCH_CHECK();
  return 0;
}
}  // namespace

// image sizes of the pipelined contraction (k6i_pipe.cuh)
inline int64_t k6i_a_image_tiles(int64_t m_total) { return (m_total + kI8M - 1) / kI8M; }
inline size_t k6i_a_image_bytes(int64_t m_total, int64_t k_pad) {
  return (size_t)k6i_a_image_tiles(m_total) * (size_t)(k_pad / kI8KChunk) * kI8AChunkBytes;
}
inline size_t k6i_b_image_bytes(int64_t n_col_tiles, int64_t k_pad) {
  return (size_t)n_col_tiles * (size_t)(k_pad / kI8KChunk) * kI8BChunkBytes;
}

// The contraction through the warp-specialised pipeline: operands re-laid as tile images first.
int launch_k6i_pipe(const K6iParams& P, unsigned char* a_img, unsigned char* b_img, cudaStream_t s) {
  const int64_t m_tiles = (P.m_total + kI8M - 1) / kI8M;
  const int n_kchunks = (int)(P.k_pad / kI8KChunk);
  const int64_t a_tiles = k6i_a_image_tiles(P.m_total);
  k6i_image_kernel<kI8M><<<blocks_for(a_tiles * kI8M * (P.k_pad / 16), 256, 148 * 16), 256, 0, s>>>(
      P.counts, P.m_total, P.k_pad, a_tiles, a_img);
  B2_LAUNCH_CHECK();
  k6i_image_kernel<kI8N><<<blocks_for(P.n_col_tiles * kI8N * (P.k_pad / 16), 256, 148 * 16), 256, 0, s>>>(
      reinterpret_cast<const unsigned char*>(P.digits), P.n_col_tiles * kI8N, P.k_pad, P.n_col_tiles, b_img);
  B2_LAUNCH_CHECK();
  const bool resident = n_kchunks <= kPipeMaxResidentChunks;
  const size_t limit = 227 * 1024;
  int n_stages = kPipeMaxStages;
  while (n_stages > 2 && k6i_pipe_smem_bytes(n_kchunks, resident, n_stages) > limit) --n_stages;
  const size_t smem = k6i_pipe_smem_bytes(n_kchunks, resident, n_stages);
  if (smem > limit) return fail(B200BOOT_EUNSUPPORTED, "k6i pipeline: shared memory");
  static PerDeviceOnce configured;
  B2_CUDA(configured.run([] {
    cudaError_t e = cudaFuncSetAttribute(k6i_pipe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(k6i_pipe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    return e;
  }));
  K6iPipeParams Q{};
  Q.a_img = a_img;
  Q.b_img = b_img;
  Q.scale = P.scale;
  Q.m_total = P.m_total;
  Q.m_tiles = m_tiles;
  Q.n_kchunks = n_kchunks;
  Q.n_stages = n_stages;
  Q.n_cols = P.n_cols;
  Q.n_col_tiles = P.n_col_tiles;
  Q.out = P.out;
  Q.out_ld = P.out_ld;
  Q.out_i32 = P.out_i32;
  Q.overflow = P.overflow;
  Q.out_divisor = P.out_divisor;
  const int grid = (int)std::min<int64_t>(sm_count(), m_tiles * P.n_col_tiles);
  if (resident)
    k6i_pipe_kernel<true><<<grid, kPipeThreads, smem, s>>>(Q);
  else
    k6i_pipe_kernel<false><<<grid, kPipeThreads, smem, s>>>(Q);
  B2_LAUNCH_CHECK();
  return 0;
}

size_t b200boot_count_gemm_ws_bytes(int64_t n_rows, int64_t n_cols, int64_t n_local) {
  if (n_rows < 1 || n_cols < 1 || n_local < 0) return 0;
  const int64_t k_pad = k6i_k_pad(n_rows), cp = k6i_cols_pad(n_cols);
  return 4096 + pad256((size_t)cp * kI8Digits * k_pad) + 3 * pad256((size_t)cp * 8) +
         pad256((size_t)std::max<int64_t>(n_local, 1) * k_pad) + pad256(k6i_a_image_bytes(std::max<int64_t>(n_local, 1), k_pad)) +
         pad256(k6i_b_image_bytes(cp / kI8ColsPerTile, k_pad));
}

int b200boot_resample_stat_columns_gemm(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                        int stat_id, uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                        const int64_t* indices, double* stat_out, double* second, void* ws,
                                        size_t ws_bytes, void* stream) {
  const int64_t n_local = resample_hi - resample_lo;
  const int ms = moment_set_of(stat_id);
  if (ms != kMomSum && ms != kMomSum2)
    return fail(B200BOOT_EUNSUPPORTED, "statistic %d is not a single-array moment statistic", stat_id);
  if (!data || !stat_out || n_rows < 1 || n_rows >= ((int64_t)1 << 25) || n_cols < 1 || ld < n_cols || n_local < 0)
    return fail(B200BOOT_EINVAL, "columns_gemm: bad arguments (n_rows must stay below 2^25)");
  const bool two = ms == kMomSum2;
  if (two && !second) return fail(B200BOOT_EINVAL, "var / std need the second-moment buffer");
  if (n_local == 0) return 0;
  if (!indices && n_rows > kSmemDataBytes / 8)
    return fail(B200BOOT_EUNSUPPORTED, "columns_gemm: n_rows=%lld beyond one cell needs materialised indices",
                (long long)n_rows);
  if ((reinterpret_cast<uintptr_t>(ws) & 255u) != 0)
    return fail(B200BOOT_EWORKSPACE, "workspace must be 256-byte aligned");
  cudaStream_t s = as_stream(stream);
  const int64_t k_pad = k6i_k_pad(n_rows), cp = k6i_cols_pad(n_cols);
  Carver cv(ws, ws_bytes);
  int32_t* flags = cv.take<int32_t>(64);           // [0] count overflow, [1] some column is non-finite
  int8_t* digits = cv.take<int8_t>((size_t)cp * kI8Digits * k_pad);
  double* scale = cv.take<double>((size_t)cp);
  double* centre = cv.take<double>((size_t)cp);
  int32_t* bad_col = cv.take<int32_t>((size_t)cp * 2);
  uint8_t* counts = cv.take<uint8_t>((size_t)n_local * k_pad);
  unsigned char* a_img = cv.take<unsigned char>(k6i_a_image_bytes(n_local, k_pad));
  unsigned char* b_img = cv.take<unsigned char>(k6i_b_image_bytes(cp / kI8ColsPerTile, k_pad));
  if (!cv.ok) return fail(B200BOOT_EWORKSPACE, "workspace too small: need %zu bytes", cv.used);
  static const bool use_pipe = getenv("B200BOOT_K6I_PIPE") == nullptr || atoi(getenv("B200BOOT_K6I_PIPE")) != 0;
  B2_CUDA(cudaMemsetAsync(flags, 0, 256, s));
  B2_CUDA(cudaMemsetAsync(bad_col, 0, (size_t)cp * 4, s));
  // the resamples' count matrix, once
  if (indices) {
    B2_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_local * k_pad, s));
    k6i_count_bytes_from_indices_kernel<<<blocks_for(n_local * n_rows, 256, 148 * 16), 256, 0, s>>>(
        indices, n_local, n_rows, k_pad, counts, flags);
    B2_LAUNCH_CHECK();
  } else {
    static PerDeviceOnce configured;
    B2_CUDA(configured.run([] {
      return cudaFuncSetAttribute(k6i_count_bytes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    }));
    const int grid = (int)std::min<int64_t>((int64_t)sm_count() * 4, n_local);
    k6i_count_bytes_kernel<<<grid, 256, (size_t)k_pad, s>>>(seed, n_rows, k_pad, resample_lo, n_local, counts, flags);
    B2_LAUNCH_CHECK();
  }
  const double* c = nullptr;
  if (two) {                                       // var / std: columns centred on their means, as K1c does
    const dim3 block(32, 8);
    k1c_column_means_kernel<<<(unsigned)((n_cols + 31) / 32), block, 0, s>>>(data, n_rows, n_cols, ld, centre);
    B2_LAUNCH_CHECK();
    c = centre;
  }
  for (int pass = 0; pass < (two ? 2 : 1); ++pass) {
    double* dst = pass == 0 ? stat_out : second;
    B2_CUDA(cudaMemsetAsync(scale, 0, (size_t)cp * 8, s));
    k6i_column_scale_kernel<<<(unsigned)((n_cols + 31) / 32), 256, 0, s>>>(data, n_rows, n_cols, ld, c, pass, scale,
                                                                          bad_col, flags + 1);
    B2_LAUNCH_CHECK();
    k6i_digits_kernel<<<blocks_for(cp * (k_pad / 16), 256, 148 * 32), 256, 0, s>>>(data, n_rows, n_cols, cp, ld, k_pad,
                                                                                  c, pass, scale, digits);
    B2_LAUNCH_CHECK();
    K6iParams P{};
    P.counts = counts;
    P.digits = digits;
    P.scale = scale;
    P.m_total = n_local;
    P.k_pad = k_pad;
    P.n_cols = n_cols;
    P.n_col_tiles = cp / kI8ColsPerTile;
    P.out = dst;
    P.out_ld = n_local;
    P.out_i32 = nullptr;
    P.overflow = flags;
    // the mean's division by N happens in the epilogue (the finisher's m[0] / cnt on the same
    // rounded sum: same bits), so mean and sum need no finisher pass over the output
    P.out_divisor = stat_id == kStatMean ? (double)n_rows : 1.0;
    int rc = use_pipe ? launch_k6i_pipe(P, a_img, b_img, s) : launch_k6i(P, s);
    if (rc) return rc;
    k6i_fixup_kernel<<<(unsigned)std::min<int64_t>((n_local * 32 + 255) / 256, (int64_t)sm_count() * 8), 256, 0, s>>>(
        data, n_rows, n_cols, ld, c, pass, bad_col, flags + 1, seed, resample_lo, n_local, indices, P.out_divisor,
        dst);
    B2_LAUNCH_CHECK();
  }
  if (stat_id == kStatSum || stat_id == kStatMean) return 0;      // finished in the contraction's epilogue
  const int64_t total = n_cols * n_local;
  k6_finish_kernel<<<blocks_for(total, 256, 148 * 16), 256, 0, s>>>(stat_id, (double)n_rows, total, stat_out,
                                                                   two ? second : nullptr);
  B2_LAUNCH_CHECK();
  return 0;
}

int b200boot_debug_i8_gemm(const uint8_t* a, const int8_t* b, int32_t* c, int64_t m, int64_t n_digit_rows,
                           int64_t k_pad, void* stream) {
  if (!a || !b || !c || m < 1 || n_digit_rows < kI8N || n_digit_rows % kI8N != 0 || k_pad < kI8KChunk ||
      k_pad % kI8KChunk != 0)
    return fail(B200BOOT_EINVAL, "debug_i8_gemm: n must be a multiple of 256 and k of 128");
  K6iParams P{};
  P.counts = a;
  P.digits = b;
  P.scale = nullptr;
  P.m_total = m;
  P.k_pad = k_pad;
  P.n_cols = 0;
  P.n_col_tiles = n_digit_rows / kI8N;
  P.out = nullptr;
  P.out_ld = 0;
  P.out_i32 = c;
  P.overflow = nullptr;
  P.out_divisor = 1.0;
  if (getenv("B200BOOT_K6I_PIPE") != nullptr && atoi(getenv("B200BOOT_K6I_PIPE")) == 0)
    return launch_k6i(P, as_stream(stream));                 // the first (unpipelined) form
  // test hook only: the images live in memory of their own for the duration of the call
  cudaStream_t s = as_stream(stream);
  unsigned char *a_img = nullptr, *b_img = nullptr;
  B2_CUDA(cudaMalloc(&a_img, k6i_a_image_bytes(m, k_pad)));
  B2_CUDA(cudaMalloc(&b_img, k6i_b_image_bytes(P.n_col_tiles, k_pad)));
  int rc = launch_k6i_pipe(P, a_img, b_img, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFree(a_img);
  cudaFree(b_img);
  if (rc) return rc;
  B2_CUDA(e);
  return 0;
}

// ---- instrumentation ---------------------------------------------------------------
int64_t b200boot_launch_count(void) { return (int64_t)launch_counter().load(std::memory_order_relaxed); }

int b200boot_profile_enable(int on) {
  if (on && !g_prof.start) {
    B2_CUDA(cudaEventCreate(&g_prof.begin));
    B2_CUDA(cudaEventCreate(&g_prof.start));
    B2_CUDA(cudaEventCreate(&g_prof.stop));
  }
  g_prof.enabled = on != 0;
  g_prof.valid = false;
  return 0;
}

int b200boot_profile_last_k1_ms(float* ms) {
  if (!ms) return fail(B200BOOT_EINVAL, "null ms");
  if (!g_prof.valid) return fail(B200BOOT_EINVAL, "no timed K1 launch recorded");
  B2_CUDA(cudaEventSynchronize(g_prof.stop));
  B2_CUDA(cudaEventElapsedTime(ms, g_prof.start, g_prof.stop));
  return 0;
}

int b200boot_profile_last_ms(float* k0_ms, float* k1_ms) {
  if (!k0_ms || !k1_ms) return fail(B200BOOT_EINVAL, "null ms");
  if (!g_prof.valid) return fail(B200BOOT_EINVAL, "no timed K1 launch recorded");
  B2_CUDA(cudaEventSynchronize(g_prof.stop));
  B2_CUDA(cudaEventElapsedTime(k0_ms, g_prof.begin, g_prof.start));
  B2_CUDA(cudaEventElapsedTime(k1_ms, g_prof.start, g_prof.stop));
  return 0;
}

// ---- host probes (CPU test-suite; same sources as the kernels) ------------------
void b200boot_host_ncdf(const double* x, double* out, int64_t n) {
  const double s2 = 1.4142135623730951;          // sqrt(2.0), as the reference's module constant
  for (int64_t i = 0; i < n; ++i) out[i] = 0.5 * (1.0 + erf(x[i] / s2));
}

void b200boot_host_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const Words4 w = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = w.w[i];
}

void b200boot_host_philox4x32_r(int rounds, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const Words4 w = rounds == 7 ? philox4x32_r<7>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1])
                               : philox4x32_r<10>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = w.w[i];
}

int64_t b200boot_host_binomial(uint64_t seed, int64_t resample, int64_t cell, int64_t n, double p) {
  const Stream s = make_stream(seed, resample, (uint32_t)cell, kPurposeBinomial);
  if (n <= 0 || p <= 0.0) return 0;
  if (p >= 1.0) return n;
  const bool flip = p > 0.5;
  const double pp = flip ? 1.0 - p : p;
  const int64_t k = ((double)n * pp >= 10.0) ? binomial_btrs(n, pp, s) : binomial_inversion(n, pp, s);
  return flip ? n - k : k;
}

int64_t b200boot_host_binomial_half(uint64_t seed, int64_t resample, int64_t cell, int64_t n) {
  if (n <= 0) return 0;
  return binomial_half(n, make_stream(seed, resample, (uint32_t)cell, kPurposeBinomial));
}

int b200boot_host_tile_counts(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample,
                              int32_t* counts) {
  if (n_rows < 1 || n_arrays < 1 || !counts) return fail(B200BOOT_EINVAL, "bad arguments");
  const Plan plan = make_plan(n_rows, n_arrays);
  tile_count_chain<int32_t>(seed, resample, plan, counts, 1);
  return 0;
}

int b200boot_host_class_counts(uint64_t seed, int64_t resample, int64_t tile, int64_t n_tile_draws,
                               int32_t* counts16) {
  if (!counts16 || tile < 0 || n_tile_draws < 0) return fail(B200BOOT_EINVAL, "bad arguments");
  class_count_tree<int32_t>(seed, resample, tile, n_tile_draws, counts16, 1);
  return 0;
}

int b200boot_host_indices(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample, int64_t* out) {
  if (n_rows < 1 || n_rows >= ((int64_t)1 << 31) || n_arrays < 1 || !out)
    return fail(B200BOOT_EINVAL, "bad arguments");
  const Plan plan = make_plan(n_rows, n_arrays);
  std::vector<int32_t> counts((size_t)plan.n_tiles);
  if (plan.n_tiles > 1)
    tile_count_chain<int32_t>(seed, resample, plan, counts.data(), 1);
  else
    counts[0] = (int32_t)n_rows;
  const uint32_t m7 = ((1u << (plan.log2_tile - 4)) - 1u) << 7;
  int64_t offset = 0;
  for (int64_t t = 0; t < plan.n_tiles; ++t) {
    const int64_t n = counts[(size_t)t];
    const int64_t sz = plan.cell_size(t);
    const int64_t row0 = t < plan.n_full ? (t << plan.log2_tile) : (plan.n_full << plan.log2_tile);
    auto put = [&](int, int64_t pos, uint32_t idx) { out[offset + pos] = row0 + (int64_t)idx; };
    if (t < plan.n_full) {
      int32_t cls[kClasses];
      class_count_tree<int32_t>(seed, resample, t, n, cls, 1);
      for (int c = 0; c < kClasses; ++c) {
        const Stream st = make_stream(seed, resample, class_cell_id(t, c), kPurposeIndex);
        for (uint32_t q = 0; (int64_t)q * kClassDrawsPerBlock < cls[c]; ++q)
          class_block<true>(st, q, m7, c, (int64_t)cls[c], put);
        offset += cls[c];
      }
    } else {
      const Stream st = make_stream(seed, resample, (uint32_t)t, kPurposeIndex);
      const uint32_t thresh = lemire_threshold((uint32_t)sz);
      for (uint32_t q = 0; (int64_t)q * 4 < n; ++q) lemire_block<true>(st, q, (uint32_t)sz, thresh, n, put);
      offset += n;
    }
  }
  return 0;
}

}  // extern "C"
