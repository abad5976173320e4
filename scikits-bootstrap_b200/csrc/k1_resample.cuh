// K0 (tile occupancies), K1 (fused draw + gather + FP64 reduce), K1i (indexed
// parity mode), K2 (index materialiser) and the moment finisher.
//
// Replaces /root/reference/src/scikits/bootstrap/bootstrap.py:429-432 (resample
// loop) and :677-680 (index generator).  No index array is materialised by K1.
//
// Execution shape of K1: persistent CTAs (one per SM, 1024 threads).  A work
// item is (cell t, chunk of resamples); the CTA bulk-copies cell t (<= 224 KB)
// into shared memory with cp.async.bulk + mbarrier, then each warp owns one
// resample at a time: its lanes generate Philox blocks, turn the words into
// row offsets inside the cell, gather from shared memory and accumulate the
// statistic's contribution vector in FP64.  A warp-shuffle butterfly produces
// the per-(cell, resample) partial, later summed over cells in fixed order.
// In a full tile each lane is tied to one shared-memory bank pair (bank class),
// so the gathers are conflict-free; see draw.cuh.
#pragma once
#include "common.cuh"
#include "draw.cuh"
#include "stats.cuh"

namespace b200boot {

constexpr int kK1Threads = 1024;
constexpr int kK1Warps = kK1Threads / 32;
constexpr int kSmemHeader = 128;  // mbarrier + work-item slot
// sm_90+/sm_100: the first 1 KB of the shared window is reserved by the system, so a
// kernel without static shared memory finds its dynamic shared memory at 0x400.  The
// FAST kernel variant bakes cell address = 0x400 + header into LDS immediates.  The host
// selects it only after a one-time probe kernel has confirmed, on this device, that the
// dynamic shared memory of a static-shared-free kernel starts at 0x400
// (probe_dynamic_smem_base_kernel); the kernel still traps if the assumption is violated.
#ifndef B2_CLASS_UNROLL
#define B2_CLASS_UNROLL 1   // measured on B200: 1 -> 82.7 ms, 2 -> 85.3, 3 -> 87.9, 4 -> 87.1 (cfg5 / 5)
#endif
constexpr int kClassUnroll = B2_CLASS_UNROLL;   // Philox blocks in flight per lane in the class loop
constexpr int kFastSmemBase = 1024 + kSmemHeader;
constexpr int kFastStride2 = 8 << 13;   // bytes between the two arrays of a 2-array full tile

__global__ void probe_dynamic_smem_base_kernel(uint32_t* out) {
  extern __shared__ __align__(128) unsigned char probe_smem[];
  if (threadIdx.x == 0) *out = smem_u32(probe_smem);
}

// ------------------------------------------------------------------ K0 --------
// one thread per resample walks both levels (inspection API, small jobs)
__global__ void __launch_bounds__(128) k0_tile_counts_kernel(uint64_t seed, int64_t resample_lo,
                                                             int64_t n_local, Plan plan,
                                                             int32_t* __restrict__ counts) {
  const int64_t rl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rl >= n_local) return;
  tile_count_chain<int32_t>(seed, resample_lo + rl, plan, counts + rl, n_local);
}

// two-kernel form: level 1 per resample, then level 2 per (group, resample) — chains of
// at most 32 steps and n_groups times the parallelism
__global__ void __launch_bounds__(128) k0_group_counts_kernel(uint64_t seed, int64_t resample_lo,
                                                              int64_t n_local, Plan plan,
                                                              int32_t* __restrict__ gcounts) {
  const int64_t rl = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rl >= n_local) return;
  group_count_chain<int32_t>(seed, resample_lo + rl, plan, gcounts + rl, n_local);
}

__global__ void __launch_bounds__(128) k0_tiles_in_groups_kernel(uint64_t seed, int64_t resample_lo,
                                                                 int64_t n_local, Plan plan,
                                                                 const int32_t* __restrict__ gcounts,
                                                                 int32_t* __restrict__ counts) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t ng = n_tile_groups(plan);
  if (e >= ng * n_local) return;
  const int64_t g = e / n_local, rl = e - g * n_local;
  const int64_t n_g = ng > 1 ? (int64_t)gcounts[e] : plan.n_rows;
  tile_chain_in_group<int32_t>(seed, resample_lo + rl, plan, g, n_g, counts + rl, n_local);
}

// Class occupancies of the full tiles: cls[(t * n_local + r) * 16 + c]; one thread
// walks the 15-step conditional-binomial chain of one (tile, resample) pair.
// Measured alternatives that were NOT faster on B200 (42 ms at cfg5 either way, see
// DESIGN.md): a per-lane state machine that lets lanes advance independently (123
// registers, 55 ms), and a warp-cooperative evaluation of the full acceptance test
// (thread efficiency 12 -> 20 of 32, same time: the kernel is bound by issue slots of
// the divergent proposal rounds, not by the FP64 pipe).
__global__ void __launch_bounds__(128) k0_class_counts_kernel(uint64_t seed, int64_t resample_lo,
                                                              int64_t n_local, int64_t n_full,
                                                              const int32_t* __restrict__ counts,
                                                              int32_t* __restrict__ cls) {
  const int64_t total = n_full * n_local;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t t = e / n_local, rl = e - t * n_local;
    class_count_chain<int32_t>(seed, resample_lo + rl, t, (int64_t)__ldg(counts + e), cls + e * kClasses, 1);
  }
}

// Two-kernel form of the class chains used by the resample path: at chain step c every
// (tile, resample) pair makes its FIRST proposal (setup + one BTRS proposal + squeeze,
// no divergence); the ~14 % of pairs whose proposal needs the full acceptance test (or
// whose occupancy is tiny) are appended to a compact list and resolved by a second
// kernel in which every lane is on the expensive path.  Step c+1 then starts from the
// updated remainders.  Values are identical to class_count_chain: the resolver simply
// replays binomial_ratio from attempt 0.
__global__ void __launch_bounds__(256) k0_class_step_propose_kernel(
    uint64_t seed, int64_t resample_lo, int64_t n_local, int64_t n_full, int c,
    const int32_t* __restrict__ counts, int32_t* __restrict__ left, int32_t* __restrict__ cls,
    int32_t* __restrict__ pending, int32_t* __restrict__ n_pending) {
  const int64_t total = n_full * n_local;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int den = kClasses - c;
  const double p = 1.0 / (double)den;
  bool defer = false;
  if (e < total) {
    const int64_t n_left = c == 0 ? (int64_t)__ldg(counts + e) : (int64_t)left[e];
    int64_t k = -1;
    if (n_left <= 0) {
      k = 0;
    } else if ((double)n_left * p < 10.0) {
      defer = true;
    } else {
      const int64_t t = e / n_local, rl = e - t * n_local;
      const Stream st = make_stream(seed, resample_lo + rl, class_cell_id(t, c), kPurposeBinomial);
      const BtrsConsts q = btrs_setup(n_left, p, p / (1.0 - p));
      double kd, inv_us, v;
      if (btrs_propose(q, stream_block(st, 0u), &kd, &inv_us, &v) > 0)
        k = (int64_t)kd;
      else
        defer = true;
    }
    if (!defer) {
      cls[e * kClasses + c] = (int32_t)k;
      left[e] = (int32_t)(n_left - k);
      if (c == kClasses - 2) cls[e * kClasses + c + 1] = (int32_t)(n_left - k);
    } else if (c == 0) {
      left[e] = (int32_t)n_left;         // the resolver reads the remainder from `left`
    }
  }
  // block-aggregated append of the deferred pairs
  __shared__ int s_base, s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  int slot = 0;
  if (defer) slot = atomicAdd(&s_count, 1);
  __syncthreads();
  if (threadIdx.x == 0 && s_count > 0) s_base = atomicAdd(n_pending, s_count);
  __syncthreads();
  if (defer) pending[s_base + slot] = (int32_t)e;      // e < 2^31 (checked by the host)
}

__global__ void __launch_bounds__(128, 8) k0_class_step_resolve_kernel(
    uint64_t seed, int64_t resample_lo, int64_t n_local, int c, int32_t* __restrict__ left,
    int32_t* __restrict__ cls, const int32_t* __restrict__ pending, const int32_t* __restrict__ n_pending) {
  const int n = *n_pending;
  const int den = kClasses - c;
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int64_t e = pending[i];
    const int64_t t = e / n_local, rl = e - t * n_local;
    const int64_t n_left = left[e];
    const Stream st = make_stream(seed, resample_lo + rl, class_cell_id(t, c), kPurposeBinomial);
    const int64_t k = binomial_ratio(n_left, 1, den, st);
    cls[e * kClasses + c] = (int32_t)k;
    left[e] = (int32_t)(n_left - k);
    if (c == kClasses - 2) cls[e * kClasses + c + 1] = (int32_t)(n_left - k);
  }
}

// ------------------------------------------------------------------ K1 --------
struct K1Params {
  const double* data[2];
  Plan plan;
  uint64_t seed;
  int64_t resample_lo;
  int64_t n_local;
  const int32_t* counts;  // [n_tiles][n_local]; null when the plan has one cell
  const int32_t* cls;     // [n_full][n_local][16] class occupancies of the full tiles
  RoundKeys rk;           // Philox round keys of `seed` (constant-bank operands)
  double* partial;        // [(t * NM + m)][n_local]
  int32_t* work_counter;
  int32_t chunk;          // resamples per work item
  int32_t n_chunks;
  int64_t n_items;
  int64_t smem_stride;    // elements between arrays in shared memory
};

// Accumulator bank of one lane: W independent FP64 chains per moment, combined in a
// fixed order (deterministic bits).
template <int MS>
struct LaneAcc {
  static constexpr int NA = MomTraits<MS>::kArrays;
  static constexpr int NM = MomTraits<MS>::kMoments;
  static constexpr int W = MomTraits<MS>::kWays;
  double acc[NM][W];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int w = 0; w < W; ++w) acc[m][w] = 0.0;
  }
  // row at byte offset `off` of the resident cell
  __device__ __forceinline__ void take(int j, const unsigned char* __restrict__ s_bytes,
                                       uint32_t stride_bytes, uint32_t off) {
    double v[NA], c[NM];
    v[0] = *reinterpret_cast<const double*>(s_bytes + off);
    if (NA == 2) v[NA - 1] = *reinterpret_cast<const double*>(s_bytes + stride_bytes + off);
    contribution<MS>(v, c);
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[m][j % W] += c[m];
  }
  // Fast addressing: `off` is the byte offset inside array 0 and the shared-window
  // address of the cell is the compile-time constant kFastSmemBase, so the load is a
  // single LDS.64 [R + imm] with no address arithmetic.
  __device__ __forceinline__ void take_fast(int j, uint32_t off) {
    double v[NA], c[NM];
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v[0]) : "r"(off), "n"(kFastSmemBase));
    if (NA == 2)
      asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v[NA - 1]) : "r"(off), "n"(kFastSmemBase + kFastStride2));
    contribution<MS>(v, c);
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[m][j % W] += c[m];
  }
  __device__ __forceinline__ void finish(double* tot) {
#pragma unroll
    for (int m = 0; m < NM; ++m) {
#pragma unroll
      for (int w = W / 2; w > 0; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; ++i) acc[m][i] += acc[m][i + w];
      tot[m] = warp_sum(acc[m][0]);
    }
  }
};

// Full tile: lane = (class c = lane & 15, half h = lane >> 4).  The lane draws the
// blocks q = h, h+2, ... of its class stream; every LDS.64 of a half-warp touches 16
// distinct bank pairs (conflict-free).  12 draws per Philox block.
template <int MS, bool FAST>
__device__ __forceinline__ void class_accumulate(const unsigned char* __restrict__ s_bytes,
                                                 uint32_t stride_bytes, const RoundKeys& rk,
                                                 const Stream& st, int64_t n_c, uint32_t m7,
                                                 uint32_t cbits, uint32_t h, double* tot) {
  LaneAcc<MS> A;
  A.clear();
  // The lane owns blocks q = h, h+2, ... of the ceil(n_c / 12) blocks of its class.  All but
  // its last one are whole and run unpredicated; the last one (whole or partial) runs in a
  // predicated epilogue that every lane of the warp enters together, so a partial block
  // never costs a separate, half-empty pass.
  const uint32_t n_blocks = (uint32_t)((n_c + kClassDrawsPerBlock - 1) / kClassDrawsPerBlock);
  const uint32_t owned = n_blocks > h ? (n_blocks - h + 1u) >> 1 : 0u;
  const uint32_t q_last = h + 2u * (owned - 1u);          // meaningful when owned > 0
#pragma unroll(kClassUnroll)
  for (uint32_t q = h; q < q_last && owned != 0u; q += 2) {
    const Words4 w = stream_block(st, q, rk);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t f[3];
      class_fields_shifted(w.w[i], m7, f);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (FAST)
          A.take_fast(3 * i + k, f[k] | cbits);
        else
          A.take(3 * i + k, s_bytes, stride_bytes, f[k] | cbits);
      }
    }
  }
  if (owned != 0u) {
    const int valid = (int)min((int64_t)kClassDrawsPerBlock, n_c - (int64_t)q_last * kClassDrawsPerBlock);
    const Words4 w = stream_block(st, q_last, rk);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint32_t f[3];
      class_fields_shifted(w.w[i], m7, f);
#pragma unroll
      for (int k = 0; k < 3; ++k)
        if (3 * i + k < valid) {
          if (FAST)
            A.take_fast(3 * i + k, f[k] | cbits);
          else
            A.take(3 * i + k, s_bytes, stride_bytes, f[k] | cbits);
        }
    }
  }
  A.finish(tot);
}

// General cell (single-cell plans and the remainder cell): lanes take the blocks
// q = lane, lane+32, ... of the cell stream; 4 Lemire draws per block.
template <int MS>
__device__ __forceinline__ void lemire_accumulate(const unsigned char* __restrict__ s_bytes,
                                                  uint32_t stride_bytes, const RoundKeys& rk,
                                                  const Stream& st, int64_t n, uint32_t size, int lane,
                                                  double* tot) {
  LaneAcc<MS> A;
  A.clear();
  auto take = [&](int j, int64_t, uint32_t idx) { A.take(j, s_bytes, stride_bytes, idx * 8u); };
  const uint32_t n_full = (uint32_t)(n / 4);
  const uint32_t thresh = lemire_threshold(size);
#pragma unroll 2
  for (uint32_t q = lane; q < n_full; q += 32) lemire_block<false>(st, q, size, thresh, n, take, &rk);
  if ((int64_t)n_full * 4 < n && (n_full & 31u) == (uint32_t)lane)
    lemire_block<true>(st, n_full, size, thresh, n, take, &rk);
  A.finish(tot);
}

template <int MS, bool FAST>
__global__ void __launch_bounds__(kK1Threads, 1) k1_resample_kernel(const K1Params P) {
  constexpr int NA = MomTraits<MS>::kArrays;
  constexpr int NM = MomTraits<MS>::kMoments;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  volatile int* s_item = reinterpret_cast<volatile int*>(smem_raw + 16);
  double* s_data = reinterpret_cast<double*>(smem_raw + kSmemHeader);
  const unsigned char* s_bytes = smem_raw + kSmemHeader;
  const uint32_t stride_bytes = (uint32_t)(P.smem_stride * 8);
  const uint32_t m7 = ((1u << (P.plan.log2_tile - 4)) - 1u) << 7;   // class row mask, pre-shifted

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  if (FAST && smem_u32(smem_raw) != 1024u) __trap();   // fast addressing assumption
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();

  uint32_t phase = 0;
  int64_t resident = -1;
  for (;;) {
    if (tid == 0) *s_item = atomicAdd(P.work_counter, 1);
    __syncthreads();  // publishes the item; every warp is done with the previous one
    const int64_t item = *s_item;
    __syncthreads();  // nobody reads the slot after this point
    if (item >= P.n_items) break;
    const int64_t t = item / P.n_chunks;
    const int64_t c = item - t * P.n_chunks;
    const int64_t sz = P.plan.cell_size(t);
    const int64_t row0 = t < P.plan.n_full ? (t << P.plan.log2_tile) : (P.plan.n_full << P.plan.log2_tile);

    if (t != resident) {
      // bulk (TMA) copy of the 16-byte-aligned body, plain loads for what is left
      bool aligned = true;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        aligned = aligned && ((reinterpret_cast<uintptr_t>(P.data[a] + row0) & 15u) == 0);
      const uint32_t body_bytes = aligned ? (uint32_t)((sz * 8) & ~(int64_t)15) : 0u;
      if (body_bytes != 0 && tid == 0) {
        fence_proxy_async();
        mbar_expect_tx(bar, body_bytes * NA);
#pragma unroll
        for (int a = 0; a < NA; ++a)
          bulk_copy_g2s(s_data + a * P.smem_stride, P.data[a] + row0, body_bytes, bar);
      }
      for (int64_t i = body_bytes / 8 + tid; i < sz; i += kK1Threads) {
#pragma unroll
        for (int a = 0; a < NA; ++a) s_data[a * P.smem_stride + i] = __ldg(P.data[a] + row0 + i);
      }
      if (body_bytes != 0) {
        mbar_wait(bar, phase);
        phase ^= 1u;
      }
      __syncthreads();
      resident = t;
    }

    const int64_t r_begin = c * P.chunk;
    const int64_t r_end = min(r_begin + (int64_t)P.chunk, P.n_local);
    const bool full_tile = t < P.plan.n_full;
    for (int64_t rl = r_begin + warp; rl < r_end; rl += kK1Warps) {
      double tot[NM];
      if (full_tile) {
        const int c = lane & (kClasses - 1);
        const int64_t n_c = (int64_t)__ldg(P.cls + (t * P.n_local + rl) * kClasses + c);
        const Stream st = make_stream(P.seed, P.resample_lo + rl, class_cell_id(t, c), kPurposeIndex);
        class_accumulate<MS, FAST>(s_bytes, stride_bytes, P.rk, st, n_c, m7, (uint32_t)c << 3,
                             (uint32_t)lane >> 4, tot);
      } else {
        const int64_t n = P.counts ? (int64_t)__ldg(P.counts + t * P.n_local + rl) : P.plan.n_rows;
        const Stream st = make_stream(P.seed, P.resample_lo + rl, (uint32_t)t, kPurposeIndex);
        lemire_accumulate<MS>(s_bytes, stride_bytes, P.rk, st, n, (uint32_t)sz, lane, tot);
      }
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < NM; ++m) P.partial[(t * NM + m) * P.n_local + rl] = tot[m];
      }
    }
  }
}

// Sums the per-cell partials in cell order and finishes the statistic (out[o][n_local] for a
// statistic with several outputs).
__global__ void __launch_bounds__(256) k1_finish_kernel(const double* __restrict__ partial,
                                                        int64_t n_tiles, int nm, int64_t n_local,
                                                        int stat, double count,
                                                        const double* __restrict__ center,
                                                        double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_local) return;
  double m[kMaxMoments];
  for (int k = 0; k < nm; ++k) {
    double acc = 0.0;
    for (int64_t t = 0; t < n_tiles; ++t) acc += partial[(t * nm + k) * n_local + r];
    m[k] = acc;
  }
  const int n_out = stat_outputs(stat);
  for (int o = 0; o < n_out; ++o)
    out[(int64_t)o * n_local + r] = finish_stat(stat_component(stat, o), m, count, center);
}

// ------------------------------------------------------------------ K1i -------
// Parity mode: index sets supplied by the caller; one warp per set.
template <int MS>
__global__ void __launch_bounds__(256) k1_indexed_kernel(const double* __restrict__ d0,
                                                         const double* __restrict__ d1,
                                                         int64_t n_rows,
                                                         const int64_t* __restrict__ indices,
                                                         int64_t n_sets,
                                                         double* __restrict__ partial) {
  constexpr int NA = MomTraits<MS>::kArrays;
  constexpr int NM = MomTraits<MS>::kMoments;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t set = warp_global; set < n_sets; set += n_warps) {
    const int64_t* idx = indices + set * n_rows;
    double acc[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[m] = 0.0;
    for (int64_t j = lane; j < n_rows; j += 32) {
      const int64_t i = idx[j];
      double v[NA], c[NM];
      v[0] = d0[i];
      if (NA == 2) v[NA - 1] = d1[i];
      contribution<MS>(v, c);
#pragma unroll
      for (int m = 0; m < NM; ++m) acc[m] += c[m];
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const double tot = warp_sum(acc[m]);
      if (lane == 0) partial[(int64_t)m * n_sets + set] = tot;
    }
  }
}

// ------------------------------------------------------------------ K2 --------
// Writes the canonical index array of each resample: cells in order, inside a
// cell the draws in stream-position order — exactly what K1 consumes.
__global__ void __launch_bounds__(256) k2_fill_indices_kernel(Plan plan, uint64_t seed,
                                                              int64_t resample_lo, int64_t n_local,
                                                              const int32_t* __restrict__ counts,
                                                              const int32_t* __restrict__ cls,
                                                              int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t m7 = ((1u << (plan.log2_tile - 4)) - 1u) << 7;
  for (int64_t rl = warp_global; rl < n_local; rl += n_warps) {
    int64_t* row = out + rl * plan.n_rows;
    int64_t offset = 0;
    for (int64_t t = 0; t < plan.n_tiles; ++t) {
      const int64_t sz = plan.cell_size(t);
      const int64_t row0 = t < plan.n_full ? (t << plan.log2_tile) : (plan.n_full << plan.log2_tile);
      auto put = [&](int, int64_t pos, uint32_t idx) { row[offset + pos] = row0 + (int64_t)idx; };
      if (t < plan.n_full) {
        for (int c = 0; c < kClasses; ++c) {
          const int64_t n_c = cls[(t * n_local + rl) * kClasses + c];
          const Stream st = make_stream(seed, resample_lo + rl, class_cell_id(t, c), kPurposeIndex);
          const uint32_t nblk = (uint32_t)((n_c + kClassDrawsPerBlock - 1) / kClassDrawsPerBlock);
          for (uint32_t q = lane; q < nblk; q += 32) class_block<true>(st, q, m7, c, n_c, put);
          offset += n_c;
        }
      } else {
        const int64_t n = counts ? (int64_t)counts[t * n_local + rl] : plan.n_rows;
        const Stream st = make_stream(seed, resample_lo + rl, (uint32_t)t, kPurposeIndex);
        const uint32_t nblk = (uint32_t)((n + 3) / 4);
        const uint32_t thresh = lemire_threshold((uint32_t)sz);
        for (uint32_t q = lane; q < nblk; q += 32)
          lemire_block<true>(st, q, (uint32_t)sz, thresh, n, put);
        offset += n;
      }
    }
  }
}

// ------------------------------------------------------------------ K2b -------
// Moving-block bootstrap indices (bootstrap.py:747-787): per resample ceil(N/L)
// block starts uniform on [0, start_bound) — Lemire draws from the kPurposeBlock
// stream — expanded to start+0..L-1, truncated to N, taken mod N when wrapping.
__global__ void __launch_bounds__(256) k2_moving_block_kernel(uint64_t seed, int64_t n_rows,
                                                              int64_t block_length, int64_t start_bound,
                                                              int wrap, int64_t resample_lo,
                                                              int64_t n_local, int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_blocks = (n_rows + block_length - 1) / block_length;
  const uint32_t thresh = lemire_threshold((uint32_t)start_bound);
  for (int64_t rl = warp_global; rl < n_local; rl += n_warps) {
    int64_t* row = out + rl * n_rows;
    const Stream st = make_stream(seed, resample_lo + rl, 0u, kPurposeBlock);
    const uint32_t n_q = (uint32_t)((n_blocks + 3) / 4);
    for (uint32_t q = lane; q < n_q; q += 32) {
      lemire_block<true>(st, q, (uint32_t)start_bound, thresh, n_blocks,
                         [&](int, int64_t b, uint32_t start) {
                           for (int64_t k = 0; k < block_length; ++k) {
                             const int64_t pos = b * block_length + k;
                             if (pos >= n_rows) break;
                             const int64_t v = (int64_t)start + k;
                             row[pos] = wrap ? v % n_rows : v;
                           }
                         });
    }
  }
}

// ------------------------------------------------------------------ K2c -------
// Sort keys of a random permutation: keys[d][i] = 64 random bits of (seed, sample d,
// position i); padding sorts last.  Sorting (key, i) pairs per sample (psort kernels)
// and keeping the first `size` row numbers gives a uniform subsample without
// replacement (bootstrap.py:717-744).
__global__ void __launch_bounds__(256) k2_shuffle_keys_kernel(uint64_t seed, int64_t n_rows,
                                                              int64_t n_pad, int64_t sample_lo,
                                                              int64_t n_local, uint64_t* __restrict__ keys,
                                                              uint32_t* __restrict__ idx) {
  const int64_t per = (n_pad + 1) / 2;                      // Philox blocks per sample (2 keys each)
  const int64_t total = per * n_local;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t d = e / per, q = e - d * per;
    const Stream st = make_stream(seed, sample_lo + d, 0u, kPurposeShuffle);
    const Words4 w = stream_block(st, (uint32_t)q);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t i = 2 * q + h;
      if (i < n_pad) {
        const uint64_t key = ((uint64_t)w.w[2 * h] << 32) | w.w[2 * h + 1];
        keys[d * n_pad + i] = i < n_rows ? (key >> 1) : ~0ull;      // top bit clear: below the padding
        idx[d * n_pad + i] = i < n_rows ? (uint32_t)i : 0xFFFFFFFFu;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k2_subsample_store_kernel(const uint32_t* __restrict__ idx,
                                                                 int64_t n_pad, int64_t size,
                                                                 int64_t n_local, int64_t* __restrict__ out) {
  const int64_t total = size * n_local;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t d = e / size, i = e - d * size;
    out[e] = (int64_t)idx[d * n_pad + i];
  }
}

__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ data,
                                                     const int64_t* __restrict__ idx, int64_t n_idx,
                                                     double* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_idx; i += stride)
    out[i] = data[idx[i]];
}

}  // namespace b200boot
