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
#ifndef B2_K1_PIPE
#define B2_K1_PIPE 0
#endif
#ifndef B2_K1_PAIR
#define B2_K1_PAIR 1    // two resamples per warp in full tiles (measured: -1.0 %)
#endif
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

// Class occupancies of the full tiles, inspection / K2 form: cls[(t * n_local + r) * 16 + c]
// as int32; one thread walks the 15-node split tree of one (tile, resample) pair.
__global__ void __launch_bounds__(128) k0_class_counts_kernel(uint64_t seed, int64_t resample_lo,
                                                              int64_t n_local, int64_t n_full,
                                                              const int32_t* __restrict__ counts,
                                                              int32_t* __restrict__ cls) {
  const int64_t total = n_full * n_local;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t t = e / n_local, rl = e - t * n_local;
    class_count_tree<int32_t>(seed, resample_lo + rl, t, (int64_t)__ldg(counts + e), cls + e * kClasses, 1);
  }
}

// Tables of the even-split sampler for every n below kHalfTableSize: the BTRS constants of
// Bin(n, 1/2) (entries below kHalfBtrsMin are never read) and log(n!).  32 + 8 bytes per
// entry, 1.25 MB: L2-resident for the tree kernel.  Built by the sampler's own functions, so
// a table read and a direct evaluation give the same bits.
constexpr int kHalfTableSize = 32768;
struct HalfTables {
  HalfConsts* consts;
  double* log_fact;
};
__global__ void __launch_bounds__(256) k0_half_table_kernel(HalfTables tab) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= kHalfTableSize) return;
  tab.consts[n] = half_setup(n < kHalfBtrsMin ? kHalfBtrsMin : n);
  tab.log_fact[n] = log_factorial((double)n);
}

// Resample-path form of the class split: a CTA takes a batch of (tile, resample) pairs, keeps
// their 16-slot records in shared memory and walks the four levels of the split tree.  At a
// level every node of the batch is independent and the work is kept dense in three passes:
//   pass 0, all nodes: the first proposal (table lookup + one BTRS proposal + squeeze), which
//           settles ~80 % of them; the others go to a shared-memory queue;
//   pass 1, the queue: the full acceptance test of that first proposal and, if it fails, the
//           second proposal with its squeeze; what is still open (~2 %) goes to a second queue;
//   pass 2, the second queue: the sampler's loop from the second attempt on.
// Each pass evaluates exactly what binomial_half evaluates for that attempt, so the values are
// those of class_count_tree.  One launch, no global scratch.  Records are uint16: a count beyond
// 65535 (never for 2^14-row tiles) saturates and raises *overflow, which the finisher turns into
// NaN statistics.
constexpr int kTreeThreads = 256;
#ifndef B2_TREE_BATCH
#define B2_TREE_BATCH 768     // measured: 256 -> 4.2 ms, 512 -> 3.5, 768 -> 3.3 (K0 of 2e4 resamples at N = 1e7)
#endif
constexpr int kTreeBatch = B2_TREE_BATCH;      // pairs per CTA pass (node numbers of a level fit 16 bits)
constexpr int kTreeRecStride = kClasses + 1;   // odd stride: slot reads of consecutive pairs spread over banks
constexpr int kTreeQueue2 = 2 * kTreeBatch;    // capacity of the second queue (expected use: 2 % of the nodes)

struct TreeLookup {
  const HalfConsts* __restrict__ consts;
  const double* __restrict__ log_fact;
  __device__ __forceinline__ HalfConsts operator()(int64_t n) const {
    if (n < kHalfTableSize) {
      const double2 lo = __ldg(reinterpret_cast<const double2*>(consts + n));
      const double2 hi = __ldg(reinterpret_cast<const double2*>(consts + n) + 1);
      return HalfConsts{lo.x, lo.y, hi.x, hi.y};
    }
    return half_setup(n);
  }
  __device__ __forceinline__ double lf(double j) const {
    return j < (double)kHalfTableSize ? __ldg(log_fact + (int)j) : log_factorial(j);
  }
};

__global__ void __launch_bounds__(kTreeThreads, 3) k0_class_tree_kernel(
    uint64_t seed, int64_t resample_lo, int64_t n_local, int64_t n_full,
    const int32_t* __restrict__ counts, HalfTables tab, uint16_t* __restrict__ cls,
    int32_t* __restrict__ overflow, const RoundKeys rk) {
  __shared__ uint16_t s_rec[kTreeBatch * kTreeRecStride];
  __shared__ uint32_t s_tile[kTreeBatch], s_res[kTreeBatch];
  __shared__ uint16_t s_q1[kTreeBatch * 8], s_q2[kTreeQueue2];
  __shared__ int s_qn[2];
  const TreeLookup look{tab.consts, tab.log_fact};
  const auto lf = [&look](double j) { return look.lf(j); };
  const int tid = threadIdx.x;
  const int64_t total = n_full * n_local;
  const int64_t n_batches = (total + kTreeBatch - 1) / kTreeBatch;
  bool sat = false;
  for (int64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const int64_t e0 = batch * kTreeBatch;
    const int np = (int)min((int64_t)kTreeBatch, total - e0);
    for (int p = tid; p < np; p += kTreeThreads) {
      const int64_t e = e0 + p;
      const int64_t t = e / n_local;
      s_tile[p] = (uint32_t)t;
      s_res[p] = (uint32_t)(e - t * n_local);
      const int32_t n_t = __ldg(counts + e);
      sat = sat || n_t > 65535;
      s_rec[p * kTreeRecStride] = (uint16_t)min(n_t, 65535);
    }
    if (tid < 2) s_qn[tid] = 0;
    __syncthreads();
#pragma unroll 1
    for (int level = 0; level < 4; ++level) {
      const int span = kClasses >> level;
      const int n_nodes = np << level;
      auto node_slot = [&](int i) { return s_rec + (i >> level) * kTreeRecStride + (i & ((1 << level) - 1)) * span; };
      auto node_stream = [&](int i) {
        const int p = i >> level;
        return make_stream(seed, resample_lo + (int64_t)s_res[p],
                           class_node_cell_id((int64_t)s_tile[p], (1 << level) + (i & ((1 << level) - 1))),
                           kPurposeBinomial);
      };
      auto settle = [&](uint16_t* slot, int64_t n, int64_t k) {
        slot[0] = (uint16_t)k;
        slot[span / 2] = (uint16_t)(n - k);
      };
      // pass 0: first proposal of every node
      for (int i = tid; i < n_nodes; i += kTreeThreads) {
        uint16_t* slot = node_slot(i);
        const int64_t n = (int64_t)*slot;
        const Words4 w = stream_block(node_stream(i), 0u, rk);
        if (n < kHalfBtrsMin) {
          settle(slot, n, binomial_half_small(n, w));
        } else {
          double k, inv_us, v;
          if (half_propose((double)n, look(n), w, &k, &inv_us, &v) > 0)
            settle(slot, n, (int64_t)k);
          else
            s_q1[atomicAdd(&s_qn[0], 1)] = (uint16_t)i;
        }
      }
      __syncthreads();
      // pass 1: full test of the first proposal, else second proposal
      const int q1n = s_qn[0];
      for (int q = tid; q < q1n; q += kTreeThreads) {
        const int i = s_q1[q];
        uint16_t* slot = node_slot(i);
        const int64_t n = (int64_t)*slot;
        const double nd = (double)n;
        const HalfConsts h = look(n);
        const Stream st = node_stream(i);
        double k, inv_us, v;
        int verdict = half_propose(nd, h, stream_block(st, 0u, rk), &k, &inv_us, &v);
        bool done = verdict == 0 && half_accept(nd, h, k, inv_us, v, lf);
        if (!done) {
          verdict = half_propose(nd, h, stream_block(st, 1u, rk), &k, &inv_us, &v);
          done = verdict > 0;
        }
        if (done) {
          settle(slot, n, (int64_t)k);
        } else {
          const int pos = atomicAdd(&s_qn[1], 1);
          if (pos < kTreeQueue2)
            s_q2[pos] = (uint16_t)i;
          else                                    // queue full (never in practice): finish here
            settle(slot, n, binomial_half_from(n, h, st, 1u, lf));
        }
      }
      __syncthreads();
      // pass 2: whatever is still open, from the second attempt on
      const int q2n = min(s_qn[1], kTreeQueue2);
      for (int q = tid; q < q2n; q += kTreeThreads) {
        const int i = s_q2[q];
        uint16_t* slot = node_slot(i);
        const int64_t n = (int64_t)*slot;
        settle(slot, n, binomial_half_from(n, look(n), node_stream(i), 1u, lf));
      }
      __syncthreads();
      if (tid < 2) s_qn[tid] = 0;
      __syncthreads();
    }
    // packed store: two classes per 32-bit word, coalesced over the batch
    uint32_t* out = reinterpret_cast<uint32_t*>(cls + e0 * kClasses);
    for (int i = tid; i < np * (kClasses / 2); i += kTreeThreads) {
      const int p = i >> 3, c = (i & 7) * 2;
      out[i] = (uint32_t)s_rec[p * kTreeRecStride + c] | ((uint32_t)s_rec[p * kTreeRecStride + c + 1] << 16);
    }
    __syncthreads();   // records are reloaded by the next batch
  }
  if (sat) atomicExch(overflow, 1);
}

// ------------------------------------------------------------------ K1 --------
struct K1Params {
  const double* data[2];
  Plan plan;
  uint64_t seed;
  int64_t resample_lo;
  int64_t n_local;
  const int32_t* counts;  // [n_tiles][n_local]; null when the plan has one cell
  const uint16_t* cls;    // [n_full][n_local][16] class occupancies of the full tiles
  RoundKeys rk;           // Philox round keys of `seed` (constant-bank operands)
  double* partial;        // [(t * NM + m)][n_local]
  int32_t* work_counter;
  int32_t chunk;          // resamples per work item
  int32_t n_chunks;
  int64_t n_items;
  int64_t smem_stride;    // elements between arrays in shared memory
  // single-cell plans: the per-resample sums are the totals, so the warp finishes the
  // statistic itself (no partial table, no finisher launch): direct_out[o * n_local + r]
  double* direct_out;
  const double* center;   // per-array means of pre-centred rows (finish_stat), may be null
  int32_t stat;
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
    // Pure asm (so the compiler schedules the twelve gathers of a block freely).  What keeps
    // it behind the barrier that publishes a freshly copied cell is a data dependency: `off`
    // contains the lane's class bits, which the kernel passes through an ordered
    // (volatile, memory-clobbering) empty asm right after that barrier (publish_token).
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v[0]) : "r"(off), "n"(kFastSmemBase));
    if (NA == 2)
      asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v[NA - 1]) : "r"(off), "n"(kFastSmemBase + kFastStride2));
    contribution<MS>(v, c);
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[m][j % W] += c[m];
  }
  // lane total -> sum over the warp (HALF: over the lane's half-warp)
  template <bool HALF = false>
  __device__ __forceinline__ void finish(double* tot) {
#pragma unroll
    for (int m = 0; m < NM; ++m) {
#pragma unroll
      for (int w = W / 2; w > 0; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; ++i) acc[m][i] += acc[m][i + w];
      double v = acc[m][0];
#pragma unroll
      for (int off = HALF ? 8 : 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      tot[m] = v;
    }
  }
};

// Returns x unchanged, but as a value the compiler must consider (re)defined here, after every
// earlier memory operation and barrier: anything computed from it cannot be hoisted above.
__device__ __forceinline__ uint32_t publish_token(uint32_t x) {
  asm volatile("" : "+r"(x) : : "memory");
  return x;
}

// Full tile: lane = (class c = lane & 15, half h = lane >> 4); every LDS.64 of a half-warp
// touches 16 distinct bank pairs (conflict-free).  12 draws per Philox block.  The lane
// draws the blocks q0, q0 + qstep, ... below n_blocks of its class stream:
//   split mode (qstep 2): both halves work on one resample, h takes every second block;
//   pair mode  (qstep 1): each half-warp owns its own resample (two resamples per warp), so
//                         the per-resample prologue / epilogue / butterfly is paid once per
//                         ~85 blocks instead of once per ~43.
template <int MS, bool FAST>
__device__ __forceinline__ void class_take12(LaneAcc<MS>& A, const Words4& w, uint32_t m7, uint32_t cbits,
                                             const unsigned char* __restrict__ s_bytes,
                                             uint32_t stride_bytes) {
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

template <int MS, bool FAST>
__device__ __forceinline__ void class_accumulate(const unsigned char* __restrict__ s_bytes,
                                                 uint32_t stride_bytes, const RoundKeys& rk,
                                                 const Stream& st, uint32_t n_c, uint32_t m7,
                                                 uint32_t cbits, uint32_t q0, uint32_t qstep,
                                                 LaneAcc<MS>& A) {
  A.clear();
  // All owned blocks but the last are whole and run unpredicated; the last one (whole or
  // partial) runs in a predicated epilogue that every lane of the warp enters together, so a
  // partial block never costs a separate, half-empty pass.
  const uint32_t n_blocks = (n_c + (uint32_t)kClassDrawsPerBlock - 1u) / (uint32_t)kClassDrawsPerBlock;
  const uint32_t owned = n_blocks > q0 ? (n_blocks - q0 + qstep - 1u) / qstep : 0u;
  const uint32_t q_last = q0 + qstep * (owned - 1u);          // meaningful when owned > 0
#if B2_K1_PIPE
  // software pipeline: the Philox rounds of block q + qstep are independent of the gathers
  // of block q, so one loop body carries both and the instruction mix of a warp is uniform
  // (multiplies, logic, loads and FP64 adds interleaved) instead of alternating between a
  // multiply-bound and a load-bound phase
  Words4 w = stream_block(st, q0, rk);
  for (uint32_t q = q0; q < q_last && owned != 0u; q += qstep) {
    const Words4 wn = stream_block(st, q + qstep, rk);
    class_take12<MS, FAST>(A, w, m7, cbits, s_bytes, stride_bytes);
    w = wn;
  }
#else
#pragma unroll(kClassUnroll)
  for (uint32_t q = q0; q < q_last && owned != 0u; q += qstep) {
    const Words4 w = stream_block(st, q, rk);
    class_take12<MS, FAST>(A, w, m7, cbits, s_bytes, stride_bytes);
  }
  const Words4 w = stream_block(st, q_last, rk);
#endif
  if (owned != 0u) {
    const int valid = (int)min((uint32_t)kClassDrawsPerBlock, n_c - q_last * (uint32_t)kClassDrawsPerBlock);
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
    // the lane's class bits (byte offset of its bank pair), tied to this point of the program
    const uint32_t cbits = publish_token((uint32_t)(lane & (kClasses - 1)) << 3);

    const int64_t r_begin = c * P.chunk;
    const int64_t r_end = min(r_begin + (int64_t)P.chunk, P.n_local);
    const bool full_tile = t < P.plan.n_full;
#if B2_K1_PAIR
    if (full_tile) {
      // two resamples per warp: half-warp h owns resample rl0 + h, lane c of it the class c
      const int c = lane & (kClasses - 1);
      const int64_t h = lane >> 4;
      for (int64_t rl0 = r_begin + 2 * warp; rl0 < r_end; rl0 += 2 * kK1Warps) {
        const int64_t rl = rl0 + h;
        const bool live = rl < r_end;
        const uint32_t n_c = live ? (uint32_t)__ldg(P.cls + (t * P.n_local + rl) * kClasses + c) : 0u;
        const Stream st = make_stream(P.seed, P.resample_lo + rl, class_cell_id(t, c), kPurposeIndex);
        LaneAcc<MS> A;
        class_accumulate<MS, FAST>(s_bytes, stride_bytes, P.rk, st, n_c, m7, cbits, 0u, 1u, A);
        double tot[NM];
        A.template finish<true>(tot);
        if (c == 0 && live) {
#pragma unroll
          for (int m = 0; m < NM; ++m) P.partial[(t * NM + m) * P.n_local + rl] = tot[m];
        }
      }
      continue;
    }
#endif
    for (int64_t rl = r_begin + warp; rl < r_end; rl += kK1Warps) {
      double tot[NM];
      if (full_tile) {
        const int c = lane & (kClasses - 1);
        const uint32_t n_c = (uint32_t)__ldg(P.cls + (t * P.n_local + rl) * kClasses + c);
        const Stream st = make_stream(P.seed, P.resample_lo + rl, class_cell_id(t, c), kPurposeIndex);
        LaneAcc<MS> A;
        class_accumulate<MS, FAST>(s_bytes, stride_bytes, P.rk, st, n_c, m7, cbits,
                                   (uint32_t)lane >> 4, 2u, A);
        A.finish(tot);
      } else {
        const int64_t n = P.counts ? (int64_t)__ldg(P.counts + t * P.n_local + rl) : P.plan.n_rows;
        const Stream st = make_stream(P.seed, P.resample_lo + rl, (uint32_t)t, kPurposeIndex);
        lemire_accumulate<MS>(s_bytes, stride_bytes, P.rk, st, n, (uint32_t)sz, lane, tot);
      }
      if (lane == 0) {
        if (P.direct_out != nullptr) {
          double m[kMaxMoments];
#pragma unroll
          for (int k = 0; k < NM; ++k) m[k] = 0.0 + tot[k];      // as the finisher's sum over one cell
          const int n_out = stat_outputs(P.stat);
          for (int o = 0; o < n_out; ++o)
            P.direct_out[(int64_t)o * P.n_local + rl] =
                finish_stat(stat_component(P.stat, o), m, (double)P.plan.n_rows, P.center);
        } else {
#pragma unroll
          for (int m = 0; m < NM; ++m) P.partial[(t * NM + m) * P.n_local + rl] = tot[m];
        }
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
                                                        double* __restrict__ out,
                                                        const int32_t* __restrict__ overflow) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_local) return;
  // a class count that did not fit its slot (see k0_class_tree_kernel): no valid statistic
  const bool poisoned = overflow != nullptr && *overflow != 0;
  double m[kMaxMoments];
  for (int k = 0; k < nm; ++k) {
    double acc = 0.0;
    for (int64_t t = 0; t < n_tiles; ++t) acc += partial[(t * nm + k) * n_local + r];
    m[k] = acc;
  }
  const int n_out = stat_outputs(stat);
  for (int o = 0; o < n_out; ++o)
    out[(int64_t)o * n_local + r] = poisoned ? nan("") : finish_stat(stat_component(stat, o), m, count, center);
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
constexpr int kK2StagePitch = kClassDrawsPerBlock + 1;   // odd pitch: the lanes' rows start in distinct banks
__global__ void __launch_bounds__(256) k2_fill_indices_kernel(Plan plan, uint64_t seed,
                                                              int64_t resample_lo, int64_t n_local,
                                                              const int32_t* __restrict__ counts,
                                                              const int32_t* __restrict__ cls,
                                                              int64_t* __restrict__ out) {
  __shared__ uint32_t s_stage[256 / 32][32 * kK2StagePitch];
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t m7 = ((1u << (plan.log2_tile - 4)) - 1u) << 7;
  // a warp per (resample, cell): long arrays have few resamples per batch and many cells, and a
  // warp per resample left the device empty (N = 1e6, 32 resamples: 58 GB/s of index writes).  The
  // cell's first position in the resample is the sum of the earlier cells' occupancies.
  const int64_t n_units = n_local * plan.n_tiles;
  for (int64_t unit = warp_global; unit < n_units; unit += n_warps) {
    const int64_t rl = unit / plan.n_tiles, t = unit - rl * plan.n_tiles;
    int64_t* row = out + rl * plan.n_rows;
    int64_t offset = 0;
    if (counts) {
      for (int64_t u = lane; u < t; u += 32) offset += (int64_t)counts[u * n_local + rl];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) offset += __shfl_xor_sync(0xffffffffu, offset, o);
    }
    const int64_t sz = plan.cell_size(t);
    const int64_t row0 = t < plan.n_full ? (t << plan.log2_tile) : (plan.n_full << plan.log2_tile);
    auto put = [&](int, int64_t pos, uint32_t idx) { row[offset + pos] = row0 + (int64_t)idx; };
    if (t < plan.n_full) {
      // A lane draws a whole block: twelve consecutive positions.  Stored from the lanes' own
      // registers that is a 96-byte stride between lanes (every store instruction touches 32
      // sectors for 8 bytes each: 0.5 TB/s of index writes); the warp's 384 draws of a round go
      // through a shared-memory transpose instead and leave as twelve 256-byte rows.
      uint32_t* stage = s_stage[threadIdx.x >> 5];
      for (int c = 0; c < kClasses; ++c) {
        const int64_t n_c = cls[(t * n_local + rl) * kClasses + c];
        const Stream st = make_stream(seed, resample_lo + rl, class_cell_id(t, c), kPurposeIndex);
        const uint32_t nblk = (uint32_t)((n_c + kClassDrawsPerBlock - 1) / kClassDrawsPerBlock);
        for (uint32_t q0 = 0; q0 < nblk; q0 += 32) {
          if (q0 + lane < nblk)
            class_block<false>(st, q0 + lane, m7, c, n_c,
                               [&](int j, int64_t, uint32_t idx) { stage[lane * kK2StagePitch + j] = idx; });
          __syncwarp();
          const int64_t base = (int64_t)q0 * kClassDrawsPerBlock;
#pragma unroll
          for (int k = 0; k < kClassDrawsPerBlock; ++k) {
            const int p = k * 32 + lane;
            if (base + p < n_c)
              row[offset + base + p] = row0 + (int64_t)stage[(p / kClassDrawsPerBlock) * kK2StagePitch + p % kClassDrawsPerBlock];
          }
          __syncwarp();
        }
        offset += n_c;
      }
    } else {
      const int64_t n = counts ? (int64_t)counts[t * n_local + rl] : plan.n_rows;
      const Stream st = make_stream(seed, resample_lo + rl, (uint32_t)t, kPurposeIndex);
      const uint32_t nblk = (uint32_t)((n + 3) / 4);
      const uint32_t thresh = lemire_threshold((uint32_t)sz);
      for (uint32_t q = lane; q < nblk; q += 32)
        lemire_block<true>(st, q, (uint32_t)sz, thresh, n, put);
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

namespace b200boot {

// ------------------------------------------------------------------ K1b -------
// Moving-block bootstrap as a resampling plan (bootstrap.py:747-787 fused with the gather and
// the statistic of bootstrap.py:429-432): per resample ceil(N / L) block starts — the very
// draws k2_moving_block_kernel materialises (kPurposeBlock stream, Lemire) — and the rows
// start .. start + L - 1 of every block, cut to N positions, modulo N when wrapping.  The
// gathers are L-contiguous, so the rows come coalesced from L2 / HBM; no index array exists.
// One warp per resample: blocks of at least 32 rows are walked by the lanes together, shorter
// blocks one per lane.  Per-resample sums never cross warps, so the bits do not depend on grid
// or shard.
template <int MS>
__global__ void __launch_bounds__(256) k1b_moving_block_kernel(const double* __restrict__ d0,
                                                               const double* __restrict__ d1, int64_t n_rows,
                                                               int64_t block_length, int64_t start_bound, int wrap,
                                                               uint64_t seed, int64_t resample_lo, int64_t n_local,
                                                               int stat, const double* __restrict__ center,
                                                               double* __restrict__ out) {
  constexpr int NA = MomTraits<MS>::kArrays;
  constexpr int NM = MomTraits<MS>::kMoments;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_blocks = (n_rows + block_length - 1) / block_length;
  const uint32_t n_q = (uint32_t)((n_blocks + 3) / 4);
  const uint32_t thresh = lemire_threshold((uint32_t)start_bound);
  for (int64_t rl = warp_global; rl < n_local; rl += n_warps) {
    const Stream st = make_stream(seed, resample_lo + rl, 0u, kPurposeBlock);
    double acc[NM];
#pragma unroll
    for (int m = 0; m < NM; ++m) acc[m] = 0.0;
    auto add_row = [&](int64_t idx) {
      double v[NA], c[NM];
      v[0] = d0[idx];
      if (NA == 2) v[NA - 1] = d1[idx];
      contribution<MS>(v, c);
#pragma unroll
      for (int m = 0; m < NM; ++m) acc[m] += c[m];
    };
    if (block_length >= 32) {
      // every lane derives the same four starts of a Philox block; the lanes share each block's rows
      for (uint32_t q = 0; q < n_q; ++q)
        lemire_block<true>(st, q, (uint32_t)start_bound, thresh, n_blocks, [&](int, int64_t b, uint32_t start) {
          const int64_t len = min(block_length, n_rows - b * block_length);
          for (int64_t k = lane; k < len; k += 32) {
            const int64_t v = (int64_t)start + k;
            add_row(wrap ? v % n_rows : v);
          }
        });
    } else {
      for (uint32_t q = lane; q < n_q; q += 32)
        lemire_block<true>(st, q, (uint32_t)start_bound, thresh, n_blocks, [&](int, int64_t b, uint32_t start) {
          const int64_t len = min(block_length, n_rows - b * block_length);
          for (int64_t k = 0; k < len; ++k) {
            const int64_t v = (int64_t)start + k;
            add_row(wrap ? v % n_rows : v);
          }
        });
    }
    double m[kMaxMoments];
#pragma unroll
    for (int k = 0; k < NM; ++k) m[k] = warp_sum(acc[k]);
    if (lane == 0) {
      const int n_out = stat_outputs(stat);
      for (int o = 0; o < n_out; ++o)
        out[(int64_t)o * n_local + rl] = finish_stat(stat_component(stat, o), m, (double)n_rows, center);
    }
  }
}

}  // namespace b200boot
