// K6i, pipelined form — the count-matrix x digit-matrix contraction of k6i_mma.cuh as a
// warp-specialised, persistent tcgen05 kernel.
//
// k6i_mma_kernel (first form) stages operands with all of the CTA's threads, has one TMEM
// accumulator and runs its epilogue between tiles: ncu shows the tensor pipe 10 % busy.  Here:
//   * both operands are first re-laid as IMAGES of their shared-memory tiles (k6i_image_kernel:
//     per (row tile, 128-byte K chunk) one contiguous block in SWIZZLE_128B order), so a stage is
//     filled by ONE 1-D bulk copy (cp.async.bulk, TMA) per operand — no tensor map, no thread
//     staging, no generic-proxy fence;
//   * roles: warp 0 = producer (one thread issues the bulk copies into a ring of stages, each with
//     a full / empty mbarrier pair), warp 1 = MMA issuer (one thread: four UMMA 128x256x32 per
//     stage, tcgen05.commit releases the stage), warps 4-19 = epilogue (four warpgroups share a
//     tile's column blocks: tcgen05.ld, exact recombination, FP64 scale, coalesced stores);
//   * two TMEM accumulators (2 x 256 columns = all of TMEM): the epilogue of tile i overlaps the
//     MMAs of tile i + 1;
//   * persistent: CTA b owns a contiguous range of (resample tile, column tile) units, so the
//     count tile stays resident across the column tiles of a resample tile (N <= 1024 + 128).
// Arithmetic, tile shapes and results are those of k6i_mma_kernel (bit-identical outputs).
//
// Measured on B200 at 1000 x 10 000 data, n = 1e4 (79 x 313 tiles of 128 x 256 x 1024, 167 per CTA):
// first form 3.45 ms; this kernel with four epilogue warps 1.40 ms (one warp per scheduler:
// latency-bound, issue 17 %); scales by shuffle instead of a global load per output 1.26 ms; sixteen
// epilogue warps 0.82-0.88 ms (tensor pipe 39-49 % busy).  Stage-elimination runs of that version
// (device time of the kernel alone): no operand copies 0.83 ms, no MMAs 0.74, no epilogue
// arithmetic 0.64, neither MMAs nor arithmetic 0.45, none of the three 0.35 ms — the barrier
// round trips of a three-stage ring (128 KB of the 227 KB hold the resident count tile) already
// cost 4.1 k cycles per tile against 4.1 k cycles of MMAs, so the stages overlap only partly.
// Tried without gain: clusters of two CTAs with the digit chunks multicast (1.03 vs 1.04 ms per
// call: L2 traffic is not what binds), column tiles taken in L2-sized blocks, streaming stores,
// int->double by the 2^52 trick instead of I2F (slower: the FP64 pipe is the busier one).  The
// next step is cta_group::2 (each CTA holds half of a digit chunk: twice the ring depth).
#pragma once
#include "k6i_mma.cuh"

namespace b200boot {

#ifndef B2_PIPE_EPI_GROUPS
#define B2_PIPE_EPI_GROUPS 4
#endif
constexpr int kPipeEpiGroups = B2_PIPE_EPI_GROUPS;        // epilogue warpgroups (4 warps each cover the 128 TMEM lanes)
constexpr int kPipeThreads = 128 + kPipeEpiGroups * 128;
constexpr int kPipeMaxStages = 6;
constexpr int kPipeMaxResidentChunks = 9;

template <int ROWS>
__global__ void __launch_bounds__(256) k6i_image_kernel(const unsigned char* __restrict__ src, int64_t src_rows,
                                                        int64_t k_pad, int64_t n_tiles,
                                                        unsigned char* __restrict__ dst) {
  const int64_t pieces = k_pad / 16;                       // 16-byte pieces per row
  const int64_t n_kchunks = k_pad / kI8KChunk;
  const int64_t total = n_tiles * ROWS * pieces;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t row = e / pieces, c = e - row * pieces;
    const int64_t t = row / ROWS, r = row - t * ROWS;
    const int64_t kc = c >> 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (row < src_rows) v = __ldg(reinterpret_cast<const uint4*>(src + row * k_pad + c * 16));
    *reinterpret_cast<uint4*>(dst + (t * n_kchunks + kc) * (int64_t)(ROWS * kI8KChunk) +
                              sw128_offset((uint32_t)r, (uint32_t)(c & 7) * 16u)) = v;
  }
}

struct K6iPipeParams {
  const unsigned char* a_img;   // [m_tiles][n_kchunks][16 KB]
  const unsigned char* b_img;   // [n_col_tiles][n_kchunks][32 KB]
  const double* scale;          // [n_col_tiles * 32]
  int64_t m_total, m_tiles;
  int32_t n_kchunks;
  int32_t n_stages;
  int64_t n_cols, n_col_tiles;
  double* out;                  // out[d * out_ld + r]
  int64_t out_ld;
  int32_t* out_i32;             // test hook: raw accumulators [m_total][n_col_tiles * 256], or null
  const int32_t* overflow;
  double out_divisor;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// A_RESIDENT: the CTA's count tile (all K chunks) stays in shared memory; stages hold B only.
template <bool A_RESIDENT>
__global__ void __launch_bounds__(kPipeThreads, 1) k6i_pipe_kernel(const K6iPipeParams P) {
  extern __shared__ unsigned char k6p_smem_raw[];
  unsigned char* smem = k6p_smem_raw + ((1024u - (smem_u32(k6p_smem_raw) & 1023u)) & 1023u);
  const int n_kchunks = P.n_kchunks;
  const int n_stages = P.n_stages;
  const uint32_t stage_bytes = A_RESIDENT ? (uint32_t)kI8BChunkBytes : (uint32_t)(kI8AChunkBytes + kI8BChunkBytes);
  unsigned char* s_a = smem;                                                   // resident count tile
  unsigned char* s_ring = smem + (A_RESIDENT ? (size_t)n_kchunks * kI8AChunkBytes : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)n_stages * stage_bytes);
  uint64_t* full = bars;                          // [n_stages]  producer -> MMA
  uint64_t* empty = bars + kPipeMaxStages;        // [n_stages]  MMA -> producer
  uint64_t* tfull = bars + 2 * kPipeMaxStages;    // [2]  MMA -> epilogue
  uint64_t* tempty = tfull + 2;                   // [2]  epilogue -> MMA
  uint64_t* afull = tempty + 2;                   // resident count tile loaded
  uint64_t* afree = afull + 1;                    // MMAs reading the resident count tile are done
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(afree + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < n_stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kPipeEpiGroups * 128);
    }
    mbar_init(afull, 1);
    mbar_init(afree, 1);
  }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // this CTA's units: unit = m_tile * n_col_tiles + col_tile, a contiguous range
  const int64_t n_units = P.m_tiles * P.n_col_tiles;
  const int64_t u_begin = n_units * blockIdx.x / gridDim.x;
  const int64_t u_end = n_units * (blockIdx.x + 1) / gridDim.x;
  auto tile_of = [&](int64_t u) { return u / P.n_col_tiles; };
  auto col_of = [&](int64_t u) { return u % P.n_col_tiles; };

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer ----------------
      uint32_t it = 0;
      int64_t cur_mt = -1;
      uint32_t a_loads = 0;
      for (int64_t u = u_begin; u < u_end; ++u) {
        const int64_t mt = tile_of(u), ct = col_of(u);
        if (A_RESIDENT && mt != cur_mt) {
          if (a_loads > 0) mbar_wait(afree, (a_loads - 1u) & 1u);        // previous tile's readers are done
          mbar_expect_tx(afull, (uint32_t)n_kchunks * kI8AChunkBytes);
          for (int kc = 0; kc < n_kchunks; ++kc)
            bulk_copy_g2s(s_a + (size_t)kc * kI8AChunkBytes,
                          P.a_img + (mt * n_kchunks + kc) * (int64_t)kI8AChunkBytes, kI8AChunkBytes, afull);
          ++a_loads;
          cur_mt = mt;
        }
        for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
          const uint32_t s = it % (uint32_t)n_stages;
          mbar_wait(&empty[s], ((it / (uint32_t)n_stages) & 1u) ^ 1u);
          unsigned char* dst = s_ring + (size_t)s * stage_bytes;
          mbar_expect_tx(&full[s], stage_bytes);
          const unsigned char* src = P.b_img + (ct * n_kchunks + kc) * (int64_t)kI8BChunkBytes;
          bulk_copy_g2s(dst, src, kI8BChunkBytes, &full[s]);
          if (!A_RESIDENT)
            bulk_copy_g2s(dst + kI8BChunkBytes, P.a_img + (mt * n_kchunks + kc) * (int64_t)kI8AChunkBytes,
                          kI8AChunkBytes, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      uint32_t it = 0, tile_i = 0, a_loads = 0;
      int64_t cur_mt = -1;
      for (int64_t u = u_begin; u < u_end; ++u, ++tile_i) {
        const int64_t mt = tile_of(u);
        if (A_RESIDENT && mt != cur_mt) {
          mbar_wait(afull, a_loads & 1u);
          ++a_loads;
          cur_mt = mt;
        }
        const uint32_t buf = tile_i & 1u;
        mbar_wait(&tempty[buf], ((tile_i >> 1) & 1u) ^ 1u);              // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * (uint32_t)kI8N;
        for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
          const uint32_t s = it % (uint32_t)n_stages;
          mbar_wait(&full[s], (it / (uint32_t)n_stages) & 1u);
          tc_fence_after();
          unsigned char* st = s_ring + (size_t)s * stage_bytes;
          const uint32_t b_addr = smem_u32(st);
          const uint32_t a_addr = A_RESIDENT ? smem_u32(s_a + (size_t)kc * kI8AChunkBytes) : smem_u32(st + kI8BChunkBytes);
#pragma unroll
          for (int kk = 0; kk < kI8KChunk / 32; ++kk)
            umma_i8(tmem_d, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), kI8InstrDesc,
                    (kc | kk) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);                                        // the stage is free once these MMAs are done
        }
        umma_commit(&tfull[buf]);                                        // accumulator complete
        if (A_RESIDENT) {
          const bool last_of_mt = (u + 1 == u_end) || tile_of(u + 1) != mt;
          if (last_of_mt) umma_commit(afree);
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    // a warp reaches TMEM lanes 32 (warp % 4) .. + 31; warpgroup g takes the column blocks g, g + G, ...
    // (one warp per scheduler would leave the epilogue latency-bound: 17 % issue, 15 k cycles per tile)
    const int w4 = warp & 3;
    const int grp = (warp - 4) >> 2;
    const bool poisoned = P.overflow != nullptr && *P.overflow != 0;
    const bool divide = P.out_divisor != 1.0;
    uint32_t tile_i = 0;
    for (int64_t u = u_begin; u < u_end; ++u, ++tile_i) {
      const int64_t mt = tile_of(u), tile = col_of(u);
      const uint32_t buf = tile_i & 1u;
      const int64_t r = mt * kI8M + w4 * 32 + lane;
      // the tile's 32 column scales: one coalesced load per tile, issued before the wait, handed
      // out by shuffles (a dependent global load per output would serialise the epilogue)
      const double my_scale = P.scale != nullptr ? __ldg(P.scale + tile * kI8ColsPerTile + lane) : 1.0;
      mbar_wait(&tfull[buf], (tile_i >> 1) & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * (uint32_t)kI8N;
#pragma unroll 1
      for (int cb = grp; cb < kI8N / 32; cb += kPipeEpiGroups) {
        uint32_t v[32];
        tmem_ld32(tmem_d + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(cb * 32), v);
        if (cb + kPipeEpiGroups >= kI8N / 32) {                          // this warp's share is in registers
          tc_fence_before();
          mbar_arrive(&tempty[buf]);
        }
        if (P.out_i32 != nullptr && r < P.m_total) {
          int32_t* o = P.out_i32 + r * (P.n_col_tiles * kI8N) + tile * kI8N + cb * 32;
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = (int32_t)v[k];
        }
#pragma unroll
        for (int c = 0; c < 32 / kI8Digits; ++c) {
          const int64_t d = tile * kI8ColsPerTile + cb * (32 / kI8Digits) + c;
          // digits 0..3 and 4..7 recombined by Horner in FP64: every intermediate is an integer
          // below 2^53 (|digit sum| <= 64 N < 2^31), so each step is exact — the values of the
          // int64 recombination of k6i_mma_kernel — then hi * 128^4 + lo with one rounding
          double lo = (double)(int32_t)v[c * kI8Digits + 3], hi = (double)(int32_t)v[c * kI8Digits + 7];
#pragma unroll
          for (int s = 2; s >= 0; --s) {
            lo = fma(lo, 128.0, (double)(int32_t)v[c * kI8Digits + s]);
            hi = fma(hi, 128.0, (double)(int32_t)v[c * kI8Digits + 4 + s]);
          }
          const double sc = __shfl_sync(0xffffffffu, my_scale, cb * (32 / kI8Digits) + c);
          // streaming stores: the statistic table (800 MB at cfg4's shape) must not evict the
          // digit images from L2, which every resample tile re-reads
          if (r < P.m_total && d < P.n_cols && P.out != nullptr) {
            double val = fma(hi, 268435456.0, lo) * sc;
            if (divide) val = val / P.out_divisor;               // x / 1 needs no division
            __stcs(P.out + d * P.out_ld + r, poisoned ? nan("") : val);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

inline size_t k6i_pipe_smem_bytes(int n_kchunks, bool resident, int n_stages) {
  const size_t a = resident ? (size_t)n_kchunks * kI8AChunkBytes : 0;
  const size_t stage = resident ? (size_t)kI8BChunkBytes : (size_t)(kI8AChunkBytes + kI8BChunkBytes);
  return 1024 + a + (size_t)n_stages * stage + (2 * kPipeMaxStages + 8) * sizeof(uint64_t) + 64;
}

}  // namespace b200boot
