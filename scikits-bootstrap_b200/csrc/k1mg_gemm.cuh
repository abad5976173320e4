// K1mg — medians / quantiles of MANY short columns through the tensor cores (cfg4: 1000 x 10 000
// table, 1e4 resamples).
//
// All columns of a resample share its row multiplicities c[r][i] (bootstrap.py:431), and the k-th
// order statistic of column d in resample r sits at the first sorted position p with
//     sum_{q <= p} c[r][perm_d[q]]  >  k                                   (perm_d: rows of column d in ascending order).
// K1m walks those cumulative sums position by position: 1e11 two-byte gathers at cfg4's shape,
// 21 ms.  Here the walk starts from a table of T = 32 prefix counts per (resample, column),
//     S[r][d][j] = #{draws of resample r whose rank in column d is below (j + 1) s},    s = ceil(N / T),
// which is a matrix product: S = C x B with C the uint8 count matrix (resamples x rows: K6i's) and
// B[(d, j)][i] = [rank_d[i] < (j + 1) s] a 0/1 int8 matrix — exact in int32 on tcgen05.mma.kind::i8.
// The product is never written out: the epilogue of each 128 x 256 tile (128 resamples x 8 columns
// x 32 thresholds) reads a (resample, column)'s 16 prefix counts from TMEM, picks the segment in
// which the wanted rank falls, and finishes with a walk over at most s positions of that segment
// (counts read from the CTA's resident count tile in shared memory, perm_d from a per-unit block
// copied next to it).  The
// gathers drop from N to ~s / 2 per (resample, column); results are identical to K1m's (same
// draws: the single-cell Lemire stream; same order statistics).
// Pipeline: k6i_pipe.cuh's (tile images + 1-D bulk copies, producer / MMA / 16 epilogue warps, two
// TMEM accumulators, persistent CTAs with a resident count tile), N <= 1024 rows.
#pragma once
#include "k1m_median.cuh"
#include "k6i_pipe.cuh"

namespace b200boot {

#ifndef B2_MG_T
#define B2_MG_T 32
#endif
constexpr int kMgT = B2_MG_T;                    // thresholds (prefix counts) per column: 16 or 32
constexpr int kMgColsPerTile = kI8N / kMgT;      // data columns per 256-row B tile
#ifndef B2_MG_WALK
#define B2_MG_WALK 4
#endif
constexpr int kMgWalk = B2_MG_WALK;              // positions fetched together in the finishing walk

// B image: tile ct holds columns 16 ct .. 16 ct + 15; B row (d_local * 16 + j), K byte i =
// [rank[i][d] < (j + 1) seg]; rows i >= n_rows and columns d >= n_cols are zero.
__global__ void __launch_bounds__(256) k1mg_indicator_image_kernel(const int32_t* __restrict__ rank, int64_t n_rows,
                                                                   int64_t n_cols, int64_t k_pad, int64_t n_col_tiles,
                                                                   int seg, unsigned char* __restrict__ b_img) {
  // thread per (column, 16 consecutive rows); consecutive threads take consecutive columns
  // (coalesced reads of a row of the rank table)
  const int64_t groups = k_pad / 16;
  const int64_t n_cols_pad = n_col_tiles * kMgColsPerTile;
  const int64_t total = n_cols_pad * groups;
  const int64_t n_kchunks = k_pad / kI8KChunk;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t g = e / n_cols_pad, d = e - g * n_cols_pad;
    int32_t rk[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int64_t i = g * 16 + t;
      rk[t] = (d < n_cols && i < n_rows) ? __ldg(rank + i * n_cols + d) : 0x7FFFFFFF;
    }
    const int64_t tile = d / kMgColsPerTile;
    const uint32_t row0 = (uint32_t)(d % kMgColsPerTile) * kMgT;
    const int64_t kc = (g * 16) >> 7;
    const uint32_t k_byte = (uint32_t)((g * 16) & 127);
    unsigned char* base = b_img + (tile * n_kchunks + kc) * (int64_t)kI8BChunkBytes;
#pragma unroll
    for (int j = 0; j < kMgT; ++j) {
      const int32_t lim = (j + 1) * seg;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int t = 0; t < 16; ++t) w[t >> 2] |= (rk[t] < lim ? 1u : 0u) << (8 * (t & 3));
      *reinterpret_cast<uint4*>(base + sw128_offset(row0 + (uint32_t)j, k_byte)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// perm16 image: per column tile the permutations of its columns as uint16 (N <= 1024 rows),
// [n_col_tiles][kMgColsPerTile][k_pad]: one bulk copy per unit puts them next to the count tile
__global__ void __launch_bounds__(256) k1mg_perm_image_kernel(const uint32_t* __restrict__ idx, int64_t n_rows,
                                                              int64_t n_pad, int64_t n_cols, int64_t k_pad,
                                                              int64_t n_col_tiles, uint16_t* __restrict__ perm_img) {
  const int64_t total = n_col_tiles * kMgColsPerTile * k_pad;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t d = e / k_pad, p = e - d * k_pad;
    // positions past the end point at row n_rows, whose count byte is zero (k_pad > n_rows), so the
    // finishing walk needs no bound check; with n_rows == k_pad there are no such positions
    perm_img[e] = (d < n_cols && p < n_rows) ? (uint16_t)__ldg(idx + d * n_pad + p) : (uint16_t)(n_rows < k_pad ? n_rows : 0);
  }
}

struct K1mgParams {
  const unsigned char* a_img;   // [m_tiles][n_kchunks][16 KB] counts
  const unsigned char* b_img;   // [n_col_tiles][n_kchunks][32 KB] indicators
  const uint16_t* perm_img;     // [n_col_tiles][kMgColsPerTile][k_pad]: row at sorted position p
  const double* sorted;         // [D][N]
  int64_t m_total, m_tiles;
  int32_t n_kchunks, n_stages;
  int64_t n_rows, n_cols, n_col_tiles;
  int64_t seg_tiles;            // column tiles per work item (see the kernel's unit order)
  int32_t seg;
  RankRule rule;
  double* out;                  // out[d * out_ld + r]
  int64_t out_ld;
  const int32_t* overflow;      // a count did not fit its byte: the output is poisoned
};

__global__ void __launch_bounds__(kPipeThreads, 1) k1mg_kernel(const K1mgParams P) {
  extern __shared__ unsigned char mg_smem_raw[];
  unsigned char* smem = mg_smem_raw + ((1024u - (smem_u32(mg_smem_raw) & 1023u)) & 1023u);
  const int n_kchunks = P.n_kchunks;
  const int n_stages = P.n_stages;
  const uint32_t stage_bytes = (uint32_t)kI8BChunkBytes;
  unsigned char* s_a = smem;                                                   // resident count tile
  unsigned char* s_ring = smem + (size_t)n_kchunks * kI8AChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)n_stages * stage_bytes +
                                               2 * (size_t)(kMgColsPerTile * n_kchunks * kI8KChunk * 2));
  uint64_t* full = bars;
  uint64_t* empty = bars + kPipeMaxStages;
  uint64_t* tfull = bars + 2 * kPipeMaxStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;
  uint64_t* afree = afull + 1;
  uint64_t* pfull = afree + 1;                         // [2] permutation block of a unit loaded
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(pfull + 2);
  volatile uint32_t* s_epi_done = s_tmem + 1;          // [2] epilogue warps past a unit (walks included), by unit parity
  const uint32_t perm_bytes = (uint32_t)(kMgColsPerTile * n_kchunks * kI8KChunk * 2);
  unsigned char* s_perm = s_ring + (size_t)n_stages * stage_bytes;             // [2][perm_bytes]

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
    mbar_init(&pfull[0], 1);
    mbar_init(&pfull[1], 1);
    s_epi_done[0] = 0u;
    s_epi_done[1] = 0u;
  }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  // Unit order.  A work item is (column segment, resample tile): the seg_tiles column tiles of the
  // segment against one count tile.  Items are numbered segment-major and dealt out round-robin
  // (CTA b takes items b, b + G, ...), so at any time all CTAs work inside one or two column
  // segments: their indicator blocks (a few MB) stay in L2 while every resample tile re-reads
  // them.  With contiguous ranges per CTA the CTAs were spread over the whole 328 MB image, which
  // then came from HBM once per resample tile (ncu: 27.9 GB of DRAM reads, 4.2 TB/s).
  const int64_t n_segs = (P.n_col_tiles + P.seg_tiles - 1) / P.seg_tiles;
  const int64_t n_items = n_segs * P.m_tiles;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------- producer ----------------
      uint32_t it = 0, ti = 0, items_done = 0;
      for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++items_done) {
        const int64_t sg = item / P.m_tiles, mt = item - sg * P.m_tiles;
        const int64_t ct0 = sg * P.seg_tiles, ct1 = min(ct0 + P.seg_tiles, P.n_col_tiles);
        if (items_done > 0) {
          mbar_wait(afree, (items_done - 1u) & 1u);                      // the MMAs reading the old count tile are done
          // ... and so are the walks of the epilogue, which read it too
          const uint32_t need = ti * (uint32_t)(kPipeEpiGroups * 4);
          while (s_epi_done[0] + s_epi_done[1] < need) __nanosleep(64);
          fence_proxy_async();
        }
        mbar_expect_tx(afull, (uint32_t)n_kchunks * kI8AChunkBytes);
        for (int kc = 0; kc < n_kchunks; ++kc)
          bulk_copy_g2s(s_a + (size_t)kc * kI8AChunkBytes, P.a_img + (mt * n_kchunks + kc) * (int64_t)kI8AChunkBytes,
                        kI8AChunkBytes, afull);
        for (int64_t ct = ct0; ct < ct1; ++ct, ++ti) {
          // the unit's permutations, double-buffered like the accumulators: buffer (unit & 1) is
          // free once every epilogue warp is past the unit before last (counted per parity: a
          // total over all units could be reached by fast warps that are already past the NEXT
          // unit while a slow one is still walking this buffer's last unit)
          if (ti >= 2) {
            const uint32_t need = (ti >> 1) * (uint32_t)(kPipeEpiGroups * 4);
            while (s_epi_done[ti & 1u] < need) __nanosleep(32);
            fence_proxy_async();
          }
          mbar_expect_tx(&pfull[ti & 1u], perm_bytes);
          bulk_copy_g2s(s_perm + (size_t)(ti & 1u) * perm_bytes,
                        reinterpret_cast<const unsigned char*>(P.perm_img) + ct * (int64_t)perm_bytes, perm_bytes,
                        &pfull[ti & 1u]);
          for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
            const uint32_t s = it % (uint32_t)n_stages;
            mbar_wait(&empty[s], ((it / (uint32_t)n_stages) & 1u) ^ 1u);
            mbar_expect_tx(&full[s], stage_bytes);
            bulk_copy_g2s(s_ring + (size_t)s * stage_bytes, P.b_img + (ct * n_kchunks + kc) * (int64_t)kI8BChunkBytes,
                          kI8BChunkBytes, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      uint32_t it = 0, tile_i = 0, items_done = 0;
      for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x, ++items_done) {
        const int64_t sg = item / P.m_tiles;
        const int64_t ct0 = sg * P.seg_tiles, ct1 = min(ct0 + P.seg_tiles, P.n_col_tiles);
        mbar_wait(afull, items_done & 1u);
        for (int64_t ct = ct0; ct < ct1; ++ct, ++tile_i) {
          const uint32_t buf = tile_i & 1u;
          mbar_wait(&tempty[buf], ((tile_i >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * (uint32_t)kI8N;
          for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
            const uint32_t s = it % (uint32_t)n_stages;
            mbar_wait(&full[s], (it / (uint32_t)n_stages) & 1u);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s_ring + (size_t)s * stage_bytes);
            const uint32_t a_addr = smem_u32(s_a + (size_t)kc * kI8AChunkBytes);
#pragma unroll
            for (int kk = 0; kk < kI8KChunk / 32; ++kk)
              umma_i8(tmem_d, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), kI8InstrDesc,
                      (kc | kk) != 0 ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          umma_commit(&tfull[buf]);
        }
        umma_commit(afree);                                              // this item's count tile has been read
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: prefix counts -> segment -> walk ----------------
    const int w4 = warp & 3;
    const int grp = (warp - 4) >> 2;
    const bool poisoned = P.overflow != nullptr && *P.overflow != 0;
    const int two = P.rule.k_hi != P.rule.k_lo ? 1 : 0;
    const int seg = P.seg;
    const uint32_t r_local = (uint32_t)(w4 * 32 + lane);
    const unsigned char* s_row = s_a + (r_local >> 3) * 1024u + (r_local & 7u) * 128u;   // this lane's row of a count chunk
    const uint32_t rx4 = (r_local & 7u) << 4;
    const int k_pad = n_kchunks * kI8KChunk;
    uint32_t tile_i = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x)
    for (int64_t sg = item / P.m_tiles, mt = item - sg * P.m_tiles, ct = sg * P.seg_tiles;
         ct < min((sg + 1) * P.seg_tiles, P.n_col_tiles); ++ct, ++tile_i) {
      const uint32_t buf = tile_i & 1u;
      const int64_t r = mt * kI8M + r_local;
      mbar_wait(&tfull[buf], (tile_i >> 1) & 1u);
      tc_fence_after();
      mbar_wait(&pfull[buf], (tile_i >> 1) & 1u);                        // this unit's permutations
      const uint16_t* perm_tile = reinterpret_cast<const uint16_t*>(s_perm + (size_t)buf * perm_bytes);
      const uint32_t tmem_d = tmem_base + buf * (uint32_t)kI8N;
      // a warp group takes the column blocks grp, grp + G, ...; the accumulator is released after
      // the last of them has been read (the walks of that block then run beside the next MMAs)
      for (int cb = grp; cb < kI8N / 32; cb += kPipeEpiGroups) {
        uint32_t v[32];
        __syncwarp();                                    // tcgen05.ld is warp-collective: the lanes' walks differ in length
        tmem_ld32(tmem_d + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(cb * 32), v);
        if (cb + kPipeEpiGroups >= kI8N / 32) {
          tc_fence_before();
          mbar_arrive(&tempty[buf]);
        }
        // the walks read the count tile of THIS unit's resample tile, which stays resident until the
        // producer has seen every epilogue warp past the unit (s_epi_done)
#pragma unroll
        for (int c = 0; c < 32 / kMgT; ++c) {
          const int64_t d = ct * kMgColsPerTile + cb * (32 / kMgT) + c;
          if (r >= P.m_total || d >= P.n_cols) continue;
          const uint32_t* S = v + c * kMgT;                              // S[j] = #draws with rank < (j + 1) seg
          const long long k_lo = P.rule.k_lo, k_hi = P.rule.k_hi;
          // first j with S[j] > k_lo, and the draws below segment j (S is non-decreasing)
          int j = 0;
          long long acc = 0;
#pragma unroll
          for (int t = 0; t < kMgT - 1; ++t) {
            const bool below = (long long)S[t] <= k_lo;
            j += below ? 1 : 0;
            acc = below ? (long long)S[t] : acc;
          }
          const uint16_t* perm = perm_tile + (size_t)(cb * (32 / kMgT) + c) * (size_t)k_pad;
          int pos_lo = -1, pos_hi = -1;
          // the walk, kMgWalk positions at a time: the permutation entries and the count bytes of a
          // group are independent loads (a position-by-position walk is a chain of two dependent
          // loads per step: 85 k cycles per tile; the lanes of a warp walk segments of different
          // lengths, so the warp pays for its longest walk).  Positions past the column's end hold
          // a row whose count byte is zero (k1mg_perm_image_kernel), so the loop needs no bound
          // checks; the byte of row i in the swizzled count tile is at
          // tile_row_base + (i >> 7) * 16 KB + ((i & 127) ^ ((row & 7) << 4)).
          for (int p0 = j * seg; p0 < k_pad && pos_hi < 0; p0 += kMgWalk) {
            uint32_t cnt[kMgWalk];
#pragma unroll
            for (int t = 0; t < kMgWalk; ++t) {
              const uint32_t i = perm[p0 + t];
              cnt[t] = s_row[((i >> 7) << 14) + ((i & 127u) ^ rx4)];
            }
#pragma unroll
            for (int t = 0; t < kMgWalk; ++t) {
              acc += cnt[t];
              pos_lo = (pos_lo < 0 && acc > k_lo) ? p0 + t : pos_lo;
              pos_hi = (pos_hi < 0 && acc > k_hi) ? p0 + t : pos_hi;
            }
          }
          const double* col = P.sorted + d * P.n_rows;
          const double a = pos_lo >= 0 ? col[pos_lo] : nan("");
          const double b = two ? (pos_hi >= 0 ? col[pos_hi] : nan("")) : a;
          __stcs(P.out + d * P.out_ld + r, poisoned ? nan("") : rank_combine(P.rule, a, b));
        }
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(const_cast<uint32_t*>(s_epi_done) + (tile_i & 1u), 1u);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- the same for longer columns (1024 < N <= 28 672): the count tile no longer fits shared memory,
// so count chunks are streamed through the ring next to the indicator chunks, and the finishing walk
// reads the permutation (uint32 [D][perm_ld]) and the count bytes (the row-major count matrix,
// [m_total][k_pad]) from L2.  Segments are longer (s = ceil(N / 32) positions), but such tables have
// few (resample, column) pairs per drawn index, and K1m spends its time on the count tables.
struct K1mgStreamParams {
  const unsigned char* a_img;   // [m_tiles][n_kchunks][16 KB] counts (tile images, for the MMAs)
  const unsigned char* b_img;   // [n_col_tiles][n_kchunks][32 KB] indicators
  const uint8_t* counts;        // [m_total][k_pad] row-major (for the walks)
  const uint32_t* perm;         // [D][perm_ld]
  int64_t perm_ld;
  const double* sorted;         // [D][N]
  int64_t m_total, m_tiles;
  int32_t n_kchunks, n_stages;
  int64_t n_rows, n_cols, n_col_tiles, seg_tiles;
  int32_t seg;
  RankRule rule;
  double* out;
  int64_t out_ld;
  const int32_t* overflow;
};

__global__ void __launch_bounds__(kPipeThreads, 1) k1mg_stream_kernel(const K1mgStreamParams P) {
  extern __shared__ unsigned char mgs_smem_raw[];
  unsigned char* smem = mgs_smem_raw + ((1024u - (smem_u32(mgs_smem_raw) & 1023u)) & 1023u);
  const int n_kchunks = P.n_kchunks;
  const int n_stages = P.n_stages;
  const uint32_t stage_bytes = (uint32_t)(kI8AChunkBytes + kI8BChunkBytes);
  unsigned char* s_ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ring + (size_t)n_stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kPipeMaxStages;
  uint64_t* tfull = bars + 2 * kPipeMaxStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tempty + 2);
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
  }
  if (warp == 0) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const int64_t n_segs = (P.n_col_tiles + P.seg_tiles - 1) / P.seg_tiles;
  const int64_t n_items = n_segs * P.m_tiles;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t sg = item / P.m_tiles, mt = item - sg * P.m_tiles;
        for (int64_t ct = sg * P.seg_tiles; ct < min((sg + 1) * P.seg_tiles, P.n_col_tiles); ++ct)
          for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
            const uint32_t s = it % (uint32_t)n_stages;
            mbar_wait(&empty[s], ((it / (uint32_t)n_stages) & 1u) ^ 1u);
            unsigned char* dst = s_ring + (size_t)s * stage_bytes;
            mbar_expect_tx(&full[s], stage_bytes);
            bulk_copy_g2s(dst, P.b_img + (ct * n_kchunks + kc) * (int64_t)kI8BChunkBytes, kI8BChunkBytes, &full[s]);
            bulk_copy_g2s(dst + kI8BChunkBytes, P.a_img + (mt * n_kchunks + kc) * (int64_t)kI8AChunkBytes,
                          kI8AChunkBytes, &full[s]);
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, tile_i = 0;
      for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int64_t sg = item / P.m_tiles;
        for (int64_t ct = sg * P.seg_tiles; ct < min((sg + 1) * P.seg_tiles, P.n_col_tiles); ++ct, ++tile_i) {
          const uint32_t buf = tile_i & 1u;
          mbar_wait(&tempty[buf], ((tile_i >> 1) & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * (uint32_t)kI8N;
          for (int kc = 0; kc < n_kchunks; ++kc, ++it) {
            const uint32_t s = it % (uint32_t)n_stages;
            mbar_wait(&full[s], (it / (uint32_t)n_stages) & 1u);
            tc_fence_after();
            const uint32_t b_addr = smem_u32(s_ring + (size_t)s * stage_bytes);
            const uint32_t a_addr = b_addr + (uint32_t)kI8BChunkBytes;
#pragma unroll
            for (int kk = 0; kk < kI8KChunk / 32; ++kk)
              umma_i8(tmem_d, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), kI8InstrDesc,
                      (kc | kk) != 0 ? 1u : 0u);
            umma_commit(&empty[s]);
          }
          umma_commit(&tfull[buf]);
        }
      }
    }
  } else if (warp >= 4) {
    const int w4 = warp & 3;
    const int grp = (warp - 4) >> 2;
    const bool poisoned = P.overflow != nullptr && *P.overflow != 0;
    const int two = P.rule.k_hi != P.rule.k_lo ? 1 : 0;
    const int seg = P.seg;
    const int n = (int)P.n_rows;
    const int64_t k_pad = (int64_t)n_kchunks * kI8KChunk;
    const uint32_t r_local = (uint32_t)(w4 * 32 + lane);
    uint32_t tile_i = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x)
    for (int64_t sg = item / P.m_tiles, mt = item - sg * P.m_tiles, ct = sg * P.seg_tiles;
         ct < min((sg + 1) * P.seg_tiles, P.n_col_tiles); ++ct, ++tile_i) {
      const uint32_t buf = tile_i & 1u;
      const int64_t r = mt * kI8M + r_local;
      const uint8_t* cnt_row = P.counts + min(r, P.m_total - 1) * k_pad;
      mbar_wait(&tfull[buf], (tile_i >> 1) & 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * (uint32_t)kI8N;
      for (int cb = grp; cb < kI8N / 32; cb += kPipeEpiGroups) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(tmem_d + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(cb * 32), v);
        if (cb + kPipeEpiGroups >= kI8N / 32) {
          tc_fence_before();
          mbar_arrive(&tempty[buf]);
        }
#pragma unroll
        for (int c = 0; c < 32 / kMgT; ++c) {
          const int64_t d = ct * kMgColsPerTile + cb * (32 / kMgT) + c;
          if (r >= P.m_total || d >= P.n_cols) continue;
          const uint32_t* S = v + c * kMgT;
          const long long k_lo = P.rule.k_lo, k_hi = P.rule.k_hi;
          int j = 0;
          long long acc = 0;
#pragma unroll
          for (int t = 0; t < kMgT - 1; ++t) {
            const bool below = (long long)S[t] <= k_lo;
            j += below ? 1 : 0;
            acc = below ? (long long)S[t] : acc;
          }
          const uint32_t* perm = P.perm + d * P.perm_ld;
          int pos_lo = -1, pos_hi = -1;
          for (int p0 = j * seg; p0 < n && pos_hi < 0; p0 += kMgWalk) {
            uint32_t cnt[kMgWalk];
#pragma unroll
            for (int t = 0; t < kMgWalk; ++t) cnt[t] = p0 + t < n ? (uint32_t)__ldg(cnt_row + __ldg(perm + p0 + t)) : 0u;
#pragma unroll
            for (int t = 0; t < kMgWalk; ++t) {
              acc += cnt[t];
              pos_lo = (pos_lo < 0 && acc > k_lo) ? p0 + t : pos_lo;
              pos_hi = (pos_hi < 0 && acc > k_hi) ? p0 + t : pos_hi;
            }
          }
          const double* col = P.sorted + d * P.n_rows;
          const double a = pos_lo >= 0 ? col[pos_lo] : nan("");
          const double b = two ? (pos_hi >= 0 ? col[pos_hi] : nan("")) : a;
          __stcs(P.out + d * P.out_ld + r, poisoned ? nan("") : rank_combine(P.rule, a, b));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

inline size_t k1mg_stream_smem_bytes(int n_stages) {
  return 1024 + (size_t)n_stages * (size_t)(kI8AChunkBytes + kI8BChunkBytes) + (2 * kPipeMaxStages + 8) * sizeof(uint64_t) + 64;
}

inline size_t k1mg_smem_bytes(int n_kchunks, int n_stages) {
  return 1024 + (size_t)n_kchunks * kI8AChunkBytes + (size_t)n_stages * kI8BChunkBytes +
         2 * (size_t)(kMgColsPerTile * n_kchunks * kI8KChunk * 2) + (2 * kPipeMaxStages + 10) * sizeof(uint64_t) + 64;
}

}  // namespace b200boot
