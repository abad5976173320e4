// K6i — the multi-column count-matrix x data contraction on the 5th-generation tensor cores,
// exactly, in integer arithmetic (tcgen05.mma kind::i8, accumulators in TMEM).
//
// For (N, D) data every column shares the index set of a resample
// (/root/reference/src/scikits/bootstrap/bootstrap.py:431), so the column sums of resample r are
//   S[r][d] = sum_i c[r][i] * x[i][d],     c[r][i] = multiplicity of row i in resample r.
// FP64 has no tcgen05 path; instead the contraction is made exact:
//   * counts are small non-negative integers: c[r][i] as uint8 (a count of 256 or more raises the
//     overflow flag and poisons the output);
//   * each column is written in fixed point relative to its own largest magnitude,
//     x[i][d] = q[i][d] * 2^(e_d - 52), |q| < 2^52 (rounding: half an ulp of the column's largest
//     value — a plain FP64 sum rounds every addend to the running sum's ulp, which is no better),
//     and q as eight balanced base-128 digits, q = sum_s dig_s * 128^s, dig_s in [-64, 63]: int8;
//   * sum_i c[r][i] * dig_s[i][d] is an int32 dot product (the counts of a resample add up to N,
//     so |.| <= 64 N: exact for N < 2^25), one tcgen05.mma.kind::i8 column per (d, s); the digit
//     sums are recombined exactly in two int64 halves, joined and scaled in FP64: two roundings
//     per output.
// Tile: 128 resamples (UMMA M) x 256 digit columns = 32 data columns (UMMA N), K = rows in
// chunks of 128 bytes (one SWIZZLE_128B atom row), four K = 32 instructions per chunk.
// Operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout by the
// CTA's own threads (16-byte chunk c of row r inside each 8-row, 1024-byte atom sits at chunk
// c ^ r), the digit matrix double-buffered so the copy of chunk j + 1 overlaps the MMAs of
// chunk j; when the count tile of the CTA's 128 resamples fits (N <= 1152) it stays resident
// for all column tiles.  Epilogue: tcgen05.ld 32 lanes x 32 columns per warp, int64 recombine,
// one FP64 multiply, coalesced stores to stat[d][r].
#pragma once
#include "common.cuh"
#include "draw.cuh"

namespace b200boot {

constexpr int kI8Threads = 256;
constexpr int kI8M = 128;                               // resamples per tile
constexpr int kI8N = 256;                               // digit columns per tile
constexpr int kI8Digits = 8;
constexpr int kI8ColsPerTile = kI8N / kI8Digits;        // 32 data columns
constexpr int kI8KChunk = 128;                          // K bytes per staged chunk
constexpr int kI8AChunkBytes = kI8M * kI8KChunk;        // 16 KB
constexpr int kI8BChunkBytes = kI8N * kI8KChunk;        // 32 KB
constexpr int kI8MaxResidentChunks = 9;                 // 9 * 16 KB + 2 * 32 KB + slack <= 227 KB
constexpr int kI8FracBits = 52;

// ---- PTX wrappers -------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, uint8 x int8 -> int32, one elected thread issues
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor of a K-major SWIZZLE_128B operand tile (cute::UMMA::SmemDescriptor):
// start address >> 4, leading byte offset 1 (unused for swizzled K-major), stride byte offset
// 1024 >> 4 between 8-row groups, descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D int32, A uint8, B int8, both K-major,
// N = 256, M = 128, dense, no saturation.
constexpr uint32_t kI8InstrDesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(kI8N >> 3) << 17) |
                                  ((uint32_t)(kI8M >> 4) << 24);

// byte offset of element (row, k_byte) inside a [rows][128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k_byte) {
  const uint32_t r = row & 7u;
  return (row >> 3) * 1024u + r * 128u + ((((k_byte >> 4) ^ r) & 7u) << 4) + (k_byte & 15u);
}

// ---- operand preparation ---------------------------------------------------------------
// Per column: largest |g(x - centre)| (g = identity or square) -> exponent; one CTA per column block.
__global__ void __launch_bounds__(256) k6i_column_scale_kernel(const double* __restrict__ data, int64_t n_rows,
                                                               int64_t n_cols, int64_t ld,
                                                               const double* __restrict__ centre, int square,
                                                               double* __restrict__ scale, int32_t* __restrict__ bad_col,
                                                               int32_t* __restrict__ any_bad) {
  // 32 columns per CTA (coalesced rows of 32 doubles), 8 row groups
  __shared__ double s_max[8][32];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t d = (int64_t)blockIdx.x * 32 + cx;
  double mx = 0.0;
  bool nonfinite = false;
  if (d < n_cols) {
    const double c = centre ? centre[d] : 0.0;
    for (int64_t i = ry; i < n_rows; i += 8) {
      double v = data[i * ld + d] - c;
      if (square) v = v * v;
      if (isfinite(v))
        mx = fmax(mx, fabs(v));
      else
        nonfinite = true;
    }
  }
  s_max[ry][cx] = mx;
  if (nonfinite) {
    atomicExch(bad_col + d, 1);          // the column is redone by the gather fix-up (k6i_fixup_kernel)
    atomicExch(any_bad, 1);
  }
  __syncthreads();
  if (ry == 0 && d < n_cols) {
    for (int k = 1; k < 8; ++k) mx = fmax(mx, s_max[k][cx]);
    // q = round(v * 2^(52 - e)) with 2^(e-1) <= max < 2^e, so |q| <= 2^52; scale = 2^(e - 52)
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);
    scale[d] = ldexp(1.0, e - kI8FracBits);
  }
}

// digits[(d * 8 + s) * k_pad + i] = digit s of q[i][d]; rows >= n_rows and columns >= n_cols are zero
__global__ void __launch_bounds__(256) k6i_digits_kernel(const double* __restrict__ data, int64_t n_rows,
                                                         int64_t n_cols, int64_t n_cols_pad, int64_t ld,
                                                         int64_t k_pad, const double* __restrict__ centre, int square,
                                                         const double* __restrict__ scale, int8_t* __restrict__ digits) {
  // thread per (column, 16 consecutive rows): 16-byte stores along K
  const int64_t groups = k_pad / 16;
  const int64_t total = n_cols_pad * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    // consecutive threads take consecutive columns (coalesced reads of a data row)
    const int64_t g = e / n_cols_pad, d = e - g * n_cols_pad;
    uint32_t out[kI8Digits][4];
#pragma unroll
    for (int s = 0; s < kI8Digits; ++s) out[s][0] = out[s][1] = out[s][2] = out[s][3] = 0u;
    if (d < n_cols) {
      const double c = centre ? centre[d] : 0.0;
      const double inv = 1.0 / scale[d];             // a power of two: exact
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int64_t i = g * 16 + j;
        if (i < n_rows) {
          double v = data[i * ld + d] - c;
          if (square) v = v * v;
          if (!isfinite(v)) v = 0.0;                  // such columns are redone by the fix-up kernel
          long long q = __double2ll_rn(v * inv);      // |q| <= 2^52
#pragma unroll
          for (int s = 0; s < kI8Digits; ++s) {
            // balanced digit in [-64, 63]
            const long long dig = ((q + 64) & 127) - 64;
            q = (q - dig) >> 7;
            out[s][j >> 2] |= ((uint32_t)(dig & 0xFF)) << (8 * (j & 3));
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < kI8Digits; ++s)
      *reinterpret_cast<uint4*>(digits + (d * kI8Digits + s) * k_pad + g * 16) =
          make_uint4(out[s][0], out[s][1], out[s][2], out[s][3]);
  }
}

// counts[r][i] (uint8, row stride k_pad) of the single-cell Lemire stream: one CTA per resample
// at a time, shared-memory word atomics.  Rows i >= n_rows stay zero.
__global__ void __launch_bounds__(256) k6i_count_bytes_kernel(uint64_t seed, int64_t n_rows, int64_t k_pad,
                                                              int64_t resample_lo, int64_t n_local,
                                                              uint8_t* __restrict__ counts,
                                                              int32_t* __restrict__ overflow) {
  extern __shared__ uint32_t s_words[];
  const int n = (int)n_rows;
  const int words = (int)(k_pad / 4);
  const uint32_t thresh = lemire_threshold((uint32_t)n);
  const int n_blocks = (n + 3) / 4;
  for (int64_t rl = blockIdx.x; rl < n_local; rl += gridDim.x) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) s_words[i] = 0u;
    __syncthreads();
    const Stream st = make_stream(seed, resample_lo + rl, 0u, kPurposeIndex);
    for (int q = threadIdx.x; q < n_blocks; q += blockDim.x)
      lemire_block<true>(st, (uint32_t)q, (uint32_t)n, thresh, (int64_t)n, [&](int, int64_t, uint32_t idx) {
        const uint32_t sh = 8u * (idx & 3u);
        const uint32_t old = atomicAdd(s_words + (idx >> 2), 1u << sh);
        if (((old >> sh) & 0xFFu) == 0xFFu) atomicExch(overflow, 1);
      });
    __syncthreads();
    uint32_t* row = reinterpret_cast<uint32_t*>(counts + rl * k_pad);
    for (int i = threadIdx.x; i < words; i += blockDim.x) row[i] = s_words[i];
    __syncthreads();
  }
}

// the same from materialised index sets (any N: the tiled plans of K2); counts zeroed by the caller
__global__ void __launch_bounds__(256) k6i_count_bytes_from_indices_kernel(const int64_t* __restrict__ idx,
                                                                           int64_t n_sets, int64_t n_rows,
                                                                           int64_t k_pad, uint8_t* __restrict__ counts,
                                                                           int32_t* __restrict__ overflow) {
  const int64_t total = n_sets * n_rows;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t r = e / n_rows;
    const int64_t i = idx[e];
    if ((uint64_t)i >= (uint64_t)n_rows) continue;                 // out-of-range entries are ignored
    const int64_t byte = r * k_pad + i;
    const uint32_t sh = 8u * (uint32_t)(byte & 3);
    const uint32_t old = atomicAdd(reinterpret_cast<uint32_t*>(counts + (byte & ~(int64_t)3)), 1u << sh);
    if (((old >> sh) & 0xFFu) == 0xFFu) atomicExch(overflow, 1);
  }
}

// ---- the contraction --------------------------------------------------------------------
struct K6iParams {
  const uint8_t* counts;     // [m_total][k_pad]
  const int8_t* digits;      // [n_cols_pad * 8][k_pad]
  const double* scale;       // [n_cols_pad]
  int64_t m_total;           // resamples (rows of `counts`)
  int64_t k_pad;             // multiple of 128
  int64_t n_cols;            // data columns written
  int64_t n_col_tiles;       // n_cols_pad / 32
  double* out;               // [n_cols][out_ld]: out[d * out_ld + r]
  int64_t out_ld;
  int32_t* out_i32;          // test hook: raw accumulators [m_total][n_col_tiles * 256], or null
  const int32_t* overflow;   // poisons the output when set
  double out_divisor;        // N for the mean (the finisher's division, same bits), else 1
};

// One staged chunk in flight: rows [row0, row0 + ROWS) x K bytes [k0, k0 + 128) of a row-major byte
// matrix, first loaded into registers (ROWS * 8 / 256 sixteen-byte pieces per thread), later
// stored into a SWIZZLE_128B tile.  Splitting the two lets the loads of chunk j + 1 fly while
// chunk j is stored, fenced and multiplied.  Rows at or beyond row_limit read as zero.
template <int ROWS>
struct ChunkRegs {
  static constexpr int kPieces = ROWS * 8 / kI8Threads;
  uint4 v[kPieces];
  __device__ __forceinline__ void load(const unsigned char* __restrict__ src, int64_t row0, int64_t row_limit,
                                       int64_t k_pad, int64_t k0) {
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int item = threadIdx.x + i * kI8Threads;
      const int row = item >> 3, c16 = item & 7;
      v[i] = make_uint4(0u, 0u, 0u, 0u);
      if (row0 + row < row_limit)
        v[i] = __ldg(reinterpret_cast<const uint4*>(src + (row0 + row) * k_pad + k0 + c16 * 16));
    }
  }
  __device__ __forceinline__ void store(unsigned char* dst) const {
#pragma unroll
    for (int i = 0; i < kPieces; ++i) {
      const int item = threadIdx.x + i * kI8Threads;
      *reinterpret_cast<uint4*>(dst + sw128_offset((uint32_t)(item >> 3), (uint32_t)(item & 7) * 16u)) = v[i];
    }
  }
};

template <bool A_RESIDENT>
__global__ void __launch_bounds__(kI8Threads, 1) k6i_mma_kernel(const K6iParams P) {
  extern __shared__ unsigned char k6i_smem_raw[];
  // 1024-byte alignment of the swizzle atoms
  unsigned char* smem = k6i_smem_raw + ((1024u - (smem_u32(k6i_smem_raw) & 1023u)) & 1023u);
  const int n_kchunks = (int)(P.k_pad / kI8KChunk);
  const int a_chunks = A_RESIDENT ? n_kchunks : 2;
  unsigned char* s_a = smem;                                        // a_chunks x 16 KB
  unsigned char* s_b = smem + (size_t)a_chunks * kI8AChunkBytes;    // 2 x 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + 2 * kI8BChunkBytes);   // [2]
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * kI8M;
  const unsigned char* digits = reinterpret_cast<const unsigned char*>(P.digits);
  const int64_t b_limit = P.n_col_tiles * kI8N;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
  }
  if (warp == 0) tmem_alloc(s_tmem, kI8N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *s_tmem;

  if (A_RESIDENT) {
    for (int j = 0; j < n_kchunks; ++j) {
      ChunkRegs<kI8M> a;
      a.load(P.counts, m0, P.m_total, P.k_pad, (int64_t)j * kI8KChunk);
      a.store(s_a + (size_t)j * kI8AChunkBytes);
    }
  }

  // the CTA's sequence of (column tile, K chunk) steps; step `it` uses buffer it & 1
  const int64_t n_my_tiles = blockIdx.y < P.n_col_tiles ? (P.n_col_tiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
  const int64_t n_steps = n_my_tiles * n_kchunks;
  auto step_tile = [&](int64_t st) { return (int64_t)blockIdx.y + (st / n_kchunks) * gridDim.y; };
  ChunkRegs<kI8N> b_next;
  ChunkRegs<kI8M> a_next;
  if (n_steps > 0) {
    b_next.load(digits, step_tile(0) * kI8N, b_limit, P.k_pad, 0);
    if (!A_RESIDENT) a_next.load(P.counts, m0, P.m_total, P.k_pad, 0);
  }
  for (int64_t st = 0; st < n_steps; ++st) {
    const uint32_t it = (uint32_t)st;
    const uint32_t buf = it & 1u;
    const int j = (int)(st % n_kchunks);
    const int64_t tile = step_tile(st);
    if (it >= 2) mbar_wait(&bars[buf], ((it >> 1) - 1u) & 1u);       // the MMAs that read this buffer are done
    b_next.store(s_b + (size_t)buf * kI8BChunkBytes);
    if (!A_RESIDENT) a_next.store(s_a + (size_t)buf * kI8AChunkBytes);
    if (st + 1 < n_steps) {                                          // loads of the next step fly from here on
      const int jn = (int)((st + 1) % n_kchunks);
      b_next.load(digits, step_tile(st + 1) * kI8N, b_limit, P.k_pad, (int64_t)jn * kI8KChunk);
      if (!A_RESIDENT) a_next.load(P.counts, m0, P.m_total, P.k_pad, (int64_t)jn * kI8KChunk);
    }
    fence_proxy_async();           // generic-proxy stores -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(s_a + (size_t)(A_RESIDENT ? j : (int)buf) * kI8AChunkBytes);
      const uint32_t b_addr = smem_u32(s_b + (size_t)buf * kI8BChunkBytes);
#pragma unroll
      for (int kk = 0; kk < kI8KChunk / 32; ++kk)
        umma_i8(tmem_d, umma_desc_sw128(a_addr + kk * 32), umma_desc_sw128(b_addr + kk * 32), kI8InstrDesc,
                (j | kk) != 0 ? 1u : 0u);
      umma_commit(&bars[buf]);     // arrives when every MMA issued so far has completed
    }
    if (j + 1 < n_kchunks) continue;
    // accumulator complete: the last commit covers all MMAs of the tile
    mbar_wait(&bars[buf], (it >> 1) & 1u);
    tc_fence_after();
    if (warp >= 4) {
      const int w4 = warp - 4;                          // TMEM lanes 32 w4 .. 32 w4 + 31
      const int64_t r = m0 + w4 * 32 + lane;
      const bool poisoned = P.overflow != nullptr && *P.overflow != 0;
#pragma unroll 1
      for (int cb = 0; cb < kI8N / 32; ++cb) {
        uint32_t v[32];
        tmem_ld32(tmem_d + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(cb * 32), v);
        if (P.out_i32 != nullptr && r < P.m_total) {
          int32_t* o = P.out_i32 + r * (P.n_col_tiles * kI8N) + tile * kI8N + cb * 32;
#pragma unroll
          for (int k = 0; k < 32; ++k) o[k] = (int32_t)v[k];
        }
#pragma unroll
        for (int c = 0; c < 32 / kI8Digits; ++c) {
          const int64_t d = tile * kI8ColsPerTile + cb * (32 / kI8Digits) + c;
          // digits 0..3 and 4..7 recombined exactly in int64 (each half below 2^52 for N < 2^25),
          // then hi * 128^4 + lo in FP64: both conversions exact, one rounding
          long long lo = 0, hi = 0;
#pragma unroll
          for (int s = 3; s >= 0; --s) {
            lo = lo * 128 + (long long)(int32_t)v[c * kI8Digits + s];
            hi = hi * 128 + (long long)(int32_t)v[c * kI8Digits + 4 + s];
          }
          if (r < P.m_total && d < P.n_cols && P.out != nullptr)
            P.out[d * P.out_ld + r] =
                poisoned ? nan("") : (fma((double)hi, 268435456.0, (double)lo) * P.scale[d]) / P.out_divisor;
        }
      }
    }
    tc_fence_before();
    __syncthreads();               // the next tile's first MMA overwrites the accumulator
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_d, kI8N);
}

// Columns holding non-finite values cannot be written in fixed point, and in a product 0 * nan is
// nan although a resample that never draws the row must stay finite (the gather x[indices] of
// bootstrap.py:431 never sees it).  Such columns (flagged by k6i_column_scale_kernel) are
// recomputed here the gather way: a warp per resample replays the index stream (or reads the
// supplied indices) and adds the rows in FP64.  Launched after every contraction; returns at once
// when no column is flagged.
__global__ void __launch_bounds__(256) k6i_fixup_kernel(const double* __restrict__ data, int64_t n_rows,
                                                        int64_t n_cols, int64_t ld,
                                                        const double* __restrict__ centre, int square,
                                                        const int32_t* __restrict__ bad_col,
                                                        const int32_t* __restrict__ any_bad, uint64_t seed,
                                                        int64_t resample_lo, int64_t n_local,
                                                        const int64_t* __restrict__ indices,
                                                        double out_divisor, double* __restrict__ out) {
  if (*any_bad == 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t thresh = indices ? 0u : lemire_threshold((uint32_t)n_rows);
  for (int64_t d = 0; d < n_cols; ++d) {
    if (bad_col[d] == 0) continue;
    const double c = centre ? centre[d] : 0.0;
    for (int64_t rl = warp_global; rl < n_local; rl += n_warps) {
      double acc = 0.0;
      auto add = [&](int64_t i) {
        double v = data[i * ld + d] - c;
        if (square) v = v * v;
        acc += v;
      };
      if (indices) {
        for (int64_t j = lane; j < n_rows; j += 32) add(indices[rl * n_rows + j]);
      } else {
        const Stream st = make_stream(seed, resample_lo + rl, 0u, kPurposeIndex);
        const uint32_t n_blocks = (uint32_t)((n_rows + 3) / 4);
        for (uint32_t q = lane; q < n_blocks; q += 32)
          lemire_block<true>(st, q, (uint32_t)n_rows, thresh, n_rows, [&](int, int64_t, uint32_t idx) { add((int64_t)idx); });
      }
      acc = warp_sum(acc);
      if (lane == 0) out[d * n_local + rl] = acc / out_divisor;
    }
  }
}

inline size_t k6i_smem_bytes(int64_t k_pad, bool resident) {
  const int64_t a_chunks = resident ? k_pad / kI8KChunk : 2;
  return 1024 + (size_t)a_chunks * kI8AChunkBytes + 2 * (size_t)kI8BChunkBytes + 64;
}

}  // namespace b200boot
