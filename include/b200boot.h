/* b200boot.h — C ABI of libb200boot.so: the sm_100a implementation of the
 * scikits.bootstrap resampling hot path.
 *
 * The reference (cgevans/scikits-bootstrap) is pure Python and has no FFI; the
 * boundary a maintainer would bind is therefore the set of numpy call sites
 * inside `src/scikits/bootstrap/bootstrap.py` that this library replaces.  Each
 * entry point below cites the reference lines it stands in for (bootstrap.py:L).
 *
 * Conventions
 *  - Plain C: pointers and sizes only.  All data pointers are DEVICE pointers
 *    unless the parameter is documented as host memory; `stream` is a
 *    cudaStream_t passed as void* (NULL = legacy default stream).
 *  - Every call returns 0 on success or a negative b200boot_status; the message
 *    of the last failure on the calling thread is b200boot_last_error().
 *    Nothing throws across the ABI.
 *  - The library never allocates caller-visible memory: scratch comes from the
 *    caller-provided workspace (`ws`, `ws_bytes`; 256-byte aligned), sized by
 *    b200boot_workspace_bytes().  No pointer is retained after the stream work
 *    completes.  Calls only enqueue work on `stream` (no host synchronisation)
 *    unless stated otherwise.
 *  - Statistic vectors are column-major: `stat[d * n_local + r]` is the value of
 *    output column d for local resample r.
 */
#ifndef B200BOOT_H
#define B200BOOT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200BOOT_ABI_VERSION 1

typedef enum b200boot_status {
  B200BOOT_OK = 0,
  B200BOOT_EINVAL = -1,       /* bad argument */
  B200BOOT_ECUDA = -2,        /* CUDA runtime error (see last_error) */
  B200BOOT_EWORKSPACE = -3,   /* workspace too small or misaligned */
  B200BOOT_EUNSUPPORTED = -4  /* statistic / shape not in the device registry */
} b200boot_status;

/* Device statistic registry (replaces the user callable at bootstrap.py:431).
 * Moment statistics are finished from per-resample sums; N is the resample size. */
typedef enum b200boot_stat {
  B200BOOT_STAT_MEAN = 0,             /* np.mean / np.average, 1 array            */
  B200BOOT_STAT_VAR = 1,              /* np.var (ddof=0), 1 array                 */
  B200BOOT_STAT_STD = 2,              /* np.std (ddof=0), 1 array                 */
  B200BOOT_STAT_RATIO_OF_MEANS = 3,   /* mean(a)/mean(b), 2 paired arrays         */
  B200BOOT_STAT_WEIGHTED_AVERAGE = 4, /* sum(x*w)/sum(w), 2 paired arrays (x, w)  */
  B200BOOT_STAT_PEARSON_R = 5,        /* corrcoef(a,b)[0,1], 2 paired arrays      */
  B200BOOT_STAT_MEDIAN = 6,           /* np.median, per column                    */
  B200BOOT_STAT_SUM = 7,              /* np.sum, 1 array                          */
  B200BOOT_STAT_SLOPE = 8,            /* np.polyfit(a, b, 1)[0], 2 paired arrays  */
  B200BOOT_STAT_INTERCEPT = 9,        /* np.polyfit(a, b, 1)[1], 2 paired arrays  */
  B200BOOT_STAT_LINEAR_FIT = 10       /* np.polyfit(a, b, 1): TWO outputs per resample (the
                                       * statistic of the reference's paired tests,
                                       * test_bootstrap.py:211-238, 286-355)     */
} b200boot_stat;

/* Comparison operators for b200boot_count (bootstrap.py:583 uses LT; pval's
 * default compfunction, bootstrap.py:793, is GT against 0). */
typedef enum b200boot_cmp {
  B200BOOT_CMP_LT = 0,
  B200BOOT_CMP_LE = 1,
  B200BOOT_CMP_GT = 2,
  B200BOOT_CMP_GE = 3,
  B200BOOT_CMP_BETWEEN = 4 /* t0 <= s <= t1 */
} b200boot_cmp;

/* How a data set of N rows is cut into shared-memory cells (tiles). */
typedef struct b200boot_plan {
  int32_t log2_tile; /* full tiles hold 2^log2_tile rows                        */
  int32_t n_arrays;
  int64_t n_full;    /* number of full power-of-two tiles (0 => single cell)   */
  int64_t rem;       /* rows in the trailing remainder cell (may be 0)         */
  int64_t n_tiles;   /* n_full + (rem > 0)                                     */
} b200boot_plan;

int b200boot_version(void);
/* Philox4x32 rounds of every stream of this build (7; 10 with -DB2_PHILOX_ROUNDS=10). */
int b200boot_philox_rounds(void);
const char* b200boot_last_error(void);

/* Tile plan for N rows of `n_arrays` paired float64 arrays. */
int b200boot_make_plan(int64_t n_rows, int n_arrays, b200boot_plan* out);

/* Upper bound on workspace bytes for any call below with these shapes. */
size_t b200boot_workspace_bytes(int stat_id, int64_t n_rows, int n_arrays, int64_t n_cols,
                                int64_t n_local, int n_alphas);

/* K1 — fused draw + gather + FP64 reduce.  Replaces the resample loop
 * bootstrap.py:429-432 together with the index generator bootstrap.py:677-680.
 * `data` is a HOST array of n_arrays device pointers to contiguous float64[N].
 * Resample ids [resample_lo, resample_hi) are this caller's shard; the value of
 * resample r depends only on (seed, r), never on the shard boundaries.
 * stat_out: float64[resample_hi - resample_lo]; for a statistic with n_out outputs
 * (B200BOOT_STAT_LINEAR_FIT: 2) float64[n_out][resample_hi - resample_lo]. */
int b200boot_resample_stat(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                           uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                           double* stat_out, void* ws, size_t ws_bytes, void* stream);

/* The same with the data possibly still in flight: `data_ready_event` (a cudaEvent_t, or
 * NULL) is recorded by the caller after its host-to-device copy of `data` on another stream.
 * The occupancy chains (which do not read the data) are enqueued first; everything that
 * reads the data waits for the event on `stream`. */
int b200boot_resample_stat_after(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                                 uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                 double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                 void* data_ready_event);

/* b200boot_resample_stat with the rows still in HOST memory (float64, contiguous; pageable or
 * pinned): the occupancy chains are enqueued first, then each array is copied to the caller's
 * device rows `data_dev[a]` (n_rows doubles) on `copy_stream` (a stream other than `stream`), and
 * only what reads the data waits for the copy — staging a pageable array and the transfer run
 * under K0.  The range must not be empty. */
int b200boot_resample_stat_host(const double* const* host_data, double* const* data_dev, int n_arrays,
                                int64_t n_rows, int stat_id, uint64_t seed, int64_t resample_lo,
                                int64_t resample_hi, double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                void* copy_stream);

/* K1i — parity mode: the caller supplies the index arrays (e.g. the reference's
 * own PCG64 draws).  Replaces bootstrap.py:431 only.  indices: int64[n_sets][N]
 * row-major, values in [0, N).  stat_out: float64[n_sets] ([n_out][n_sets]). */
int b200boot_resample_stat_indexed(const double* const* data, int n_arrays, int64_t n_rows,
                                   int stat_id, const int64_t* indices, int64_t n_sets,
                                   double* stat_out, void* ws, size_t ws_bytes, void* stream);

/* K1c — column-batched moment statistics (mean / sum / var / std) over data
 * float64[N][n_cols] row-major (ld = row stride): all columns share one index set per
 * resample (bootstrap.py:431 on 2-D data).  Available when a row-major block of W >= 4
 * columns fits in shared memory (b200boot_columns_block_width > 0, i.e. N <= 7168);
 * column d of the result equals the 1-D run on column d.
 * stat_out: float64[n_cols][n_local]; ws: b200boot_workspace_bytes(stat, N, 1, n_cols, ...). */
int b200boot_columns_block_width(int stat_id, int64_t n_rows);
int b200boot_resample_stat_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                   int stat_id, uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                   double* stat_out, void* ws, size_t ws_bytes, void* stream);
/* K3c — closed-form jackknife of those statistics for every column (any N):
 * out float64[n_cols][8] as b200boot_jackknife. */
int b200boot_jackknife_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, int stat_id,
                               double* out, void* stream);

/* Which order statistics make a rank statistic.  np.median is the mean of the two middle
 * order statistics; np.quantile / np.percentile (method 'linear') interpolate between
 * ranks floor(h) and floor(h) + 1, h = (N - 1) q, with numpy's _lerp.  The caller fixes
 * ranks and weight (it knows numpy's rounding of h) for a sample of N rows and for a
 * leave-one-out sample of N - 1 rows (used by the jackknife entry points only).
 * Every median entry point takes `rule`; NULL means np.median. */
typedef struct b200boot_rank_rule {
  int64_t k_lo, k_hi;          /* 0-based ranks, k_lo <= k_hi <= k_lo + 1, k_hi < N            */
  double gamma;                /* weight of the upper rank (mode 1)                           */
  int64_t loo_k_lo, loo_k_hi;  /* the same for N - 1 rows                                     */
  double loo_gamma;
  int32_t mode;                /* 0: (lo + hi) / 2 as np.median; 1: numpy 'linear' interpolation */
  int32_t reserved;
} b200boot_rank_rule;

/* K6 — multi-column moment statistics as (count matrix) x (data) GEMM.  For (N, D) data the
 * column sums of resample r are sum_i c[r][i] * x[i][:], c[r][i] = multiplicity of row i in
 * resample r (all columns share the index set, bootstrap.py:431): one FP64 GEMM
 * C[n x N] * X[N x D] replaces the gather of bootstrap.py:429-432 for every column at once.
 * The GEMM is a plain library DGEMM issued by the caller between these two calls;
 * the library supplies the count matrix (multiplicities of exactly the indices
 * b200boot_fill_indices yields, N <= 28 672) and the finisher. */
int b200boot_gemm_supported(int stat_id, int64_t n_rows);
/* out: float64[resample_hi - resample_lo][N] row-major */
int b200boot_count_matrix(uint64_t seed, int64_t n_rows, int64_t resample_lo, int64_t resample_hi,
                          double* out, void* stream);
/* The count matrix of caller-supplied index sets, int64[n_sets][N] (e.g. from b200boot_fill_indices
 * for data longer than one cell): out float64[n_sets][N], overwritten. */
int b200boot_counts_from_indices(const int64_t* indices, int64_t n_sets, int64_t n_rows, double* out,
                                 void* stream);
/* s1: float64[n_cols][n_local] = X^T C^T (in place -> the statistic); s2: the same for the squares of
 * the column-centred data (B200BOOT_STAT_VAR / _STD, where s1 is of the centred data too), else NULL */
int b200boot_finish_columns(int stat_id, int64_t n_rows, int64_t n_cols, int64_t n_local, double* s1,
                            const double* s2, void* stream);

/* K1m — column-batched order statistic: data float64[N][n_cols] row-major
 * (ld = row stride in elements); every column shares one index set per
 * resample, as x[indices] does for 2-D data at bootstrap.py:431.
 * stat_out: float64[n_cols][n_local] (column-major as everywhere).
 * jack_out (optional, may be NULL): float64[n_cols][8] as b200boot_jackknife_median_columns —
 * the columns are sorted for the resampling anyway, so the leave-one-out summary
 * (bootstrap.py:595-615) comes with one more small launch instead of a second sort. */
int b200boot_resample_median_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                     uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                     double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                     const b200boot_rank_rule* rule, double* jack_out);

/* K1m in parity mode: index sets supplied by the caller, int64[n_sets][N]. */
int b200boot_resample_median_columns_indexed(const double* data, int64_t n_rows, int64_t n_cols,
                                             int64_t ld, const int64_t* indices, int64_t n_sets,
                                             double* stat_out, void* ws, size_t ws_bytes, void* stream,
                                             const b200boot_rank_rule* rule);

/* K1mt — median of a long 1-D array in rank space (n_rows beyond the 28 672-row
 * resident limit of K1m).  `sorted_vals` is the data sorted ascending (e.g. by
 * b200boot_sort_columns); the resample is sorted_vals[indices] with `indices` exactly
 * what b200boot_fill_indices yields for (seed, n_rows, 1 array).  Only the tile that
 * holds the middle rank is drawn: O(n_tiles + tile) work per resample.
 * stat_out: float64[resample_hi - resample_lo]. */
int b200boot_resample_median_sorted(const double* sorted_vals, int64_t n_rows, uint64_t seed,
                                    int64_t resample_lo, int64_t resample_hi, double* stat_out, void* ws,
                                    size_t ws_bytes, void* stream, const b200boot_rank_rule* rule);
/* K1mr — rank statistics of LONG (N, D) data in row space (bootstrap.py:431 gathers whole rows:
 * all columns of a resample share the index set).  b200boot_rank_columns sorts every column once:
 * sorted_out float64[D][N] ascending, rank_out int32[N][D] = position of row i in column d's
 * order; ws: b200boot_rank_columns_ws_bytes (more workspace = more columns per sorting pass).
 * b200boot_median_rows_indexed then takes index sets int64[n_sets][N] — what
 * b200boot_fill_indices yields, or the caller's own (out-of-range entries are ignored) — and
 * finds each column's order statistics exactly by a two-level counting select over the drawn
 * rows' ranks: stat_out[d * out_ld + set].  1 <= N <= 2^24. */
size_t b200boot_rank_columns_ws_bytes(int64_t n_rows, int64_t n_cols);
int b200boot_rank_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, double* sorted_out,
                          int32_t* rank_out, void* ws, size_t ws_bytes, void* stream);
/* The same pair with an explicit row stride of the rank table, rank_out int32[N][rank_ld],
 * rank_ld >= n_cols (entries beyond n_cols are never interpreted): with rank_ld a multiple of
 * four the select reads a row's group of four ranks with one 128-bit load. */
int b200boot_rank_columns_ld(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, double* sorted_out,
                             int32_t* rank_out, int64_t rank_ld, void* ws, size_t ws_bytes, void* stream);
int b200boot_median_rows_indexed_ld(const double* sorted_vals, const int32_t* rank, int64_t rank_ld, int64_t n_rows,
                                    int64_t n_cols, const int64_t* indices, int64_t n_sets, double* stat_out,
                                    int64_t out_ld, void* stream, const b200boot_rank_rule* rule);
int b200boot_median_rows_supported(int64_t n_rows);
int b200boot_median_rows_indexed(const double* sorted_vals, const int32_t* rank, int64_t n_rows, int64_t n_cols,
                                 const int64_t* indices, int64_t n_sets, double* stat_out, int64_t out_ld,
                                 void* stream, const b200boot_rank_rule* rule);
/* NaN propagation for rank statistics: np.median / np.quantile of a sample holding a NaN is NaN
 * (so the reference's statistic is NaN for every resample that draws a NaN row), while the rank
 * kernels order NaNs last.  Columns with NaNs are flagged (flags int32[n_cols + 1], the last entry
 * = any) and patched afterwards: out[set] = NaN for every index set int64[n_sets][N] that holds a
 * row i with src[i * stride] NaN; rows of a [n_cols][width] table (the jackknife summaries) of
 * flagged columns become NaN. */
int b200boot_nan_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t* flags,
                         void* stream);
int b200boot_nan_fixup(const double* src, int64_t stride, int64_t n_rows, const int64_t* indices, int64_t n_sets,
                       double* out, void* stream);
int b200boot_nan_poison_rows(double* table, int64_t n_cols, int width, const int32_t* flags, void* stream);
/* A whole ci() of a rank statistic (np.median, or a quantile through `rule`) of ONE resident
 * column (N <= 28 672) in one library call: (host rows staged,) the column sorted, its jackknife
 * summary, the resamples (K1m1) and the single-CTA tail of b200boot_ci_tail; returns after the
 * stream has been synchronised with the 32-double result block in result_host
 * ([25] = 1 when the column holds NaNs: numpy's NaN rule applies and the caller takes the general
 * path).  data: device rows (n_rows doubles; the destination of the copy when host_data is given);
 * result_dev: 48 doubles; ws: b200boot_workspace_bytes(B200BOOT_STAT_MEDIAN, n_rows, 1, 1, n_samples, 1). */
int b200boot_ci_median_small_supported(int64_t n_rows, int64_t n_samples, int n_alphas);
int b200boot_ci_median_small(const double* data, const double* host_data, double* staging, int64_t n_rows,
                             uint64_t seed, int64_t n_samples, const double* alphas, int n_alphas, int bca,
                             int need_jack, const b200boot_rank_rule* rule, double* stat_out, double* result_dev,
                             double* result_host, void* ws, size_t ws_bytes, void* stream);

/* K3m on a sorted 1-D array of any length: out float64[8] as b200boot_jackknife. */
int b200boot_jackknife_median_sorted(const double* sorted_vals, int64_t n_rows, double* out, void* stream,
                                     const b200boot_rank_rule* rule);

/* K2 — materialises exactly the indices K1 draws; backs the Python
 * `bootstrap_indices` iterator (bootstrap.py:667-680).
 * out: int64[resample_hi - resample_lo][N] row-major.  `n_arrays` selects the
 * same tile plan K1 uses for that many paired arrays. */
int b200boot_fill_indices(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample_lo,
                          int64_t resample_hi, int64_t* out, void* ws, size_t ws_bytes,
                          void* stream);

/* K2b — moving-block bootstrap index sets (bootstrap.py:747-787): ceil(N/L) block
 * starts per resample, uniform on [0, N) when `wrap` else on [0, N-L) as numpy's
 * half-open draw gives; indices start+0..L-1 (mod N when wrapping), truncated to N.
 * out: int64[resample_hi - resample_lo][N]. */
int b200boot_fill_moving_block(uint64_t seed, int64_t n_rows, int64_t block_length, int wrap,
                               int64_t resample_lo, int64_t resample_hi, int64_t* out, void* stream);

/* K1b — the moving-block bootstrap as a fused resampling plan: the block starts that
 * b200boot_fill_moving_block materialises, the L-contiguous gathers and the moment statistic in
 * one kernel (no index array).  Arguments as b200boot_resample_stat plus block_length / wrap;
 * ws as for b200boot_jackknife. */
int b200boot_resample_stat_moving_block(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                                        uint64_t seed, int64_t block_length, int wrap, int64_t resample_lo,
                                        int64_t resample_hi, double* stat_out, void* ws, size_t ws_bytes,
                                        void* stream);

/* K2c — subsamples without replacement (bootstrap.py:717-744): per sample the first
 * `size` entries of a uniform random permutation of 0..N-1 (argsort of 63-bit Philox
 * keys).  out: int64[sample_hi - sample_lo][size]; ws: 12 bytes per padded row entry. */
int b200boot_fill_subsample(uint64_t seed, int64_t n_rows, int64_t size, int64_t sample_lo,
                            int64_t sample_hi, int64_t* out, void* ws, size_t ws_bytes, void* stream);

/* Small calls (rows x arrays <= 28672 in one resident cell, a scalar moment statistic,
 * n_samples <= 32768, <= 8 levels): the whole ci() of bootstrap.py:222-504 as ONE kernel launch.
 * Every CTA keeps the rows in shared memory; the last CTA of the grid runs the closed-form
 * jackknife (bootstrap.py:595-615) while the others resample (the single-cell Lemire stream of
 * b200boot_resample_stat, bit-equal statistics); the CTA that finishes last counts stat < ostat
 * (bootstrap.py:583), forms the BCa levels and ranks (bootstrap.py:609-629, :456) and selects the
 * order statistics exactly (bootstrap.py:485-490; a monotone bucket map + exact ranking inside the
 * bucket, radix descent for degenerate vectors).  The 32-double result block
 *   [0..7] order statistics   [8..15] the ranks used   [16] #{stat < ostat}   [17..24] jackknife
 *   summary as b200boot_jackknife writes it
 * is stored to result_dev and, when `result_host` is pinned memory the device can address, straight
 * into it by the kernel (otherwise by one device-to-host copy; NULL: not at all).
 * The ranks are speculative (device evaluation of the reference's nppf / ncdf): the caller
 * recomputes them on the host from [16] and [17..24] and selects again if they differ.
 * stat_out: float64[n_samples] device; result_dev: 48 doubles device, ZERO before the first call
 * (a ticket counter lives behind the results and is left at zero by every call); ws: 4096 bytes. */
/* cudaStreamSynchronize on the caller's stream (the host wait that follows b200boot_ci_small) */
int b200boot_stream_synchronize(void* stream);
int b200boot_ci_small_supported(int stat_id, int n_arrays, int64_t n_rows, int64_t n_samples, int n_alphas);
int b200boot_ci_small(const double* const* data, int n_arrays, int64_t n_rows, int stat_id, uint64_t seed,
                      int64_t n_samples, const double* alphas, int n_alphas, int bca, int need_jack,
                      double* stat_out, double* result_dev, double* result_host, void* ws, size_t ws_bytes,
                      void* stream);
/* pval of a short job (bootstrap.py:790-866) through the same single kernel: the statistic vector
 * and, instead of levels and order statistics, #{stat op t0 [, t1]} (op a b200boot_cmp) in
 * result[16].  data: device rows; host_data (may be NULL): the rows are still on the host and are
 * staged through `staging` into `data` first.  Returns after the stream has been synchronised. */
int b200boot_pval_small(const double* const* data, const double* const* host_data, double* const* staging,
                        int n_arrays, int64_t n_rows, int stat_id, uint64_t seed, int64_t n_samples, int cmp_op,
                        double t0, double t1, double* stat_out, double* result_dev, double* result_host, void* ws,
                        size_t ws_bytes, void* stream);

/* The tail of b200boot_ci_small as its own launch, for a statistic vector of any origin (K1 of
 * the general path, a gathered sharded vector), n_samples <= 2^21: #{stat < jack[0]}
 * (bootstrap.py:583), speculative BCa levels and ranks, exact order statistics — one CTA, the
 * 32-double result block of b200boot_ci_small (jack: the 8 doubles b200boot_jackknife wrote, on
 * the device; may be NULL when bca == 0).  result_dev: 32 doubles; result_host as above. */
int b200boot_ci_tail_supported(int64_t n_samples, int n_alphas);
int b200boot_ci_tail(const double* stat, int64_t n_samples, const double* jack, const double* alphas, int n_alphas,
                     int bca, double* result_dev, double* result_host, void* stream);
/* The same from HOST rows (float64, contiguous): each array is copied into the caller's pinned
 * `staging[a]`, from there to the caller's device rows `data_dev[a]` (both >= n_rows doubles) on
 * `stream`, the kernel follows, and the call returns after cudaStreamSynchronize: result_host
 * (required) holds the result block. */
int b200boot_ci_small_host(const double* const* host_data, double* const* staging, double* const* data_dev,
                           int n_arrays, int64_t n_rows, int stat_id, uint64_t seed, int64_t n_samples,
                           const double* alphas, int n_alphas, int bca, int need_jack, double* stat_out,
                           double* result_dev, double* result_host, void* ws, size_t ws_bytes, void* stream);

/* multi='independent' with a two-sample statistic f(s_a(a), s_b(b)) (bootstrap.py:438-444,
 * 601-607, 708-714).  f is one of b200boot_op. */
enum b200boot_op { B200BOOT_OP_SUB = 0, B200BOOT_OP_ADD = 1, B200BOOT_OP_DIV = 2, B200BOOT_OP_MUL = 3,
                   B200BOOT_OP_LOG_RATIO = 4 };
/* out[i] = f(a[i], b[i]): the two arrays' resample statistics combined */
int b200boot_combine(int op, const double* a, const double* b, double* out, int64_t n, void* stream);
/* K3 as vectors: values[i] = statistic of the data without row i (moment statistics; closed
 * forms), summary[8] as b200boot_jackknife.  ws as for b200boot_jackknife. */
int b200boot_jackknife_values(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                              double* values, double* summary, void* ws, size_t ws_bytes, void* stream);
/* the same for rank statistics of an ascending array: values[r] = statistic without sorted row r */
int b200boot_jackknife_values_sorted(const double* sorted_vals, int64_t n_rows, double* values,
                                     double* summary, void* stream, const b200boot_rank_rule* rule);
/* pooled jackknife of f(s_a, s_b): the n_a + n_b values f(loo_a[j], s_b), f(s_a, loo_b[j]);
 * out[8] = {f(s_a, s_b), their mean, sum d^2, sum d^3, acceleration, 0, 0, 0}.
 * s_a, s_b: device scalars.  ws: 64 KB. */
int b200boot_jackknife_pooled(int op, const double* loo_a, int64_t n_a, const double* s_a,
                              const double* loo_b, int64_t n_b, const double* s_b, double* out,
                              void* ws, size_t ws_bytes, void* stream);

/* method='abc' (bootstrap.py:507-567) for np.average: the reference's weighted means with numpy's
 * own arithmetic (single IEEE operations, pairwise summation order), so that its finite
 * differences — rounding noise included — are reproduced:
 *   tp[i] = average(x, p0 + ep (e_i - p0)), tm[i] = average(x, p0 - ep (e_i - p0)), p0 = 1/N;
 *   out[v] = average(x, weights[v][:]) for explicit weight vectors float64[n_vectors][N];
 *   weights = NULL, n_vectors = 1: out[0] = np.mean(x), numpy's pairwise sum over N. */
int b200boot_abc_identity_averages(const double* x, int64_t n_rows, double ep, double* tp, double* tm,
                                   void* stream);
int b200boot_weighted_averages(const double* x, int64_t n_rows, const double* weights, int64_t n_vectors,
                               double* out, void* stream);

/* Occupancies of the multinomial chain (K0), for inspection and tests:
 * counts int32[n_tiles][n_local]; class_counts (may be NULL) int32[n_full][n_local][16],
 * the split of every full tile's draws over its 16 shared-memory bank classes.
 * Only meaningful when the plan has > 1 tile. */
int b200boot_tile_counts(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample_lo,
                         int64_t resample_hi, int32_t* counts, int32_t* class_counts, void* stream);

/* The same occupancies as the resample path computes them (two-level tile chains, class split
 * tree level by level): class_counts uint16[n_full][n_local][16].  Values equal
 * b200boot_tile_counts'.  ws: b200boot_tile_counts_packed_ws_bytes(n_rows, n_arrays, n_local). */
size_t b200boot_tile_counts_packed_ws_bytes(int64_t n_rows, int n_arrays, int64_t n_local);
int b200boot_tile_counts_packed(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample_lo,
                                int64_t resample_hi, int32_t* counts, uint16_t* class_counts,
                                void* ws, size_t ws_bytes, void* stream);

/* Plain gather out[i] = data[idx[i]] (bit-exactness probe for bootstrap.py:431). */
int b200boot_gather(const double* data, int64_t n_rows, const int64_t* idx, int64_t n_idx,
                    double* out, void* stream);

/* K3 — closed-form leave-one-out pass.  Replaces the O(N^2) loop
 * bootstrap.py:595-599 and the moments at bootstrap.py:609-615.
 * out (device) float64[8]: [0] statistic on the full data (bootstrap.py:580),
 * [1] mean of the jackknife statistics, [2] sum d^2, [3] sum d^3,
 * [4] acceleration a = [3] / (6 * [2]^1.5); d = jmean - jstat_i.
 * A statistic with n_out outputs fills float64[n_out][8]. */
int b200boot_jackknife(const double* const* data, int n_arrays, int64_t n_rows, int stat_id,
                       double* out, void* ws, size_t ws_bytes, void* stream);

/* Column-batched K3 for the median: out float64[n_cols][8]. */
int b200boot_jackknife_median_columns(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                      double* out, void* ws, size_t ws_bytes, void* stream,
                                      const b200boot_rank_rule* rule);

/* K4 — count reduction per column: counts[d] = #{ r : stat[d][r] cmp t0[d] (,t1[d]) }.
 * Replaces np.sum(stat < ostat, axis=0) at bootstrap.py:583 and the pval mean
 * at bootstrap.py:861-866.  counts: int64[n_cols], overwritten. */
int b200boot_count(const double* stat, int64_t n_local, int64_t n_cols, int cmp_op,
                   const double* t0, const double* t1, int64_t* counts, void* stream);

/* K5 — exact k-th order statistics without sorting; replaces stat.sort(axis=0)
 * (bootstrap.py:445) + the picks at bootstrap.py:485-490.  NaNs order last.
 * ranks: int64[n_alphas][n_cols] GLOBAL 0-based ranks; out: float64[n_alphas][n_cols].
 * The select is an 8-pass MSB radix descent.  Single-GPU callers use
 * b200boot_select; sharded callers run begin / (hist, all-reduce hist, pick) x 8 / end,
 * all-reducing the int64 histogram buffer returned by b200boot_select_hist_buffer
 * between hist and pick. */
int b200boot_select(const double* stat, int64_t n_local, int64_t n_cols, const int64_t* ranks,
                    int n_alphas, double* out, void* ws, size_t ws_bytes, void* stream);
int b200boot_select_begin(const int64_t* ranks, int n_alphas, int64_t n_cols, void* ws,
                          size_t ws_bytes, void* stream);
int b200boot_select_hist(const double* stat, int64_t n_local, int64_t n_cols, int n_alphas,
                         int pass, void* ws, size_t ws_bytes, void* stream);
int b200boot_select_pick(int n_alphas, int64_t n_cols, int pass, void* ws, size_t ws_bytes,
                         void* stream);
int b200boot_select_end(int n_alphas, int64_t n_cols, double* out, void* ws, size_t ws_bytes,
                        void* stream);
/* Device pointer + element count of the histogram buffer inside `ws`. */
int b200boot_select_hist_buffer(int n_alphas, int64_t n_cols, void* ws, size_t ws_bytes,
                                int64_t** hist, int64_t* n_elems);

/* Ascending in-place sort of every column (NaNs last); only for return_dist=True
 * (bootstrap.py:445, 501-504).  stat: float64[n_cols][n]. */
int b200boot_sort_columns(double* stat, int64_t n, int64_t n_cols, void* ws, size_t ws_bytes,
                          void* stream);

/* K6i — (N, D) moment statistics as count matrix x data on the tensor cores, exactly
 * (tcgen05.mma kind::i8: uint8 counts x eight int8 base-128 digits of every column in fixed
 * point, int32 accumulators in TMEM, int64 recombination; csrc/k6i_mma.cuh).  mean / sum / var /
 * std of every column for resamples [lo, hi): stat_out float64[D][hi - lo]; `second` (same shape)
 * is scratch for var / std.  The counts are those of the single-cell Lemire stream
 * (indices = NULL, n_rows within one cell) or of the materialised index sets `indices`
 * int64[hi - lo][n_rows] (any n_rows < 2^25: the int32 digit sums stay exact).  Columns holding non-finite values are
 * recomputed the gather way inside the call (a resample that never draws the row stays finite).
 * ws: b200boot_count_gemm_ws_bytes(n_rows, n_cols, hi - lo). */
size_t b200boot_count_gemm_ws_bytes(int64_t n_rows, int64_t n_cols, int64_t n_local);
int b200boot_resample_stat_columns_gemm(const double* data, int64_t n_rows, int64_t n_cols, int64_t ld,
                                        int stat_id, uint64_t seed, int64_t resample_lo, int64_t resample_hi,
                                        const int64_t* indices, double* stat_out, double* second, void* ws,
                                        size_t ws_bytes, void* stream);
/* test hook of the same kernel: c[m][n] (int32) = a[m][k] (uint8) . b[n][k]^T (int8), k a multiple
 * of 128, n a multiple of 256 */
int b200boot_debug_i8_gemm(const uint8_t* a, const int8_t* b, int32_t* c, int64_t m, int64_t n_digit_rows,
                           int64_t k_pad, void* stream);

/* Instrumentation for bench.py: cumulative number of kernels this library has
 * launched (host counter), and optional CUDA-event timing of the K1 launch on the
 * caller's stream.  b200boot_profile_last_k1_ms synchronises on the stop event. */
int64_t b200boot_launch_count(void);
int b200boot_profile_enable(int on);
int b200boot_profile_last_k1_ms(float* ms);
/* the occupancy kernels (K0) and the K1 launch of the calling thread's last resample call */
int b200boot_profile_last_ms(float* k0_ms, float* k1_ms);

/* The reference's normal cdf (bootstrap.py:72-76: np.vectorize of 0.5 * (1 + math.erf(x / sqrt(2))))
 * over an array on the HOST, with the C library's erf — the function math.erf calls — so the
 * values equal the reference's bit for bit without a Python-level loop (a BCa tail over 10^4
 * columns evaluates 2 x 10^4 levels).  x, out: host float64[n]; may alias. */
void b200boot_host_ncdf(const double* x, double* out, int64_t n);

/* Host-side probes built from the same __host__ __device__ sources as the
 * kernels (no GPU needed); used by the CPU test-suite only. */
void b200boot_host_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);   /* 10 rounds */
void b200boot_host_philox4x32_r(int rounds /* 7 or 10 */, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int64_t b200boot_host_binomial(uint64_t seed, int64_t resample, int64_t cell, int64_t n,
                               double p);
/* Bin(n, 1/2) as the nodes of the bank-class split tree draw it (popcount below 20, BTRS above) */
int64_t b200boot_host_binomial_half(uint64_t seed, int64_t resample, int64_t cell, int64_t n);
int b200boot_host_tile_counts(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample,
                              int32_t* counts /* [n_tiles] */);
int b200boot_host_class_counts(uint64_t seed, int64_t resample, int64_t tile, int64_t n_tile_draws,
                               int32_t* counts16);
int b200boot_host_indices(uint64_t seed, int64_t n_rows, int n_arrays, int64_t resample,
                          int64_t* out /* [n_rows] */);

#ifdef __cplusplus
}
#endif
#endif /* B200BOOT_H */
