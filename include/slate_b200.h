/* slate_b200.h -- C ABI of libslate_b200.so
 *
 * B200-native (sm_100a) replacement for the device BLAS / tile-kernel layer that
 * SLATE's Target::Devices hot path calls:
 *
 *   seam 1 (vendor BLAS):  blas::batch::gemm / herk / syrk / trsm, blas::herk,
 *                          blas::swap, lapack::potrf(queue)
 *                          -> cublas{D,Z}gemmBatched, cublas?trsmBatched, per-tile
 *                             cublas?syrk/herk, cusolverDn?potrf, cublas?swap
 *                          (reference: blaspp/src/device_batch_gemm.cc:27-155,
 *                           device_batch_herk.cc:30-75, device_batch_trsm.cc:27-130,
 *                           cublas_wrappers.cc:1623-1811, lapackpp/src/cuda/cuda_potrf.cc,
 *                           src/internal/internal_swap.cc:674-688)
 *   seam 2 (tile kernels): namespace slate::device in
 *                          include/slate/internal/device.hh:92-281 (src/cuda/.cu files)
 *
 * Conventions (identical to the reference's boundary):
 *   - plain pointers and sizes, no C++ / torch types;
 *   - every pointer is a DEVICE pointer unless the name says host; pointer arrays
 *     (T* const*) are device arrays of device pointers, as for cublas*Batched and
 *     slate::device::*;
 *   - enums are the blaspp character codes: layout 'C'|'R', op 'N'|'T'|'C',
 *     uplo 'L'|'U'|'G', diag 'N'|'U', side 'L'|'R', norm 'M'|'O'|'I'|'F';
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on
 *     that stream, never synchronises, never allocates tile memory;
 *   - the current CUDA device must be the one that owns the pointers
 *     (the C++ shim calls cudaSetDevice(queue.device()) first, as the reference does);
 *   - return value: 0 on success, a negative SB200_E* code for argument errors,
 *     or a positive cudaError_t from the launch.  No exceptions cross the ABI.
 *   - m, n, k, batch == 0 are quick returns (reference: device_geadd.cu:127-129).
 *
 * Type suffixes: s = float, d = double, c = complex<float>, z = complex<double>
 * (complex passed as interleaved (re, im) pairs, i.e. pointer-compatible with
 * std::complex / cuComplex).
 */
#ifndef SLATE_B200_H
#define SLATE_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_VERSION 100

enum {
    SB200_OK          =  0,
    SB200_EINVAL      = -1,   /* bad enum / negative size / ld too small          */
    SB200_ENOTSUP     = -2,   /* combination not implemented                       */
    SB200_ENOMEM      = -3,
    SB200_ENODEV      = -4,   /* no CUDA device: the library has no CPU fallback   */
    SB200_ENCCL       = -5
};

typedef void* sb200_stream_t;          /* cudaStream_t */
typedef struct { float  re, im; } sb200_c32;
typedef struct { double re, im; } sb200_c64;

int         sb200_version(void);
const char* sb200_strerror(int code);
/* number of visible CUDA devices, or SB200_ENODEV */
int         sb200_device_count(void);
/* number of kernels launched by this library since load (all threads); bench.py's gpu_launches */
int64_t     sb200_launch_count(void);

/* ===========================================================================
 * Type families.  Every kernel entry point exists for the four SLATE scalar types; the
 * declarations below are generated with an X-macro: SB200_FOR_TYPES(M) expands
 * M(suffix, scalar type, real type).
 * =========================================================================== */
#define SB200_FOR_TYPES(M) \
    M(s, float,     float)  \
    M(d, double,    double) \
    M(c, sb200_c32, float)  \
    M(z, sb200_c64, double)

/* ---------------------------------------------------------------------------
 * Batched tile GEMM:  C_t = alpha * op(A_t) * op(B_t) + beta * C_t,  t < batch
 * replaces blas::batch::gemm fixed-size path -> cublas?gemmBatched
 * (blaspp/src/device_batch_gemm.cc:76-130; call sites src/internal/internal_gemm.cc:498-504,
 *  internal_herk.cc:510-516).  FP64 real and complex run on the FP64 tensor-core MMA (DMMA).
 *   _batched : dA/dB/dC are DEVICE arrays of device pointers (as cublas?gemmBatched)
 *   _strided : tile t is base + t*stride (batch == 1: a single GEMM; replaces blas::gemm(queue),
 *              blaspp/src/device_gemm.cc -> cublas?gemm)
 * ------------------------------------------------------------------------- */
#define SB200_DECL_GEMM(X, T, R) \
int sb200_gemm_batched_##X(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k, \
                           T alpha, const T* const* dA, int64_t lda, \
                           const T* const* dB, int64_t ldb, \
                           T beta, T* const* dC, int64_t ldc, \
                           int64_t batch, sb200_stream_t stream); \
int sb200_gemm_strided_##X(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k, \
                           T alpha, const T* dA, int64_t lda, int64_t strideA, \
                           const T* dB, int64_t ldb, int64_t strideB, \
                           T beta, T* dC, int64_t ldc, int64_t strideC, \
                           int64_t batch, sb200_stream_t stream);
SB200_FOR_TYPES(SB200_DECL_GEMM)

/* Same operation with per-array element offsets (A_t = dA[t] + offA, ...): lets the
 * host runtime address sub-blocks of resident tiles without rebuilding pointer arrays. */
int sb200_gemm_batched_off_d(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                             double alpha, const double* const* dA, int64_t offA, int64_t lda,
                             const double* const* dB, int64_t offB, int64_t ldb,
                             double beta, double* const* dC, int64_t offC, int64_t ldc,
                             int64_t batch, sb200_stream_t stream);

/* ---------------------------------------------------------------------------
 * Mixed-precision trailing update on the 5th-generation tensor cores (tcgen05.mma, accumulators in
 * TMEM): FP32 emulated by three TF32 MMAs per product (hi*hi + hi*lo + lo*hi).  This is the
 * contraction of the LOW-PRECISION factorisation inside gesv_mixed (src/gesv_mixed.cc:106-300,
 * where the reference calls internal::gemm<Devices, float> -> cublasSgemmBatched).
 *
 * Operands are PACKED once per panel into the tensor core's canonical shared-memory layout and
 * then streamed by the GEMM with 1-D TMA bulk copies:
 *   side 'A': the m x k operand op(X) (op 'N': X is m x k; 'T': X is k x m), units of 128 rows;
 *   side 'B': the k x n operand op(X) packed as its n x k transpose, units of 256 rows.
 * sb200_tf32x3_packed_bytes(side, rows, k) = bytes of one packed operand (rows = m or n).
 *   C_t = alpha * A_t * B_t + beta * C_t,  C_t column-major m x n (ldc), FP32.
 * ------------------------------------------------------------------------- */
size_t sb200_tf32x3_packed_bytes(int side, int64_t rows, int64_t k);
int sb200_tf32x3_pack_batched_s(int side, int op, int64_t rows, int64_t k,
                                const float* const* dX, int64_t ldx, void* const* dPacked,
                                int64_t batch, sb200_stream_t stream);
int sb200_gemm_tf32x3_packed_s(int64_t m, int64_t n, int64_t k, float alpha,
                               const void* const* dApacked, const void* const* dBpacked,
                               float beta, float* const* dC, int64_t ldc,
                               int64_t batch, sb200_stream_t stream);

/* ---------------------------------------------------------------------------
 * Batched HERK / SYRK on the stored triangle:
 *   C_t = alpha * op(A_t) * op(A_t)^H + beta * C_t   (herk; alpha, beta real; imag(diag) := 0)
 *   C_t = alpha * op(A_t) * op(A_t)^T + beta * C_t   (syrk; alpha, beta of the scalar type)
 * replaces blas::batch::herk/syrk = a LOOP of per-tile cublas?herk/?syrk on forked
 * streams (blaspp/src/device_batch_herk.cc:57-73, device_batch_syrk.cc) and the single-tile
 * blas::herk / blas::syrk (blaspp/src/device_herk.cc, device_syrk.cc; call site
 * src/internal/internal_herk.cc:380-385) with one launch.
 * op = 'N': A_t is n-by-k;  op = 'C'/'T': A_t is k-by-n.  For real types herk == syrk.
 * ------------------------------------------------------------------------- */
#define SB200_DECL_HERK(X, T, R) \
int sb200_herk_batched_##X(int layout, int uplo, int op, int64_t n, int64_t k, \
                           R alpha, const T* const* dA, int64_t lda, \
                           R beta, T* const* dC, int64_t ldc, \
                           int64_t batch, sb200_stream_t stream); \
int sb200_syrk_batched_##X(int layout, int uplo, int op, int64_t n, int64_t k, \
                           T alpha, const T* const* dA, int64_t lda, \
                           T beta, T* const* dC, int64_t ldc, \
                           int64_t batch, sb200_stream_t stream); \
int sb200_herk_##X(int layout, int uplo, int op, int64_t n, int64_t k, \
                   R alpha, const T* dA, int64_t lda, R beta, T* dC, int64_t ldc, sb200_stream_t stream); \
int sb200_syrk_##X(int layout, int uplo, int op, int64_t n, int64_t k, \
                   T alpha, const T* dA, int64_t lda, T beta, T* dC, int64_t ldc, sb200_stream_t stream);
SB200_FOR_TYPES(SB200_DECL_HERK)

/* ---------------------------------------------------------------------------
 * Batched TRSM with ONE triangular tile shared by the batch (SLATE replicates the
 * A pointer count times: src/internal/internal_trsm.cc:225-249):
 *   side 'L':  B_t <- alpha * op(A)^{-1} * B_t      side 'R':  B_t <- alpha * B_t * op(A)^{-1}
 * replaces blas::batch::trsm -> cublas?trsmBatched (blaspp/src/device_batch_trsm.cc:27-130).
 * `dA` is the device pointer of the single na-by-na triangular tile (na = m for 'L', n for 'R').
 * `work` is device scratch of at least sb200_trsm_work_bytes(...) bytes (the inverted
 * diagonal blocks); it may be NULL, in which case an internal per-device scratch is used.
 * ------------------------------------------------------------------------- */
size_t sb200_trsm_work_bytes_d(int side, int64_t m, int64_t n);
size_t sb200_trsm_work_bytes(int dtype /* 's','d','c','z' */, int side, int64_t m, int64_t n);
#define SB200_DECL_TRSM(X, T, R) \
int sb200_trsm_batched_##X(int layout, int side, int uplo, int op, int diag, \
                           int64_t m, int64_t n, T alpha, \
                           const T* dA, int64_t lda, \
                           T* const* dB, int64_t ldb, \
                           int64_t batch, void* work, sb200_stream_t stream);
SB200_FOR_TYPES(SB200_DECL_TRSM)

/* ---------------------------------------------------------------------------
 * Cholesky factorisation of one diagonal tile on the device, LAPACK info in *dinfo
 * (device int; 0 = ok, j > 0 = leading minor j not positive definite).
 * replaces lapack::potrf(uplo, n, dA, ldda, dinfo, queue) -> cusolverDn?potrf
 * (lapackpp/src/cuda/cuda_potrf.cc; call site src/internal/internal_potrf.cc:72-78).
 * ------------------------------------------------------------------------- */
size_t sb200_potrf_work_bytes_d(int64_t n);
#define SB200_DECL_POTRF(X, T, R) \
int sb200_potrf_tile_##X(int uplo, int64_t n, T* dA, int64_t lda, \
                         int* dinfo, void* work, sb200_stream_t stream);
SB200_FOR_TYPES(SB200_DECL_POTRF)

/* ---------------------------------------------------------------------------
 * Fused row interchanges for getrf: applies all `npiv` pivots of one panel to a whole
 * block row/column set in ONE launch.
 * replaces the per-pivot-row cublas?swap loop of internal::permuteRows<Devices>
 * (src/internal/internal_swap.cc:674-688: nb launches per block column per step).
 *
 * The block column is given as `mt` tiles stacked vertically; dTiles[t + j*mt] is tile t of
 * block column j (tile_mb rows each, ncols columns, RowMajor with leading dimension ld,
 * or ColMajor when layout == 'C').  Pivot i swaps global row i of the stack with row
 * (piv_tile[i], piv_off[i]) -- the reference's Pivot{tileIndex, elementOffset}
 * (include/slate/types.hh:84-105) -- applied in order i = 0..npiv-1 (forward) or reversed.
 * ------------------------------------------------------------------------- */
#define SB200_DECL_PERMUTE_ROWS(X, T, R) \
int sb200_permute_rows_##X(int layout, int forward, int64_t npiv, \
                           const int64_t* d_piv_tile, const int64_t* d_piv_off, \
                           T* const* dTiles, int64_t mt, int64_t ncolblocks, \
                           int64_t tile_mb, int64_t ncols, int64_t ld, \
                           sb200_stream_t stream);
SB200_FOR_TYPES(SB200_DECL_PERMUTE_ROWS)

/* ---------------------------------------------------------------------------
 * Seam 2: memory-bound tile kernels (slate::device::*, include/slate/internal/device.hh:92-281).
 * Tiles are m-by-n, column-major, leading dim ld.  `_batched` entries take DEVICE arrays of
 * device pointers (the caller uploads them, as for the reference: src/internal/internal_geadd.cc:163-165);
 * the un-suffixed single-tile entries take the tile pointer itself.
 * uplo: 'G' whole tile (ge*), 'L' / 'U' trapezoid (tz*).
 * ------------------------------------------------------------------------- */
#define SB200_DECL_TILE_OPS(X, T, R) \
/* B = alpha*A + beta*B       device::geadd / batch::geadd / tzadd (src/cuda/device_geadd.cu:121-146,234-262; device_tzadd.cu) */ \
int sb200_geadd_##X(int64_t m, int64_t n, T alpha, const T* dA, int64_t lda, T beta, T* dB, int64_t ldb, sb200_stream_t stream); \
int sb200_geadd_batched_##X(int64_t m, int64_t n, T alpha, const T* const* dA, int64_t lda, \
                            T beta, T* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream); \
int sb200_tzadd_batched_##X(int uplo, int64_t m, int64_t n, T alpha, const T* const* dA, int64_t lda, \
                            T beta, T* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream); \
/* A *= numer/denom            device::gescale / batch::gescale / batch::tzscale (device_gescale.cu, device_tzscale.cu) */ \
int sb200_gescale_##X(int64_t m, int64_t n, T numer, T denom, T* dA, int64_t lda, sb200_stream_t stream); \
int sb200_gescale_batched_##X(int64_t m, int64_t n, T numer, T denom, \
                              T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream); \
int sb200_tzscale_batched_##X(int uplo, int64_t m, int64_t n, R numer, R denom, \
                              T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream); \
/* A_ij *= R_i * C_j           device::gescale_row_col_batch (device_gescale_row_col.cu:150-210); \
 * equed 'R' rows only, 'C' cols only, 'B' both; scale vectors of the scalar type (_t) or real type (_r) */ \
int sb200_gescale_row_col_batched_##X(int equed, int64_t m, int64_t n, \
                                      const T* const* dR, const T* const* dC, \
                                      T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream); \
int sb200_gescale_row_col_real_batched_##X(int equed, int64_t m, int64_t n, \
                                      const R* const* dR, const R* const* dC, \
                                      T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream); \
/* offdiag -> A_ij (i != j), diag -> A_ii   device::geset / tzset + batch:: (device_geset.cu, device_tzset.cu) */ \
int sb200_geset_##X(int uplo, int64_t m, int64_t n, T offdiag, T diag, T* dA, int64_t lda, sb200_stream_t stream); \
int sb200_geset_batched_##X(int64_t m, int64_t n, T offdiag, T diag, \
                            T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream); \
int sb200_tzset_batched_##X(int uplo, int64_t m, int64_t n, T offdiag, T diag, \
                            T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream); \
/* transposes                   device::transpose / transpose_batch (src/cuda/device_transpose.cu): \
 * in-place square (n-by-n) and out-of-place (A m-by-n -> AT n-by-m); conj != 0 conjugates */ \
int sb200_transpose_inplace_##X(int conj, int64_t n, T* dA, int64_t lda, sb200_stream_t stream); \
int sb200_transpose_##X(int conj, int64_t m, int64_t n, const T* dA, int64_t lda, T* dAT, int64_t ldat, sb200_stream_t stream); \
int sb200_transpose_inplace_batched_##X(int conj, int64_t n, T* const* dA, int64_t lda, \
                                        int64_t batch, sb200_stream_t stream); \
int sb200_transpose_batched_##X(int conj, int64_t m, int64_t n, const T* const* dA, int64_t lda, \
                                T* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream); \
/* per-tile norms               device::genorm / henorm / synorm / synormOffdiag / trnorm (src/cuda/device_*norm.cu) \
 * norm 'M' max, 'O' (or '1', lapack::Norm::One's own character) one, 'I' inf, 'F' frobenius; scope 'M' matrix (per-tile partial results), \
 * 'C' columns (norm 'M' only: per-column max, used by colNorms). \
 * values layout as the reference (device_genorm.cu:373-445): ldv >= 1 (max), n (one), m (inf), 2 (fro: scale, sumsq); \
 * tile t writes values[t*ldv ...].  NaN-propagating max (device_util.cuh:22-25). */ \
int sb200_genorm_batched_##X(int norm, int scope, int64_t m, int64_t n, \
                             const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream); \
int sb200_henorm_batched_##X(int norm, int uplo, int64_t n, \
                             const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream); \
int sb200_synorm_batched_##X(int norm, int uplo, int64_t n, \
                             const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream); \
int sb200_synorm_offdiag_batched_##X(int norm, int64_t m, int64_t n, \
                                     const T* const* dA, int64_t lda, \
                                     R* values, int64_t ldv, int64_t batch, sb200_stream_t stream); \
int sb200_trnorm_batched_##X(int norm, int uplo, int diag, int64_t m, int64_t n, \
                             const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
SB200_FOR_TYPES(SB200_DECL_TILE_OPS)

/* precision / domain converting copies B = A   device::gecopy / tzcopy (src/cuda/device_gecopy.cu:75-235,
 * device_tzcopy.cu): suffix = <source><destination>; every pair the reference instantiates. */
#define SB200_FOR_COPY_PAIRS(M) \
    M(ss, float, float)         M(sd, float, double)        M(dd, double, double)       M(ds, double, float) \
    M(cc, sb200_c32, sb200_c32) M(cz, sb200_c32, sb200_c64) M(zz, sb200_c64, sb200_c64) M(zc, sb200_c64, sb200_c32) \
    M(sc, float, sb200_c32)     M(dz, double, sb200_c64)
#define SB200_DECL_COPY(XY, S, D) \
int sb200_gecopy_batched_##XY(int64_t m, int64_t n, const S* const* dA, int64_t lda, \
                              D* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream); \
int sb200_tzcopy_batched_##XY(int uplo, int64_t m, int64_t n, const S* const* dA, int64_t lda, \
                              D* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
SB200_FOR_COPY_PAIRS(SB200_DECL_COPY)

/* ---------------------------------------------------------------------------
 * Host runtime (C++ inside the library, exposed as opaque handles): a tile matrix
 * resident in HBM, 2-D block-cyclic over the ranks of a process grid, and the
 * drivers that schedule the trailing-matrix update with lookahead on CUDA streams.
 * Mirrors slate::Matrix / HermitianMatrix + slate::gemm / potrf / getrf for
 * Target::Devices (include/slate/Matrix.hh, src/potrf.cc:22-210, src/getrf.cc:22-244,
 * src/gemmC.cc:39-202).  See slate_b200/csrc/runtime.hh.
 * ------------------------------------------------------------------------- */
typedef struct sb200_grid_s*   sb200_grid_t;
typedef struct sb200_matrix_s* sb200_matrix_t;

/* Process grid p x q (column-major rank order, as slate::GridOrder::Col).  `nccl_unique_id`
 * is the 128-byte ncclUniqueId produced by sb200_grid_unique_id() on rank 0 and distributed
 * by the caller (bench.py uses torch.distributed); NULL when p*q == 1. */
int sb200_grid_unique_id(void* out_id_128);
int sb200_grid_create(int p, int q, int rank, const void* nccl_unique_id, sb200_grid_t* out);
int sb200_grid_destroy(sb200_grid_t g);
/* Tile broadcasts of one step in one go -- the hook for BaseMatrix::tileBcast / listBcast / listBcastMT
 * (include/slate/BaseMatrix.hh:1889-2140: per tile an MPI broadcast to the ranks that need it, then a
 * host-to-device copy).  Every rank of the grid calls it with the SAME list: range t is `bytes[t]` bytes that live at
 * `src[t]` on rank `roots[t]` (a tile, or any number of tiles that are contiguous in the root's pool) and arrive at
 * `dst[t]` on every rank (`src[t]` is read on the root only; the root's `dst[t]` may equal `src[t]`).  All arrays are
 * HOST arrays of `count` entries, the pointers in them DEVICE pointers.  Ranges of >= 4 MiB travel as scatter +
 * in-place all-gather over NVLink / NVSwitch, smaller ones as grouped NCCL broadcasts; asynchronous on `stream`.
 * A 1 x 1 grid returns at once (the reference's single-rank early exit, BaseMatrix.hh:2006). */
int sb200_bcast_tiles(sb200_grid_t g, int64_t count, const void* const* src, void* const* dst,
                      const size_t* bytes, const int* roots, sb200_stream_t stream);

/* kind: 'G' general m-by-n, 'H' Hermitian/symmetric n-by-n lower-stored.
 * layout: 'C' column-major tiles (gemm, potrf), 'R' row-major tiles (getrf on devices,
 * src/getrf.cc:51-55).  Tiles are nb-by-nb, ld = tile rows (cols for 'R'). */
int sb200_matrix_create_d(sb200_grid_t g, int kind, int layout, int64_t m, int64_t n, int64_t nb,
                          sb200_matrix_t* out);
int sb200_matrix_destroy(sb200_matrix_t A);
/* fill with the reference generator (Philox-2x64 on global indices; matgen/random.cc:54-159):
 * kind_code 0 = rand (uniform [0,1)), 1 = rand_dominant (+n on the diagonal). On the device. */
int sb200_matrix_generate_d(sb200_matrix_t A, int kind_code, int64_t seed, sb200_stream_t stream);
/* copy local tiles from/to a host column-major m-by-n array holding the GLOBAL matrix
 * (only locally owned tiles are touched); pinned or pageable. */
int sb200_matrix_from_host_d(sb200_matrix_t A, const double* hA, int64_t lda, sb200_stream_t stream);
int sb200_matrix_to_host_d(sb200_matrix_t A, double* hA, int64_t lda, sb200_stream_t stream);
/* the same for the packed LOCAL tiles only: `htiles` holds sb200_matrix_local_tiles(A) tiles of
 * nb*nb elements (ld = nb) in the matrix's pool order -- local block column, then local block row --
 * i.e. the caller-owned tile storage of Matrix::insertLocalTiles (include/slate/Matrix.hh:631-662). */
int sb200_matrix_from_host_local_d(sb200_matrix_t A, const double* htiles, sb200_stream_t stream);
int sb200_matrix_to_host_local_d(sb200_matrix_t A, double* htiles, sb200_stream_t stream);
int sb200_matrix_copy_d(sb200_matrix_t dst, sb200_matrix_t src, sb200_stream_t stream);
int64_t sb200_matrix_local_tiles(sb200_matrix_t A);

/* options common to the drivers (NULL = the defaults).  Honour-or-reject, never a silent ignore:
 *   lookahead: sb200_potrf_* honours 1 .. 8 (depth of its lookahead stream; 0 or NULL options = the library's tuned depth;
 *              the factor is bitwise independent of it); every other driver has a fixed depth of 1 and returns
 *              SB200_ENOTSUP for lookahead > 1;
 *   pivot_threshold: the LU panel always takes the largest candidate; != 1.0 returns SB200_ENOTSUP;
 *   out-of-range values return SB200_EINVAL.
 * inner_blocking only re-associates the reference panel's rank-ib updates (it does not enter the pivot rule,
 * src/internal/Tile_getrf.hh:160-447); the GPU panel has its own blocking and accepts any value >= 1. */
typedef struct {
    int64_t lookahead;        /* slate::Option::Lookahead, default 1 (src/potrf.cc:41-42, src/getrf.cc:38-43) */
    int64_t inner_blocking;   /* slate::Option::InnerBlocking (getrf panel), default 16  */
    double  pivot_threshold;  /* slate::Option::PivotThreshold, default 1.0              */
    int     reserved[8];
} sb200_options_t;

/* C = alpha A B + beta C            (slate::gemm -> gemmC, src/gemmC.cc)        */
int sb200_gemm_d(double alpha, sb200_matrix_t A, sb200_matrix_t B, double beta, sb200_matrix_t C,
                 const sb200_options_t* opts);
/* A = L L^T, lower                  (slate::potrf, src/potrf.cc); returns info via *info (host) */
int sb200_potrf_d(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info);
/* P A = L U, partial pivoting       (slate::getrf, src/getrf.cc).  pivots: host array of
 * 2*min(m,n) int64 (tileIndex, elementOffset) pairs relative to each panel, as slate::Pivots. */
int sb200_getrf_d(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
/* complex LU: the pivot of a column is the first strict maximum of cabs1 = |re| + |im| (src/internal/Tile_getrf.hh:196-237),
 * the column is scaled by the complex reciprocal of the pivot (:334-361).  1 x 1 grid (SB200_ENOTSUP on p x q grids). */
int sb200_getrf_z(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
int sb200_getrf_c(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
/* LU without pivoting (slate::getrf_nopiv, src/getrf_nopiv.cc:  A = L U, unit lower L); info = first zero pivot + 1 */
int sb200_getrf_nopiv_d(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info);
int sb200_getrf_nopiv_s(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info);
int sb200_getrf_nopiv_z(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info);     /* complex: 1 x 1 grid */
int sb200_getrf_nopiv_c(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info);
/* LU with tournament pivoting, CALU (slate::getrf_tntpiv, src/getrf_tntpiv.cc:22-395; Option::MethodLU = CALU of
 * slate::lu_factor, src/getrf.cc:324-329).  Per panel the process rows of the grid hold a tournament (local LU, binary
 * tree of pairwise LUs of the stacked candidate rows, src/internal/internal_getrf_tntpiv.cc:357-640); the rows below
 * the diagonal tile are solved against U_kk.  pivots as for sb200_getrf_d.  Every diagonal tile has to be square (the
 * reference throws otherwise, include/slate/TriangularMatrix.hh:459): SB200_ENOTSUP for other shapes. */
int sb200_getrf_tntpiv_d(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
int sb200_getrf_tntpiv_s(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
int sb200_getrf_tntpiv_z(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);     /* complex: 1 x 1 grid */
int sb200_getrf_tntpiv_c(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
/* device time of the last driver call on this matrix, milliseconds (CUDA events) */
double sb200_last_driver_ms(sb200_matrix_t A);
/* summed device time of the panel-stream critical work (diagonal factor / LU panel + solve +
 * broadcast) of the last driver call, milliseconds: the part the lookahead has to hide */
double sb200_last_driver_panel_ms(sb200_matrix_t A);
/* out4 = { driver ms, summed ms of the trailing-update GEMM launches (events on their stream),
 *          algorithmic flops of those launches, number of those launches } */
int sb200_last_driver_stats(sb200_matrix_t A, double* out4);
/* FP64 pipe peak probe (roofline denominator; MEASURED_PEAKS.json has no FP64 figure):
 * kind 0 = DMMA.8x8x4, 1 = DFMA; launches ctas_per_sm x SMs CTAs; *flops = work of the launch. */
int sb200_fp64_peak_probe(int kind, int iters, int ctas_per_sm, double* d_scratch, double* flops,
                          sb200_stream_t stream);
/* SM partition of the factorisation drivers (csrc/sm_partition.cu): with SB200_CHAIN_SMS = c > 0 (default: 4 on a
 * multi-rank grid, 0 on one rank) the panel chain -- the diagonal-tile factor the reference runs through
 * cusolverDn?potrf (src/internal/internal_potrf.cc:57-81) -- gets c whole SMs of its own (CUDA green context) and
 * every other stream the remaining ones.  Reports the split the device grants; SB200_ENOTSUP if it cannot. */
int sb200_sm_partition_probe(int chain_sms, int* sm_chain, int* sm_rest);
/* The drivers keep their device workspaces (panel rings, pointer plans, scratch) in a grow-only per-device cache
 * between calls -- the reference keeps its workspace tiles in the matrix's memory pool the same way
 * (include/slate/internal/MatrixStorage.hh:483-520, Memory::alloc / free).  This frees every idle block. */
int sb200_release_workspaces(void);

/* ---------------------------------------------------------------------------
 * Round-1 widening of the host runtime: all four scalar types, herk, the solve path and the
 * mixed-precision drivers.  The matrix handle carries its element type; the un-suffixed data
 * movement entries below work for every type (host buffers hold elements of that type).
 * ------------------------------------------------------------------------- */
int sb200_matrix_dtype(sb200_matrix_t A);                 /* 's', 'd', 'c' or 'z' */
int sb200_matrix_generate(sb200_matrix_t A, int kind_code, int64_t seed, sb200_stream_t stream);
int sb200_matrix_from_host(sb200_matrix_t A, const void* hA, int64_t lda, sb200_stream_t stream);
int sb200_matrix_to_host(sb200_matrix_t A, void* hA, int64_t lda, sb200_stream_t stream);
int sb200_matrix_from_host_local(sb200_matrix_t A, const void* htiles, sb200_stream_t stream);
int sb200_matrix_to_host_local(sb200_matrix_t A, void* htiles, sb200_stream_t stream);
int sb200_matrix_copy(sb200_matrix_t dst, sb200_matrix_t src, sb200_stream_t stream);
/* ScaLAPACK-style local arrays (Matrix::fromScaLAPACK, include/slate/Matrix.hh:75-99): `local` is this rank's
 * column-major local array (leading dimension lld >= local rows) of the 2-D block-cyclic distribution with square
 * nb x nb blocks on the matrix's p x q grid (column-major rank order): tile (i, j) sits at local offset
 * ((i / p) * nb, (j / q) * nb).  Gathers the locally owned (stored) tiles into the HBM tile pool / scatters them
 * back; on_device != 0: `local` is device memory.  ncols_local = columns the caller's array holds; SB200_EINVAL when
 * lld or ncols_local is smaller than this rank's numroc rows / columns (nothing is copied). */
int sb200_matrix_from_scalapack(sb200_matrix_t A, const void* local, int64_t lld, int64_t ncols_local, int on_device, sb200_stream_t stream);
int sb200_matrix_to_scalapack(sb200_matrix_t A, void* local, int64_t lld, int64_t ncols_local, int on_device, sb200_stream_t stream);

/* Probe-vector product with the tiles this rank stores (residual checks at bench size without a second n x n
 * matrix; the reference tester forms ||B - A X|| with gemm / hemm, test/test_posv.cc:336-342, test/test_gesv.cc:371-377):
 *   y += op(part(A)) x     op 'N' | 'C';   part 'G' whole matrix, 'L' lower triangle (diag 'U': unit), 'U' upper triangle,
 *                          'H' Hermitian matrix from its stored lower tiles;   use_abs != 0: |a_ij| instead of a_ij.
 * x, y: device vectors of the matrix's element type holding the WHOLE vector (length n resp. m, swapped for 'C');
 * contributions are added atomically, so the caller zeroes y and sums it over the ranks of the grid. */
int sb200_matrix_probe_mv(sb200_matrix_t A, int part, int op, int diag, int use_abs, const void* x, void* y, sb200_stream_t stream);

/* host-only description of the 2-D block-cyclic tile map (no GPU needed):
 * tileRank (include/slate/func.hh:96-104, GridOrder::Col), the number of tiles a rank stores, and the
 * slot of tile (i, j) in its owner's packed local buffer (sb200_matrix_{from,to}_host_local order). */
int     sb200_tile_rank(int p, int q, int64_t i, int64_t j);
int64_t sb200_local_tile_count(int kind, int p, int q, int rank, int64_t m, int64_t n, int64_t nb);
int64_t sb200_local_tile_index(int kind, int p, int q, int64_t m, int64_t n, int64_t nb, int64_t i, int64_t j);

#define SB200_DECL_RUNTIME(X, T, R) \
int sb200_matrix_create_##X(sb200_grid_t g, int kind, int layout, int64_t m, int64_t n, int64_t nb, sb200_matrix_t* out); \
/* C = alpha A B + beta C                       slate::gemm -> gemmC (src/gemmC.cc:39-202) */ \
int sb200_gemm_##X(T alpha, sb200_matrix_t A, sb200_matrix_t B, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* C = alpha A A^H + beta C, C Hermitian lower  slate::herk (src/herk.cc:25-162; real types: syrk) */ \
int sb200_herk_mat_##X(R alpha, sb200_matrix_t A, R beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* C = alpha A B^H + conj(alpha) B A^H + beta C, C Hermitian lower   slate::her2k (src/her2k.cc:27-170; real types: syr2k) */ \
int sb200_her2k_mat_##X(T alpha, sb200_matrix_t A, sb200_matrix_t B, R beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* (complex-)symmetric rank-k / rank-2k updates, no conjugation, C symmetric lower (kind 'H' storage: lower tiles) \
 * slate::syrk (src/syrk.cc), slate::syr2k (src/syr2k.cc); for real types the same as herk / her2k */ \
int sb200_syrk_mat_##X(T alpha, sb200_matrix_t A, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
int sb200_syr2k_mat_##X(T alpha, sb200_matrix_t A, sb200_matrix_t B, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* A = L L^H, lower                             slate::potrf (src/potrf.cc:22-210) */ \
int sb200_potrf_##X(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info); \
/* the same, streaming every finished block column into the packed host buffer `htiles` (order and size of \
 * sb200_matrix_to_host_local; pinned memory makes the copies overlap the factorisation) */ \
int sb200_potrf_to_host_local_##X(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info, void* htiles); \
/* the same with the INPUT streaming in as well: the matrix is read from the packed host buffer `htiles_in` (layout of \
 * sb200_matrix_from_host_local) in chunks of block columns while earlier chunks are being factored; `htiles_out` may \
 * be NULL (the factor then stays on the device only).  One rank; bitwise the same factor as sb200_potrf. */ \
int sb200_potrf_stream_##X(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info, const void* htiles_in, void* htiles_out); \
/* B <- A^{-1} B from the Cholesky factor       slate::potrs (src/potrs.cc:54-77); 1 x 1 grid this round */ \
int sb200_potrs_##X(sb200_matrix_t A, sb200_matrix_t B, const sb200_options_t* opts); \
/* C = alpha A X + beta C, A Hermitian lower, Side::Left   slate::hemm (src/hemmC.cc); 1 x 1 grid */ \
int sb200_hemm_##X(T alpha, sb200_matrix_t A, sb200_matrix_t Xm, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* C = alpha A X + beta C, A (complex-)symmetric lower, Side::Left, no conjugation   slate::symm (src/symm.cc); 1 x 1 grid */ \
int sb200_symm_##X(T alpha, sb200_matrix_t A, sb200_matrix_t X, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* the same with slate::hemm's / slate::symm's Side argument: side 'R' = C = alpha X A + beta C (X and C m x n, A n x n); \
 * 'L' forwards to the entry points above.  1 x 1 grid */ \
int sb200_hemm_side_##X(int side, T alpha, sb200_matrix_t A, sb200_matrix_t Xm, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
int sb200_symm_side_##X(int side, T alpha, sb200_matrix_t A, sb200_matrix_t X, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* B = alpha op(A) B (side 'L') or B = alpha B op(A) (side 'R'), A lower triangular (lower tiles of a kind 'H' matrix), \
 * uplo 'L', op 'N' | 'T' | 'C' (the transposed views slate::trmm takes, src/trmm.cc:61-75), diag 'N' | 'U'; \
 * uplo 'U': SB200_ENOTSUP; 1 x 1 grid */ \
int sb200_trmm_##X(int side, int uplo, int op, int diag, T alpha, sb200_matrix_t A, sb200_matrix_t B, const sb200_options_t* opts); \
/* C = alpha op(A) op(B) + beta C with the (conjugate-)transposed views slate::gemm is handed (opA / opB 'N' | 'T' | 'C'; \
 * A is stored k x m when opA != 'N', B n x k when opB != 'N').  'N','N' forwards to sb200_gemm_X; other pairs: 1 x 1 grid */ \
int sb200_gemm_op_##X(int opA, int opB, T alpha, sb200_matrix_t A, sb200_matrix_t B, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* the rank-k / rank-2k updates handed (conjugate-)transposed views (slate::herk / her2k: op 'C', real types also 'T'; \
 * slate::syrk / syr2k: op 'T'): A (and B) are stored k x n, C = alpha A^H A + beta C etc.  op 'N' forwards to the \
 * _mat entry points (grids); other ops: 1 x 1 grid, SB200_ENOTSUP for the op the routine does not take */ \
int sb200_herk_op_##X(int op, R alpha, sb200_matrix_t A, R beta, sb200_matrix_t C, const sb200_options_t* opts); \
int sb200_her2k_op_##X(int op, T alpha, sb200_matrix_t A, sb200_matrix_t B, R beta, sb200_matrix_t C, const sb200_options_t* opts); \
int sb200_syrk_op_##X(int op, T alpha, sb200_matrix_t A, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
int sb200_syr2k_op_##X(int op, T alpha, sb200_matrix_t A, sb200_matrix_t B, T beta, sb200_matrix_t C, const sb200_options_t* opts); \
/* B = alpha op(A)^{-1} B (side 'L') or B = alpha B op(A)^{-1} (side 'R') at matrix level, A triangular: the lower tiles of a \
 * kind 'H' matrix (uplo 'L') or the lower / upper triangle of a general square matrix such as an LU factor; op 'N' | 'T' | 'C', \
 * diag 'N' | 'U'.  slate::trsm / triangular_solve (src/trsm.cc -> work::trsm, src/work/work_trsm.cc:24-387); 1 x 1 grid */ \
int sb200_trsm_mat_##X(int side, int uplo, int op, int diag, T alpha, sb200_matrix_t A, sb200_matrix_t B, const sb200_options_t* opts); \
/* slate::norm(norm, A) (src/norm.cc): norm 'M' max, 'O' | '1' one, 'I' inf, 'F' Frobenius; A general, or lower tiles read as a \
 * Hermitian (flavour 'H') or symmetric (flavour 'S') matrix; per-tile kernels + host combination as the reference's \
 * internal::norm<Devices> (src/internal/internal_genorm.cc:440-560).  1 x 1 grid */ \
int sb200_norm_##X(int norm, int flavour, sb200_matrix_t A, double* value); \
/* norm(Norm::Inf, A), A general or Hermitian   slate::norm (src/norm.cc); 1 x 1 grid */ \
int sb200_norm_inf_##X(sb200_matrix_t A, double* value);
SB200_FOR_TYPES(SB200_DECL_RUNTIME)

/* FP32 factorisations of the mixed-precision solvers; the _tc05 variants run the trailing-matrix
 * update on the tcgen05 FP32-emulated (3 x TF32) tensor-core kernel (gemm_tc05.cu). getrf_*_s: 1 x 1 grid. */
int sb200_potrf_tc05_s(sb200_matrix_t A, const sb200_options_t* opts, int64_t* info);
int sb200_getrf_s(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
int sb200_getrf_tc05_s(sb200_matrix_t A, int64_t* pivots, const sb200_options_t* opts, int64_t* info);
/* B <- A^{-1} B from the LU factors and pivots   slate::getrs (src/getrs.cc:25-66); 1 x 1 grid */
int sb200_getrs_d(sb200_matrix_t A, const int64_t* pivots, sb200_matrix_t B, const sb200_options_t* opts);
int sb200_getrs_s(sb200_matrix_t A, const int64_t* pivots, sb200_matrix_t B, const sb200_options_t* opts);
int sb200_getrs_z(sb200_matrix_t A, const int64_t* pivots, sb200_matrix_t B, const sb200_options_t* opts);
int sb200_getrs_c(sb200_matrix_t A, const int64_t* pivots, sb200_matrix_t B, const sb200_options_t* opts);
/* op(A) X = B with the factors of A: slate::getrs handed a (conjugate-)transposed view (src/getrs.cc:97-112); op 'N' | 'T' | 'C';
 * the handle carries the element type.  'N' as sb200_getrs_X; other ops: 1 x 1 grid */
int sb200_getrs_op(int op, sb200_matrix_t A, const int64_t* pivots, sb200_matrix_t B, const sb200_options_t* opts);

/* slate::posv_mixed / gesv_mixed <double, float> (src/posv_mixed.cc:111-297, src/gesv_mixed.cc:106-300):
 * factor a float copy of A (tcgen05 trailing update), solve, refine in FP64 until
 * max|r_j| <= max|x_j| * ||A||_inf * tolerance for every column j.
 * *iter: >= 0 refinement iterations; -3 low-precision factorisation failed; -(max_iterations+1) not converged
 * (then, with use_fallback_solver, A is factored and the system solved in FP64 -- A is overwritten).
 * timers_ms8 (optional): total, factor_lo, solve_lo, residual_hi, add_hi, factor_hi, solve_hi, norm+convert. */
typedef struct {
    int64_t max_iterations;       /* Option::MaxIterations, default 30; < 0 -> default      */
    double  tolerance;            /* Option::Tolerance, <= 0 -> eps * sqrt(n)               */
    int     use_fallback_solver;  /* Option::UseFallbackSolver, default 1                   */
    int     reserved[5];
} sb200_mixed_options_t;
int sb200_posv_mixed_d(sb200_matrix_t A, sb200_matrix_t B, sb200_matrix_t X, const sb200_mixed_options_t* mo,
                       int* iter, int64_t* info, double* timers_ms8);
int sb200_gesv_mixed_d(sb200_matrix_t A, int64_t* pivots, sb200_matrix_t B, sb200_matrix_t X,
                       const sb200_mixed_options_t* mo, int* iter, int64_t* info, double* timers_ms8);
/* <complex<double>, complex<float>> (the second pair of explicit instantiations, src/gesv_mixed.cc:303-316): complex<float>
 * factorisation, complex<double> refinement; same conventions.  1 x 1 grid. */
int sb200_posv_mixed_z(sb200_matrix_t A, sb200_matrix_t B, sb200_matrix_t X, const sb200_mixed_options_t* mo,
                       int* iter, int64_t* info, double* timers_ms8);
int sb200_gesv_mixed_z(sb200_matrix_t A, int64_t* pivots, sb200_matrix_t B, sb200_matrix_t X,
                       const sb200_mixed_options_t* mo, int* iter, int64_t* info, double* timers_ms8);

#ifdef __cplusplus
}
#endif
#endif /* SLATE_B200_H */
