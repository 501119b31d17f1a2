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

/* ---------------------------------------------------------------------------
 * Batched tile GEMM:  C_t = alpha * op(A_t) * op(B_t) + beta * C_t,  t < batch
 * replaces blas::batch::gemm fixed-size path -> cublas?gemmBatched
 * (blaspp/src/device_batch_gemm.cc:76-130; call sites src/internal/internal_gemm.cc:498-504,
 *  internal_herk.cc:510-516).  FP64 real/complex run on the FP64 tensor-core MMA (DMMA).
 * ------------------------------------------------------------------------- */
int sb200_gemm_batched_d(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                         double alpha, const double* const* dA, int64_t lda,
                         const double* const* dB, int64_t ldb,
                         double beta, double* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);
int sb200_gemm_batched_z(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                         sb200_c64 alpha, const sb200_c64* const* dA, int64_t lda,
                         const sb200_c64* const* dB, int64_t ldb,
                         sb200_c64 beta, sb200_c64* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);
int sb200_gemm_batched_s(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                         float alpha, const float* const* dA, int64_t lda,
                         const float* const* dB, int64_t ldb,
                         float beta, float* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);
int sb200_gemm_batched_c(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                         sb200_c32 alpha, const sb200_c32* const* dA, int64_t lda,
                         const sb200_c32* const* dB, int64_t ldb,
                         sb200_c32 beta, sb200_c32* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);

/* Same operation with per-array element offsets (A_t = dA[t] + offA, ...): lets the
 * host runtime address sub-blocks of resident tiles without rebuilding pointer arrays. */
int sb200_gemm_batched_off_d(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                             double alpha, const double* const* dA, int64_t offA, int64_t lda,
                             const double* const* dB, int64_t offB, int64_t ldb,
                             double beta, double* const* dC, int64_t offC, int64_t ldc,
                             int64_t batch, sb200_stream_t stream);

/* ---------------------------------------------------------------------------
 * Batched HERK / SYRK on the stored triangle:
 *   C_t = alpha * op(A_t) * op(A_t)^H + beta * C_t   (herk; alpha, beta real; imag(diag) := 0)
 *   C_t = alpha * op(A_t) * op(A_t)^T + beta * C_t   (syrk)
 * replaces blas::batch::herk/syrk = a LOOP of per-tile cublas?herk/?syrk on forked
 * streams (blaspp/src/device_batch_herk.cc:57-73) and the single-tile blas::herk
 * (src/internal/internal_herk.cc:380-385) with one launch.
 * op = 'N': A_t is n-by-k;  op = 'C'/'T': A_t is k-by-n.
 * ------------------------------------------------------------------------- */
int sb200_herk_batched_d(int layout, int uplo, int op, int64_t n, int64_t k,
                         double alpha, const double* const* dA, int64_t lda,
                         double beta, double* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);
int sb200_herk_batched_z(int layout, int uplo, int op, int64_t n, int64_t k,
                         double alpha, const sb200_c64* const* dA, int64_t lda,
                         double beta, sb200_c64* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);
int sb200_syrk_batched_d(int layout, int uplo, int op, int64_t n, int64_t k,
                         double alpha, const double* const* dA, int64_t lda,
                         double beta, double* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);
int sb200_herk_batched_s(int layout, int uplo, int op, int64_t n, int64_t k,
                         float alpha, const float* const* dA, int64_t lda,
                         float beta, float* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream);

/* ---------------------------------------------------------------------------
 * Batched TRSM with ONE triangular tile shared by the batch (SLATE replicates the
 * A pointer count times: src/internal/internal_trsm.cc:225-249):
 *   side 'L':  B_t <- alpha * op(A)^{-1} * B_t      side 'R':  B_t <- alpha * B_t * op(A)^{-1}
 * replaces blas::batch::trsm -> cublas?trsmBatched (blaspp/src/device_batch_trsm.cc:27-130).
 * `dA` is the device pointer of the single na-by-na triangular tile (na = m for 'L', n for 'R').
 * `work` is device scratch of at least sb200_trsm_work_bytes_X(...) bytes (the inverted
 * diagonal blocks); it may be NULL, in which case an internal per-device scratch is used.
 * ------------------------------------------------------------------------- */
size_t sb200_trsm_work_bytes_d(int side, int64_t m, int64_t n);
int sb200_trsm_batched_d(int layout, int side, int uplo, int op, int diag,
                         int64_t m, int64_t n, double alpha,
                         const double* dA, int64_t lda,
                         double* const* dB, int64_t ldb,
                         int64_t batch, void* work, sb200_stream_t stream);
int sb200_trsm_batched_s(int layout, int side, int uplo, int op, int diag,
                         int64_t m, int64_t n, float alpha,
                         const float* dA, int64_t lda,
                         float* const* dB, int64_t ldb,
                         int64_t batch, void* work, sb200_stream_t stream);

/* ---------------------------------------------------------------------------
 * Cholesky factorisation of one diagonal tile on the device, LAPACK info in *dinfo
 * (device int; 0 = ok, j > 0 = leading minor j not positive definite).
 * replaces lapack::potrf(uplo, n, dA, ldda, dinfo, queue) -> cusolverDn?potrf
 * (lapackpp/src/cuda/cuda_potrf.cc; call site src/internal/internal_potrf.cc:72-78).
 * ------------------------------------------------------------------------- */
int sb200_potrf_tile_d(int uplo, int64_t n, double* dA, int64_t lda,
                       int* dinfo, void* work, sb200_stream_t stream);
int sb200_potrf_tile_s(int uplo, int64_t n, float* dA, int64_t lda,
                       int* dinfo, void* work, sb200_stream_t stream);
size_t sb200_potrf_work_bytes_d(int64_t n);

/* ---------------------------------------------------------------------------
 * Fused row interchanges for getrf on RowMajor tiles: applies all `npiv` pivots of
 * one panel to a whole block row/column set in ONE launch.
 * replaces the per-pivot-row cublas?swap loop of internal::permuteRows<Devices>
 * (src/internal/internal_swap.cc:674-688: nb launches per block column per step).
 *
 * The block column is given as `mt` tiles stacked vertically; dTiles[t + j*mt] is tile t of
 * block column j (tile_rows[t] rows each, ncols columns, RowMajor with leading dimension ld,
 * or ColMajor when layout == 'C').  Pivot i swaps global row i of the stack with row
 * (piv_tile[i], piv_off[i]) -- the reference's Pivot{tileIndex, elementOffset}
 * (include/slate/types.hh:84-105) -- applied in order i = 0..npiv-1 (forward) or reversed.
 * ------------------------------------------------------------------------- */
int sb200_permute_rows_d(int layout, int forward, int64_t npiv,
                         const int64_t* d_piv_tile, const int64_t* d_piv_off,
                         double* const* dTiles, int64_t mt, int64_t ncolblocks,
                         int64_t tile_mb, int64_t ncols, int64_t ld,
                         sb200_stream_t stream);

/* ---------------------------------------------------------------------------
 * Seam 2: memory-bound tile kernels (slate::device::*, include/slate/internal/device.hh).
 * All batched over device pointer arrays; tile = m-by-n, column-major, leading dim ld.
 * ------------------------------------------------------------------------- */
/* B = alpha*A + beta*B                     device::batch::geadd  (src/cuda/device_geadd.cu:215-260) */
int sb200_geadd_batched_d(int64_t m, int64_t n, double alpha, const double* const* dA, int64_t lda,
                          double beta, double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_geadd_batched_s(int64_t m, int64_t n, float alpha, const float* const* dA, int64_t lda,
                          float beta, float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_geadd_batched_z(int64_t m, int64_t n, sb200_c64 alpha, const sb200_c64* const* dA, int64_t lda,
                          sb200_c64 beta, sb200_c64* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
/* A *= numer/denom                         device::batch::gescale (src/cuda/device_gescale.cu:200-240) */
int sb200_gescale_batched_d(int64_t m, int64_t n, double numer, double denom,
                            double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
int sb200_gescale_batched_s(int64_t m, int64_t n, float numer, float denom,
                            float* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
int sb200_gescale_batched_z(int64_t m, int64_t n, sb200_c64 numer, sb200_c64 denom,
                            sb200_c64* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
/* A_ij *= R_i * C_j (equilibration)        device::gescale_row_col_batch (device_gescale_row_col.cu:150-210)
 * equed: 'R' rows only, 'C' cols only, 'B' both.  dR/dC: device arrays of pointers to scale vectors. */
int sb200_gescale_row_col_batched_d(int equed, int64_t m, int64_t n,
                                    const double* const* dR, const double* const* dC,
                                    double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
/* offdiag -> A_ij (i != j), diag -> A_ii    device::batch::geset   (src/cuda/device_geset.cu:180-220) */
int sb200_geset_batched_d(int64_t m, int64_t n, double offdiag, double diag,
                          double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
int sb200_geset_batched_s(int64_t m, int64_t n, float offdiag, float diag,
                          float* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
int sb200_geset_batched_z(int64_t m, int64_t n, sb200_c64 offdiag, sb200_c64 diag,
                          sb200_c64* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
/* precision-converting copy B = A           device::gecopy          (src/cuda/device_gecopy.cu:75-115) */
int sb200_gecopy_batched_dd(int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_gecopy_batched_ds(int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_gecopy_batched_sd(int64_t m, int64_t n, const float* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_gecopy_batched_ss(int64_t m, int64_t n, const float* const* dA, int64_t lda,
                            float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_gecopy_batched_zz(int64_t m, int64_t n, const sb200_c64* const* dA, int64_t lda,
                            sb200_c64* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_gecopy_batched_zc(int64_t m, int64_t n, const sb200_c64* const* dA, int64_t lda,
                            sb200_c32* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_gecopy_batched_cz(int64_t m, int64_t n, const sb200_c32* const* dA, int64_t lda,
                            sb200_c64* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
/* trapezoid variants (uplo 'L'|'U')          device::tzset/tzadd/tzcopy/tzscale (src/cuda/device_tz*.cu) */
int sb200_tzset_batched_d(int uplo, int64_t m, int64_t n, double offdiag, double diag,
                          double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
int sb200_tzadd_batched_d(int uplo, int64_t m, int64_t n, double alpha, const double* const* dA, int64_t lda,
                          double beta, double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_tzscale_batched_d(int uplo, int64_t m, int64_t n, double numer, double denom,
                            double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream);
int sb200_tzcopy_batched_dd(int uplo, int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_tzcopy_batched_ds(int uplo, int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
int sb200_tzcopy_batched_sd(int uplo, int64_t m, int64_t n, const float* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream);
/* transposes                                 device::transpose / transpose_batch (src/cuda/device_transpose.cu)
 * in-place square (n-by-n) and out-of-place rectangular (A m-by-n -> AT n-by-m); conj != 0 conjugates (complex) */
int sb200_transpose_inplace_batched_d(int64_t n, double* const* dA, int64_t lda,
                                      int64_t batch, sb200_stream_t stream);
int sb200_transpose_batched_d(int64_t m, int64_t n, const double* const* dA, int64_t lda,
                              double* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream);
int sb200_transpose_inplace_batched_z(int conj, int64_t n, sb200_c64* const* dA, int64_t lda,
                                      int64_t batch, sb200_stream_t stream);
int sb200_transpose_batched_z(int conj, int64_t m, int64_t n, const sb200_c64* const* dA, int64_t lda,
                              sb200_c64* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream);
int sb200_transpose_inplace_batched_s(int64_t n, float* const* dA, int64_t lda,
                                      int64_t batch, sb200_stream_t stream);
int sb200_transpose_batched_s(int64_t m, int64_t n, const float* const* dA, int64_t lda,
                              float* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream);
/* per-tile norms                             device::genorm / henorm / synorm / trnorm (src/cuda/device_*norm.cu)
 * norm 'M' max, 'O' one, 'I' inf, 'F' frobenius; scope 'M' matrix (per-tile partial results),
 * 'C' columns (norm 'M' only: per-column max, used by colNorms).
 * values layout as the reference (device_genorm.cu:373-445): ldv >= 1 (max), n (one), m (inf), 2 (fro: scale, sumsq);
 * tile t writes values[t*ldv ...].  NaN-propagating max (device_util.cuh:22-25). */
int sb200_genorm_batched_d(int norm, int scope, int64_t m, int64_t n,
                           const double* const* dA, int64_t lda,
                           double* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
int sb200_genorm_batched_s(int norm, int scope, int64_t m, int64_t n,
                           const float* const* dA, int64_t lda,
                           float* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
int sb200_genorm_batched_z(int norm, int scope, int64_t m, int64_t n,
                           const sb200_c64* const* dA, int64_t lda,
                           double* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
int sb200_henorm_batched_d(int norm, int uplo, int64_t n,
                           const double* const* dA, int64_t lda,
                           double* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
int sb200_henorm_batched_z(int norm, int uplo, int64_t n,
                           const sb200_c64* const* dA, int64_t lda,
                           double* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
int sb200_synorm_batched_d(int norm, int uplo, int64_t n,
                           const double* const* dA, int64_t lda,
                           double* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
int sb200_synorm_offdiag_batched_d(int norm, int64_t m, int64_t n,
                                   const double* const* dA, int64_t lda,
                                   double* values, int64_t ldv, int64_t batch, sb200_stream_t stream);
int sb200_trnorm_batched_d(int norm, int uplo, int diag, int64_t m, int64_t n,
                           const double* const* dA, int64_t lda,
                           double* values, int64_t ldv, int64_t batch, sb200_stream_t stream);

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
int sb200_matrix_copy_d(sb200_matrix_t dst, sb200_matrix_t src, sb200_stream_t stream);
int64_t sb200_matrix_local_tiles(sb200_matrix_t A);

/* options common to the drivers */
typedef struct {
    int64_t lookahead;        /* slate::Option::Lookahead, default 1                     */
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
/* device time of the last driver call on this matrix, milliseconds (CUDA events) */
double sb200_last_driver_ms(sb200_matrix_t A);
/* out4 = { driver ms, summed ms of the trailing-update GEMM launches (events on their stream),
 *          algorithmic flops of those launches, number of those launches } */
int sb200_last_driver_stats(sb200_matrix_t A, double* out4);
/* FP64 pipe peak probe (roofline denominator; MEASURED_PEAKS.json has no FP64 figure):
 * kind 0 = DMMA.8x8x4, 1 = DFMA; launches ctas_per_sm x SMs CTAs; *flops = work of the launch. */
int sb200_fp64_peak_probe(int kind, int iters, int ctas_per_sm, double* d_scratch, double* flops,
                          sb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SLATE_B200_H */
