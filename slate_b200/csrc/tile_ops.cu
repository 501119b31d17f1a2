// tile_ops.cu -- memory-bound batched tile kernels (seam 2: namespace slate::device of
// include/slate/internal/device.hh:92-281; reference kernels src/cuda/device_{geadd,gecopy,
// gescale,gescale_row_col,geset,tzadd,tzcopy,tzscale,tzset,transpose}.cu).
//
// The reference launches ONE CTA of min(1024, m) threads per tile, one thread per row walking the
// columns (written "for compute capability <= 7.5", device_geadd.cu:134): a 2 MiB tile is served
// by a single SM.  Here the grid is (column chunks) x (tiles): every warp streams whole columns
// (one 256-byte line per warp instruction, 4 independent accesses in flight per lane), so even
// a single tile is spread over 64+ CTAs and a batch saturates HBM.  All kernels are pure streaming:
// algorithmic bytes = traffic.
#include "common.cuh"
#include "scalar_ops.cuh"

namespace sb200 {

// precision / domain conversions of device::gecopy (device_gecopy.cu:22-39 copy_a2b)
template <typename D, typename S> __device__ inline D convert(S a);
template <> __device__ inline float  convert<float, float>(float a) { return a; }
template <> __device__ inline double convert<double, double>(double a) { return a; }
template <> __device__ inline float  convert<float, double>(double a) { return float(a); }
template <> __device__ inline double convert<double, float>(float a) { return double(a); }
template <> __device__ inline cuDoubleComplex convert<cuDoubleComplex, cuDoubleComplex>(cuDoubleComplex a) { return a; }
template <> __device__ inline cuFloatComplex  convert<cuFloatComplex, cuFloatComplex>(cuFloatComplex a) { return a; }
template <> __device__ inline cuFloatComplex  convert<cuFloatComplex, cuDoubleComplex>(cuDoubleComplex a) { return make_cuFloatComplex(float(a.x), float(a.y)); }
template <> __device__ inline cuDoubleComplex convert<cuDoubleComplex, cuFloatComplex>(cuFloatComplex a) { return make_cuDoubleComplex(a.x, a.y); }
template <> __device__ inline cuDoubleComplex convert<cuDoubleComplex, double>(double a) { return make_cuDoubleComplex(a, 0.0); }
template <> __device__ inline cuFloatComplex  convert<cuFloatComplex, float>(float a) { return make_cuFloatComplex(a, 0.f); }

// ----------------------------------------------------------------------------- streaming skeleton
// grid.x = column chunks, grid.y = tiles (strided), 8 warps per CTA, one warp per column at a
// time, lanes stride the rows with a 4-deep unroll so >= 4 independent loads are in flight.
// mask: 0 = whole tile, 1 = lower trapezoid (i >= j), 2 = upper trapezoid (i <= j).
constexpr int TILE_WARPS = 8;

template <typename F>
__global__ void __launch_bounds__(TILE_WARPS * 32)
tile_foreach_kernel(int m, int n, int batch, int mask, F f)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = blockIdx.y; t < batch; t += gridDim.y) {
        for (int j = blockIdx.x * TILE_WARPS + warp; j < n; j += gridDim.x * TILE_WARPS) {
            int i0 = 0, i1 = m;
            if (mask == 1) i0 = j;
            if (mask == 2) i1 = min(m, j + 1);
            // start on a 32-row boundary so that accesses stay aligned to 256-byte lines
            int i = (i0 & ~31) + lane;
            #pragma unroll 4
            for (; i < i1; i += 32)
                if (i >= i0) f(t, i, j);
        }
    }
}

template <typename F>
static int launch_foreach(int64_t m, int64_t n, int64_t batch, int mask, F f, cudaStream_t s)
{
    if (m < 0 || n < 0 || batch < 0) return SB200_EINVAL;
    if (m == 0 || n == 0 || batch == 0) return SB200_OK;
    if (m > 0x7fffffff || n > 0x7fffffff || batch > 0x7fffffff) return SB200_EINVAL;
    const unsigned gx = unsigned(std::min<int64_t>(ceil_div(n, TILE_WARPS), 1024));
    const unsigned gy = unsigned(std::min<int64_t>(batch, 65535));
    tile_foreach_kernel<<<dim3(gx, gy), TILE_WARPS * 32, 0, s>>>(int(m), int(n), int(batch), mask, f);
    return launch_status();
}

// ----------------------------------------------------------------------------- functors
template <typename T> struct AddOp {          // B = alpha A + beta B
    const T* const* A; T* const* B; int64_t lda, ldb; T alpha, beta;
    __device__ void operator()(int t, int i, int j) const {
        T* b = B[t] + i + j * ldb;
        *b = add(mul(alpha, A[t][i + j * lda]), mul(beta, *b));
    }
};
template <typename T> struct ScaleOp {        // A *= numer / denom
    T* const* A; int64_t lda; T mult;
    __device__ void operator()(int t, int i, int j) const { T* a = A[t] + i + j * lda; *a = mul(*a, mult); }
};
template <typename T> struct ScaleRowColOp {  // A_ij *= R_i C_j
    T* const* A; int64_t lda; const T* const* R; const T* const* C; int use_r, use_c;
    __device__ void operator()(int t, int i, int j) const {
        T* a = A[t] + i + j * lda;
        T v = *a;
        if (use_r) v = mul(v, R[t][i]);
        if (use_c) v = mul(v, C[t][j]);
        *a = v;
    }
};
template <typename T> struct SetOp {          // offdiag / diag fill
    T* const* A; int64_t lda; T offdiag, diag;
    __device__ void operator()(int t, int i, int j) const { A[t][i + j * lda] = (i == j) ? diag : offdiag; }
};
template <typename S, typename D> struct CopyOp {   // B = convert(A)
    const S* const* A; D* const* B; int64_t lda, ldb;
    __device__ void operator()(int t, int i, int j) const { B[t][i + j * ldb] = convert<D, S>(A[t][i + j * lda]); }
};

// ----------------------------------------------------------------------------- transposes
// Out-of-place: AT (n x m) = A^T (A m x n).  32 x 32 shared tiles (+1 pad), both sides coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
transpose_oop_kernel(int m, int n, const T* const* A, int64_t lda, T* const* AT, int64_t ldat, int conj, int batch)
{
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int bm = (m + 31) / 32, bn = (n + 31) / 32;
    for (int t = blockIdx.y; t < batch; t += gridDim.y) {
        const T* a = A[t];
        T* at = AT[t];
        for (int b = blockIdx.x; b < bm * bn; b += gridDim.x) {
            const int i0 = (b % bm) * 32, j0 = (b / bm) * 32;
            #pragma unroll
            for (int r = 0; r < 32; r += 8) {
                const int i = i0 + tx, j = j0 + ty + r;
                if (i < m && j < n) tile[ty + r][tx] = a[i + j * lda];
            }
            __syncthreads();
            #pragma unroll
            for (int r = 0; r < 32; r += 8) {
                const int j = j0 + tx, i = i0 + ty + r;          // AT(j, i)
                if (i < m && j < n) { T v = tile[tx][ty + r]; at[j + i * ldat] = conj ? conj_(v) : v; }
            }
            __syncthreads();
        }
    }
}

// In-place square: swap 32x32 block pairs (bi > bj) through shared memory; diagonal blocks alone.
template <typename T>
__global__ void __launch_bounds__(256)
transpose_inplace_kernel(int n, T* const* A, int64_t lda, int conj, int batch)
{
    __shared__ T t1[32][33], t2[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int nbk = (n + 31) / 32;
    const int npairs = nbk * (nbk + 1) / 2;
    for (int t = blockIdx.y; t < batch; t += gridDim.y) {
        T* a = A[t];
        for (int pidx = blockIdx.x; pidx < npairs; pidx += gridDim.x) {
            // pidx -> (bi >= bj) in the lower block triangle
            int bi = int((sqrtf(8.f * pidx + 1.f) - 1.f) * 0.5f);
            while (bi * (bi + 1) / 2 > pidx) --bi;
            while ((bi + 1) * (bi + 2) / 2 <= pidx) ++bi;
            const int bj = pidx - bi * (bi + 1) / 2;
            const int i0 = bi * 32, j0 = bj * 32;
            #pragma unroll
            for (int r = 0; r < 32; r += 8) {
                int i = i0 + tx, j = j0 + ty + r;
                if (i < n && j < n) t1[ty + r][tx] = a[i + j * lda];          // block (bi, bj)
                i = j0 + tx; j = i0 + ty + r;
                if (bi != bj && i < n && j < n) t2[ty + r][tx] = a[i + j * lda];   // block (bj, bi)
            }
            __syncthreads();
            #pragma unroll
            for (int r = 0; r < 32; r += 8) {
                // (bj, bi) <- (bi, bj)^T
                int i = j0 + tx, j = i0 + ty + r;
                if (i < n && j < n) { T v = t1[tx][ty + r]; a[i + j * lda] = conj ? conj_(v) : v; }
                i = i0 + tx; j = j0 + ty + r;
                if (bi != bj && i < n && j < n) { T v = t2[tx][ty + r]; a[i + j * lda] = conj ? conj_(v) : v; }
            }
            __syncthreads();
        }
    }
}

template <typename T>
static int launch_transpose_oop(int conj, int64_t m, int64_t n, const T* const* A, int64_t lda,
                                T* const* AT, int64_t ldat, int64_t batch, cudaStream_t s)
{
    if (m < 0 || n < 0 || batch < 0 || lda < m || ldat < n) return SB200_EINVAL;
    if (m == 0 || n == 0 || batch == 0) return SB200_OK;
    const int64_t blocks = ceil_div(m, 32) * ceil_div(n, 32);
    transpose_oop_kernel<T><<<dim3(unsigned(std::min<int64_t>(blocks, 4096)), unsigned(std::min<int64_t>(batch, 65535))), 256, 0, s>>>(
        int(m), int(n), A, lda, AT, ldat, conj, int(batch));
    return launch_status();
}

template <typename T>
static int launch_transpose_inplace(int conj, int64_t n, T* const* A, int64_t lda, int64_t batch, cudaStream_t s)
{
    if (n < 0 || batch < 0 || lda < n) return SB200_EINVAL;
    if (n == 0 || batch == 0) return SB200_OK;
    const int64_t nbk = ceil_div(n, 32), pairs = nbk * (nbk + 1) / 2;
    transpose_inplace_kernel<T><<<dim3(unsigned(std::min<int64_t>(pairs, 4096)), unsigned(std::min<int64_t>(batch, 65535))), 256, 0, s>>>(
        int(n), A, lda, conj, int(batch));
    return launch_status();
}

// ----------------------------------------------------------------------------- typed entry helpers
template <typename T>
static int geadd_t(int mask, int64_t m, int64_t n, T alpha, const T* const* A, int64_t lda, T beta,
                   T* const* B, int64_t ldb, int64_t batch, cudaStream_t s)
{
    if (lda < m || ldb < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, AddOp<T>{A, B, lda, ldb, alpha, beta}, s);
}
template <typename T>
static int gescale_t(int mask, int64_t m, int64_t n, T numer, T denom, T* const* A, int64_t lda, int64_t batch, cudaStream_t s)
{
    if (lda < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, ScaleOp<T>{A, lda, divide(numer, denom)}, s);
}
template <typename T>
static int geset_t(int mask, int64_t m, int64_t n, T offdiag, T diag, T* const* A, int64_t lda, int64_t batch, cudaStream_t s)
{
    if (lda < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, SetOp<T>{A, lda, offdiag, diag}, s);
}
template <typename S, typename D>
static int gecopy_t(int mask, int64_t m, int64_t n, const S* const* A, int64_t lda, D* const* B, int64_t ldb, int64_t batch, cudaStream_t s)
{
    if (lda < m || ldb < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, CopyOp<S, D>{A, B, lda, ldb}, s);
}

static inline int mask_of(int uplo) { return uplo == 'L' ? 1 : (uplo == 'U' ? 2 : -1); }
static inline cuDoubleComplex z(sb200_c64 v) { return make_cuDoubleComplex(v.re, v.im); }

} // namespace sb200

using namespace sb200;
#define ST cudaStream_t(stream)
typedef const cuDoubleComplex* const* zcpp;
typedef cuDoubleComplex* const* zpp;
typedef const cuFloatComplex* const* ccpp;
typedef cuFloatComplex* const* cpp_;

extern "C" {

// ---- geadd
int sb200_geadd_batched_d(int64_t m, int64_t n, double alpha, const double* const* dA, int64_t lda,
                          double beta, double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return geadd_t<double>(0, m, n, alpha, dA, lda, beta, dB, ldb, batch, ST); }
int sb200_geadd_batched_s(int64_t m, int64_t n, float alpha, const float* const* dA, int64_t lda,
                          float beta, float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return geadd_t<float>(0, m, n, alpha, dA, lda, beta, dB, ldb, batch, ST); }
int sb200_geadd_batched_z(int64_t m, int64_t n, sb200_c64 alpha, const sb200_c64* const* dA, int64_t lda,
                          sb200_c64 beta, sb200_c64* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return geadd_t<cuDoubleComplex>(0, m, n, z(alpha), zcpp(dA), lda, z(beta), zpp(dB), ldb, batch, ST); }
int sb200_tzadd_batched_d(int uplo, int64_t m, int64_t n, double alpha, const double* const* dA, int64_t lda,
                          double beta, double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ if (mask_of(uplo) < 0) return SB200_EINVAL; return geadd_t<double>(mask_of(uplo), m, n, alpha, dA, lda, beta, dB, ldb, batch, ST); }

// ---- gescale
int sb200_gescale_batched_d(int64_t m, int64_t n, double numer, double denom,
                            double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return gescale_t<double>(0, m, n, numer, denom, dA, lda, batch, ST); }
int sb200_gescale_batched_s(int64_t m, int64_t n, float numer, float denom,
                            float* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return gescale_t<float>(0, m, n, numer, denom, dA, lda, batch, ST); }
int sb200_gescale_batched_z(int64_t m, int64_t n, sb200_c64 numer, sb200_c64 denom,
                            sb200_c64* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return gescale_t<cuDoubleComplex>(0, m, n, z(numer), z(denom), zpp(dA), lda, batch, ST); }
int sb200_tzscale_batched_d(int uplo, int64_t m, int64_t n, double numer, double denom,
                            double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ if (mask_of(uplo) < 0) return SB200_EINVAL; return gescale_t<double>(mask_of(uplo), m, n, numer, denom, dA, lda, batch, ST); }

int sb200_gescale_row_col_batched_d(int equed, int64_t m, int64_t n,
                                    const double* const* dR, const double* const* dC,
                                    double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{
    if (equed != 'R' && equed != 'C' && equed != 'B') return SB200_EINVAL;
    if (lda < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, 0,
                          ScaleRowColOp<double>{dA, lda, dR, dC, equed != 'C', equed != 'R'}, ST);
}

// ---- geset / tzset
int sb200_geset_batched_d(int64_t m, int64_t n, double offdiag, double diag,
                          double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return geset_t<double>(0, m, n, offdiag, diag, dA, lda, batch, ST); }
int sb200_geset_batched_s(int64_t m, int64_t n, float offdiag, float diag,
                          float* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return geset_t<float>(0, m, n, offdiag, diag, dA, lda, batch, ST); }
int sb200_geset_batched_z(int64_t m, int64_t n, sb200_c64 offdiag, sb200_c64 diag,
                          sb200_c64* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return geset_t<cuDoubleComplex>(0, m, n, z(offdiag), z(diag), zpp(dA), lda, batch, ST); }
int sb200_tzset_batched_d(int uplo, int64_t m, int64_t n, double offdiag, double diag,
                          double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ if (mask_of(uplo) < 0) return SB200_EINVAL; return geset_t<double>(mask_of(uplo), m, n, offdiag, diag, dA, lda, batch, ST); }

// ---- gecopy / tzcopy
int sb200_gecopy_batched_dd(int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return gecopy_t<double, double>(0, m, n, dA, lda, dB, ldb, batch, ST); }
int sb200_gecopy_batched_ds(int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return gecopy_t<double, float>(0, m, n, dA, lda, dB, ldb, batch, ST); }
int sb200_gecopy_batched_sd(int64_t m, int64_t n, const float* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return gecopy_t<float, double>(0, m, n, dA, lda, dB, ldb, batch, ST); }
int sb200_gecopy_batched_ss(int64_t m, int64_t n, const float* const* dA, int64_t lda,
                            float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return gecopy_t<float, float>(0, m, n, dA, lda, dB, ldb, batch, ST); }
int sb200_gecopy_batched_zz(int64_t m, int64_t n, const sb200_c64* const* dA, int64_t lda,
                            sb200_c64* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return gecopy_t<cuDoubleComplex, cuDoubleComplex>(0, m, n, zcpp(dA), lda, zpp(dB), ldb, batch, ST); }
int sb200_gecopy_batched_zc(int64_t m, int64_t n, const sb200_c64* const* dA, int64_t lda,
                            sb200_c32* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return gecopy_t<cuDoubleComplex, cuFloatComplex>(0, m, n, zcpp(dA), lda, cpp_(dB), ldb, batch, ST); }
int sb200_gecopy_batched_cz(int64_t m, int64_t n, const sb200_c32* const* dA, int64_t lda,
                            sb200_c64* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ return gecopy_t<cuFloatComplex, cuDoubleComplex>(0, m, n, ccpp(dA), lda, zpp(dB), ldb, batch, ST); }
int sb200_tzcopy_batched_dd(int uplo, int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ if (mask_of(uplo) < 0) return SB200_EINVAL; return gecopy_t<double, double>(mask_of(uplo), m, n, dA, lda, dB, ldb, batch, ST); }
int sb200_tzcopy_batched_ds(int uplo, int64_t m, int64_t n, const double* const* dA, int64_t lda,
                            float* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ if (mask_of(uplo) < 0) return SB200_EINVAL; return gecopy_t<double, float>(mask_of(uplo), m, n, dA, lda, dB, ldb, batch, ST); }
int sb200_tzcopy_batched_sd(int uplo, int64_t m, int64_t n, const float* const* dA, int64_t lda,
                            double* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream)
{ if (mask_of(uplo) < 0) return SB200_EINVAL; return gecopy_t<float, double>(mask_of(uplo), m, n, dA, lda, dB, ldb, batch, ST); }

// ---- transposes
int sb200_transpose_inplace_batched_d(int64_t n, double* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return launch_transpose_inplace<double>(0, n, dA, lda, batch, ST); }
int sb200_transpose_batched_d(int64_t m, int64_t n, const double* const* dA, int64_t lda,
                              double* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream)
{ return launch_transpose_oop<double>(0, m, n, dA, lda, dAT, ldat, batch, ST); }
int sb200_transpose_inplace_batched_z(int conj, int64_t n, sb200_c64* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return launch_transpose_inplace<cuDoubleComplex>(conj, n, zpp(dA), lda, batch, ST); }
int sb200_transpose_batched_z(int conj, int64_t m, int64_t n, const sb200_c64* const* dA, int64_t lda,
                              sb200_c64* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream)
{ return launch_transpose_oop<cuDoubleComplex>(conj, m, n, zcpp(dA), lda, zpp(dAT), ldat, batch, ST); }
int sb200_transpose_inplace_batched_s(int64_t n, float* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream)
{ return launch_transpose_inplace<float>(0, n, dA, lda, batch, ST); }
int sb200_transpose_batched_s(int64_t m, int64_t n, const float* const* dA, int64_t lda,
                              float* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream)
{ return launch_transpose_oop<float>(0, m, n, dA, lda, dAT, ldat, batch, ST); }

} // extern "C"
