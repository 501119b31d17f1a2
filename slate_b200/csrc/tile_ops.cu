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
// time, lanes stride the rows 8 deep: all 8 loads of a lane are issued before its first store.
// mask: 0 = whole tile, 1 = lower trapezoid (i >= j), 2 = upper trapezoid (i <= j).
constexpr int TILE_WARPS = 8;

template <typename F>
__global__ void __launch_bounds__(TILE_WARPS * 32)
tile_foreach_kernel(int m, int n, int batch, int mask, F f)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int UN = 8;
    for (int t = blockIdx.y; t < batch; t += gridDim.y) {
        const auto ft = f.bind(t);                  // tile pointers resolved once per tile, not once per element
        for (int j = blockIdx.x * TILE_WARPS + warp; j < n; j += gridDim.x * TILE_WARPS) {
            int i0 = 0, i1 = m;
            if (mask == 1) i0 = j;
            if (mask == 2) i1 = min(m, j + 1);
            // start on a 32-row boundary so that accesses stay aligned to 256-byte lines.  All UN loads of a lane are
            // issued before the first store: a store followed by a load through the same pointer type cannot be
            // reordered by the compiler, and with one load in flight per lane the read-modify-write kernels
            // (gescale, gecopy) sat at 0.74-0.81 of the HBM peak while geadd, with two, reached it.
            for (int ib = (i0 & ~31) + lane; ib < i1; ib += 32 * UN) {
                typename F::Val v[UN];
                #pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = ib + 32 * u;
                    if (i >= i0 && i < i1) v[u] = ft.load(i, j);
                }
                #pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = ib + 32 * u;
                    if (i >= i0 && i < i1) ft.store(i, j, v[u]);
                }
            }
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
// A tile operand is either a device pointer array (batched entry points) or one tile pointer
// (the reference's single-tile device::geadd / gescale / geset / tzset / transpose).
template <typename T> struct Ptrs {
    T* const* arr; T* one;
    __device__ __forceinline__ T* operator[](int t) const { return arr ? arr[t] : one; }
};
template <typename T> static Ptrs<T> ptrs(T* const* arr) { return Ptrs<T>{arr, nullptr}; }
template <typename T> static Ptrs<T> one(T* p) { return Ptrs<T>{nullptr, p}; }

// Every functor binds to a tile (`bind(t)`: pointers resolved once) and splits an element update into the value(s)
// it reads (`load`) and the write (`store`).
template <typename T> struct AddOp {          // B = alpha A + beta B
    Ptrs<const T> A; Ptrs<T> B; int64_t lda, ldb; T alpha, beta;
    struct Val { T a, b; };
    struct Bound {
        const T* a; T* b; int64_t lda, ldb; T alpha, beta;
        __device__ __forceinline__ Val load(int i, int j) const { return Val{a[i + j * lda], b[i + j * ldb]}; }
        __device__ __forceinline__ void store(int i, int j, const Val& v) const { b[i + j * ldb] = add(mul(alpha, v.a), mul(beta, v.b)); }
    };
    __device__ __forceinline__ Bound bind(int t) const { return Bound{A[t], B[t], lda, ldb, alpha, beta}; }
};
template <typename T> struct ScaleOp {        // A *= numer / denom
    Ptrs<T> A; int64_t lda; T mult;
    using Val = T;
    struct Bound {
        T* a; int64_t lda; T mult;
        __device__ __forceinline__ T load(int i, int j) const { return a[i + j * lda]; }
        __device__ __forceinline__ void store(int i, int j, const T& v) const { a[i + j * lda] = mul(v, mult); }
    };
    __device__ __forceinline__ Bound bind(int t) const { return Bound{A[t], lda, mult}; }
};
template <typename T, typename S> struct ScaleRowColOp {  // A_ij *= R_i C_j  (S = T or real(T))
    Ptrs<T> A; int64_t lda; const S* const* R; const S* const* C; int use_r, use_c;
    struct Val { T a; S r; };
    struct Bound {
        T* a; int64_t lda; const S* r; const S* c; int use_r, use_c;
        __device__ __forceinline__ Val load(int i, int j) const { return Val{a[i + j * lda], use_r ? r[i] : S()}; }
        __device__ __forceinline__ void store(int i, int j, const Val& v) const {
            T x = v.a;
            if (use_r) x = mul(x, convert<T, S>(v.r));
            if (use_c) x = mul(x, convert<T, S>(c[j]));
            a[i + j * lda] = x;
        }
    };
    __device__ __forceinline__ Bound bind(int t) const { return Bound{A[t], lda, use_r ? R[t] : nullptr, use_c ? C[t] : nullptr, use_r, use_c}; }
};
template <typename T> struct SetOp {          // offdiag / diag fill
    Ptrs<T> A; int64_t lda; T offdiag, diag;
    struct Val {};
    struct Bound {
        T* a; int64_t lda; T offdiag, diag;
        __device__ __forceinline__ Val load(int, int) const { return Val{}; }
        __device__ __forceinline__ void store(int i, int j, const Val&) const { a[i + j * lda] = (i == j) ? diag : offdiag; }
    };
    __device__ __forceinline__ Bound bind(int t) const { return Bound{A[t], lda, offdiag, diag}; }
};
template <typename S, typename D> struct CopyOp {   // B = convert(A)
    Ptrs<const S> A; Ptrs<D> B; int64_t lda, ldb;
    using Val = S;
    struct Bound {
        const S* a; D* b; int64_t lda, ldb;
        __device__ __forceinline__ S load(int i, int j) const { return a[i + j * lda]; }
        __device__ __forceinline__ void store(int i, int j, const S& v) const { b[i + j * ldb] = convert<D, S>(v); }
    };
    __device__ __forceinline__ Bound bind(int t) const { return Bound{A[t], B[t], lda, ldb}; }
};

// ----------------------------------------------------------------------------- transposes
// Out-of-place: AT (n x m) = A^T (A m x n).  32 x 32 shared tiles (+1 pad), both sides coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
transpose_oop_kernel(int m, int n, Ptrs<const T> A, int64_t lda, Ptrs<T> AT, int64_t ldat, int conj, int batch)
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
transpose_inplace_kernel(int n, Ptrs<T> A, int64_t lda, int conj, int batch)
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
static int launch_transpose_oop(int conj, int64_t m, int64_t n, Ptrs<const T> A, int64_t lda,
                                Ptrs<T> AT, int64_t ldat, int64_t batch, cudaStream_t s)
{
    if (m < 0 || n < 0 || batch < 0 || lda < m || ldat < n) return SB200_EINVAL;
    if (m == 0 || n == 0 || batch == 0) return SB200_OK;
    const int64_t blocks = ceil_div(m, 32) * ceil_div(n, 32);
    transpose_oop_kernel<T><<<dim3(unsigned(std::min<int64_t>(blocks, 4096)), unsigned(std::min<int64_t>(batch, 65535))), 256, 0, s>>>(
        int(m), int(n), A, lda, AT, ldat, conj, int(batch));
    return launch_status();
}

template <typename T>
static int launch_transpose_inplace(int conj, int64_t n, Ptrs<T> A, int64_t lda, int64_t batch, cudaStream_t s)
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
static int geadd_t(int mask, int64_t m, int64_t n, T alpha, Ptrs<const T> A, int64_t lda, T beta,
                   Ptrs<T> B, int64_t ldb, int64_t batch, cudaStream_t s)
{
    if (mask < 0 || lda < m || ldb < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, AddOp<T>{A, B, lda, ldb, alpha, beta}, s);
}
template <typename T>
static int gescale_t(int mask, int64_t m, int64_t n, T numer, T denom, Ptrs<T> A, int64_t lda, int64_t batch, cudaStream_t s)
{
    if (mask < 0 || lda < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, ScaleOp<T>{A, lda, divide(numer, denom)}, s);
}
template <typename T>
static int geset_t(int mask, int64_t m, int64_t n, T offdiag, T diag, Ptrs<T> A, int64_t lda, int64_t batch, cudaStream_t s)
{
    if (mask < 0 || lda < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, SetOp<T>{A, lda, offdiag, diag}, s);
}
template <typename S, typename D>
static int gecopy_t(int mask, int64_t m, int64_t n, Ptrs<const S> A, int64_t lda, Ptrs<D> B, int64_t ldb, int64_t batch, cudaStream_t s)
{
    if (mask < 0 || lda < m || ldb < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, mask, CopyOp<S, D>{A, B, lda, ldb}, s);
}
template <typename T, typename S>
static int gescale_row_col_t(int equed, int64_t m, int64_t n, const S* const* R, const S* const* C,
                             Ptrs<T> A, int64_t lda, int64_t batch, cudaStream_t s)
{
    if (equed != 'R' && equed != 'C' && equed != 'B') return SB200_EINVAL;
    if (lda < m) return SB200_EINVAL;
    return launch_foreach(m, n, batch, 0, ScaleRowColOp<T, S>{A, lda, R, C, equed != 'C', equed != 'R'}, s);
}

// uplo code -> mask: 'G' whole tile, 'L' lower trapezoid, 'U' upper trapezoid
static inline int mask_of(int uplo) { return uplo == 'G' ? 0 : (uplo == 'L' ? 1 : (uplo == 'U' ? 2 : -1)); }
static inline int tz_mask_of(int uplo) { return uplo == 'L' ? 1 : (uplo == 'U' ? 2 : -1); }

// ABI scalar/pointer types -> CUDA types (layout-compatible)
template <typename A> struct Cu { using type = A; };
template <> struct Cu<sb200_c32> { using type = cuFloatComplex; };
template <> struct Cu<sb200_c64> { using type = cuDoubleComplex; };
static inline float  cv(float v) { return v; }
static inline double cv(double v) { return v; }
static inline cuFloatComplex  cv(sb200_c32 v) { return make_cuFloatComplex(v.re, v.im); }
static inline cuDoubleComplex cv(sb200_c64 v) { return make_cuDoubleComplex(v.re, v.im); }

} // namespace sb200

using namespace sb200;
#define ST cudaStream_t(stream)
#define CU(T) typename Cu<T>::type
#define PA(T, p)  ptrs(reinterpret_cast<Cu<T>::type* const*>(p))
#define PCA(T, p) ptrs(reinterpret_cast<const Cu<T>::type* const*>(p))
#define P1(T, p)  one(reinterpret_cast<Cu<T>::type*>(p))
#define PC1(T, p) one(reinterpret_cast<const Cu<T>::type*>(p))

extern "C" {

#define SB200_DEF_TILE_OPS(X, T, R) \
int sb200_geadd_##X(int64_t m, int64_t n, T alpha, const T* dA, int64_t lda, T beta, T* dB, int64_t ldb, sb200_stream_t stream) \
{ return geadd_t<Cu<T>::type>(0, m, n, cv(alpha), PC1(T, dA), lda, cv(beta), P1(T, dB), ldb, 1, ST); } \
int sb200_geadd_batched_##X(int64_t m, int64_t n, T alpha, const T* const* dA, int64_t lda, \
                            T beta, T* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream) \
{ return geadd_t<Cu<T>::type>(0, m, n, cv(alpha), PCA(T, dA), lda, cv(beta), PA(T, dB), ldb, batch, ST); } \
int sb200_tzadd_batched_##X(int uplo, int64_t m, int64_t n, T alpha, const T* const* dA, int64_t lda, \
                            T beta, T* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream) \
{ return geadd_t<Cu<T>::type>(tz_mask_of(uplo), m, n, cv(alpha), PCA(T, dA), lda, cv(beta), PA(T, dB), ldb, batch, ST); } \
int sb200_gescale_##X(int64_t m, int64_t n, T numer, T denom, T* dA, int64_t lda, sb200_stream_t stream) \
{ return gescale_t<Cu<T>::type>(0, m, n, cv(numer), cv(denom), P1(T, dA), lda, 1, ST); } \
int sb200_gescale_batched_##X(int64_t m, int64_t n, T numer, T denom, \
                              T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream) \
{ return gescale_t<Cu<T>::type>(0, m, n, cv(numer), cv(denom), PA(T, dA), lda, batch, ST); } \
int sb200_tzscale_batched_##X(int uplo, int64_t m, int64_t n, R numer, R denom, \
                              T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream) \
{ return gescale_t<Cu<T>::type>(tz_mask_of(uplo), m, n, from_real<Cu<T>::type>(numer), from_real<Cu<T>::type>(denom), PA(T, dA), lda, batch, ST); } \
int sb200_gescale_row_col_batched_##X(int equed, int64_t m, int64_t n, const T* const* dR, const T* const* dC, \
                                      T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream) \
{ return gescale_row_col_t<Cu<T>::type, Cu<T>::type>(equed, m, n, reinterpret_cast<const Cu<T>::type* const*>(dR), \
      reinterpret_cast<const Cu<T>::type* const*>(dC), PA(T, dA), lda, batch, ST); } \
int sb200_gescale_row_col_real_batched_##X(int equed, int64_t m, int64_t n, const R* const* dR, const R* const* dC, \
                                      T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream) \
{ return gescale_row_col_t<Cu<T>::type, R>(equed, m, n, dR, dC, PA(T, dA), lda, batch, ST); } \
int sb200_geset_##X(int uplo, int64_t m, int64_t n, T offdiag, T diag, T* dA, int64_t lda, sb200_stream_t stream) \
{ return geset_t<Cu<T>::type>(mask_of(uplo), m, n, cv(offdiag), cv(diag), P1(T, dA), lda, 1, ST); } \
int sb200_geset_batched_##X(int64_t m, int64_t n, T offdiag, T diag, \
                            T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream) \
{ return geset_t<Cu<T>::type>(0, m, n, cv(offdiag), cv(diag), PA(T, dA), lda, batch, ST); } \
int sb200_tzset_batched_##X(int uplo, int64_t m, int64_t n, T offdiag, T diag, \
                            T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream) \
{ return geset_t<Cu<T>::type>(tz_mask_of(uplo), m, n, cv(offdiag), cv(diag), PA(T, dA), lda, batch, ST); } \
int sb200_transpose_inplace_##X(int conj, int64_t n, T* dA, int64_t lda, sb200_stream_t stream) \
{ return launch_transpose_inplace<Cu<T>::type>(conj, n, P1(T, dA), lda, 1, ST); } \
int sb200_transpose_##X(int conj, int64_t m, int64_t n, const T* dA, int64_t lda, T* dAT, int64_t ldat, sb200_stream_t stream) \
{ return launch_transpose_oop<Cu<T>::type>(conj, m, n, PC1(T, dA), lda, P1(T, dAT), ldat, 1, ST); } \
int sb200_transpose_inplace_batched_##X(int conj, int64_t n, T* const* dA, int64_t lda, int64_t batch, sb200_stream_t stream) \
{ return launch_transpose_inplace<Cu<T>::type>(conj, n, PA(T, dA), lda, batch, ST); } \
int sb200_transpose_batched_##X(int conj, int64_t m, int64_t n, const T* const* dA, int64_t lda, \
                                T* const* dAT, int64_t ldat, int64_t batch, sb200_stream_t stream) \
{ return launch_transpose_oop<Cu<T>::type>(conj, m, n, PCA(T, dA), lda, PA(T, dAT), ldat, batch, ST); }
SB200_FOR_TYPES(SB200_DEF_TILE_OPS)

#define SB200_DEF_COPY(XY, S, D) \
int sb200_gecopy_batched_##XY(int64_t m, int64_t n, const S* const* dA, int64_t lda, \
                              D* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream) \
{ return gecopy_t<Cu<S>::type, Cu<D>::type>(0, m, n, PCA(S, dA), lda, PA(D, dB), ldb, batch, ST); } \
int sb200_tzcopy_batched_##XY(int uplo, int64_t m, int64_t n, const S* const* dA, int64_t lda, \
                              D* const* dB, int64_t ldb, int64_t batch, sb200_stream_t stream) \
{ return gecopy_t<Cu<S>::type, Cu<D>::type>(tz_mask_of(uplo), m, n, PCA(S, dA), lda, PA(D, dB), ldb, batch, ST); }
SB200_FOR_COPY_PAIRS(SB200_DEF_COPY)

} // extern "C"
