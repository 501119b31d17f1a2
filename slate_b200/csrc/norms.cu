// norms.cu -- per-tile partial norms (seam 2: device::genorm / henorm / synorm / synormOffdiag /
// trnorm; reference src/cuda/device_{genorm,henorm,synorm,trnorm}.cu and device_util.cuh).
//
// Output layout is the reference's (device_genorm.cu:373-445): per tile t
//   Max  -> values[t]            (ldv = 1), NaN-propagating (device_util.cuh:22-25)
//   One  -> values[t*ldv + j]    column sums        Inf -> values[t*ldv + i]   row sums
//   Fro  -> values[2t] = scale, values[2t+1] = sumsq  with scale^2 * sumsq = sum |a|^2
//   scope Columns + Max -> values[t*ldv + j] = max_i |a_ij|   (ge_col_norms_max_kernel)
// The reference uses shared-memory tree reductions and, for Frobenius, a serial combine on
// thread 0 (device_genorm.cu:268-277).  Here: one CTA (512 threads) per tile, warp-shuffle
// reductions, deterministic combine order, columns walked by warps (coalesced 256-byte lines)
// and rows by threads (coalesced across the CTA).  HBM-bound: every lane keeps 8 unconditional
// loads in flight (the shape / diagonal classification is applied to the loaded values), and the
// Frobenius accumulation caches 1 / scale so that the common case is one multiply and one FMA.
// Measured on B200 (1024 tiles of 512 x 512 FP64): max 5.3, one 6.0, inf 5.1, fro 3.7 TB/s of 6.5.
#include "common.cuh"
#include "scalar_ops.cuh"

namespace sb200 {


__device__ inline float  abs_(float a) { return fabsf(a); }
__device__ inline double abs_(double a) { return fabs(a); }
__device__ inline float  abs_(cuFloatComplex a) { return cuCabsf(a); }
__device__ inline double abs_(cuDoubleComplex a) { return cuCabs(a); }
// Hermitian diagonal: only the real part counts (device_henorm.cu uses abs(real(a_jj)))
__device__ inline float  abs_real(float a) { return fabsf(a); }
__device__ inline double abs_real(double a) { return fabs(a); }
__device__ inline float  abs_real(cuFloatComplex a) { return fabsf(a.x); }
__device__ inline double abs_real(cuDoubleComplex a) { return fabs(a.x); }

template <typename R> __device__ inline R max_nan(R a, R b)
{
    return (a != a) ? a : ((b != b) ? b : (a > b ? a : b));
}

// (scale, sumsq) accumulation, LAPACK lassq style (device_util.cuh:243-268 add_sumsq / combine_sumsq)
template <typename R> __device__ inline void add_sumsq(R& scale, R& sumsq, R absx)
{
    if (absx != absx) { scale = absx; return; }                 // NaN poisons the result
    if (scale != scale) return;
    if (absx == R(0)) return;
    if (scale < absx) { const R r = scale / absx; sumsq = R(1) + sumsq * r * r; scale = absx; }
    else              { const R r = absx / scale; sumsq += r * r; }
}
// per-thread streaming variant: `inv` caches 1 / scale so that the common case (|x| <= scale) is one
// multiply and one FMA; the division only happens when the running maximum changes.  The value kept is
// the same (scale, sumsq) pair up to rounding of x / scale vs x * (1 / scale).
__device__ inline float  fma_r(float a, float b, float c)    { return fmaf(a, b, c); }
__device__ inline double fma_r(double a, double b, double c) { return ::fma(a, b, c); }
template <typename R> struct SumSq {
    R scale = R(0), sumsq = R(1), inv = R(0);
    // N values at once: bmax = their maximum (0 if none counts), nan = one of them is NaN
    template <int N> __device__ inline void add_batch(const R (&v)[N], const R (&wgt)[N], R bmax, bool nan)
    {
        if (nan) { R t = R(0); for (int u = 0; u < N; ++u) t += v[u]; scale = t; return; }      // NaN poisons the result
        if (scale != scale || bmax == R(0)) return;
        if (scale < bmax) { const R r = scale / bmax; sumsq = sumsq * r * r; scale = bmax; inv = R(1) / bmax; }
        #pragma unroll
        for (int u = 0; u < N; ++u) { const R r = v[u] * inv; sumsq = fma_r(r * wgt[u], r, sumsq); }
    }
    __device__ inline void add(R absx)
    {
        if (absx != absx) { scale = absx; return; }             // NaN poisons the result
        if (scale != scale) return;
        if (absx == R(0)) return;
        if (scale < absx) { const R r = scale / absx; sumsq = R(1) + sumsq * r * r; scale = absx; inv = R(1) / absx; }
        else              { const R r = absx * inv; sumsq = fma_r(r, r, sumsq); }
    }
};

template <typename R> __device__ inline void combine_sumsq(R& scale, R& sumsq, R scale2, R sumsq2)
{
    if (scale2 != scale2) { scale = scale2; return; }
    if (scale != scale) return;
    if (scale2 == R(0)) return;
    if (scale == R(0)) { scale = scale2; sumsq = sumsq2; return; }
    if (scale > scale2) { const R r = scale2 / scale; sumsq += sumsq2 * r * r; }
    else                { const R r = scale / scale2; sumsq = sumsq * r * r + sumsq2; scale = scale2; }
}

template <typename R> __device__ inline R warp_sum(R v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename R> __device__ inline R warp_max_nan(R v)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Element classes for the structured tiles.
//   shape 0: general;  1: lower trapezoid (i >= j);  2: upper trapezoid (i <= j)
//   sym  0: plain (ge / tr);  1: Hermitian/symmetric diagonal tile stored in `shape`
//            (off-diagonal entries count for both (i,j) and (j,i); herm: diag uses |real|)
//   unit: triangular with implicit unit diagonal
struct NormCfg { int shape, sym, herm, unit; };

__device__ inline bool in_shape(const NormCfg& c, int i, int j)
{
    return c.shape == 0 || (c.shape == 1 ? i >= j : i <= j);
}

// |a_ij| as the norm counts it, from an ALREADY LOADED value (the loads themselves stay unconditional /
// predicated only on the bounds, so that the compiler issues UN of them back to back)
template <typename T>
__device__ inline typename RealOf<T>::type abs_loaded(const NormCfg& c, T raw, int i, int j)
{
    using R = typename RealOf<T>::type;
    if (i == j && c.unit) return R(1);
    if (i == j && c.herm) return abs_real(raw);
    return abs_(raw);
}

constexpr int NORM_THREADS = 512;

// mode: 'M' max, 'O' one (column sums [+ row part for sym]), 'I' inf (row sums), 'F' fro,
//       'C' column maxima, 'B' both column sums (values[0..n)) and row sums (values[n..n+m))
template <typename T>
__global__ void __launch_bounds__(NORM_THREADS)
norm_kernel(int mode, NormCfg cfg, int m, int n, const T* const* A, int64_t lda,
            typename RealOf<T>::type* values, int64_t ldv)
{
    using R = typename RealOf<T>::type;
    const T* a = A[blockIdx.x];
    R* out = values + int64_t(blockIdx.x) * ldv;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int NW = NORM_THREADS / 32;
    __shared__ R red[2 * NW];

    constexpr int UN = 8;                       // independent loads in flight per lane (HBM latency hiding)
    if (mode == 'M' || mode == 'F') {
        R vmax = 0;
        SumSq<R> ss;
        for (int j = warp; j < n; j += NW) {
            for (int i0 = lane; i0 < m; i0 += 32 * UN) {
                T raw[UN];
                #pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = i0 + 32 * u;
                    raw[u] = (i < m) ? a[i + int64_t(j) * lda] : zero_of<T>();
                }
                if (mode == 'M') {
                    // NaN-propagating max (device_util.cuh:22-25) as a plain max + one sticky NaN per batch
                    R bm = R(0), bnan = R(0);
                    #pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int i = i0 + 32 * u;
                        const bool on = i < m && in_shape(cfg, i, j);
                        const R v = on ? abs_loaded(cfg, raw[u], i, j) : R(0);
                        if (v != v) bnan = v;
                        bm = v > bm ? v : bm;
                    }
                    vmax = (bnan != bnan) ? bnan : max_nan(vmax, bm);
                }
                else {
                    // Frobenius: ONE rescale decision per batch of UN loaded values (the branchy per-element lassq
                    // update made this kernel instruction-bound: 3.7 of 6.5 TB/s), then UN straight multiply + FMA
                    R v[UN], wgt[UN], bmax = R(0);
                    bool nan = false;
                    #pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int i = i0 + 32 * u;
                        const bool on = i < m && in_shape(cfg, i, j);
                        v[u] = on ? abs_loaded(cfg, raw[u], i, j) : R(0);
                        wgt[u] = (cfg.sym && i != j) ? R(2) : R(1);          // mirrored entry counts twice
                        nan |= v[u] != v[u];
                        bmax = v[u] > bmax ? v[u] : bmax;
                    }
                    ss.add_batch(v, wgt, bmax, nan);
                }
            }
        }
        R scale = ss.scale, sumsq = ss.sumsq;
        if (mode == 'M') {
            vmax = warp_max_nan(vmax);
            if (lane == 0) red[warp] = vmax;
            __syncthreads();
            if (warp == 0) {
                R v = lane < NW ? red[lane] : R(0);
                v = warp_max_nan(v);
                if (lane == 0) out[0] = v;
            }
        }
        else {
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const R s2 = __shfl_xor_sync(0xffffffffu, scale, o), q2 = __shfl_xor_sync(0xffffffffu, sumsq, o);
                combine_sumsq(scale, sumsq, s2, q2);     // only lane 0's (deterministic) result is used
            }
            if (lane == 0) { red[2 * warp] = scale; red[2 * warp + 1] = sumsq; }
            __syncthreads();
            if (tid == 0) {
                R s = red[0], q = red[1];
                for (int w = 1; w < NW; ++w) combine_sumsq(s, q, red[2 * w], red[2 * w + 1]);
                out[0] = s; out[1] = q;
            }
        }
        return;
    }

    if (mode == 'O' || mode == 'C' || mode == 'B') {
        // column pass: one warp per column
        for (int j = warp; j < n; j += NW) {
            R acc = 0;
            for (int i0 = lane; i0 < m; i0 += 32 * UN) {
                T raw[UN];
                #pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = i0 + 32 * u;
                    raw[u] = (i < m) ? a[i + int64_t(j) * lda] : zero_of<T>();
                }
                #pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int i = i0 + 32 * u;
                    if (i >= m || ! in_shape(cfg, i, j)) continue;
                    const R v = abs_loaded(cfg, raw[u], i, j);
                    acc = (mode == 'C') ? max_nan(acc, v) : acc + v;
                }
            }
            acc = (mode == 'C') ? warp_max_nan(acc) : warp_sum(acc);
            if (lane == 0) out[j] = acc;
        }
    }
    if (mode == 'I' || mode == 'B' || (mode == 'O' && cfg.sym)) {
        if (mode == 'O') __syncthreads();          // column sums of this CTA are written
        // row pass: one thread per row (coalesced across the CTA), deterministic order in j
        for (int i = tid; i < m; i += NORM_THREADS) {
            R acc = 0;
            for (int j0 = 0; j0 < n; j0 += UN) {
                T raw[UN];
                #pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int j = j0 + u;
                    raw[u] = (j < n) ? a[i + int64_t(j) * lda] : zero_of<T>();
                }
                #pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int j = j0 + u;
                    if (j >= n || ! in_shape(cfg, i, j) || (mode == 'O' && i == j)) continue;   // sym: diagonal already in the column sum
                    acc += abs_loaded(cfg, raw[u], i, j);
                }
            }
            if (mode == 'I') out[i] = acc;
            else if (mode == 'B') out[n + i] = acc;
            else out[i] += acc;                                 // sym one-norm: mirrored part of column i
        }
    }
}

template <typename T>
static int launch_norm(int mode, NormCfg cfg, int64_t m, int64_t n, const T* const* A, int64_t lda,
                       typename RealOf<T>::type* values, int64_t ldv, int64_t batch, cudaStream_t s)
{
    using R = typename RealOf<T>::type;
    if (m < 0 || n < 0 || batch < 0) return SB200_EINVAL;
    if (batch == 0) return SB200_OK;
    if (m > 0x7fffffff || n > 0x7fffffff || batch > 0x7fffffff || lda < m) return SB200_EINVAL;
    int64_t need = 1;
    if (mode == 'O' || mode == 'C') need = n;
    if (mode == 'I') need = m;
    if (mode == 'F') need = 2;
    if (mode == 'B') need = m + n;
    if (ldv < need) return SB200_EINVAL;
    if (m == 0 || n == 0) {
        // empty tiles: zero result (reference: device_genorm.cu:393-395)
        cudaError_t e = cudaMemsetAsync(values, 0, size_t(batch) * ldv * sizeof(R), s);
        return e == cudaSuccess ? SB200_OK : int(e);
    }
    norm_kernel<T><<<unsigned(batch), NORM_THREADS, 0, s>>>(mode, cfg, int(m), int(n), A, lda, values, ldv);
    return launch_status();
}

static int norm_mode(int norm, int scope)
{
    if (norm == '1') norm = 'O';        // lapack::Norm::One is the character '1' (blaspp to_char); 'O' is accepted too
    if (scope == 'C') return norm == 'M' ? 'C' : -1;
    if (scope != 'M') return -1;
    return (norm == 'M' || norm == 'O' || norm == 'I' || norm == 'F') ? norm : -1;
}

// Hermitian / symmetric DIAGONAL tiles: One == Inf (device_henorm.cu:300-345)
static int he_mode(int norm) { return norm == 'I' ? 'O' : norm_mode(norm, 'M'); }
static bool is_one_or_inf(int norm) { return norm == 'O' || norm == '1' || norm == 'I'; }

template <typename A> struct Cu { using type = A; };
template <> struct Cu<sb200_c32> { using type = cuFloatComplex; };
template <> struct Cu<sb200_c64> { using type = cuDoubleComplex; };

} // namespace sb200

using namespace sb200;
#define ST cudaStream_t(stream)
#define CPP(T, p) reinterpret_cast<const Cu<T>::type* const*>(p)

extern "C" {

#define SB200_DEF_NORMS(X, T, R) \
int sb200_genorm_batched_##X(int norm, int scope, int64_t m, int64_t n, const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream) \
{ \
    const int mode = norm_mode(norm, scope); \
    if (mode < 0) return SB200_ENOTSUP; \
    return launch_norm<Cu<T>::type>(mode, NormCfg{0, 0, 0, 0}, m, n, CPP(T, dA), lda, values, ldv, batch, ST); \
} \
int sb200_henorm_batched_##X(int norm, int uplo, int64_t n, const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream) \
{ \
    if (! valid_uplo(uplo)) return SB200_EINVAL; \
    if (he_mode(norm) < 0) return SB200_ENOTSUP; \
    return launch_norm<Cu<T>::type>(he_mode(norm), NormCfg{uplo == 'L' ? 1 : 2, 1, 1, 0}, n, n, CPP(T, dA), lda, values, ldv, batch, ST); \
} \
int sb200_synorm_batched_##X(int norm, int uplo, int64_t n, const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream) \
{ \
    if (! valid_uplo(uplo)) return SB200_EINVAL; \
    if (he_mode(norm) < 0) return SB200_ENOTSUP; \
    return launch_norm<Cu<T>::type>(he_mode(norm), NormCfg{uplo == 'L' ? 1 : 2, 1, 0, 0}, n, n, CPP(T, dA), lda, values, ldv, batch, ST); \
} \
/* full off-diagonal tile of a symmetric matrix: column sums in values[0..n), row sums in \
 * values[n..n+m)  (synorm_offdiag_one_kernel, device_synorm.cu:381-431) */ \
int sb200_synorm_offdiag_batched_##X(int norm, int64_t m, int64_t n, const T* const* dA, int64_t lda, \
                                     R* values, int64_t ldv, int64_t batch, sb200_stream_t stream) \
{ \
    if (! is_one_or_inf(norm)) return SB200_ENOTSUP; \
    return launch_norm<Cu<T>::type>('B', NormCfg{0, 0, 0, 0}, m, n, CPP(T, dA), lda, values, ldv, batch, ST); \
} \
int sb200_trnorm_batched_##X(int norm, int uplo, int diag, int64_t m, int64_t n, const T* const* dA, int64_t lda, \
                             R* values, int64_t ldv, int64_t batch, sb200_stream_t stream) \
{ \
    if (! valid_uplo(uplo) || ! valid_diag(diag)) return SB200_EINVAL; \
    const int mode = norm_mode(norm, 'M'); \
    if (mode < 0) return SB200_ENOTSUP; \
    return launch_norm<Cu<T>::type>(mode, NormCfg{uplo == 'L' ? 1 : 2, 0, 0, diag == 'U'}, m, n, CPP(T, dA), lda, values, ldv, batch, ST); \
}
SB200_FOR_TYPES(SB200_DEF_NORMS)

} // extern "C"
