// factor_small.cu -- small latency-bound kernels on the critical path of a panel:
//   * potrf_diag_kernel : Cholesky of one IB x IB (IB = 64) diagonal block in shared memory,
//                         also emits inv(L) so the following block solve is a plain GEMM
//   * trtri_diag_kernel : inverts the IB x IB diagonal blocks of a triangular tile
// and the host compositions built on them and on the batched tile GEMM:
//   * sb200_potrf_tile_{d,s}   (replaces cusolverDn?potrf: lapackpp/src/cuda/cuda_potrf.cc,
//                               call site src/internal/internal_potrf.cc:57-81)
//   * sb200_trsm_batched_{d,s} (replaces cublas?trsmBatched: blaspp/src/device_batch_trsm.cc:27-130,
//                               call site src/internal/internal_trsm.cc:132-262)
// Block algorithm: invert the diagonal IB-blocks once, then block substitution where every
// step is a batched GEMM over all tiles of the block row/column (FP64: DMMA tensor cores) --
// the approach MAGMA / cuBLAS take for large trsm; backward error is governed by the
// conditioning of the IB x IB diagonal blocks only.
#include "gemm_dmma.cuh"
#include "scalar_ops.cuh"
#include "diag64.cuh"
#include <cstdlib>
#include <type_traits>

namespace sb200 {

// potrf_tile_fused.cu (opt-in one-launch tile Cholesky, SB200_TILE_FUSED); declared again in runtime_internal.hh
constexpr int FUSED_NOT_TAKEN_ = -1000001;
int potrf_tile_fused_d(int n, double* A, int lda, int* dinfo, int info_base, int variant, cudaStream_t stream);
int potrf_tile_fused_s(int n, float* A, int lda, int* dinfo, int info_base, int variant, cudaStream_t stream);
// opt-in one-launch panel solve B <- alpha B L^{-T} (SB200_TRSM_FUSED=1), W = inverted diagonal blocks
int trsm_rlt_fused_d(int m, int na, double alpha, const double* Tm, int ldt, const double* W, double* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream);
int trsm_rlt_fused_s(int m, int na, float alpha, const float* Tm, int ldt, const float* W, float* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream);
// opt-in direct substitution for a small triangle, na <= 64 (SB200_TRSM_FUSED bit 2)
int trsm_lln_small_d(int na, int n, double alpha, bool unit, const double* Tm, int ldt, double* const* dB, int64_t offB,
                     int ldb, int batch, cudaStream_t stream);
int trsm_lln_small_s(int na, int n, float alpha, bool unit, const float* Tm, int ldt, float* const* dB, int64_t offB,
                     int ldb, int batch, cudaStream_t stream);
// opt-in one-launch LU row solve B <- alpha L^{-1} B (SB200_TRSM_FUSED bit 1)
int trsm_lln_fused_d(int na, int n, double alpha, const double* Tm, int ldt, const double* W, double* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream);
int trsm_lln_fused_s(int na, int n, float alpha, const float* Tm, int ldt, const float* W, float* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream);

constexpr int IB = 64;
template <typename T> constexpr size_t small_smem() { return 2 * IB * (IB + 1) * sizeof(T); }

// ---------------------------------------------------------------------------------------------
// One CTA (256 threads) factors A (nv x nv, nv <= 64, lower, column-major, lda) in shared
// memory; writes L back (strict upper part untouched) and inv(L) (dense IB x IB, ld = IB,
// zero above the diagonal) to Winv.  On a non-positive pivot at column j writes
// *info = info_base + j + 1 (first failure only) and stops.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
potrf_diag_kernel(T* __restrict__ A, int lda, int nv, T* __restrict__ Winv,
                  int* __restrict__ info, int info_base)
{
    using R = typename RealOf<T>::type;
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    T (*L)[IB + 1] = reinterpret_cast<T (*)[IB + 1]>(smem_dyn);
    T (*X)[IB + 1] = L + IB;
    __shared__ int fail;
    const int tid = threadIdx.x;
    if (tid == 0) fail = 0;
    for (int e = tid; e < IB * IB; e += 256) {
        const int i = e % IB, j = e / IB;
        L[i][j] = (i < nv && j < nv && i >= j) ? A[i + int64_t(j) * lda]
                                               : (i == j ? from_real<T>(R(1)) : zero_of<T>());
    }
    __syncthreads();
    if (*info != 0) return;          // an earlier block already failed: leave the tile alone

    for (int j = 0; j < nv; ++j) {
        const R d = real_of(L[j][j]);        // Hermitian: the imaginary part of the diagonal is ignored
        if (!(d > R(0))) {                   // also catches NaN
            if (tid == 0) { fail = j + 1; }
        }
        __syncthreads();
        if (fail) break;
        const R r = sqrt(d);
        __syncthreads();
        for (int i = j + tid; i < nv; i += 256) L[i][j] = (i == j) ? from_real<T>(r) : div_real(L[i][j], r);
        __syncthreads();
        const int rem = nv - j - 1;
        for (int e = tid; e < rem * rem; e += 256) {
            const int i = j + 1 + e % rem, c = j + 1 + e / rem;
            if (i >= c) L[i][c] = sub(L[i][c], mul(L[i][j], conj_(L[c][j])));
        }
        __syncthreads();
    }
    if (fail) {
        if (tid == 0 && *info == 0) *info = info_base + fail;
        return;
    }
    for (int e = tid; e < nv * nv; e += 256) {
        const int i = e % nv, j = e / nv;
        if (i >= j) A[i + int64_t(j) * lda] = L[i][j];
    }
    // inverse by forward substitution, one column per thread (columns >= nv: identity)
    if (tid < IB) {
        const int j = tid;
        for (int i = 0; i < IB; ++i) X[i][j] = zero_of<T>();
        X[j][j] = divide(from_real<T>(R(1)), L[j][j]);
        for (int i = j + 1; i < IB; ++i) {
            T s = zero_of<T>();
            for (int l = j; l < i; ++l) fma_acc(s, L[i][l], X[l][j]);
            X[i][j] = divide(neg(s), L[i][i]);
        }
    }
    __syncthreads();
    for (int e = tid; e < IB * IB; e += 256) {
        const int i = e % IB, j = e / IB;
        Winv[e] = X[i][j];
    }
}

// ---------------------------------------------------------------------------------------------
// Invert the diagonal IB-blocks of a triangular na x na tile T (column-major, ldt).  Block b is
// written to W + b*IB*IB (ld = IB), zero in the other triangle, padded with identity when the
// last block is ragged.  lower != 0: T lower triangular; unit != 0: unit diagonal.
// ---------------------------------------------------------------------------------------------
// Batched form (gridDim.y > 1): tile y is Tarr[y], its blocks go to W + y * gridDim.x * IB * IB; the LAST tile
// (y == gridDim.y - 1) has na_last rows, the others na.  Blocks wholly outside a ragged tile become identity.
template <typename T>
__global__ void __launch_bounds__(256)
trtri_diag_kernel(const T* __restrict__ Tm, int ldt, int na, int lower, int unit, T* __restrict__ W,
                  const T* const* __restrict__ Tarr = nullptr, int na_last = 0)
{
    using R = typename RealOf<T>::type;
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    T (*L)[IB + 1] = reinterpret_cast<T (*)[IB + 1]>(smem_dyn);   // always handled as LOWER:
    T (*X)[IB + 1] = L + IB;                                       // upper blocks are transposed in
    const int b = blockIdx.x, tid = threadIdx.x;
    if (Tarr) {
        Tm = Tarr[blockIdx.y];
        W += int64_t(blockIdx.y) * gridDim.x * IB * IB;
        if (blockIdx.y == gridDim.y - 1) na = na_last;
    }
    const int o = b * IB;
    const int nv = max(0, min(IB, na - o));
    for (int e = tid; e < IB * IB; e += 256) {
        const int i = e % IB, j = e / IB;
        T v = (i == j) ? from_real<T>(R(1)) : zero_of<T>();
        if (i < nv && j < nv) {
            if (i > j)                v = lower ? Tm[o + i + int64_t(o + j) * ldt] : Tm[o + j + int64_t(o + i) * ldt];
            else if (i == j && !unit) v = Tm[o + i + int64_t(o + j) * ldt];
        }
        L[i][j] = v;
    }
    __syncthreads();
    if (tid < IB) {
        const int j = tid;
        for (int i = 0; i < IB; ++i) X[i][j] = zero_of<T>();
        X[j][j] = divide(from_real<T>(R(1)), L[j][j]);
        for (int i = j + 1; i < IB; ++i) {
            T s = zero_of<T>();
            for (int l = j; l < i; ++l) fma_acc(s, L[i][l], X[l][j]);
            X[i][j] = divide(neg(s), L[i][i]);
        }
    }
    __syncthreads();
    T* Wb = W + int64_t(b) * IB * IB;
    for (int e = tid; e < IB * IB; e += 256) {
        const int i = e % IB, j = e / IB;
        Wb[e] = lower ? X[i][j] : X[j][i];      // inverse of the transpose = transpose of the inverse
    }
}

// ---------------------------------------------------------------------------------------------
// Fast variants for the real types (the critical path of every dpotrf / dgetrf step).
// 64 threads, one matrix ROW (Cholesky) or one COLUMN of the inverse per thread, held in
// registers; the only shared-memory traffic is the broadcast of one column per elimination step,
// and there is ONE __syncthreads per column (the kernels above need four and walk the trailing
// block with div/mod indexing).  Same results up to rounding: the update is formed as
// a_ic - (a_ij / d) * a_cj with the un-normalised column, L_ij = a_ij / sqrt(d) at the end.
// ---------------------------------------------------------------------------------------------
template <typename R> constexpr size_t fast_smem() { return (size_t(IB) * (IB + 1) + IB) * sizeof(R); }

// X = inv(L) for a 64 x 64 lower-triangular L held in shared memory as Ls[j * IB + i] = L_ij
// (entries above the diagonal are never read), rd[i] = 1 / L_ii.  Thread j = threadIdx.x owns
// column j: forward substitution in axpy form, so the only dependent chain is x_i -> x_{i+1}.
template <typename R>
__device__ __forceinline__ void inv_lower_column(const R* __restrict__ Ls, const R* __restrict__ rd,
                                                 int j, R (&x)[IB])
{
    #pragma unroll
    for (int k = 0; k < IB; ++k) x[k] = (k == j) ? R(1) : R(0);
    #pragma unroll
    for (int i = 0; i < IB; ++i) {
        x[i] *= rd[i];
        const R xi = x[i];
        #pragma unroll
        for (int k = i + 1; k < IB; ++k) x[k] = fma(-Ls[i * IB + k], xi, x[k]);
    }
}

// stage the columns held in registers through a padded buffer and write W[i + j*IB] = X_ij
// (transpose_out: W[i + j*IB] = X_ji) with coalesced stores.  All 64 threads must call it.
template <typename R>
__device__ __forceinline__ void store_inverse(R* __restrict__ Xs, const R (&x)[IB], int tid,
                                              R* __restrict__ W, bool transpose_out)
{
    __syncthreads();                               // everybody is done reading Ls (aliased by Xs)
    #pragma unroll
    for (int i = 0; i < IB; ++i) Xs[i * (IB + 1) + tid] = x[i];          // Xs[i][j = tid]
    __syncthreads();
    if (! transpose_out) {
        #pragma unroll 8
        for (int j = 0; j < IB; ++j) W[tid + j * IB] = Xs[tid * (IB + 1) + j];
    }
    else {
        #pragma unroll 8
        for (int j = 0; j < IB; ++j) W[tid + j * IB] = Xs[j * (IB + 1) + tid];
    }
}

// RSQ (the default since round 2; SB200_DIAG_RSQRT=0 for the divided form): one rsqrt per column instead of two
// divisions and a square root -- the dependent chain of software FP64 divisions is what this kernel's 109 us were made
// of; L_ij = a_ij * rinv, update with (a_ij * rinv^2), diag = d * rinv: <= 2-3 ulp from the
// divided form.  Measured (profiles/r02b_pytest_gpu_tail.txt, r2b): the nb = 512 tile 566 -> 458 us, parity tests green.
// The warp-synchronous (474 us) and multi-warp shared-memory (665 us) variants tried in round 2 lost and were deleted.
template <typename R, bool RSQ = false>
__global__ void __launch_bounds__(IB, 1)
potrf_diag_fast_kernel(R* __restrict__ A, int lda, int nv, R* __restrict__ Winv,
                       int* __restrict__ info, int info_base)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    R* Ls = reinterpret_cast<R*>(smem_dyn);        // [IB*(IB+1)]: columns (stride IB), later padded staging
    R* rd = Ls + IB * (IB + 1);                    // [IB] reciprocal diagonal
    const int i = threadIdx.x;                     // this thread's row
    if (*info != 0) return;                        // an earlier block already failed: leave the tile alone
    R a[IB];
    #pragma unroll
    for (int c = 0; c < IB; ++c)
        a[c] = (i < nv && c < nv) ? (c <= i ? A[i + int64_t(c) * lda] : R(0)) : (i == c ? R(1) : R(0));
    int fail = 0;
    R rdiag = R(1);
    #pragma unroll
    for (int j = 0; j < IB; ++j) {
        Ls[j * IB + i] = a[j];
        __syncthreads();
        const R d = Ls[j * IB + j];
        if (fail == 0 && !(d > R(0))) fail = j + 1;          // also catches NaN; uniform over the CTA
        if constexpr (RSQ) {
            const R rinv = rsqrt(d);
            const R w = a[j] * (rinv * rinv);
            #pragma unroll
            for (int c = j + 1; c < IB; ++c) a[c] = fma(-w, Ls[j * IB + c], a[c]);
            a[j] = (i == j) ? d * rinv : a[j] * rinv;
            if (i == j) rdiag = rinv;
        }
        else {
        const R w = a[j] / d;
        #pragma unroll
        for (int c = j + 1; c < IB; ++c) a[c] = fma(-w, Ls[j * IB + c], a[c]);
        const R r = sqrt(d);
        a[j] = (i == j) ? r : a[j] / r;
        if (i == j) rdiag = R(1) / r;
        }
    }
    if (fail) {
        if (i == 0 && *info == 0) *info = info_base + fail;
        return;
    }
    #pragma unroll
    for (int c = 0; c < IB; ++c)
        if (c <= i && i < nv) A[i + int64_t(c) * lda] = a[c];
    // L (final) into shared memory for the inverse; rows above the diagonal are never read
    __syncthreads();
    #pragma unroll
    for (int c = 0; c < IB; ++c) Ls[c * IB + i] = a[c];
    rd[i] = rdiag;
    __syncthreads();
    R x[IB];
    inv_lower_column<R>(Ls, rd, i, x);
    store_inverse<R>(Ls, x, i, Winv, false);
}

// fast trtri of the diagonal IB-blocks (real types): same contract as trtri_diag_kernel
template <typename R>
__global__ void __launch_bounds__(IB, 1)
trtri_diag_fast_kernel(const R* __restrict__ Tm, int ldt, int na, int lower, int unit, R* __restrict__ W,
                       const R* const* __restrict__ Tarr = nullptr, int na_last = 0)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    R* Ls = reinterpret_cast<R*>(smem_dyn);
    R* rd = Ls + IB * (IB + 1);
    const int b = blockIdx.x, tid = threadIdx.x;
    if (Tarr) {
        Tm = Tarr[blockIdx.y];
        W += int64_t(blockIdx.y) * gridDim.x * IB * IB;
        if (blockIdx.y == gridDim.y - 1) na = na_last;
    }
    const int o = b * IB;
    const int nv = max(0, min(IB, na - o));
    // always handled as LOWER: upper blocks are transposed in.  Thread = row i of the lower block.
    #pragma unroll 8
    for (int j = 0; j < IB; ++j) {
        const int i = tid;
        R v = (i == j) ? R(1) : R(0);
        if (i < nv && j < nv) {
            if (i > j)                v = lower ? Tm[o + i + int64_t(o + j) * ldt] : Tm[o + j + int64_t(o + i) * ldt];
            else if (i == j && !unit) v = Tm[o + i + int64_t(o + j) * ldt];
        }
        Ls[j * IB + i] = v;
        if (i == j) rd[i] = R(1) / v;
    }
    __syncthreads();
    R x[IB];
    inv_lower_column<R>(Ls, rd, tid, x);
    // inverse of the transpose = transpose of the inverse
    store_inverse<R>(Ls, x, tid, W + int64_t(b) * IB * IB, lower == 0);
}


template <typename T> struct IsRealType { static constexpr bool value = false; };
template <> struct IsRealType<float>  { static constexpr bool value = true; };
template <> struct IsRealType<double> { static constexpr bool value = true; };

template <typename T>
static int launch_potrf_diag(T* A, int lda, int nv, T* Winv, int* info, int info_base, cudaStream_t stream)
{
    if constexpr (IsRealType<T>::value) {
        const char* e = getenv("SB200_DIAG_RSQRT");                  // per call: a test switches it
        const bool rsq = ! e || atoi(e) != 0;
        if (rsq) potrf_diag_fast_kernel<T, true><<<1, IB, fast_smem<T>(), stream>>>(A, lda, nv, Winv, info, info_base);
        else     potrf_diag_fast_kernel<T, false><<<1, IB, fast_smem<T>(), stream>>>(A, lda, nv, Winv, info, info_base);
    }
    else
        potrf_diag_kernel<T><<<1, 256, small_smem<T>(), stream>>>(A, lda, nv, Winv, info, info_base);
    return launch_status();
}

template <typename T>
static int launch_trtri_diag(int nblk, const T* Tm, int ldt, int na, int lower, int unit, T* W, cudaStream_t stream)
{
    if constexpr (IsRealType<T>::value) {
        trtri_diag_fast_kernel<T><<<nblk, IB, fast_smem<T>(), stream>>>(Tm, ldt, na, lower, unit, W);
    }
    else
        trtri_diag_kernel<T><<<nblk, 256, small_smem<T>(), stream>>>(Tm, ldt, na, lower, unit, W);
    return launch_status();
}

// inverted diagonal IB-blocks of `ntiles` triangular tiles in ONE launch: W[tile][block][IB*IB], nblk = ceil(na / IB)
// blocks per tile (the last tile has na_last <= na rows)
template <typename T>
static int launch_trtri_diag_batched(int ntiles, const T* const* Tarr, int ldt, int na, int na_last, int lower, int unit,
                                     T* W, cudaStream_t stream)
{
    const int nblk = int(ceil_div(na, IB));
    if (ntiles <= 0 || nblk <= 0) return SB200_OK;
    if constexpr (IsRealType<T>::value) {
        trtri_diag_fast_kernel<T><<<dim3(nblk, ntiles), IB, fast_smem<T>(), stream>>>(nullptr, ldt, na, lower, unit, W, Tarr, na_last);
    }
    else
        trtri_diag_kernel<T><<<dim3(nblk, ntiles), 256, small_smem<T>(), stream>>>(nullptr, ldt, na, lower, unit, W, Tarr, na_last);
    return launch_status();
}

template <typename T>
static void small_kernels_init()
{
    static thread_local bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (done[dev & 63]) return;
    cudaFuncSetAttribute(potrf_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(small_smem<T>()));
    cudaFuncSetAttribute(trtri_diag_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(small_smem<T>()));
    if constexpr (IsRealType<T>::value) {
        cudaFuncSetAttribute(potrf_diag_fast_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fast_smem<T>()));
        cudaFuncSetAttribute(potrf_diag_fast_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fast_smem<T>()));
        cudaFuncSetAttribute(trtri_diag_fast_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fast_smem<T>()));
    }
    done[dev & 63] = true;
}

// per-device scratch for callers that pass work == NULL
static void* device_scratch(size_t bytes)
{
    struct Slot { void* p = nullptr; size_t n = 0; };
    static thread_local Slot slots[64];
    int dev = 0;
    cudaGetDevice(&dev);
    Slot& s = slots[dev & 63];
    if (s.n < bytes) {
        if (s.p) cudaFree(s.p);
        s.p = nullptr; s.n = 0;
        if (cudaMalloc(&s.p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        s.n = bytes;
    }
    return s.p;
}

template <typename T>
static GemmParamsT<T> gp(int m, int n, int k, T alpha, T beta, int batch)
{
    GemmParamsT<T> p{};
    p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta; p.batch = batch; p.tri = 0;
    return p;
}

// column-major block solve over a batch of B tiles with ONE triangular tile T (na x na):
//   right: B_t <- alpha B_t op(T)^{-1}   (B_t is m x na)
//   left : B_t <- alpha op(T)^{-1} B_t   (B_t is na x n)
// In-place GEMM steps are safe because one CTA covers the whole 64-wide aliased dimension.
template <typename T>
int trsm_colmajor(bool left, bool lower, int op, bool unit, int m, int n, T alpha,
                  const T* Tm, int ldt, T* const* dB, int64_t offB, int ldb, int batch,
                  T* W, cudaStream_t stream)
{
    const int na = left ? m : n;
    const int nblk = int(ceil_div(na, IB));
    small_kernels_init<T>();
    if constexpr (IsRealType<T>::value) {
        // SB200_TRSM_FUSED (default 7, measured r2a): bit 2 = a small triangle (na <= 64: the U12 solves inside the LU panel)
        // by direct substitution in one launch, no inversion kernel
        if ((switch_value(SW_TRSM_FUSED) & 4) && left && lower && op == 'N' && na <= 32) {      // 13 us vs 24 us (na = 32); slower than the inverse path at 64
            if constexpr (std::is_same<T, double>::value) return trsm_lln_small_d(na, n, alpha, unit, Tm, ldt, dB, offB, ldb, batch, stream);
            else                                          return trsm_lln_small_s(na, n, alpha, unit, Tm, ldt, dB, offB, ldb, batch, stream);
        }
    }
    int st = launch_trtri_diag<T>(nblk, Tm, ldt, na, lower ? 1 : 0, unit ? 1 : 0, W, stream);
    if (st) return st;
    if constexpr (IsRealType<T>::value) {
        // bits 0 / 1 (potrf_tile_fused.cu): one launch after the inverses for
        //   bit 0: the Cholesky panel solve (Right, Lower, Trans, NonUnit)
        //   bit 1: the LU row solve (Left, Lower, NoTrans, Unit / NonUnit)      (bit 2: see above)
        // read per call so that a test can switch it
        const int fused = switch_value(SW_TRSM_FUSED);
        if ((fused & 1) && ! left && lower && op != 'N' && ! unit && na > IB) {
            if constexpr (std::is_same<T, double>::value) return trsm_rlt_fused_d(m, na, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
            else                                          return trsm_rlt_fused_s(m, na, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
        }
        if ((fused & 2) && left && lower && op == 'N' && na > IB) {
            if constexpr (std::is_same<T, double>::value) return trsm_lln_fused_d(na, n, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
            else                                          return trsm_lln_fused_s(na, n, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
        }
    }
    const bool trans = (op != 'N');
    const bool eff_lower = (lower != trans);        // op(T) as a math matrix
    const int opT = trans ? op : 'N';          // 'C' conjugates (complex)

    for (int s = 0; s < nblk; ++s) {
        // right/upper and left/lower sweep forward; right/lower and left/upper sweep backward
        const bool forward = left ? eff_lower : !eff_lower;
        const int j = forward ? s : nblk - 1 - s;
        const int jo = j * IB, jv = min(IB, na - jo);
        const int r0 = forward ? 0 : jo + jv;            // already-solved block range [r0, r1)
        const int r1 = forward ? jo : na;
        const int rk = r1 - r0;
        T scale = alpha;
        if (!left) {
            if (rk > 0) {
                // B_j <- alpha B_j - X[:, r0:r1] * M[r0:r1, j],  M = op(T)
                GemmParamsT<T> p = gp<T>(m, jv, rk, from_real<T>(-1), alpha, batch);
                p.A = dB; p.offA = offB + int64_t(r0) * ldb; p.lda = ldb;
                p.B0 = trans ? Tm + jo + int64_t(r0) * ldt : Tm + r0 + int64_t(jo) * ldt;
                p.ldb = ldt; p.strideB = 0;
                p.C = dB; p.offC = offB + int64_t(jo) * ldb; p.ldc = ldb;
                if ((st = launch_gemm<T>('N', opT, p, stream))) return st;
                scale = from_real<T>(1);
            }
            // B_j <- scale * B_j * op(Winv_j)
            GemmParamsT<T> p = gp<T>(m, jv, jv, scale, zero_of<T>(), batch);
            p.A = dB; p.offA = offB + int64_t(jo) * ldb; p.lda = ldb;
            p.B0 = W + int64_t(j) * IB * IB; p.ldb = IB; p.strideB = 0;
            p.C = dB; p.offC = offB + int64_t(jo) * ldb; p.ldc = ldb;
            if ((st = launch_gemm<T>('N', opT, p, stream))) return st;
        }
        else {
            if (rk > 0) {
                // B_j <- alpha B_j - M[j, r0:r1] * X[r0:r1, :]
                GemmParamsT<T> p = gp<T>(jv, n, rk, from_real<T>(-1), alpha, batch);
                p.A0 = trans ? Tm + r0 + int64_t(jo) * ldt : Tm + jo + int64_t(r0) * ldt;
                p.lda = ldt; p.strideA = 0;
                p.B = dB; p.offB = offB + r0; p.ldb = ldb;
                p.C = dB; p.offC = offB + jo; p.ldc = ldb;
                if ((st = launch_gemm<T>(opT, 'N', p, stream))) return st;
                scale = from_real<T>(1);
            }
            // B_j <- scale * op(Winv_j) * B_j
            GemmParamsT<T> p = gp<T>(jv, n, jv, scale, zero_of<T>(), batch);
            p.A0 = W + int64_t(j) * IB * IB; p.lda = IB; p.strideA = 0;
            p.B = dB; p.offB = offB + jo; p.ldb = ldb;
            p.C = dB; p.offC = offB + jo; p.ldc = ldb;
            if ((st = launch_gemm<T>(opT, 'N', p, stream))) return st;
        }
    }
    return SB200_OK;
}

// lower Cholesky of one n x n tile, blocked by IB; W >= IB*IB elements
template <typename T>
int potrf_tile_lower(int n, T* A, int lda, int* dinfo, int info_base, T* W, cudaStream_t stream, int fused_dflt)
{
    int st;
    if constexpr (IsRealType<T>::value) {
        // the whole tile in one launch (potrf_tile_fused.cu): 0.67 ms against 0.57 ms for the launch chain below on an
        // idle device, but 0.69 ms against 1.1 ms on the chain's own SM partition (fused_dflt = 1 from the driver,
        // profiles/r02d_greenctx_chain_partition.jsonl); the environment switch overrides, read per call
        const char* fe = getenv(SW_TILE_FUSED.env);
        const int fused = fe ? atoi(fe) : (fused_dflt >= 0 ? fused_dflt : SW_TILE_FUSED.dflt);
        if (fused > 0 && n > IB) {
            if constexpr (std::is_same<T, double>::value) st = potrf_tile_fused_d(n, A, lda, dinfo, info_base, fused, stream);
            else                                          st = potrf_tile_fused_s(n, A, lda, dinfo, info_base, fused, stream);
            if (st != FUSED_NOT_TAKEN_) return st;
        }
    }
    small_kernels_init<T>();
    for (int jo = 0; jo < n; jo += IB) {
        const int jv = min(IB, n - jo);
        T* Ajj = A + jo + int64_t(jo) * lda;
        if ((st = launch_potrf_diag<T>(Ajj, lda, jv, W, dinfo, info_base + jo, stream))) return st;
        const int rest = n - jo - jv;
        if (rest <= 0) break;
        T* Pnl = A + (jo + jv) + int64_t(jo) * lda;          // rest x jv block below the diagonal
        // panel <- panel * inv(L_jj)^H   (in place)
        GemmParamsT<T> p = gp<T>(rest, jv, jv, from_real<T>(1), zero_of<T>(), 1);
        p.A0 = Pnl; p.lda = lda; p.B0 = W; p.ldb = IB; p.C0 = Pnl; p.ldc = lda;
        if ((st = launch_gemm<T>('N', 'C', p, stream))) return st;
        // trailing lower triangle -= panel panel^H
        GemmParamsT<T> q = gp<T>(rest, rest, jv, from_real<T>(-1), from_real<T>(1), 1);
        q.A0 = Pnl; q.lda = lda; q.B0 = Pnl; q.ldb = lda;
        q.C0 = A + (jo + jv) + int64_t(jo + jv) * lda; q.ldc = lda; q.tri = 1; q.herk = 1;
        if ((st = launch_gemm<T>('N', 'C', q, stream))) return st;
    }
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------
// Triangular tile solve for FEW right-hand sides (the solve path: potrs / getrs with nrhs ~ 10;
// reference call site work::trsm -> internal::trsm, src/work/work_trsm.cc:60-230).
// Same arithmetic as trsm_colmajor (block substitution with the inverted IB x IB diagonal blocks),
// but the whole na x na triangle is walked inside ONE CTA per <= TS_NC columns of one B tile, instead
// of 2 GEMM launches per diagonal block: the solve path is a chain of nt dependent tile solves, so
// launch latency, not bandwidth, is what it pays for.
//   mode 0: lower, NoTrans   (forward;  right-looking: thread-per-row axpy updates, coalesced down T's columns)
//   mode 1: lower, Trans/Conj (backward; left-looking: warp-per-row dot products down T's columns)
//   mode 2: upper, NoTrans   (backward; right-looking)
// Winv: the nblk inverted diagonal blocks of T (IB x IB dense each, from the trtri kernels).
// ---------------------------------------------------------------------------------------------
constexpr int TS_NC = 8;            // right-hand-side columns per CTA
constexpr int TS_THREADS = 256;

// 16-byte global loads of `n16` consecutive chunks into a register array of T
__device__ __forceinline__ void ld16(const float* p, float* o)   { const float4 v = *reinterpret_cast<const float4*>(p); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void ld16(const double* p, double* o) { const double2 v = *reinterpret_cast<const double2*>(p); o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ void ld16(const cuFloatComplex* p, cuFloatComplex* o)
{
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = make_cuFloatComplex(v.x, v.y); o[1] = make_cuFloatComplex(v.z, v.w);
}
__device__ __forceinline__ void ld16(const cuDoubleComplex* p, cuDoubleComplex* o) { o[0] = *p; }

template <typename T>
__global__ void __launch_bounds__(TS_THREADS)
trsm_small_kernel(const T* __restrict__ Tm, int ldt, const T* __restrict__ Winv, T* const* __restrict__ Barr,
                  int64_t offB, int ldb, int na, int n, int mode, int conj)
{
    constexpr int QB = sizeof(T) <= 8 ? 32 : 16;         // loads in flight per thread in the update
    constexpr int PER16 = 16 / int(sizeof(T));           // elements per 16-byte load
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    T* Xs = reinterpret_cast<T*>(smem_dyn);          // [TS_NC][na]   the right-hand sides / solution
    T* V  = Xs + size_t(TS_NC) * na;                 // [TS_NC][IB]   block staging
    T* Ws = V + TS_NC * IB;                          // [IB][IB]      inverted diagonal block of this step
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * TS_NC;
    const int nc = min(TS_NC, n - c0);
    T* __restrict__ B = Barr[blockIdx.y] + offB + int64_t(c0) * ldb;
    const int nblk = (na + IB - 1) / IB;

    {   // pull the stored triangle of T towards L2 in one go: the walk below is a chain of dependent round trips
        constexpr int PER_LINE = 128 / int(sizeof(T));
        const int lines_per_col = (na + PER_LINE - 1) / PER_LINE;
        for (int e = tid; e < lines_per_col * na; e += TS_THREADS) {
            const int col = e / lines_per_col, r0 = (e - col * lines_per_col) * PER_LINE;
            const bool need = (mode == 2) ? (r0 <= col) : (r0 + PER_LINE > col);
            if (need) asm volatile("prefetch.global.L2 [%0];" :: "l"(Tm + r0 + int64_t(col) * ldt));
        }
    }
    for (int e = tid; e < TS_NC * na; e += TS_THREADS) {
        const int c = e / na, r = e - c * na;
        Xs[e] = (c < nc) ? B[r + int64_t(c) * ldb] : zero_of<T>();
    }
    __syncthreads();

    for (int s = 0; s < nblk; ++s) {
        const int b = (mode == 0) ? s : nblk - 1 - s;
        const int o = b * IB, nv = min(IB, na - o);
        {   // stage the inverted diagonal block (independent coalesced loads: one memory round trip)
            const T* __restrict__ Wg = Winv + int64_t(b) * IB * IB;
            #pragma unroll
            for (int e = 0; e < IB * IB / TS_THREADS; ++e) Ws[tid + e * TS_THREADS] = Wg[tid + e * TS_THREADS];
        }
        for (int e = tid; e < TS_NC * nv; e += TS_THREADS) {
            const int c = e / nv, r = e - c * nv;
            V[c * IB + r] = Xs[c * na + o + r];
        }
        __syncthreads();
        // y = op(Winv_b) v  (Winv_b is dense IB x IB with zeros in the other triangle)
        for (int e = tid; e < TS_NC * nv; e += TS_THREADS) {
            const int c = e / nv, r = e - c * nv;
            T y = zero_of<T>();
            if (mode == 1) {
                for (int q = r; q < nv; ++q) {           // op(W)(r, q) = [conj] W(q, r), W lower
                    T w = Ws[q + r * IB];
                    if (conj) w = conj_(w);
                    fma_acc(y, w, V[c * IB + q]);
                }
            }
            else if (mode == 0) { for (int q = 0; q <= r; ++q) fma_acc(y, Ws[r + q * IB], V[c * IB + q]); }
            else                { for (int q = r; q < nv; ++q) fma_acc(y, Ws[r + q * IB], V[c * IB + q]); }
            Xs[c * na + o + r] = y;
        }
        __syncthreads();
        // right-looking update of the rows not solved yet:  x(i, :) -= op(T)(i, o : o + nv) y
        //   mode 0 / 2: op(T)(i, o + q) = T(i, o + q): thread-per-row, coalesced down T's columns
        //   mode 1    : op(T)(i, o + q) = [conj] T(o + q, i): thread-per-row reads a CONTIGUOUS segment of column i
        const int i_lo = (mode == 0) ? o + nv : 0, i_hi = (mode == 0) ? na : o;
        for (int i = i_lo + tid; i < i_hi; i += TS_THREADS) {
            T acc[TS_NC];
            #pragma unroll
            for (int c = 0; c < TS_NC; ++c) acc[c] = zero_of<T>();
            if (mode != 1) {
                const T* __restrict__ row = Tm + i + int64_t(o) * ldt;
                for (int q0 = 0; q0 < nv; q0 += QB) {
                    T t[QB];
                    #pragma unroll
                    for (int u = 0; u < QB; ++u) t[u] = (q0 + u < nv) ? row[int64_t(q0 + u) * ldt] : zero_of<T>();
                    #pragma unroll
                    for (int u = 0; u < QB; ++u) {
                        if (q0 + u >= nv) continue;
                        #pragma unroll
                        for (int c = 0; c < TS_NC; ++c) fma_acc(acc[c], t[u], Xs[c * na + o + q0 + u]);
                    }
                }
            }
            else {
                const T* __restrict__ seg = Tm + o + int64_t(i) * ldt;
                const bool vec = (nv % PER16 == 0) && ((reinterpret_cast<uintptr_t>(seg) & 15) == 0);
                for (int q0 = 0; q0 < nv; q0 += QB) {
                    T t[QB];
                    if (vec) {
                        #pragma unroll
                        for (int u = 0; u < QB; u += PER16) {
                            if (q0 + u < nv) ld16(seg + q0 + u, t + u);
                            else { 
                                #pragma unroll
                                for (int w = 0; w < PER16; ++w) t[u + w] = zero_of<T>();
                            }
                        }
                    }
                    else {
                        #pragma unroll
                        for (int u = 0; u < QB; ++u) t[u] = (q0 + u < nv) ? seg[q0 + u] : zero_of<T>();
                    }
                    #pragma unroll
                    for (int u = 0; u < QB; ++u) {
                        if (q0 + u >= nv) continue;
                        const T tv = conj ? conj_(t[u]) : t[u];
                        #pragma unroll
                        for (int c = 0; c < TS_NC; ++c) fma_acc(acc[c], tv, Xs[c * na + o + q0 + u]);
                    }
                }
            }
            #pragma unroll
            for (int c = 0; c < TS_NC; ++c) Xs[c * na + i] = sub(Xs[c * na + i], acc[c]);
        }
        __syncthreads();
    }
    for (int e = tid; e < nc * na; e += TS_THREADS) {
        const int c = e / na, r = e - c * na;
        B[r + int64_t(c) * ldb] = Xs[e];
    }
}

template <typename T> constexpr size_t ts_smem(int na) { return (size_t(TS_NC) * na + size_t(TS_NC) * IB + size_t(IB) * IB) * sizeof(T); }

// Inverted diagonal blocks of every diagonal tile of a sweep, one launch (see launch_trtri_diag_batched).
template <typename T>
int trtri_diag_all(int ntiles, const T* const* Tarr, int ldt, int na, int na_last, bool lower, bool unit, T* W, cudaStream_t stream)
{
    small_kernels_init<T>();
    return launch_trtri_diag_batched<T>(ntiles, Tarr, ldt, na, na_last, lower ? 1 : 0, unit ? 1 : 0, W, stream);
}

// B_t <- op(T)^{-1} B_t for `batch` tiles of n columns (left side).  Returns SB200_ENOTSUP when the case is not
// served by the small kernel (upper + transposed, or a triangle too tall for shared memory): callers fall back to
// trsm_colmajor.
template <typename T>
int trsm_small(bool lower, int op, int na, int n, const T* Tm, int ldt, const T* Winv, T* const* dB, int64_t offB,
               int ldb, int batch, cudaStream_t stream)
{
    if (na <= 0 || n <= 0 || batch <= 0) return SB200_OK;
    const bool trans = (op != 'N');
    if (! lower && trans) return SB200_ENOTSUP;
    const size_t smem = ts_smem<T>(na);
    if (smem > 200 * 1024) return SB200_ENOTSUP;
    static thread_local size_t attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_set[dev & 63] < smem) {
        cudaError_t e = cudaFuncSetAttribute(trsm_small_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(200 * 1024));
        if (e != cudaSuccess) return int(e);
        attr_set[dev & 63] = 200 * 1024;
    }
    const int mode = lower ? (trans ? 1 : 0) : 2;
    trsm_small_kernel<T><<<dim3(unsigned(ceil_div(n, TS_NC)), unsigned(batch)), TS_THREADS, smem, stream>>>(
        Tm, ldt, Winv, dB, offB, ldb, na, n, mode, op == 'C' ? 1 : 0);
    return launch_status();
}

#define SB200_INST_SMALL(T) \
    template int trtri_diag_all<T>(int, const T* const*, int, int, int, bool, bool, T*, cudaStream_t); \
    template int trsm_small<T>(bool, int, int, int, const T*, int, const T*, T* const*, int64_t, int, int, cudaStream_t);
SB200_INST_SMALL(float)
SB200_INST_SMALL(double)
SB200_INST_SMALL(cuFloatComplex)
SB200_INST_SMALL(cuDoubleComplex)

// the drivers (runtime.cu, solve.cu, getrf.cu) use these for every scalar type
#define SB200_INST_FACTOR(T) \
    template int trsm_colmajor<T>(bool, bool, int, bool, int, int, T, const T*, int, T* const*, int64_t, int, int, T*, cudaStream_t); \
    template int potrf_tile_lower<T>(int, T*, int, int*, int, T*, cudaStream_t, int);
SB200_INST_FACTOR(float)
SB200_INST_FACTOR(double)
SB200_INST_FACTOR(cuFloatComplex)
SB200_INST_FACTOR(cuDoubleComplex)

// double-precision entry points used by the runtime
int trsm_colmajor_d(bool left, bool lower, int op, bool unit, int m, int n, double alpha,
                    const double* T, int ldt, double* const* dB, int64_t offB, int ldb, int batch,
                    double* W, cudaStream_t stream)
{
    return trsm_colmajor<double>(left, lower, op, unit, m, n, alpha, T, ldt, dB, offB, ldb, batch, W, stream);
}
int potrf_tile_lower_d(int n, double* A, int lda, int* dinfo, int info_base, double* W, cudaStream_t stream)
{
    return potrf_tile_lower<double>(n, A, lda, dinfo, info_base, W, stream, -1);
}

template <typename T>
static int trsm_batched_t(int layout, int side, int uplo, int op, int diag, int64_t m, int64_t n, T alpha,
                          const T* dA, int64_t lda, T* const* dB, int64_t ldb, int64_t batch, void* work,
                          cudaStream_t stream)
{
    if (! valid_layout(layout) || ! valid_side(side) || ! valid_uplo(uplo) || ! valid_op(op) || ! valid_diag(diag))
        return SB200_EINVAL;
    if (m < 0 || n < 0 || batch < 0) return SB200_EINVAL;
    if (m == 0 || n == 0 || batch == 0) return SB200_OK;
    if (m > 0x7fffffff || n > 0x7fffffff || lda > 0x7fffffff || ldb > 0x7fffffff || batch > 0x7fffffff)
        return SB200_EINVAL;
    bool left = (side == 'L'), lower = (uplo == 'L');
    if (layout == 'R') {
        // row-major B (m x n) is column-major B^T (n x m): flip side and uplo, keep op
        // (the reference does the same: blaspp/src/device_batch_trsm.cc:82-87)
        left = !left; lower = !lower; std::swap(m, n);
    }
    const int64_t na = left ? m : n;
    if (lda < na || ldb < m) return SB200_EINVAL;
    const size_t wbytes = size_t(ceil_div(na, IB)) * IB * IB * sizeof(T);
    T* W = static_cast<T*>(work);
    if (! W) {
        W = static_cast<T*>(device_scratch(wbytes));
        if (! W) return SB200_ENOMEM;
    }
    return trsm_colmajor<T>(left, lower, op, diag == 'U', int(m), int(n), alpha, dA, int(lda),
                            dB, 0, int(ldb), int(batch), W, stream);
}

template <typename T>
static int potrf_tile_t(int uplo, int64_t n, T* dA, int64_t lda, int* dinfo, void* work, cudaStream_t stream)
{
    if (! valid_uplo(uplo) || n < 0 || lda < (n > 1 ? n : 1) || n > 0x7fffffff || lda > 0x7fffffff)
        return SB200_EINVAL;
    if (uplo != 'L') return SB200_ENOTSUP;    // SLATE's potrf driver works on the lower triangle (src/potrf.cc:230-240)
    cudaError_t e = cudaMemsetAsync(dinfo, 0, sizeof(int), stream);
    if (e != cudaSuccess) return int(e);
    if (n == 0) return SB200_OK;
    T* W = static_cast<T*>(work);
    if (! W) {
        W = static_cast<T*>(device_scratch(size_t(IB) * IB * sizeof(T)));
        if (! W) return SB200_ENOMEM;
    }
    return potrf_tile_lower<T>(int(n), dA, int(lda), dinfo, 0, W, stream, -1);
}

template <typename A> struct Cu { using type = A; };
template <> struct Cu<sb200_c32> { using type = cuFloatComplex; };
template <> struct Cu<sb200_c64> { using type = cuDoubleComplex; };
static inline float  cv(float v) { return v; }
static inline double cv(double v) { return v; }
static inline cuFloatComplex  cv(sb200_c32 v) { return make_cuFloatComplex(v.re, v.im); }
static inline cuDoubleComplex cv(sb200_c64 v) { return make_cuDoubleComplex(v.re, v.im); }

} // namespace sb200

using namespace sb200;

extern "C" {

size_t sb200_trsm_work_bytes(int dtype, int side, int64_t m, int64_t n)
{
    const int64_t na = (side == 'L') ? m : n;
    const size_t es = dtype == 's' ? 4 : (dtype == 'd' || dtype == 'c') ? 8 : 16;
    return size_t(ceil_div(na > 0 ? na : 1, IB)) * IB * IB * es;
}
size_t sb200_trsm_work_bytes_d(int side, int64_t m, int64_t n) { return sb200_trsm_work_bytes('d', side, m, n); }
size_t sb200_potrf_work_bytes_d(int64_t n) { (void) n; return size_t(IB) * IB * sizeof(double); }

#define SB200_DEF_FACTOR(X, T, R) \
int sb200_trsm_batched_##X(int layout, int side, int uplo, int op, int diag, int64_t m, int64_t n, T alpha, \
                           const T* dA, int64_t lda, T* const* dB, int64_t ldb, \
                           int64_t batch, void* work, sb200_stream_t stream) \
{ return trsm_batched_t<Cu<T>::type>(layout, side, uplo, op, diag, m, n, cv(alpha), \
      reinterpret_cast<const Cu<T>::type*>(dA), lda, reinterpret_cast<Cu<T>::type* const*>(dB), ldb, batch, work, cudaStream_t(stream)); } \
int sb200_potrf_tile_##X(int uplo, int64_t n, T* dA, int64_t lda, int* dinfo, void* work, sb200_stream_t stream) \
{ return potrf_tile_t<Cu<T>::type>(uplo, n, reinterpret_cast<Cu<T>::type*>(dA), lda, dinfo, work, cudaStream_t(stream)); }
SB200_FOR_TYPES(SB200_DEF_FACTOR)

} // extern "C"
