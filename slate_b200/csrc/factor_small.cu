// factor_small.cu -- small latency-bound kernels on the critical path of a panel:
//   * potrf_diag_kernel : Cholesky of one IB x IB (IB = 64) diagonal block in shared memory,
//                         also emits inv(L) so the following block solve is a plain GEMM
//   * trtri_diag_kernel : inverts the IB x IB diagonal blocks of a triangular tile
// and the host compositions built on them and on the DMMA GEMM:
//   * sb200_potrf_tile_d   (replaces cusolverDnDpotrf: lapackpp/src/cuda/cuda_potrf.cc,
//                           call site src/internal/internal_potrf.cc:57-81)
//   * sb200_trsm_batched_d (replaces cublasDtrsmBatched: blaspp/src/device_batch_trsm.cc:27-130,
//                           call site src/internal/internal_trsm.cc:132-262)
// Block algorithm: invert the diagonal IB-blocks once, then block substitution where every
// step is a batched DMMA GEMM over all tiles of the block row/column (the approach MAGMA /
// cuBLAS take for large trsm; backward error is governed by cond of the IB x IB blocks only).
#include "gemm_dmma.cuh"

namespace sb200 {

constexpr int IB = 64;
constexpr size_t SMALL_SMEM = 2 * IB * (IB + 1) * sizeof(double);

// opt the two small kernels in to > 48 KB of dynamic shared memory (once per device)
static void small_kernels_init();

// ---------------------------------------------------------------------------------------------
// One CTA (256 threads) factors A (nv x nv, nv <= 64, lower, column-major, lda) in shared
// memory; writes L back (strict upper part untouched) and inv(L) (dense IB x IB, ld = IB,
// zero above the diagonal) to Winv.  On a non-positive pivot at column j writes
// *info = info_base + j + 1 (first failure only) and stops.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
potrf_diag_kernel(double* __restrict__ A, int lda, int nv, double* __restrict__ Winv,
                  int* __restrict__ info, int info_base)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    double (*L)[IB + 1] = reinterpret_cast<double (*)[IB + 1]>(smem_dyn);
    double (*X)[IB + 1] = L + IB;
    __shared__ int fail;
    const int tid = threadIdx.x;
    if (tid == 0) fail = 0;
    for (int e = tid; e < IB * IB; e += 256) {
        const int i = e % IB, j = e / IB;
        L[i][j] = (i < nv && j < nv && i >= j) ? A[i + int64_t(j) * lda] : (i == j ? 1.0 : 0.0);
    }
    __syncthreads();
    if (*info != 0) return;          // an earlier block already failed: leave the tile alone

    for (int j = 0; j < nv; ++j) {
        const double d = L[j][j];
        if (!(d > 0.0)) {            // also catches NaN
            if (tid == 0) { fail = j + 1; }
        }
        __syncthreads();
        if (fail) break;
        const double r = sqrt(d);
        __syncthreads();
        // scale column j
        for (int i = j + tid; i < nv; i += 256) L[i][j] = (i == j) ? r : L[i][j] / r;
        __syncthreads();
        // rank-1 update of the trailing lower triangle
        const int rem = nv - j - 1;
        for (int e = tid; e < rem * rem; e += 256) {
            const int i = j + 1 + e % rem, c = j + 1 + e / rem;
            if (i >= c) L[i][c] -= L[i][j] * L[c][j];
        }
        __syncthreads();
    }
    if (fail) {
        if (tid == 0 && *info == 0) *info = info_base + fail;
        return;
    }
    // write L back
    for (int e = tid; e < nv * nv; e += 256) {
        const int i = e % nv, j = e / nv;
        if (i >= j) A[i + int64_t(j) * lda] = L[i][j];
    }
    // inverse by forward substitution, one column per thread (columns >= nv: identity)
    if (tid < IB) {
        const int j = tid;
        for (int i = 0; i < IB; ++i) X[i][j] = 0.0;
        X[j][j] = 1.0 / L[j][j];
        for (int i = j + 1; i < IB; ++i) {
            double s = 0.0;
            for (int l = j; l < i; ++l) s = fma(L[i][l], X[l][j], s);
            X[i][j] = -s / L[i][i];
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
__global__ void __launch_bounds__(256)
trtri_diag_kernel(const double* __restrict__ T, int ldt, int na, int lower, int unit,
                  double* __restrict__ W)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    double (*L)[IB + 1] = reinterpret_cast<double (*)[IB + 1]>(smem_dyn);   // always handled as LOWER:
    double (*X)[IB + 1] = L + IB;                                           // upper blocks are transposed in
    const int b = blockIdx.x, tid = threadIdx.x;
    const int o = b * IB;
    const int nv = min(IB, na - o);
    for (int e = tid; e < IB * IB; e += 256) {
        const int i = e % IB, j = e / IB;
        double v = (i == j) ? 1.0 : 0.0;
        if (i < nv && j < nv) {
            if (lower) { if (i > j) v = T[o + i + int64_t(o + j) * ldt]; else if (i == j && !unit) v = T[o + i + int64_t(o + j) * ldt]; }
            else       { if (i > j) v = T[o + j + int64_t(o + i) * ldt]; else if (i == j && !unit) v = T[o + i + int64_t(o + j) * ldt]; }
        }
        L[i][j] = v;
    }
    __syncthreads();
    if (tid < IB) {
        const int j = tid;
        for (int i = 0; i < IB; ++i) X[i][j] = 0.0;
        X[j][j] = 1.0 / L[j][j];
        for (int i = j + 1; i < IB; ++i) {
            double s = 0.0;
            for (int l = j; l < i; ++l) s = fma(L[i][l], X[l][j], s);
            X[i][j] = -s / L[i][i];
        }
    }
    __syncthreads();
    double* Wb = W + int64_t(b) * IB * IB;
    for (int e = tid; e < IB * IB; e += 256) {
        const int i = e % IB, j = e / IB;
        Wb[e] = lower ? X[i][j] : X[j][i];      // inverse of the transpose = transpose of the inverse
    }
}

static void small_kernels_init()
{
    static thread_local bool done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (done[dev & 63]) return;
    cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMALL_SMEM));
    cudaFuncSetAttribute(trtri_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMALL_SMEM));
    done[dev & 63] = true;
}

// ---------------------------------------------------------------------------------------------
// per-device scratch for callers that pass work == NULL
// ---------------------------------------------------------------------------------------------
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

static GemmParamsD gp(int m, int n, int k, double alpha, double beta, int batch)
{
    GemmParamsD p{};
    p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta; p.batch = batch; p.tri = 0;
    return p;
}

// column-major block solve over a batch of B tiles with ONE triangular tile T (na x na):
//   right: B_t <- alpha B_t op(T)^{-1}   (B_t is m x na)
//   left : B_t <- alpha op(T)^{-1} B_t   (B_t is na x n)
int trsm_colmajor_d(bool left, bool lower, int op, bool unit, int m, int n, double alpha,
                    const double* T, int ldt, double* const* dB, int64_t offB, int ldb, int batch,
                    double* W, cudaStream_t stream)
{
    const int na = left ? m : n;
    const int nblk = int(ceil_div(na, IB));
    small_kernels_init();
    trtri_diag_kernel<<<nblk, 256, SMALL_SMEM, stream>>>(T, ldt, na, lower ? 1 : 0, unit ? 1 : 0, W);
    int st = launch_status();
    if (st) return st;
    const bool trans = (op != 'N');
    const bool eff_lower = (lower != trans);        // op(T) as a math matrix
    const int opT = trans ? 'T' : 'N';

    for (int s = 0; s < nblk; ++s) {
        // right/upper and left/lower sweep forward; right/lower and left/upper sweep backward
        const bool forward = left ? eff_lower : !eff_lower;
        const int j = forward ? s : nblk - 1 - s;
        const int jo = j * IB, jv = min(IB, na - jo);
        const int r0 = forward ? 0 : jo + jv;            // already-solved block range [r0, r1)
        const int r1 = forward ? jo : na;
        const int rk = r1 - r0;
        double scale = alpha;
        if (!left) {
            if (rk > 0) {
                // B_j <- alpha B_j - X[:, r0:r1] * M[r0:r1, j],  M = op(T)
                GemmParamsD p = gp(m, jv, rk, -1.0, alpha, batch);
                p.A = dB; p.offA = offB + int64_t(r0) * ldb; p.lda = ldb;
                p.B0 = trans ? T + jo + int64_t(r0) * ldt : T + r0 + int64_t(jo) * ldt;
                p.ldb = ldt; p.strideB = 0;
                p.C = dB; p.offC = offB + int64_t(jo) * ldb; p.ldc = ldb;
                if ((st = launch_gemm_d('N', opT, p, stream))) return st;
                scale = 1.0;
            }
            // B_j <- scale * B_j * op(Winv_j)   (in place: one CTA column covers all jv <= 64 columns)
            GemmParamsD p = gp(m, jv, jv, scale, 0.0, batch);
            p.A = dB; p.offA = offB + int64_t(jo) * ldb; p.lda = ldb;
            p.B0 = W + int64_t(j) * IB * IB; p.ldb = IB; p.strideB = 0;
            p.C = dB; p.offC = offB + int64_t(jo) * ldb; p.ldc = ldb;
            if ((st = launch_gemm_d('N', opT, p, stream))) return st;
        }
        else {
            if (rk > 0) {
                // B_j <- alpha B_j - M[j, r0:r1] * X[r0:r1, :]
                GemmParamsD p = gp(jv, n, rk, -1.0, alpha, batch);
                p.A0 = trans ? T + r0 + int64_t(jo) * ldt : T + jo + int64_t(r0) * ldt;
                p.lda = ldt; p.strideA = 0;
                p.B = dB; p.offB = offB + r0; p.ldb = ldb;
                p.C = dB; p.offC = offB + jo; p.ldc = ldb;
                if ((st = launch_gemm_d(opT, 'N', p, stream))) return st;
                scale = 1.0;
            }
            // B_j <- scale * op(Winv_j) * B_j   (in place: one CTA row covers all jv <= 64 rows)
            GemmParamsD p = gp(jv, n, jv, scale, 0.0, batch);
            p.A0 = W + int64_t(j) * IB * IB; p.lda = IB; p.strideA = 0;
            p.B = dB; p.offB = offB + jo; p.ldb = ldb;
            p.C = dB; p.offC = offB + jo; p.ldc = ldb;
            if ((st = launch_gemm_d(opT, 'N', p, stream))) return st;
        }
    }
    return SB200_OK;
}

// lower Cholesky of one n x n tile, blocked by IB; W >= 2*IB*IB doubles + 1 pointer slot
int potrf_tile_lower_d(int n, double* A, int lda, int* dinfo, int info_base, double* W, cudaStream_t stream)
{
    int st;
    small_kernels_init();
    for (int jo = 0; jo < n; jo += IB) {
        const int jv = min(IB, n - jo);
        double* Ajj = A + jo + int64_t(jo) * lda;
        potrf_diag_kernel<<<1, 256, SMALL_SMEM, stream>>>(Ajj, lda, jv, W, dinfo, info_base + jo);
        if ((st = launch_status())) return st;
        const int rest = n - jo - jv;
        if (rest <= 0) break;
        double* Pnl = A + (jo + jv) + int64_t(jo) * lda;          // rest x jv block below the diagonal
        // panel <- panel * inv(L_jj)^T   (in place)
        GemmParamsD p = gp(rest, jv, jv, 1.0, 0.0, 1);
        p.A0 = Pnl; p.lda = lda; p.B0 = W; p.ldb = IB; p.C0 = Pnl; p.ldc = lda;
        if ((st = launch_gemm_d('N', 'T', p, stream))) return st;
        // trailing lower triangle -= panel panel^T
        GemmParamsD q = gp(rest, rest, jv, -1.0, 1.0, 1);
        q.A0 = Pnl; q.lda = lda; q.B0 = Pnl; q.ldb = lda;
        q.C0 = A + (jo + jv) + int64_t(jo + jv) * lda; q.ldc = lda; q.tri = 1;
        if ((st = launch_gemm_d('N', 'T', q, stream))) return st;
    }
    return SB200_OK;
}

} // namespace sb200

using namespace sb200;

extern "C" {

size_t sb200_trsm_work_bytes_d(int side, int64_t m, int64_t n)
{
    const int64_t na = (side == 'L') ? m : n;
    return size_t(ceil_div(na > 0 ? na : 1, IB)) * IB * IB * sizeof(double);
}

int sb200_trsm_batched_d(int layout, int side, int uplo, int op, int diag,
                         int64_t m, int64_t n, double alpha,
                         const double* dA, int64_t lda,
                         double* const* dB, int64_t ldb,
                         int64_t batch, void* work, sb200_stream_t stream)
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
    double* W = static_cast<double*>(work);
    if (! W) {
        W = static_cast<double*>(device_scratch(sb200_trsm_work_bytes_d(left ? 'L' : 'R', m, n)));
        if (! W) return SB200_ENOMEM;
    }
    return trsm_colmajor_d(left, lower, op, diag == 'U', int(m), int(n), alpha, dA, int(lda),
                           dB, 0, int(ldb), int(batch), W, cudaStream_t(stream));
}

size_t sb200_potrf_work_bytes_d(int64_t n) { (void) n; return size_t(IB) * IB * sizeof(double); }

int sb200_potrf_tile_d(int uplo, int64_t n, double* dA, int64_t lda,
                       int* dinfo, void* work, sb200_stream_t stream)
{
    if (! valid_uplo(uplo) || n < 0 || lda < (n > 1 ? n : 1) || n > 0x7fffffff || lda > 0x7fffffff)
        return SB200_EINVAL;
    if (uplo != 'L') return SB200_ENOTSUP;    // SLATE's potrf driver works on the lower triangle (src/potrf.cc:230-240)
    cudaError_t e = cudaMemsetAsync(dinfo, 0, sizeof(int), cudaStream_t(stream));
    if (e != cudaSuccess) return int(e);
    if (n == 0) return SB200_OK;
    double* W = static_cast<double*>(work);
    if (! W) {
        W = static_cast<double*>(device_scratch(sb200_potrf_work_bytes_d(n)));
        if (! W) return SB200_ENOMEM;
    }
    return potrf_tile_lower_d(int(n), dA, int(lda), dinfo, 0, W, cudaStream_t(stream));
}

} // extern "C"
