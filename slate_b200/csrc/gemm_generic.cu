// gemm_generic.cu -- type-generic batched tile GEMM / HERK for float, complex<float> and
// complex<double> (same boundary as the FP64 DMMA kernel: blas::batch::gemm / herk).
//
// Scope note: FP64 real is the headline path and runs on the tensor cores (gemm_dmma.cuh).
// This kernel is the correct-first SIMT path for the other three types: shared-memory tiles
// 64 x 64 x 16, 256 threads, 4 x 4 register micro-tiles, conflict-free strided mapping.
// FP32 is what the reference's gesv_mixed factors in (src/gesv_mixed.cc:106-300); complex
// double serves zgemm / zherk.  (tcgen05 / complex-DMMA versions are listed as next steps in
// DESIGN.md.)
#include "gemm_dmma.cuh"
#include "scalar_ops.cuh"
#include <cstdlib>

namespace sb200 {

constexpr int GBM = 64, GBN = 64, GBK = 16;

template <typename T>
__device__ inline T load_op(const T* __restrict__ X, int64_t ld, int op, int r, int c)
{
    // element (r, c) of op(X)
    if (op == 'N') return X[r + int64_t(c) * ld];
    const T v = X[c + int64_t(r) * ld];
    return op == 'C' ? conj_(v) : v;
}

template <typename T>
__global__ void __launch_bounds__(256)
gemm_generic_kernel(const GemmParamsT<T> p, int opA, int opB)
{
    __shared__ T As[GBK][GBM + 1];
    __shared__ T Bs[GBK][GBN + 1];
    const int tiles_m = (p.m + GBM - 1) / GBM, tiles_n = (p.n + GBN - 1) / GBN;
    const int per = tiles_m * tiles_n;
    const int t = blockIdx.x / per, r = blockIdx.x - t * per;
    const int m0 = (r % tiles_m) * GBM, n0 = (r / tiles_m) * GBN;
    if (p.tri == 1 && n0 >= m0 + GBM) return;
    if (p.tri == 2 && m0 >= n0 + GBN) return;
    const T* __restrict__ A = (p.A ? p.A[t] : p.A0 + int64_t(t) * p.strideA) + p.offA;
    const T* __restrict__ B = (p.B ? p.B[t] : p.B0 + int64_t(t) * p.strideB) + p.offB;
    T* __restrict__ C = (p.C ? p.C[t] : p.C0 + int64_t(t) * p.strideC) + p.offC;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    T acc[4][4];
    #pragma unroll
    for (int i = 0; i < 4; ++i)
        #pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = zero_of<T>();

    for (int k0 = 0; k0 < p.k; k0 += GBK) {
        // stage op(A)(m0.., k0..) and op(B)(k0.., n0..); the fastest-varying thread index follows
        // the contiguous direction of the stored operand
        for (int e = tid; e < GBM * GBK; e += 256) {
            int i, l;
            if (opA == 'N') { i = e % GBM; l = e / GBM; } else { l = e % GBK; i = e / GBK; }
            As[l][i] = (m0 + i < p.m && k0 + l < p.k) ? load_op(A, p.lda, opA, m0 + i, k0 + l) : zero_of<T>();
        }
        for (int e = tid; e < GBN * GBK; e += 256) {
            int j, l;
            if (opB == 'N') { l = e % GBK; j = e / GBK; } else { j = e % GBN; l = e / GBN; }
            Bs[l][j] = (n0 + j < p.n && k0 + l < p.k) ? load_op(B, p.ldb, opB, k0 + l, n0 + j) : zero_of<T>();
        }
        __syncthreads();
        #pragma unroll
        for (int l = 0; l < GBK; ++l) {
            T a[4], b[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[l][tx + 16 * i];
            #pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[l][ty + 16 * j];
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) fma_acc(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }

    const bool use_beta = ! is_zero(p.beta);
    T cv[4][4];
    #pragma unroll
    for (int j = 0; j < 4; ++j)
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = m0 + tx + 16 * i, col = n0 + ty + 16 * j;
            bool ok = row < p.m && col < p.n;
            if (p.tri == 1) ok = ok && row >= col;
            if (p.tri == 2) ok = ok && row <= col;
            cv[i][j] = (ok && use_beta) ? C[row + int64_t(col) * p.ldc] : zero_of<T>();
        }
    #pragma unroll
    for (int j = 0; j < 4; ++j)
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = m0 + tx + 16 * i, col = n0 + ty + 16 * j;
            bool ok = row < p.m && col < p.n;
            if (p.tri == 1) ok = ok && row >= col;
            if (p.tri == 2) ok = ok && row <= col;
            if (ok) {
                T v = mul(p.alpha, acc[i][j]);
                if (use_beta) v = add(v, mul(p.beta, cv[i][j]));
                if (p.herk && row == col) v = real_part_only(v);
                C[row + int64_t(col) * p.ldc] = v;
            }
        }
}

template <typename T>
static int launch_generic(int opA, int opB, const GemmParamsT<T>& p, cudaStream_t stream)
{
    if (p.m <= 0 || p.n <= 0 || p.batch <= 0) return SB200_OK;
    const int64_t grid = ceil_div(p.m, GBM) * ceil_div(p.n, GBN) * int64_t(p.batch);
    if (grid > 0x7fffffffLL) return SB200_EINVAL;
    gemm_generic_kernel<T><<<unsigned(grid), 256, 0, stream>>>(p, opA, opB);
    return launch_status();
}

int launch_gemm_d(int opA, int opB, GemmParamsD p, cudaStream_t stream);

template <> int launch_gemm<double>(int opA, int opB, GemmParamsT<double> p, cudaStream_t s) { return launch_gemm_d(opA, opB, p, s); }
template <> int launch_gemm<float>(int opA, int opB, GemmParamsT<float> p, cudaStream_t s) { return launch_generic(opA, opB, p, s); }
template <> int launch_gemm<cuFloatComplex>(int opA, int opB, GemmParamsT<cuFloatComplex> p, cudaStream_t s) { return launch_generic(opA, opB, p, s); }
int launch_gemm_z(int opA, int opB, GemmParamsT<cuDoubleComplex> p, cudaStream_t stream);     // gemm_zdmma.cu
template <> int launch_gemm<cuDoubleComplex>(int opA, int opB, GemmParamsT<cuDoubleComplex> p, cudaStream_t s)
{
    // complex128 runs on the FP64 tensor cores; SB200_ZGEMM_SIMT=1 selects the SIMT kernel (A/B debugging only)
    static const bool simt = [] { const char* e = getenv("SB200_ZGEMM_SIMT"); return e && atoi(e) != 0; }();
    return simt ? launch_generic(opA, opB, p, s) : launch_gemm_z(opA, opB, p, s);
}

} // namespace sb200
