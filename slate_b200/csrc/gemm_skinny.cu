// gemm_skinny.cu -- batched tile GEMM with a SKINNY right operand: C_t = alpha op(A_t) B_t + beta C_t with
// n <= 16 columns (the solve path: nrhs ~ 10 right-hand sides; reference call sites internal::gemm inside
// work::trsm, src/work/work_trsm.cc:150-230, hemm / gemmA for the residual of the mixed solvers,
// src/posv_mixed.cc:203-209, src/gesv_mixed.cc:201-206, src/gemm.cc:18-21).
//
// With n = 10 a tensor-core tile kernel computes a 64-wide N block and throws 84 % of it away; this
// path is HBM-bound instead: every A tile is streamed exactly once (algorithmic bytes = m k s per tile),
// B (k x n) sits in shared memory, C is m x n.
//   op(A) = A      : 64 rows x 4 k-slices per CTA, lanes (row, slice) so that A is read in full 32-byte
//                    sectors down its columns; the 4 slices are combined with two shuffles
//   op(A) = A^T/A^H: one warp per output row = one stored COLUMN of A, lanes along k (contiguous), 16
//                    accumulators per lane combined by shuffle reduction
#include "gemm_dmma.cuh"
#include "scalar_ops.cuh"
#include <cstdlib>

namespace sb200 {

constexpr int SK_NC = SKINNY_MAX_N; // max columns of B / C (16)
constexpr int SK_KCH = 128;        // k chunk of B staged in shared memory
constexpr int SK_ROWS = 64;        // rows of C per CTA, op(A) = A
constexpr int SK_ROWS_T = 16;      // rows of C per CTA, op(A) = A^T / A^H (2 rows per warp: 32 accumulators per lane)
constexpr int SK_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(SK_THREADS)
gemm_skinny_kernel(const GemmParamsT<T> p, int opA)
{
    __shared__ T Bs[SK_NC][SK_KCH];          // B chunk, transposed: Bs[c][kk]
    const int t = blockIdx.y;
    const int i0 = blockIdx.x * (opA == 'N' ? SK_ROWS : SK_ROWS_T);
    const T* __restrict__ A = (p.A ? p.A[t] : p.A0 + int64_t(t) * p.strideA) + p.offA;
    const T* __restrict__ B = (p.B ? p.B[t] : p.B0 + int64_t(t) * p.strideB) + p.offB;
    T* __restrict__ C = (p.C ? p.C[t] : p.C0 + int64_t(t) * p.strideC) + p.offC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = p.n;

    T acc[SK_NC];
    #pragma unroll
    for (int c = 0; c < SK_NC; ++c) acc[c] = zero_of<T>();

    if (opA == 'N') {
        const int r = tid >> 2, ks = tid & 3;             // row within the CTA, k-slice
        const int row = i0 + r;
        const bool live = row < p.m;
        for (int k0 = 0; k0 < p.k; k0 += SK_KCH) {
            const int kc = min(SK_KCH, p.k - k0);
            __syncthreads();
            for (int e = tid; e < SK_NC * SK_KCH; e += SK_THREADS) {
                const int c = e / SK_KCH, kk = e - c * SK_KCH;
                Bs[c][kk] = (c < n && kk < kc) ? B[(k0 + kk) + int64_t(c) * p.ldb] : zero_of<T>();
            }
            __syncthreads();
            if (live) {
                const T* __restrict__ arow = A + row + int64_t(k0) * p.lda;
                for (int kk = ks; kk < kc; kk += 4 * 8) {          // 8 independent loads per round trip
                    T a[8];
                    #pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = (kk + 4 * u < kc) ? arow[int64_t(kk + 4 * u) * p.lda] : zero_of<T>();
                    #pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (kk + 4 * u >= kc) continue;
                        #pragma unroll
                        for (int c = 0; c < SK_NC; ++c) fma_acc(acc[c], a[u], Bs[c][kk + 4 * u]);
                    }
                }
            }
        }
        #pragma unroll
        for (int c = 0; c < SK_NC; ++c) {
            acc[c] = add(acc[c], shfl_xor_t(acc[c], 1));
            acc[c] = add(acc[c], shfl_xor_t(acc[c], 2));
        }
        if (live && ks == 0) {
            const bool use_beta = ! is_zero(p.beta);
            #pragma unroll
            for (int c = 0; c < SK_NC; ++c) {
                if (c >= n) break;
                T v = mul(p.alpha, acc[c]);
                if (use_beta) v = add(v, mul(p.beta, C[row + int64_t(c) * p.ldc]));
                C[row + int64_t(c) * p.ldc] = v;
            }
        }
        return;
    }

    // op(A) = A^T or A^H: output row i = stored column i of A (k contiguous)
    const bool cj = (opA == 'C');
    constexpr int RPW = SK_ROWS_T / (SK_THREADS / 32);    // 2 rows per warp
    T res[RPW][SK_NC];
    #pragma unroll
    for (int j = 0; j < RPW; ++j)
        #pragma unroll
        for (int c = 0; c < SK_NC; ++c) res[j][c] = zero_of<T>();
    for (int k0 = 0; k0 < p.k; k0 += SK_KCH) {
        const int kc = min(SK_KCH, p.k - k0);
        __syncthreads();
        for (int e = tid; e < SK_NC * SK_KCH; e += SK_THREADS) {
            const int c = e / SK_KCH, kk = e - c * SK_KCH;
            Bs[c][kk] = (c < n && kk < kc) ? B[(k0 + kk) + int64_t(c) * p.ldb] : zero_of<T>();
        }
        __syncthreads();
        #pragma unroll
        for (int j = 0; j < RPW; ++j) {
            const int row = i0 + warp + (SK_THREADS / 32) * j;
            if (row >= p.m) continue;                                     // warp-uniform
            const T* __restrict__ acol = A + k0 + int64_t(row) * p.lda;
            T a[SK_KCH / 32];                                     // the whole chunk of this row: 4 loads in flight
            #pragma unroll
            for (int u = 0; u < SK_KCH / 32; ++u) a[u] = (lane + 32 * u < kc) ? acol[lane + 32 * u] : zero_of<T>();
            #pragma unroll
            for (int u = 0; u < SK_KCH / 32; ++u) {
                if (lane + 32 * u >= kc) continue;
                const T av = cj ? conj_(a[u]) : a[u];
                #pragma unroll
                for (int c = 0; c < SK_NC; ++c) fma_acc(res[j][c], av, Bs[c][lane + 32 * u]);
            }
        }
    }
    const bool use_beta = ! is_zero(p.beta);
    #pragma unroll
    for (int j = 0; j < RPW; ++j) {
        const int row = i0 + warp + (SK_THREADS / 32) * j;
        if (row >= p.m) continue;
        #pragma unroll
        for (int c = 0; c < SK_NC; ++c) {
            T v = res[j][c];
            #pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) v = add(v, shfl_xor_t(v, sh));
            if (lane == 0 && c < n) {
                T w = mul(p.alpha, v);
                if (use_beta) w = add(w, mul(p.beta, C[row + int64_t(c) * p.ldc]));
                C[row + int64_t(c) * p.ldc] = w;
            }
        }
    }
}

static bool skinny_enabled()
{
    static const bool on = [] { const char* e = getenv("SB200_SKINNY"); return ! (e && atoi(e) == 0); }();
    return on;
}

// true when the skinny kernel serves this problem (n <= 16, B not transposed, no triangle mask)
template <typename T>
bool gemm_skinny_applies(int opB, const GemmParamsT<T>& p)
{
    return skinny_enabled() && opB == 'N' && p.n >= 1 && p.n <= SK_NC && p.tri == 0 && p.k >= 1;
}

template <typename T>
int launch_gemm_skinny(int opA, const GemmParamsT<T>& p, cudaStream_t stream)
{
    if (p.m <= 0 || p.n <= 0 || p.batch <= 0) return SB200_OK;
    for (int b0 = 0; b0 < p.batch; b0 += 65535) {
        GemmParamsT<T> q = p;
        q.batch = std::min(65535, p.batch - b0);
        if (q.A) q.A += b0; else q.A0 += int64_t(b0) * q.strideA;
        if (q.B) q.B += b0; else q.B0 += int64_t(b0) * q.strideB;
        if (q.C) q.C += b0; else q.C0 += int64_t(b0) * q.strideC;
        const int rows = (opA == 'N') ? SK_ROWS : SK_ROWS_T;
        gemm_skinny_kernel<T><<<dim3(unsigned(ceil_div(p.m, rows)), unsigned(q.batch)), SK_THREADS, 0, stream>>>(q, opA);
        const int st = launch_status();
        if (st) return st;
    }
    return SB200_OK;
}

#define SB200_INST_SKINNY(T) \
    template bool gemm_skinny_applies<T>(int, const GemmParamsT<T>&); \
    template int launch_gemm_skinny<T>(int, const GemmParamsT<T>&, cudaStream_t);
SB200_INST_SKINNY(float)
SB200_INST_SKINNY(double)
SB200_INST_SKINNY(cuFloatComplex)
SB200_INST_SKINNY(cuDoubleComplex)

} // namespace sb200
