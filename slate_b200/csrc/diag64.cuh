// diag64.cuh -- 64 x 64 Cholesky and triangular inverse held in SHARED memory and worked on by the WHOLE CTA
// (rolled loops, a few hundred instructions), for the diagonal-block steps of the tile factor / solve kernels.
//
// Why (profiles/r01e_launches_potrf_n2048_per_grid.txt + cuobjdump): the round-1 kernels keep one matrix row /
// column per thread in registers (64 threads, everything unrolled): potrf_diag_fast_kernel<double> is 14 520 SASS
// instructions and takes 109 us, trtri_diag_fast_kernel<double> 4 272 instructions and 41.5 us -- 15-19 cycles per
// instruction in both, i.e. two warps of straight-line code with nothing to hide shared-memory / instruction-fetch
// latency behind; the arithmetic itself (2 016 FMAs per thread) is ~2 us.  Here the same O(64^3) work is spread over
// all threads of the CTA with two barriers per column, so every latency is overlapped by 4-8 warps.
//
// Used by the one-launch tile Cholesky (potrf_tile_fused.cu), which the potrf driver takes when the chain runs on its
// own SM partition.  As stand-alone replacements of the register kernels they lost (665 us against 555 us per nb = 512
// tile, profiles/r02a_bench_tile_chain_kernels.jsonl) and that use was deleted.
//
// Layout: column-major with leading dimension LD (padded), element (r, c) at M[c * LD + r].
#pragma once
#include "common.cuh"

namespace sb200 {

__device__ __forceinline__ double rsqrt_of(double x) { return rsqrt(x); }
__device__ __forceinline__ float  rsqrt_of(float x)  { return rsqrtf(x); }

// Right-looking Cholesky of the lower triangle of As (destroyed); L goes to Ls (entries above the diagonal are
// NOT written), rd[j] = 1 / L(j,j).  All NT threads must call it; contains barriers.  Returns the first
// non-positive (or NaN) pivot column + 1, or 0 -- the same value in every thread.
//   RSQ = false:  L(j,j) = sqrt(d), column scaled by the reciprocal 1 / L(j,j)     (LAPACK potf2's operations)
//   RSQ = true :  one rsqrt per column, L(j,j) = d * rsqrt(d)                      (<= 2-3 ulp from the above)
template <typename R, int NT, int LD, bool RSQ>
__device__ __forceinline__ int chol64_smem(R* __restrict__ As, R* __restrict__ Ls, R* __restrict__ rd, int tid)
{
    static_assert(NT % 16 == 0 && NT >= 64, "thread layout: 16 threads along the rows");
    constexpr int TC = NT / 16;
    const int tr = tid & 15, tc = tid >> 4;
    int fail = 0;
    for (int j = 0; j < 64; ++j) {
        const R d = As[j * LD + j];
        if (fail == 0 && !(d > R(0))) fail = j + 1;
        R diag, rinv;
        if constexpr (RSQ) { rinv = rsqrt_of(d); diag = d * rinv; }
        else               { diag = sqrt(d); rinv = R(1) / diag; }
        for (int r = j + tid; r < 64; r += NT) Ls[j * LD + r] = (r == j) ? diag : As[j * LD + r] * rinv;
        if (tid == 0) rd[j] = rinv;
        __syncthreads();
        // A(r, c) -= L(r, j) L(c, j),  j < c <= r
        for (int c = j + 1 + tc; c < 64; c += TC) {
            const R lc = Ls[j * LD + c];
            for (int r = c + tr; r < 64; r += 16) As[c * LD + r] = fma(-Ls[j * LD + r], lc, As[c * LD + r]);
        }
        __syncthreads();
    }
    return fail;
}

// X = L^{-1} for the lower-triangular L in Ls (entries above the diagonal are never read; the diagonal enters only
// through rd[i] = 1 / L(i,i), so a unit triangle is rd = 1).  Xs receives the FULL 64 x 64 inverse (exact zeros above the
// diagonal).  Forward substitution in axpy form on all 64 columns at once.  All NT threads; contains barriers.
template <typename R, int NT, int LD>
__device__ __forceinline__ void inv64_smem(const R* __restrict__ Ls, const R* __restrict__ rd, R* __restrict__ Xs, int tid)
{
    static_assert(NT % 16 == 0 && NT >= 64, "thread layout: 16 threads along the rows");
    constexpr int TC = NT / 16;
    const int tr = tid & 15, tc = tid >> 4;
    for (int e = tid; e < 64 * 64; e += NT) {
        const int r = e & 63, c = e >> 6;
        Xs[c * LD + r] = (r == c) ? R(1) : R(0);
    }
    __syncthreads();
    for (int i = 0; i < 64; ++i) {
        if (tid <= i) Xs[tid * LD + i] *= rd[i];                 // row i is final: X(i, j), j <= i
        __syncthreads();
        for (int j = tc; j <= i; j += TC) {                      // rows below: X(r, j) -= L(r, i) X(i, j)
            const R xij = Xs[j * LD + i];
            for (int r = i + 1 + tr; r < 64; r += 16) Xs[j * LD + r] = fma(-Ls[i * LD + r], xij, Xs[j * LD + r]);
        }
        __syncthreads();
    }
}

} // namespace sb200
