// potrf_tile_fused.cu -- the lower Cholesky of ONE diagonal tile (n <= 1024) in ONE launch.
// Taken by the potrf driver when the chain runs on its own SM partition (sm_partition.cu: 0.69 ms per nb = 512 tile
// against 1.1 ms for the launch chain of factor_small.cu there; on an idle whole device the launch chain wins, 0.56
// against 0.66 ms), or with SB200_TILE_FUSED=1 | 2.  Validated on B200 in round 2 (tests/test_zzz_gpu_round2_candidates.py).
//
// Why (profiles/r01e_launches_potrf_n2048_per_grid.txt): the diagonal tile of every potrf
// step (reference: internal::potrf<Devices> -> cusolverDn?potrf, src/internal/internal_potrf.cc:57-81) is
// 8 x (64-block Cholesky kernel + 2 small GEMM launches) = 24 dependent launches, 1.3 ms uncontended and 1.9 ms
// next to the trailing update; at 8 GPUs that chain (128 tiles at n = 65536) is longer than the trailing update.
//
// Design: one CTA per 64-row block, LEFT-looking, rows pipelined through release/acquire flags in global memory
// (no grid barrier, no cooperative launch: a CTA only ever waits for CTAs with a smaller index):
//   CTA r, for b = 0 .. r-1:   S = A(r,b) - sum_{c<b} L(r,c) L(b,c)^T     needs row b complete   (rowcnt[b] >= b)
//                              L(r,b) = S inv(L(b,b))^T                    needs W_b = inv(L(b,b)) (diagf[b])
//                              publish rowcnt[r] = b + 1
//   then                       D = A(r,r) - sum_{c<r} L(r,c) L(r,c)^T ;  L(r,r) = chol(D) ;  W_r = inv(L(r,r)) ;
//                              publish diagf[r]
// The only work between "W_{r-1} published" and "W_r published" is one 64^3 product for L(r,r-1), one for its
// contribution to D, the 64 x 64 Cholesky and the inverse: everything else of row r (S for b = r-1 and the part of D
// that does not involve L(r,r-1)) is computed while CTA r-1 factors its diagonal block.
// FP64 products run on the FP64 tensor-core MMA (DMMA.8x8x4, fragment conventions of gemm_dmma.cuh), FP32 products
// (the FP32 factorisation of posv_mixed) on FP32 FMAs with a 4 x 8 register tile per thread; operands are staged
// through padded shared memory; the 64 x 64 Cholesky + inverse run in shared memory on all threads (diag64.cuh).
// Failure (block not positive definite): info = info_base + column + 1 as the default path; the failing CTA
// publishes FAILED on its flags and every later CTA leaves on seeing it.
//
// The same product micro-kernels serve the other opt-in one-launch kernels further down (SB200_TRSM_FUSED):
// trsm_rlt_fused_kernel (Cholesky panel solve), trsm_lln_fused_kernel (LU row solve), trsm_lln_small_kernel (small
// triangle, direct substitution).  CPU transcriptions of all of them: scratch/emulate_fused.py.
#include "common.cuh"
#include "diag64.cuh"
#include "runtime_internal.hh"
#include <cstdlib>

namespace sb200 {

namespace {

constexpr int FB = 64;              // block size (== IB of factor_small.cu)
constexpr int FLD = FB + 4;         // padded shared leading dimension: conflict-free DMMA fragment loads
constexpr int FT = 128;             // threads per CTA (4 warps).  128 x <= 255 registers and
                                    // 104 KB of shared memory leave room for ONE trailing-update CTA on the same SM, so a
                                    // tile CTA needs a single CTA slot to free up, not a whole SM
constexpr int FMAXB = 16;           // at most 16 row blocks (n <= 1024)
constexpr unsigned F_FAILED = 0x40000000u;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// all FT threads: wait until *p >= target (FAILED is larger than any target); returns the value seen
__device__ __forceinline__ unsigned wait_flag(const unsigned* p, unsigned target, unsigned* s_v)
{
    if (threadIdx.x == 0) {
        unsigned v;
        const long long t0 = clock64();
        while ((v = ld_acquire_u32(p)) < target) { __nanosleep(40); spin_watchdog(t0); }
        *s_v = v;
    }
    __syncthreads();
    const unsigned v = *s_v;
    __syncthreads();
    return v;
}

// dst[k * FLD + i] = src[i + k * lds] for i < rv (rows beyond rv are zero-filled), 64 x 64, L2 loads
template <typename R>
__device__ __forceinline__ void load_block(R* __restrict__ dst, const R* src, int lds, int rv)
{
    const int i = threadIdx.x & (FB - 1), k0 = threadIdx.x >> 6;          // k0 in {0, 1}
    #pragma unroll
    for (int half = 0; half < 2; ++half) {                                 // 2 x 16 loads in flight per thread
        R v[FB / 4];
        #pragma unroll
        for (int t = 0; t < FB / 4; ++t) {
            const int k = k0 + 2 * (half * (FB / 4) + t);
            v[t] = (i < rv) ? __ldcg(src + i + int64_t(k) * lds) : R(0);
        }
        #pragma unroll
        for (int t = 0; t < FB / 4; ++t) dst[(k0 + 2 * (half * (FB / 4) + t)) * FLD + i] = v[t];
    }
}

// The 64 x 64 x 64 product micro-kernel of a 128-thread CTA: acc(row, col) += sum_k X(row, k) Y(col, k) with
// Xs[k * FLD + row], Ys[k * FLD + col]; 32 accumulators per thread, element e of thread `tid` is (row(e), col(e)).
template <typename R> struct Prod;

template <> struct Prod<double> {          // DMMA.8x8x4: 4 warps, 2 x 2 warp tiles of 32 x 32 (4 x 4 MMA tiles)
    int wm, wn, lr, lc;
    __device__ __forceinline__ explicit Prod(int tid)
    {
        const int warp = tid >> 5, lane = tid & 31;
        lr = lane >> 2; lc = lane & 3; wm = (warp & 1) * 32; wn = (warp >> 1) * 32;
    }
    // e = (i * 4 + j) * 2 + h
    __device__ __forceinline__ int row(int e) const { return wm + (e >> 3) * 8 + lr; }
    __device__ __forceinline__ int col(int e) const { return wn + ((e >> 1) & 3) * 8 + 2 * lc + (e & 1); }
    __device__ __forceinline__ void mma(double (&acc)[32], const double* __restrict__ Xs, const double* __restrict__ Ys) const
    {
        const double* cA = Xs + lc * FLD + wm + lr;
        const double* cB = Ys + lc * FLD + wn + lr;
        #pragma unroll 4
        for (int k4 = 0; k4 < FB / 4; ++k4) {
            double a[4], b[4];
            #pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = cA[k4 * 4 * FLD + i * 8];
            #pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = cB[k4 * 4 * FLD + j * 8];
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[(i * 4 + j) * 2], acc[(i * 4 + j) * 2 + 1], a[i], b[j]);
        }
    }
};

template <> struct Prod<float> {           // FP32 FMA: thread (tx, ty) of 16 x 8 owns rows 4 tx .. +3, columns 8 ty .. +7
    int tx, ty;
    __device__ __forceinline__ explicit Prod(int tid) { tx = tid & 15; ty = tid >> 4; }
    // e = i * 8 + j
    __device__ __forceinline__ int row(int e) const { return tx * 4 + (e >> 3); }
    __device__ __forceinline__ int col(int e) const { return ty * 8 + (e & 7); }
    __device__ __forceinline__ void mma(float (&acc)[32], const float* __restrict__ Xs, const float* __restrict__ Ys) const
    {
        const float* cA = Xs + tx * 4;
        const float* cB = Ys + ty * 8;
        #pragma unroll 8
        for (int k = 0; k < FB; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(cA + k * FLD);
            const float4 b0 = *reinterpret_cast<const float4*>(cB + k * FLD);
            const float4 b1 = *reinterpret_cast<const float4*>(cB + k * FLD + 4);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            #pragma unroll
            for (int i = 0; i < 4; ++i)
                #pragma unroll
                for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(a[i], b[j], acc[i * 8 + j]);
        }
    }
};

template <typename R>
__device__ __forceinline__ void zero_acc(R (&acc)[32])
{
    #pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = R(0);
}

template <typename R, bool RSQ>
__global__ void __launch_bounds__(FT, 1)
potrf_tile_fused_kernel(R* __restrict__ A, int lda, int n, int* __restrict__ info, int info_base,
                        R* __restrict__ Wg, unsigned* __restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    R* Xs = reinterpret_cast<R*>(smem_dyn);                // operand X / Cholesky columns + staging
    R* Ys = Xs + FB * FLD;                                 // operand Y / W_b
    R* Cs = Ys + FB * FLD;                                 // S (operand of the solve) / D (input of the Cholesky)
    __shared__ unsigned s_v;
    unsigned* rowcnt = flags;                              // rowcnt[r] = number of final blocks L(r, 0 .. cnt-1)
    unsigned* diagf = flags + FMAXB;                       // diagf[r]  = 1 once L(r,r) and W_r are final

    const int r = blockIdx.x, nblk = gridDim.x;
    const int tid = threadIdx.x;
    const Prod<R> pr(tid);
    const int rv = min(FB, n - r * FB);                    // valid rows (= columns of the diagonal block) of this row block
    R* Arow = A + r * FB;                                  // row block r, column 0

    if (*reinterpret_cast<volatile int*>(info) != 0) {     // an earlier tile / block already failed: leave the tile alone
        if (tid == 0) { st_release_u32(&rowcnt[r], F_FAILED); st_release_u32(&diagf[r], F_FAILED); }
        return;
    }

    R acc[32], accD[32];
    zero_acc(accD);

    for (int b = 0; b < r; ++b) {
        // ---- S = A(r,b) - sum_{c<b} L(r,c) L(b,c)^T
        if (b > 0) {
            const unsigned v = wait_flag(&rowcnt[b], unsigned(b), &s_v);
            if (v & F_FAILED) { if (tid == 0) { st_release_u32(&rowcnt[r], F_FAILED); st_release_u32(&diagf[r], F_FAILED); } return; }
        }
        zero_acc(acc);
        for (int c = 0; c < b; ++c) {
            load_block<R>(Xs, Arow + int64_t(c) * FB * lda, lda, rv);              // L(r,c): written by this CTA
            load_block<R>(Ys, A + b * FB + int64_t(c) * FB * lda, lda, FB);        // L(b,c): final (rowcnt[b] >= b)
            __syncthreads();
            pr.mma(acc, Xs, Ys);
            __syncthreads();
        }
        R* Ab = Arow + int64_t(b) * FB * lda;                                     // A(r,b) as given, L(r,b) on return
        #pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int row = pr.row(e), col = pr.col(e);
            const R o = (row < rv) ? __ldcg(Ab + row + int64_t(col) * lda) : R(0);
            Cs[col * FLD + row] = o - acc[e];                                     // rows >= rv: 0 - 0
        }
        if (b == r - 1) {
            // the part of D that does not need L(r,r-1), while CTA r-1 is still factoring its diagonal block
            for (int c = 0; c < b; ++c) {
                load_block<R>(Xs, Arow + int64_t(c) * FB * lda, lda, rv);
                __syncthreads();
                pr.mma(accD, Xs, Xs);
                __syncthreads();
            }
        }
        // ---- L(r,b) = S W_b^T
        {
            const unsigned v = wait_flag(&diagf[b], 1u, &s_v);
            if (v & F_FAILED) { if (tid == 0) { st_release_u32(&rowcnt[r], F_FAILED); st_release_u32(&diagf[r], F_FAILED); } return; }
        }
        load_block<R>(Ys, Wg + int64_t(b) * FB * FB, FB, FB);
        __syncthreads();                                                           // Cs and Ys complete
        zero_acc(acc);
        pr.mma(acc, Cs, Ys);
        #pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int row = pr.row(e), col = pr.col(e);
            if (row < rv) Ab[row + int64_t(col) * lda] = acc[e];
            if (b == r - 1) Xs[col * FLD + row] = acc[e];                          // operand of the last update of D
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_u32(&rowcnt[r], unsigned(b + 1));
        if (b == r - 1) pr.mma(accD, Xs, Xs);
    }

    // ---- D = A(r,r) - sum_c L(r,c) L(r,c)^T, identity-padded to 64 x 64 (only the lower triangle is used)
    {
        const R* Ad = Arow + int64_t(r) * FB * lda;
        __syncthreads();                                   // everybody is done with Xs as an operand; Cs free
        #pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int row = pr.row(e), col = pr.col(e);
            R v;
            if (row < rv && col < rv) v = (col <= row) ? __ldcg(Ad + row + int64_t(col) * lda) - accD[e] : R(0);
            else                      v = (row == col) ? R(1) : R(0);
            Cs[col * FLD + row] = v;
        }
        __syncthreads();
    }

    // ---- 64 x 64 Cholesky + inverse in shared memory, all threads (diag64.cuh): D is in Cs, L goes to Xs, W_r to Ys
    R* rd = Cs + FB * FLD;                                 // [FB] reciprocal diagonal (extra FB elements of shared memory)
    const int fail = chol64_smem<R, FT, FLD, RSQ>(Cs, Xs, rd, tid);
    if (fail) {
        if (tid == 0) {
            if (*reinterpret_cast<volatile int*>(info) == 0) *info = info_base + r * FB + fail;
            st_release_u32(&rowcnt[r], F_FAILED);
            st_release_u32(&diagf[r], F_FAILED);
        }
        return;
    }
    if (r + 1 < nblk) {                                    // W_r first: the rows below are waiting for it
        inv64_smem<R, FT, FLD>(Xs, rd, Ys, tid);
        R* W = Wg + int64_t(r) * FB * FB;
        for (int e = tid; e < FB * FB; e += FT) W[e] = Ys[(e >> 6) * FLD + (e & (FB - 1))];       // W(i, j) at i + j * FB
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_u32(&diagf[r], 1u);
    }
    {
        R* Ad = Arow + int64_t(r) * FB * lda;              // L(r,r) itself is only read after the kernel
        for (int e = tid; e < FB * FB; e += FT) {
            const int row = e & (FB - 1), col = e >> 6;
            if (col <= row && row < rv) Ad[row + int64_t(col) * lda] = Xs[col * FLD + row];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Panel solve of the Cholesky step in ONE launch (SB200_TRSM_FUSED bit 0, on by default):
//   B_t <- alpha B_t L^{-T}      (Right, Lower, Trans, NonUnit;  reference: internal::trsm<Devices>,
//                                 src/internal/internal_trsm.cc:132-262 -> cublas?trsmBatched)
// The rows of B are independent, so one CTA takes 64 rows of one B tile through the whole block substitution
//   for j = 0 .. nblk-1:   X_j = (alpha B_j - sum_{c<j} X_c L(j,c)^T) W_j^T,      W_j = inv(L(j,j)) from trtri_diag
// with the product micro-kernel of the tile Cholesky above -- the default path (trsm_colmajor) is the same
// arithmetic as 2 launches per 64-column block (16 dependent launches for nb = 512).
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(FT, 1)
trsm_rlt_fused_kernel(int m, int na, R alpha, const R* __restrict__ Tm, int ldt, const R* __restrict__ W,
                      R* const* __restrict__ dB, int64_t offB, int ldb)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    R* Xs = reinterpret_cast<R*>(smem_dyn);
    R* Ys = Xs + FB * FLD;
    R* Cs = Ys + FB * FLD;
    const int tid = threadIdx.x;
    const Prod<R> pr(tid);
    const int r0 = blockIdx.x * FB;                        // first row of this CTA inside the tile
    const int rv = min(FB, m - r0);
    if (rv <= 0) return;
    R* Brow = dB[blockIdx.y] + offB + r0;
    const int nblk = (na + FB - 1) / FB;
    R acc[32];
    for (int j = 0; j < nblk; ++j) {
        const int jv = min(FB, na - j * FB);
        zero_acc(acc);
        for (int c = 0; c < j; ++c) {
            load_block<R>(Xs, Brow + int64_t(c) * FB * ldb, ldb, rv);              // X_c: written by this CTA
            load_block<R>(Ys, Tm + j * FB + int64_t(c) * FB * ldt, ldt, jv);       // L(j,c), rows >= jv zero
            __syncthreads();
            pr.mma(acc, Xs, Ys);
            __syncthreads();
        }
        R* Bj = Brow + int64_t(j) * FB * ldb;
        #pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int row = pr.row(e), col = pr.col(e);
            const R o = (row < rv && col < jv) ? __ldcg(Bj + row + int64_t(col) * ldb) : R(0);
            Cs[col * FLD + row] = alpha * o - acc[e];
        }
        load_block<R>(Ys, W + int64_t(j) * FB * FB, FB, FB);                       // W_j (identity-padded by trtri_diag)
        __syncthreads();
        zero_acc(acc);
        pr.mma(acc, Cs, Ys);
        #pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int row = pr.row(e), col = pr.col(e);
            if (row < rv && col < jv) Bj[row + int64_t(col) * ldb] = acc[e];
        }
        __syncthreads();                                   // X_j visible to the whole CTA (global), Cs / Ys free
    }
}

// dst[k * FLD + j] = src[k + j * lds] for j < cv (columns beyond cv are zero-filled): the 64 x 64 block is stored
// transposed, so that an operand which is contiguous along k in memory gets the [k][column] layout the product wants
template <typename R>
__device__ __forceinline__ void load_block_t(R* __restrict__ dst, const R* src, int lds, int cv)
{
    const int k = threadIdx.x & (FB - 1), j0 = threadIdx.x >> 6;          // j0 in {0, 1}
    #pragma unroll
    for (int half = 0; half < 2; ++half) {
        R v[FB / 4];
        #pragma unroll
        for (int t = 0; t < FB / 4; ++t) {
            const int j = j0 + 2 * (half * (FB / 4) + t);
            v[t] = (j < cv) ? __ldcg(src + k + int64_t(j) * lds) : R(0);
        }
        #pragma unroll
        for (int t = 0; t < FB / 4; ++t) dst[k * FLD + j0 + 2 * (half * (FB / 4) + t)] = v[t];
    }
}

// ---------------------------------------------------------------------------------------------
// Row solve of the LU step in ONE launch (SB200_TRSM_FUSED bit 1, on by default):
//   B_t <- alpha L^{-1} B_t      (Left, Lower, NoTrans, Unit or NonUnit: the diagonal only enters through W)
// The columns of B are independent: one CTA takes 64 columns of one B tile through
//   for j = 0 .. nblk-1:   X_j = W_j (alpha B_j - sum_{c<j} L(j,c) X_c)
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(FT, 1)
trsm_lln_fused_kernel(int na, int n, R alpha, const R* __restrict__ Tm, int ldt, const R* __restrict__ W,
                      R* const* __restrict__ dB, int64_t offB, int ldb)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    R* Xs = reinterpret_cast<R*>(smem_dyn);
    R* Ys = Xs + FB * FLD;
    R* Cs = Ys + FB * FLD;
    const int tid = threadIdx.x;
    const Prod<R> pr(tid);
    const int c0 = blockIdx.x * FB;                        // first column of this CTA inside the tile
    const int cv = min(FB, n - c0);
    if (cv <= 0) return;
    R* Bcol = dB[blockIdx.y] + offB + int64_t(c0) * ldb;
    const int nblk = (na + FB - 1) / FB;
    R acc[32];
    for (int j = 0; j < nblk; ++j) {
        const int jv = min(FB, na - j * FB);
        zero_acc(acc);
        for (int c = 0; c < j; ++c) {
            load_block<R>(Xs, Tm + j * FB + int64_t(c) * FB * ldt, ldt, jv);       // L(j,c)(row, k), rows >= jv zero
            load_block_t<R>(Ys, Bcol + c * FB, ldb, cv);                          // X_c(k, col): written by this CTA
            __syncthreads();
            pr.mma(acc, Xs, Ys);
            __syncthreads();
        }
        R* Bj = Bcol + j * FB;
        #pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int row = pr.row(e), col = pr.col(e);
            const R o = (row < jv && col < cv) ? __ldcg(Bj + row + int64_t(col) * ldb) : R(0);
            Cs[row * FLD + col] = alpha * o - acc[e];                             // S(k = row, col)
        }
        load_block<R>(Xs, W + int64_t(j) * FB * FB, FB, FB);                       // W_j(row, k)
        __syncthreads();
        zero_acc(acc);
        pr.mma(acc, Xs, Cs);
        #pragma unroll
        for (int e = 0; e < 32; ++e) {
            const int row = pr.row(e), col = pr.col(e);
            if (row < jv && col < cv) Bj[row + int64_t(col) * ldb] = acc[e];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Left / Lower / NoTrans solve with a SMALL triangle (na <= 64) by direct substitution, one launch
// (SB200_TRSM_FUSED bit 2, on by default): the U12 = L11^-1 A12 steps inside the recursive
// LU panel (w1 = 32 or 64; 12 of the 15 updates of an nb = 512 panel) are today an inversion kernel with a 64-step
// dependent chain (~35 us) plus a GEMM launch.  Here: L and a 64-column slab of B in shared memory, axpy-form
// substitution with rolled loops on 256 threads (one barrier per row of the triangle, two if the diagonal is not unit).
// ---------------------------------------------------------------------------------------------
constexpr int SL_COLS = 64;         // columns of B per CTA
constexpr int SL_THREADS = 256;     // rolled loops over shared memory, 8 warps (see diag64.cuh for why not thread-per-column)

template <typename R>
__global__ void __launch_bounds__(SL_THREADS)
trsm_lln_small_kernel(int na, int n, R alpha, int unit, const R* __restrict__ Tm, int ldt,
                      R* const* __restrict__ dB, int64_t offB, int ldb)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    R* Ls = reinterpret_cast<R*>(smem_dyn);        // Ls[k * 65 + i] = L(i, k), i > k;  Ls[k * 65 + k] = 1 / L(k, k) (1 if unit)
    R* Bs = Ls + 64 * 65;                          // Bs[c * 65 + i] = B(i, c0 + c)
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * SL_COLS;
    const int cv = min(SL_COLS, n - c0);
    R* B = dB[blockIdx.y] + offB + int64_t(c0) * ldb;
    for (int e = tid; e < na * na; e += SL_THREADS) {
        const int i = e % na, k = e / na;
        if (i > k)       Ls[k * 65 + i] = Tm[i + int64_t(k) * ldt];
        else if (i == k) Ls[k * 65 + i] = unit ? R(1) : R(1) / Tm[i + int64_t(k) * ldt];
    }
    for (int e = tid; e < na * cv; e += SL_THREADS) {
        const int i = e % na, c = e / na;
        Bs[c * 65 + i] = alpha * B[i + int64_t(c) * ldb];
    }
    __syncthreads();
    const int ti = tid & 15, tcol = tid >> 4;      // 16 threads along the rows, 16 along the columns
    for (int k = 0; k < na; ++k) {
        if (! unit) {
            if (tid < cv) Bs[tid * 65 + k] *= Ls[k * 65 + k];           // x_k = b_k / L(k,k)
            __syncthreads();
        }
        for (int c = tcol; c < cv; c += SL_THREADS / 16) {
            const R xk = Bs[c * 65 + k];
            for (int i = k + 1 + ti; i < na; i += 16) Bs[c * 65 + i] = fma(-Ls[k * 65 + i], xk, Bs[c * 65 + i]);
        }
        __syncthreads();
    }
    for (int e = tid; e < na * cv; e += SL_THREADS) {
        const int i = e % na, c = e / na;
        B[i + int64_t(c) * ldb] = Bs[c * 65 + i];
    }
}

template <typename R> constexpr size_t fused_smem() { return (size_t(3) * FB * FLD + FB) * sizeof(R); }

template <typename R>
int potrf_tile_fused_t(int n, R* A, int lda, int* dinfo, int info_base, int variant, cudaStream_t stream)
{
    const int nblk = int(ceil_div(n, FB));
    if (nblk < 2 || nblk > FMAXB) return FUSED_NOT_TAKEN;
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (! attr_done[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(potrf_tile_fused_kernel<R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fused_smem<R>())));
        CUDA_TRY(cudaFuncSetAttribute(potrf_tile_fused_kernel<R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fused_smem<R>())));
        // keep freed stream-ordered allocations in the pool (the workspace below is allocated per launch)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
        attr_done[dev & 63] = true;
    }
    // stream-ordered workspace: the inverted diagonal blocks W_0 .. W_{nblk-2} and the flags
    const size_t wbytes = size_t(nblk) * FB * FB * sizeof(R);
    void* ws = nullptr;
    CUDA_TRY(cudaMallocAsync(&ws, wbytes + 2 * FMAXB * sizeof(unsigned), stream));
    unsigned* flags = reinterpret_cast<unsigned*>(static_cast<char*>(ws) + wbytes);
    cudaError_t e = cudaMemsetAsync(flags, 0, 2 * FMAXB * sizeof(unsigned), stream);
    int st = (e == cudaSuccess) ? SB200_OK : int(e);
    if (st == SB200_OK) {
        if (variant == 2)
            potrf_tile_fused_kernel<R, true><<<nblk, FT, fused_smem<R>(), stream>>>(A, lda, n, dinfo, info_base, static_cast<R*>(ws), flags);
        else
            potrf_tile_fused_kernel<R, false><<<nblk, FT, fused_smem<R>(), stream>>>(A, lda, n, dinfo, info_base, static_cast<R*>(ws), flags);
        st = launch_status();
    }
    cudaFreeAsync(ws, stream);
    return st;
}


// B_t <- alpha B_t L^{-T} for `batch` m x na tiles; W = inverted diagonal 64-blocks of L (trtri_diag layout)
template <typename R>
int trsm_rlt_fused_t(int m, int na, R alpha, const R* Tm, int ldt, const R* W, R* const* dB, int64_t offB, int ldb,
                     int batch, cudaStream_t stream)
{
    if (m <= 0 || na <= 0 || batch <= 0) return SB200_OK;
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (! attr_done[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(trsm_rlt_fused_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fused_smem<R>())));
        attr_done[dev & 63] = true;
    }
    const dim3 grid(unsigned(ceil_div(m, FB)), unsigned(batch));
    trsm_rlt_fused_kernel<R><<<grid, FT, fused_smem<R>(), stream>>>(m, na, alpha, Tm, ldt, W, dB, offB, ldb);
    return launch_status();
}

// B_t <- alpha L^{-1} B_t for `batch` na x n tiles; W = inverted diagonal 64-blocks of L (unit or not: trtri_diag)
template <typename R>
int trsm_lln_fused_t(int na, int n, R alpha, const R* Tm, int ldt, const R* W, R* const* dB, int64_t offB, int ldb,
                     int batch, cudaStream_t stream)
{
    if (n <= 0 || na <= 0 || batch <= 0) return SB200_OK;
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (! attr_done[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(trsm_lln_fused_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fused_smem<R>())));
        attr_done[dev & 63] = true;
    }
    const dim3 grid(unsigned(ceil_div(n, FB)), unsigned(batch));
    trsm_lln_fused_kernel<R><<<grid, FT, fused_smem<R>(), stream>>>(na, n, alpha, Tm, ldt, W, dB, offB, ldb);
    return launch_status();
}

template <typename R>
int trsm_lln_small_t(int na, int n, R alpha, bool unit, const R* Tm, int ldt, R* const* dB, int64_t offB, int ldb,
                     int batch, cudaStream_t stream)
{
    if (n <= 0 || na <= 0 || batch <= 0) return SB200_OK;
    if (na > 64) return SB200_EINVAL;
    const dim3 grid(unsigned(ceil_div(n, SL_COLS)), unsigned(batch));
    constexpr size_t smem = size_t(2) * 64 * 65 * sizeof(R);
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (! attr_done[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(trsm_lln_small_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        attr_done[dev & 63] = true;
    }
    trsm_lln_small_kernel<R><<<grid, SL_THREADS, smem, stream>>>(na, n, alpha, unit ? 1 : 0, Tm, ldt, dB, offB, ldb);
    return launch_status();
}

} // namespace

// return FUSED_NOT_TAKEN when the variant does not apply (the caller then runs the default path)
int potrf_tile_fused_d(int n, double* A, int lda, int* dinfo, int info_base, int variant, cudaStream_t stream)
{
    return potrf_tile_fused_t<double>(n, A, lda, dinfo, info_base, variant, stream);
}
int potrf_tile_fused_s(int n, float* A, int lda, int* dinfo, int info_base, int variant, cudaStream_t stream)
{
    return potrf_tile_fused_t<float>(n, A, lda, dinfo, info_base, variant, stream);
}

int trsm_rlt_fused_d(int m, int na, double alpha, const double* Tm, int ldt, const double* W, double* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream)
{
    return trsm_rlt_fused_t<double>(m, na, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
}
int trsm_rlt_fused_s(int m, int na, float alpha, const float* Tm, int ldt, const float* W, float* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream)
{
    return trsm_rlt_fused_t<float>(m, na, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
}
int trsm_lln_fused_d(int na, int n, double alpha, const double* Tm, int ldt, const double* W, double* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream)
{
    return trsm_lln_fused_t<double>(na, n, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
}
int trsm_lln_fused_s(int na, int n, float alpha, const float* Tm, int ldt, const float* W, float* const* dB,
                     int64_t offB, int ldb, int batch, cudaStream_t stream)
{
    return trsm_lln_fused_t<float>(na, n, alpha, Tm, ldt, W, dB, offB, ldb, batch, stream);
}
int trsm_lln_small_d(int na, int n, double alpha, bool unit, const double* Tm, int ldt, double* const* dB, int64_t offB,
                     int ldb, int batch, cudaStream_t stream)
{
    return trsm_lln_small_t<double>(na, n, alpha, unit, Tm, ldt, dB, offB, ldb, batch, stream);
}
int trsm_lln_small_s(int na, int n, float alpha, bool unit, const float* Tm, int ldt, float* const* dB, int64_t offB,
                     int ldb, int batch, cudaStream_t stream)
{
    return trsm_lln_small_t<float>(na, n, alpha, unit, Tm, ldt, dB, offB, ldb, batch, stream);
}

} // namespace sb200
