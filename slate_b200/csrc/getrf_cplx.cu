// getrf_cplx.cu -- base block of the LU panel for complex<float> / complex<double>.
//
// Reference: src/internal/Tile_getrf.hh:160-447 for complex scalar_t: the pivot of a column is the first strict
// maximum of cabs1(a) = |re a| + |im a| (:196-237), starting from the diagonal entry; the column below it is scaled by
// the complex reciprocal 1 / pivot unless cabs1(pivot) < safe_min (:334-361); an exactly zero pivot sets info and leaves
// the column unscaled (:362-366).
//
// Same cooperative scheme as getrf_base_kernel of getrf.cu (a w <= 32 column block of the panel resident in shared
// memory, rows spread over the CTAs, one grid barrier per column, every CTA picks the same winner), kept as a separate
// kernel so that the validated real-type kernels and their register allocation stay exactly as measured.  A complex<double>
// block keeps 384 rows per CTA (197 KB of shared memory) instead of 768.  The rest of the panel -- recursion, U12 solves,
// rank-w updates on the DMMA / SIMT GEMM, fused interchanges -- is the type-generic host code of getrf.cu.
#include "runtime_internal.hh"
#include "getrf_internal.hh"
#include <cooperative_groups.h>
#include <cfloat>
#include <climits>

namespace cg = cooperative_groups;

namespace sb200 {

namespace {

__device__ __forceinline__ float  abs1(cuFloatComplex a)  { return fabsf(a.x) + fabsf(a.y); }
__device__ __forceinline__ double abs1(cuDoubleComplex a) { return fabs(a.x) + fabs(a.y); }
template <typename R> __device__ __forceinline__ R tiny_real();
template <> __device__ __forceinline__ float  tiny_real<float>()  { return FLT_MIN; }
template <> __device__ __forceinline__ double tiny_real<double>() { return DBL_MIN; }

// NOPIV (getrf_nopiv): no candidate is proposed, so the diagonal entry is the pivot of every column
template <typename T, bool NOPIV>
__global__ void __launch_bounds__(PTHREADS)
getrf_base_cplx_kernel(const BaseArgs<T> a)
{
    using R = typename RealOf<T>::type;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char blk_raw[];
    T* blk = reinterpret_cast<T*>(blk_raw);          // [w][RP]
    __shared__ T s_prow[PW], s_drow[PW];
    __shared__ R s_val[PTHREADS / 32];
    __shared__ int s_row[PTHREADS / 32];
    __shared__ int s_p, s_w;
    R* gval = reinterpret_cast<R*>(a.gval);          // the candidates' cabs1 values are real
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RP = a.rows_per | 1;
    const int r_begin = a.c0 + b * a.rows_per;
    const int r_end = min(r_begin + a.rows_per, a.m_p);
    const int nr = max(r_end - r_begin, 0);
    const int nb = a.nb;
    const int GS = gridDim.x;                        // slots of the exchange arrays (one per CTA of the launch)

    for (int c = 0; c < a.w; ++c)
        for (int lr = tid; lr < nr; lr += PTHREADS) {
            const int r = r_begin + lr;
            blk[c * RP + lr] = a.tiles[r / nb][(r % nb) + int64_t(a.c0 + c) * nb];
        }
    __syncthreads();

    for (int j = 0; j < a.w; ++j) {
        const int d = a.c0 + j;                    // panel row of the diagonal entry
        const int par = j & 1;
        // ---- local candidate: first maximum of cabs1 over this CTA's rows below the diagonal
        R best = R(-1);
        int brow = INT_MAX;
        if constexpr (! NOPIV) {
        for (int lr = tid; lr < nr; lr += PTHREADS) {
            const int r = r_begin + lr;
            if (r > d) {
                const R v = abs1(blk[j * RP + lr]);
                if (v > best) { best = v; brow = r; }      // rows ascend per thread: first max kept
            }
        }
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const R ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
        }
        if (lane == 0) { s_val[warp] = best; s_row[warp] = brow; }
        __syncthreads();
        if (warp == 0) {
            best = lane < PTHREADS / 32 ? s_val[lane] : R(-1);
            brow = lane < PTHREADS / 32 ? s_row[lane] : INT_MAX;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const R ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
                if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
            }
            if (lane == 0) { gval[par * GS + b] = best; a.grow[par * GS + b] = brow; s_p = brow; }
        }
        __syncthreads();
        if (tid < a.w) {
            const int cr = s_p;
            if (cr != INT_MAX) a.gcand[(int64_t(par) * GS + b) * PW + tid] = blk[tid * RP + (cr - r_begin)];
            if (d >= r_begin && d < r_end) a.gdiag[par * PW + tid] = blk[tid * RP + (d - r_begin)];
        }
        __threadfence();
        grid.sync();

        // ---- every CTA picks the same winner: diagonal first, then strictly larger candidates
        if (warp == 0) {
            R bv = R(-1);
            int br = INT_MAX, bw = -1;
            for (int c = lane; c < GS; c += 32) {
                const R v = gval[par * GS + c];
                const int r = a.grow[par * GS + c];
                if (v > bv || (v == bv && r < br)) { bv = v; br = r; bw = c; }
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const R ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int orow = __shfl_xor_sync(0xffffffffu, br, o);
                const int ow = __shfl_xor_sync(0xffffffffu, bw, o);
                if (ov > bv || (ov == bv && orow < br)) { bv = ov; br = orow; bw = ow; }
            }
            if (lane == 0) {
                const R dv = abs1(a.gdiag[par * PW + j]);
                if (bv > dv) { s_p = br; s_w = bw; }        // strict: the diagonal wins ties (and NaN)
                else         { s_p = d;  s_w = -1; }
            }
        }
        __syncthreads();
        const int p = s_p;
        if (tid < a.w) {
            s_drow[tid] = a.gdiag[par * PW + tid];
            s_prow[tid] = (p == d) ? s_drow[tid] : a.gcand[(int64_t(par) * GS + s_w) * PW + tid];
        }
        __syncthreads();
        if (p != d && tid < a.w) {
            if (p >= r_begin && p < r_end) blk[tid * RP + (p - r_begin)] = s_drow[tid];
            if (d >= r_begin && d < r_end) blk[tid * RP + (d - r_begin)] = s_prow[tid];
        }
        if (b == 0 && tid == 0) {
            a.piv_tile[d] = p / nb;
            a.piv_off[d] = p % nb;
            if (a.rowmap && p != d) { const int t = a.rowmap[d]; a.rowmap[d] = a.rowmap[p]; a.rowmap[p] = t; }
        }
        if (a.kw_wide > 0 && b == GS - 1 && p != d) {
            // the row-less CTA applies the interchange to the panel columns outside [c0, c0 + w)
            T* rd_ = a.tiles[d / nb] + (d % nb);
            T* rp_ = a.tiles[p / nb] + (p % nb);
            for (int c = tid; c < a.kw_wide; c += PTHREADS)
                if (c < a.c0 || c >= a.c0 + a.w) {
                    const T t0 = rd_[int64_t(c) * nb], t1 = rp_[int64_t(c) * nb];
                    rd_[int64_t(c) * nb] = t1;
                    rp_[int64_t(c) * nb] = t0;
                }
        }
        __syncthreads();
        const T pv = s_prow[j];
        if (is_zero(pv)) {
            if (b == 0 && tid == 0 && *a.info == 0) *a.info = a.info_base + d + 1;
        }
        else {
            const bool use_rcp = abs1(pv) >= tiny_real<R>();
            const T rcp = divide(from_real<T>(R(1)), pv);
            for (int lr = tid; lr < nr; lr += PTHREADS) {
                const int r = r_begin + lr;
                if (r > d) {
                    T l = blk[j * RP + lr];
                    l = use_rcp ? mul(l, rcp) : divide(l, pv);
                    blk[j * RP + lr] = l;
                    const T ml = neg(l);
                    for (int c = j + 1; c < a.w; ++c) {
                        T acc = blk[c * RP + lr];
                        fma_acc(acc, ml, s_prow[c]);
                        blk[c * RP + lr] = acc;
                    }
                }
            }
        }
        __syncthreads();
    }
    for (int c = 0; c < a.w; ++c)
        for (int lr = tid; lr < nr; lr += PTHREADS) {
            const int r = r_begin + lr;
            a.tiles[r / nb][(r % nb) + int64_t(a.c0 + c) * nb] = blk[c * RP + lr];
        }
}

} // namespace

// one cooperative launch of `grid` CTAs (the caller's BaseArgs: getrf.cu panel_base_wide / getrf_panel_v1)
template <typename T>
int launch_base_cplx(BaseArgs<T>& a, int grid, size_t smem, bool nopiv, cudaStream_t s)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    static thread_local bool attr_done[64] = {};
    if (! attr_done[dev & 63]) {
        const int max_smem = int(PW * (panel_rows_max<T>() | 1) * sizeof(T));
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_cplx_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_cplx_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_done[dev & 63] = true;
    }
    void* args[] = {&a};
    void* fn = nopiv ? reinterpret_cast<void*>(getrf_base_cplx_kernel<T, true>) : reinterpret_cast<void*>(getrf_base_cplx_kernel<T, false>);
    const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(PTHREADS), args, smem, s);
    if (e != cudaSuccess) return int(e);
    return launch_status();
}

template int launch_base_cplx<cuFloatComplex>(BaseArgs<cuFloatComplex>&, int, size_t, bool, cudaStream_t);
template int launch_base_cplx<cuDoubleComplex>(BaseArgs<cuDoubleComplex>&, int, size_t, bool, cudaStream_t);

} // namespace sb200
