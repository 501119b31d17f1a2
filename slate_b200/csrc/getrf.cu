// getrf.cu -- LU with partial pivoting: GPU panel, fused row interchanges, driver.
//
// Reference: src/getrf.cc:22-244 (driver), src/internal/internal_getrf.cc:21-121 +
// src/internal/Tile_getrf.hh:160-447 (panel, ALWAYS on the host CPU in the reference: tiles are
// pulled device->host and pushed back at every step), src/internal/internal_swap.cc:510-805
// (permuteRows<Devices>: one cublas?swap launch per pivot row per block column).
//
// B200-first redesign:
//   * The panel is factored ON THE GPU.  A cooperative kernel keeps a whole m x 32 column block
//     resident in shared memory, rows spread over up to 148 CTAs (<= 768 rows each); per
//     column there is ONE grid-wide barrier: every CTA publishes its local |max| candidate row,
//     all CTAs redundantly pick the winner, swap, scale and rank-1 update from shared memory.
//     HBM traffic of a block = one read + one write.  Blocks are chained right-looking with
//     the DMMA GEMM (trsm + rank-32 update), as the reference's ib-blocked panel does.
//   * Pivot rule = the reference's (Tile_getrf.hh:196-289): start from the diagonal entry, take
//     the first strictly larger |a_ij| scanning rows in increasing order (ties keep the lowest
//     row; equals the reference run with one panel thread); scale by the reciprocal unless
//     |pivot| < safe_min; exact zero pivot -> info = column+1 and the column is left unscaled.
//   * Row interchanges of a whole panel are applied to a block-column range in ONE launch.
//   * Column-major tiles throughout (the reference converts to row-major for its swap BLAS
//     calls, src/getrf.cc:51-55); the fused swap kernel takes either layout.
#include "runtime_internal.hh"
#include "getrf_internal.hh"
#include <cooperative_groups.h>
#include <cfloat>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace cg = cooperative_groups;

namespace sb200 {

template <typename T> __device__ __forceinline__ T tiny_of();
template <> __device__ __forceinline__ float  tiny_of<float>()  { return FLT_MIN; }
template <> __device__ __forceinline__ double tiny_of<double>() { return DBL_MIN; }
__device__ __forceinline__ float  fma_t(float a, float b, float c)    { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }


// ------------------------------------------------------------------------------------------
// Fused row interchanges: pivot jj in [j0, j1) swaps stack row jj with stack row
// piv_tile[jj]*nb + piv_off[jj]; applied in order (forward) or reversed.  One thread per matrix
// column; `tiles[t + jt*ldt]` is tile t of the stack in block column jt.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
laswp_kernel(T* const* __restrict__ tiles, int64_t ldt, int mb, int nb, int ld, int colmajor,
             const int64_t* __restrict__ piv_tile, const int64_t* __restrict__ piv_off,
             int j0, int j1, int forward, int64_t col_lo, int64_t col_hi)
{
    extern __shared__ int s_piv[];               // (j1 - j0) panel-relative target rows
    for (int e = threadIdx.x; e < j1 - j0; e += blockDim.x)
        s_piv[e] = int(piv_tile[j0 + e] * mb + piv_off[j0 + e]);
    __syncthreads();
    const int64_t gc = col_lo + int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (gc >= col_hi) return;
    const int64_t jt = gc / nb;
    const int cc = int(gc - jt * nb);
    T* const* stack = tiles + jt * ldt;
    const int64_t coff = colmajor ? int64_t(cc) * ld : cc;
    const int64_t rstr = colmajor ? 1 : ld;
    const int cnt = j1 - j0;
    for (int s = 0; s < cnt; ++s) {
        const int e = forward ? s : cnt - 1 - s;
        const int r1 = j0 + e, r2 = s_piv[e];
        if (r1 == r2) continue;
        T* p1 = stack[r1 / mb] + int64_t(r1 % mb) * rstr + coff;
        T* p2 = stack[r2 / mb] + int64_t(r2 % mb) * rstr + coff;
        const T t = *p1; *p1 = *p2; *p2 = t;
    }
}

template <typename T>
int launch_laswp(T* const* tiles, int64_t ldt, int mb, int nb, int ld, int colmajor,
                        const int64_t* piv_tile, const int64_t* piv_off, int j0, int j1, int forward,
                        int64_t col_lo, int64_t col_hi, cudaStream_t s)
{
    if (col_hi <= col_lo || j1 <= j0) return SB200_OK;
    const int64_t cols = col_hi - col_lo;
    laswp_kernel<T><<<unsigned(ceil_div(cols, 256)), 256, size_t(j1 - j0) * sizeof(int), s>>>(
        tiles, ldt, mb, nb, ld, colmajor, piv_tile, piv_off, j0, j1, forward, col_lo, col_hi);
    return launch_status();
}

// ------------------------------------------------------------------------------------------
// Cooperative panel base block: columns [c0, c0+w) of the panel (tile stack `tiles`, panel rows
// [c0, m_p) active), rows_per rows per CTA held in shared memory.
// ------------------------------------------------------------------------------------------
// BaseArgs<T>: getrf_internal.hh (shared with the complex base kernel, getrf_cplx.cu)

// NOPIV = true (getrf_nopiv, src/getrf_nopiv.cc): no candidate is ever proposed, so the diagonal entry is the pivot of
// every column and the interchange logic below degenerates to the identity; everything else (the diagonal row's
// broadcast, zero-pivot info, scaling, rank-1 update) is the same code.  SURVEY section 8(f) item 2.
template <typename T, bool NOPIV = false>
__global__ void __launch_bounds__(PTHREADS)
getrf_base_kernel(const BaseArgs<T> a)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char blk_raw[];
    T* blk = reinterpret_cast<T*>(blk_raw);          // [w][RP]
    __shared__ T s_prow[PW], s_drow[PW];
    __shared__ T s_val[PTHREADS / 32];
    __shared__ int    s_row[PTHREADS / 32];
    __shared__ int    s_p, s_w;
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int RP = a.rows_per | 1;
    const int r_begin = a.c0 + b * a.rows_per;
    const int r_end = min(r_begin + a.rows_per, a.m_p);
    const int nr = max(r_end - r_begin, 0);
    const int nb = a.nb;

    for (int c = 0; c < a.w; ++c)
        for (int lr = tid; lr < nr; lr += PTHREADS) {
            const int r = r_begin + lr;
            blk[c * RP + lr] = a.tiles[r / nb][(r % nb) + int64_t(a.c0 + c) * nb];
        }
    __syncthreads();

    for (int j = 0; j < a.w; ++j) {
        const int d = a.c0 + j;                    // panel row of the diagonal entry
        const int par = j & 1;
        // ---- local candidate: first maximum of |a| over this CTA's rows below the diagonal
        T best = T(-1);
        int brow = INT_MAX;
        if constexpr (! NOPIV) {
        for (int lr = tid; lr < nr; lr += PTHREADS) {
            const int r = r_begin + lr;
            if (r > d) {
                const T v = fabs(blk[j * RP + lr]);
                if (v > best) { best = v; brow = r; }      // rows ascend per thread: first max kept
            }
        }
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
        }
        if (lane == 0) { s_val[warp] = best; s_row[warp] = brow; }
        __syncthreads();
        if (warp == 0) {
            best = lane < PTHREADS / 32 ? s_val[lane] : T(-1);
            brow = lane < PTHREADS / 32 ? s_row[lane] : INT_MAX;
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const T ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
                if (ov > best || (ov == best && orow < brow)) { best = ov; brow = orow; }
            }
            if (lane == 0) { a.gval[par * G + b] = best; a.grow[par * G + b] = brow; s_p = brow; }
        }
        __syncthreads();
        if (tid < a.w) {
            const int cr = s_p;
            if (cr != INT_MAX) a.gcand[(int64_t(par) * G + b) * PW + tid] = blk[tid * RP + (cr - r_begin)];
            if (d >= r_begin && d < r_end) a.gdiag[par * PW + tid] = blk[tid * RP + (d - r_begin)];
        }
        __threadfence();
        grid.sync();

        // ---- every CTA picks the same winner: diagonal first, then strictly larger candidates
        if (warp == 0) {
            T bv = T(-1);
            int br = INT_MAX, bw = -1;
            for (int c = lane; c < G; c += 32) {
                const T v = a.gval[par * G + c];
                const int r = a.grow[par * G + c];
                if (v > bv || (v == bv && r < br)) { bv = v; br = r; bw = c; }
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const T ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int orow = __shfl_xor_sync(0xffffffffu, br, o);
                const int ow = __shfl_xor_sync(0xffffffffu, bw, o);
                if (ov > bv || (ov == bv && orow < br)) { bv = ov; br = orow; bw = ow; }
            }
            if (lane == 0) {
                const T dv = fabs(a.gdiag[par * PW + j]);
                if (bv > dv) { s_p = br; s_w = bw; }        // strict: the diagonal wins ties (and NaN)
                else         { s_p = d;  s_w = -1; }
            }
        }
        __syncthreads();
        const int p = s_p;
        if (tid < a.w) {
            s_drow[tid] = a.gdiag[par * PW + tid];
            s_prow[tid] = (p == d) ? s_drow[tid] : a.gcand[(int64_t(par) * G + s_w) * PW + tid];
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
        if (a.kw_wide > 0 && b == G - 1 && p != d) {
            // thread t always handles panel columns t, t + PTHREADS, ...: the interchanges of one column
            // are applied in pivot order by one thread (rows d and p of later pivots may coincide)
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
        if (pv == T(0)) {
            if (b == 0 && tid == 0 && *a.info == 0) *a.info = a.info_base + d + 1;
        }
        else {
            const bool use_rcp = fabs(pv) >= tiny_of<T>();
            const T rcp = T(1) / pv;
            for (int lr = tid; lr < nr; lr += PTHREADS) {
                const int r = r_begin + lr;
                if (r > d) {
                    T l = blk[j * RP + lr];
                    l = use_rcp ? l * rcp : l / pv;
                    blk[j * RP + lr] = l;
                    for (int c = j + 1; c < a.w; ++c)
                        blk[c * RP + lr] = fma_t(-l, s_prow[c], blk[c * RP + lr]);
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

// set by the getrf_nopiv entry points around their driver call (the drivers construct their PanelScratch on the
// calling thread); read once per driver call by PanelScratch::init
static thread_local bool g_getrf_nopiv = false;
// likewise for getrf_tntpiv: 0 = partial pivoting, > 0 = participants of the tournament
static thread_local int g_getrf_tnt = 0;

int tnt_ranks_for(const Grid& g)
{
    const char* e = getenv("SB200_TNT_RANKS");
    const int r = e ? atoi(e) : 0;
    return r > 0 ? r : std::max(g.p, 1);
}

PanelScratch::~PanelScratch() { if (raw) ws_cache_put(raw); }

int PanelScratch::init()
{
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    max_ctas = sms;
    const size_t G = size_t(sms);
    const size_t v3_bytes = base_v3_scratch_bytes(sms);
    const size_t bytes = 2 * G * 8 + 2 * G * 8 + 2 * G * PW * 16 + 2 * PW * 16 + 16 * 64 * 64 * 16 + 64 + 16 + v3_bytes;      // 16: complex<double> elements
    raw = ws_cache_get(bytes);
    if (! raw) return SB200_ENOMEM;
    char* p = static_cast<char*>(raw);
    gval = reinterpret_cast<double*>(p); p += 2 * G * 8;
    grow = reinterpret_cast<int*>(p);    p += 2 * G * 8;
    gcand = reinterpret_cast<double*>(p); p += 2 * G * PW * 16;
    gdiag = reinterpret_cast<double*>(p); p += 2 * PW * 16;
    W = reinterpret_cast<double*>(p); p += 16 * 64 * 64 * 16;
    bar = reinterpret_cast<unsigned*>(p); p += 64;
    p = static_cast<char*>(raw) + (size_t(p - static_cast<char*>(raw)) + 15) / 16 * 16;      // 16-byte vector accesses
    v3_buf = reinterpret_cast<unsigned long long*>(p);
    nopiv = g_getrf_nopiv;
    tnt_ranks = g_getrf_tnt;
    use_v3 = ! nopiv && switch_value(SW_PANEL_V3) != 0;
    if (use_v3) {
        CUDA_TRY(cudaMemset(v3_buf, 0, v3_bytes));             // tag 0 is never used by a launch
        SB_TRY(base_v3_init());
    }
    v3_gen = 0;
    static thread_local bool attr_done[64] = {};
    if (! attr_done[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(PW * (PROWS_MAX | 1) * sizeof(double))));
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(PW * (PROWS_MAX | 1) * sizeof(float))));
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_kernel<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(PW * (PROWS_MAX | 1) * sizeof(double))));
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(PW * (PROWS_MAX | 1) * sizeof(float))));
        attr_done[dev & 63] = true;
    }
    return SB200_OK;
}

// Factor the panel given as a stack of `ntile` tiles (device pointer array `stack`, nb x nb, ld =
// nb; last tile has m_p - (ntile-1)*nb rows), kw columns, diag_len = min(m_p, kw) pivots.
// Panel algorithm selector: SB200_PANEL=1 right-looking 32-column blocks with laswp launches between
// them; SB200_PANEL=2 (default) recursive panel, interchanges applied to the whole panel width inside
// the base kernel.  Both give the pivot sequence of unblocked partial pivoting.
static int panel_version()
{
    const char* e = getenv("SB200_PANEL");
    return e ? atoi(e) : 2;
}

// one base-block launch of the cooperative-groups kernel (getrf_nopiv, SB200_PANEL=1, SB200_PANEL_V3=0; the default
// path is getrf_base_v3.cu)
template <typename T>
static int launch_base(BaseArgs<T>& a, int grid, size_t smem, PanelScratch& ps, cudaStream_t s)
{
    if constexpr (IsComplex<T>::value) {
        return launch_base_cplx<T>(a, grid, smem, ps.nopiv, s);
    }
    else {
    cudaError_t e;
    if (ps.nopiv) {                                             // getrf_nopiv: the barrier kernel without a pivot search
        void* args[] = {&a};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(getrf_base_kernel<T, true>), dim3(grid), dim3(PTHREADS), args, smem, s);
    }
    else {
        void* args[] = {&a};
        e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(getrf_base_kernel<T>), dim3(grid), dim3(PTHREADS), args, smem, s);
    }
    if (e != cudaSuccess) return int(e);
    return launch_status();
    }
}

namespace {
template <typename T>
struct PanelCtx {
    T* const* stack; T* tile0; int ntile, nb, m_p, kw;
    int64_t* piv_tile; int64_t* piv_off; int* dinfo; int info_base;
    PanelScratch* ps; cudaStream_t s; int* rowmap; PhaseTimer* pt;
};
}

// one cooperative launch: columns [c0, c0+w) over panel rows [c0, m_p), interchanges applied to all
// kw panel columns by the extra (row-less) CTA
template <typename T>
static int panel_base_wide(const PanelCtx<T>& x, int c0, int w, int upd_c0 = -1)
{
    if constexpr (! IsComplex<T>::value) {             // the one-round kernels are real-type kernels
    if (x.ps->use_v3) {
        x.pt->begin("pnl_base", x.s);
        SB_TRY(launch_base_v3<T>(x.stack, x.nb, x.m_p, c0, w, x.kw, x.piv_tile, x.piv_off, x.dinfo, x.info_base,
                                 x.rowmap, *x.ps, x.s, upd_c0));
        x.pt->end(x.s);
        return SB200_OK;
    }
    }
    constexpr int rows_max = panel_rows_max<T>();
    const int active = x.m_p - c0;
    const int ctas = x.ps->max_ctas - 1;                       // one SM is kept for the interchange CTA
    int rows_per = std::max(int(ceil_div(active, ctas)), std::min(active, rows_max));
    rows_per = std::max(rows_per, PW);
    if (rows_per > rows_max) return SB200_ENOTSUP;
    const int G = int(ceil_div(active, rows_per));
    BaseArgs<T> a{x.stack, x.nb, x.m_p, c0, w, rows_per, x.piv_tile, x.piv_off,
                  reinterpret_cast<T*>(x.ps->gval), x.ps->grow, reinterpret_cast<T*>(x.ps->gcand),
                  reinterpret_cast<T*>(x.ps->gdiag), x.dinfo, x.info_base, x.rowmap, x.kw,
                  nullptr};
    const size_t smem = size_t(w) * (rows_per | 1) * sizeof(T);
    x.pt->begin("pnl_base", x.s);
    SB_TRY(launch_base<T>(a, G + 1, smem, *x.ps, x.s));
    x.pt->end(x.s);
    return SB200_OK;
}

// columns [cc, cc+n2) of the panel, given that columns [c0, c0+w1) are factored:
//   U12 = L11^{-1} A12 (rows c0..c0+w1 of the top tile), then A22 -= L21 U12 on rows [c0+w1, m_p)
template <typename T>
static int panel_update(const PanelCtx<T>& x, int c0, int w1, int cc, int n2)
{
    if (n2 <= 0 || w1 <= 0) return SB200_OK;
    const int nb = x.nb;
    const T ONE = from_real<T>(1), MINUS_ONE = from_real<T>(-1);
    x.pt->begin("pnl_trsm", x.s);
    SB_TRY(trsm_colmajor<T>(true, true, 'N', true, w1, n2, ONE, x.tile0 + c0 + int64_t(c0) * nb, nb,
                           x.stack, c0 + int64_t(cc) * nb, nb, 1, reinterpret_cast<T*>(x.ps->W), x.s));
    x.pt->end(x.s);
    x.pt->begin("pnl_gemm", x.s);
    const T* U12 = x.tile0 + c0 + int64_t(cc) * nb;
    const int r0 = c0 + w1;                                   // first row of A22 (inside the top tile)
    const int top_rows = std::min(nb, x.m_p) - r0;
    if (top_rows > 0) {
        GemmParamsT<T> p{};
        p.m = top_rows; p.n = n2; p.k = w1; p.alpha = MINUS_ONE; p.beta = ONE; p.batch = 1;
        p.A0 = x.tile0 + r0 + int64_t(c0) * nb; p.lda = nb;
        p.B0 = U12; p.ldb = nb;
        p.C0 = x.tile0 + r0 + int64_t(cc) * nb; p.ldc = nb;
        SB_TRY(launch_gemm<T>('N', 'N', p, x.s));
    }
    const int full = (x.m_p % nb == 0) ? x.ntile - 1 : x.ntile - 2;      // full-height tiles below tile 0
    if (full > 0) {
        GemmParamsT<T> p{};
        p.m = nb; p.n = n2; p.k = w1; p.alpha = MINUS_ONE; p.beta = ONE; p.batch = full;
        p.A = x.stack + 1; p.offA = int64_t(c0) * nb; p.lda = nb;
        p.B0 = U12; p.ldb = nb; p.strideB = 0;
        p.C = x.stack + 1; p.offC = int64_t(cc) * nb; p.ldc = nb;
        SB_TRY(launch_gemm<T>('N', 'N', p, x.s));
    }
    if (x.ntile > 1 && x.m_p % nb != 0) {
        GemmParamsT<T> p{};
        p.m = x.m_p % nb; p.n = n2; p.k = w1; p.alpha = MINUS_ONE; p.beta = ONE; p.batch = 1;
        p.A = x.stack + (x.ntile - 1); p.offA = int64_t(c0) * nb; p.lda = nb;
        p.B0 = U12; p.ldb = nb;
        p.C = x.stack + (x.ntile - 1); p.offC = int64_t(cc) * nb; p.ldc = nb;
        SB_TRY(launch_gemm<T>('N', 'N', p, x.s));
    }
    x.pt->end(x.s);
    return SB200_OK;
}

// recursive LU of columns [c0, c0+w); on entry they carry every update of the columns left of c0 and
// every interchange chosen so far; on return so do ALL panel columns (base kernels swap panel-wide)
template <typename T>
static int panel_recurse(const PanelCtx<T>& x, int c0, int w)
{
    if (w <= PW) return panel_base_wide<T>(x, c0, w);
    int w1 = int(ceil_div(w / 2, PW)) * PW;
    if (w1 >= w) w1 = w - PW;
    SB_TRY(panel_recurse<T>(x, c0, w1));
    // a 32-column block that follows a 32-column block takes that block's update inside its own launch
    if constexpr (! IsComplex<T>::value) {
        if (base_v3_can_fuse(*x.ps, x.m_p, c0 + w1, w1, w - w1)) return panel_base_wide<T>(x, c0 + w1, w - w1, c0);
    }
    SB_TRY(panel_update<T>(x, c0, w1, c0 + w1, w - w1));
    return panel_recurse<T>(x, c0 + w1, w - w1);
}

template <typename T>
static int getrf_panel_v2(T* const* stack, T* tile0, int ntile, int nb, int m_p, int kw,
                          int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                          PanelScratch& ps, cudaStream_t s, int* rowmap, PhaseTimer* ph)
{
    PhaseTimer off_timer;
    off_timer.on = false;
    PanelCtx<T> x{stack, tile0, ntile, nb, m_p, kw, piv_tile, piv_off, dinfo, info_base, &ps, s, rowmap,
               ph ? ph : &off_timer};
    const int diag_len = std::min(m_p, kw);
    SB_TRY(panel_recurse<T>(x, 0, diag_len));
    // wide last panel (m_p < kw): the columns right of the square part only get U = L^{-1} A
    return panel_update<T>(x, 0, diag_len, diag_len, kw - diag_len);
}

template <typename T>
static int getrf_panel_v1(T* const* stack, T* tile0, int ntile, int nb, int m_p, int kw,
                  int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                  PanelScratch& ps, cudaStream_t s, int* rowmap, PhaseTimer* ph)
{
    PhaseTimer off_timer;
    off_timer.on = false;
    PhaseTimer& pt = ph ? *ph : off_timer;
    const int diag_len = std::min(m_p, kw);
    const T ONE = from_real<T>(1), MINUS_ONE = from_real<T>(-1);
    constexpr int rows_max = panel_rows_max<T>();
    for (int c0 = 0; c0 < diag_len; c0 += PW) {
        const int w = std::min(PW, diag_len - c0);
        const int active = m_p - c0;
        int rows_per = std::max(int(ceil_div(active, ps.max_ctas)), std::min(active, rows_max));
        rows_per = std::max(rows_per, PW);
        if (rows_per > rows_max) return SB200_ENOTSUP;       // panel taller than 148 * 768 rows
        const int G = int(ceil_div(active, rows_per));
        BaseArgs<T> a{stack, nb, m_p, c0, w, rows_per, piv_tile, piv_off,
                      reinterpret_cast<T*>(ps.gval), ps.grow, reinterpret_cast<T*>(ps.gcand),
                      reinterpret_cast<T*>(ps.gdiag), dinfo, info_base, rowmap, 0, nullptr};
        const size_t smem = size_t(w) * (rows_per | 1) * sizeof(T);
        pt.begin("pnl_base", s);
        SB_TRY(launch_base<T>(a, G, smem, ps, s));
        pt.end(s);
        // interchanges of this block applied to the rest of the panel (left and right of the block)
        pt.begin("pnl_laswp", s);
        SB_TRY(launch_laswp<T>(stack, 0, nb, nb, nb, 1, piv_tile, piv_off, c0, c0 + w, 1, 0, c0, s));
        SB_TRY(launch_laswp<T>(stack, 0, nb, nb, nb, 1, piv_tile, piv_off, c0, c0 + w, 1, c0 + w, kw, s));
        pt.end(s);
        const int rest = kw - c0 - w;
        if (rest <= 0) continue;
        // U12 = L11^{-1} A12  (top tile, rows c0..c0+w)
        pt.begin("pnl_trsm", s);
        SB_TRY(trsm_colmajor<T>(true, true, 'N', true, w, rest, ONE, tile0 + c0 + int64_t(c0) * nb, nb,
                               stack, c0 + int64_t(c0 + w) * nb, nb, 1, reinterpret_cast<T*>(ps.W), s));
        pt.end(s);
        pt.begin("pnl_gemm", s);
        // A22 -= L21 U12
        const T* U12 = tile0 + c0 + int64_t(c0 + w) * nb;
        const int top_rows = std::min(nb, m_p) - (c0 + w);
        if (top_rows > 0) {
            GemmParamsT<T> p{};
            p.m = top_rows; p.n = rest; p.k = w; p.alpha = MINUS_ONE; p.beta = ONE; p.batch = 1;
            p.A0 = tile0 + (c0 + w) + int64_t(c0) * nb; p.lda = nb;
            p.B0 = U12; p.ldb = nb;
            p.C0 = tile0 + (c0 + w) + int64_t(c0 + w) * nb; p.ldc = nb;
            SB_TRY(launch_gemm<T>('N', 'N', p, s));
        }
        const int full = (m_p % nb == 0) ? ntile - 1 : ntile - 2;      // full-height tiles below tile 0
        if (full > 0) {
            GemmParamsT<T> p{};
            p.m = nb; p.n = rest; p.k = w; p.alpha = MINUS_ONE; p.beta = ONE; p.batch = full;
            p.A = stack + 1; p.offA = int64_t(c0) * nb; p.lda = nb;
            p.B0 = U12; p.ldb = nb; p.strideB = 0;
            p.C = stack + 1; p.offC = int64_t(c0 + w) * nb; p.ldc = nb;
            SB_TRY(launch_gemm<T>('N', 'N', p, s));
        }
        if (ntile > 1 && m_p % nb != 0) {
            GemmParamsT<T> p{};
            p.m = m_p % nb; p.n = rest; p.k = w; p.alpha = MINUS_ONE; p.beta = ONE; p.batch = 1;
            p.A = stack + (ntile - 1); p.offA = int64_t(c0) * nb; p.lda = nb;
            p.B0 = U12; p.ldb = nb;
            p.C = stack + (ntile - 1); p.offC = int64_t(c0 + w) * nb; p.ldc = nb;
            SB_TRY(launch_gemm<T>('N', 'N', p, s));
        }
        pt.end(s);
    }
    return SB200_OK;
}

template <typename T>
int getrf_panel(T* const* stack, T* tile0, int ntile, int nb, int m_p, int kw,
                int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                PanelScratch& ps, cudaStream_t s, int* rowmap, PhaseTimer* ph)
{
    return (panel_version() == 1 ? getrf_panel_v1<T> : getrf_panel_v2<T>)(
        stack, tile0, ntile, nb, m_p, kw, piv_tile, piv_off, dinfo, info_base, ps, s, rowmap, ph);
}

int getrf_panel_d(double* const* stack, double* tile0, int ntile, int nb, int m_p, int kw,
                  int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                  PanelScratch& ps, cudaStream_t s, int* rowmap, PhaseTimer* ph)
{
    return getrf_panel<double>(stack, tile0, ntile, nb, m_p, kw, piv_tile, piv_off, dinfo, info_base, ps, s, rowmap, ph);
}

// ------------------------------------------------------------------------------------------
// driver, one rank (the p x q variant in getrf_dist.cu gathers the panel on the diagonal owner).
// T = double: the FP64 path.  T = float: the low-precision factorisation of gesv_mixed
// (src/gesv_mixed.cc:171-176); with use_tc05 its trailing / lookahead GEMMs run on the tcgen05
// FP32-emulated kernel, L(:,k) and U(k,:) being split-packed once per step.
// ------------------------------------------------------------------------------------------
template <typename T>
int getrf_driver_t(Matrix& A, int64_t* pivots_out, int64_t* info_out, bool use_tc05)
{
    Grid& g = *A.g;
    if (A.kind != 'G' || A.layout != 'C' || A.dtype != TypeChar<T>::value) return SB200_EINVAL;
    constexpr bool is_float = std::is_same<T, float>::value;
    if (use_tc05 && ! is_float) return SB200_EINVAL;
    if (g.size() > 1) return SB200_ENOTSUP;
    CUDA_TRY(cudaDeviceSynchronize());
    const int64_t mt = A.mt, nt = A.nt, nb = A.nb;
    const int ld = int(nb);
    const int64_t kt = std::min(mt, nt);
    const int64_t mn = std::min(A.m, A.n);
    if (mn == 0) { if (info_out) *info_out = 0; return SB200_OK; }
    const T ONE = from_real<T>(1), MINUS_ONE = from_real<T>(-1);

    // packed operands of the tcgen05 path: L(i,k) in slot [k & 1][i], U(k,j) in slot [k & 1][j]
    DevBuf packA, packB;
    const size_t pa_bytes = use_tc05 ? tc05_packed_bytes('A', nb, nb) : 0;
    const size_t pb_bytes = use_tc05 ? tc05_packed_bytes('B', nb, nb) : 0;
    if (use_tc05) {
        SB_TRY(packA.alloc(size_t(2) * mt * pa_bytes));
        SB_TRY(packB.alloc(size_t(2) * nt * pb_bytes));
    }
    auto pkA = [&](int64_t i, int64_t k) { return packA.as<unsigned char>() + ((k & 1) * mt + i) * pa_bytes; };
    auto pkB = [&](int64_t j, int64_t k) { return packB.as<unsigned char>() + ((k & 1) * nt + j) * pb_bytes; };

    // tile pointer tables: column-major (stacks of a block column) and row-major (block rows)
    std::vector<T*> tbl(size_t(mt * nt)), tblT(size_t(mt * nt));
    for (int64_t j = 0; j < nt; ++j)
        for (int64_t i = 0; i < mt; ++i) {
            tbl[size_t(i + j * mt)] = A.tile_as<T>(i, j);
            tblT[size_t(j + i * nt)] = A.tile_as<T>(i, j);
        }
    // SB200_GEMM_BT (default 1, measured r2a; double, not the tcgen05 path): the row U(k, k+2:) is
    // transposed once per step after its solve, so that the trailing update runs as 'N','T' (both operands staged by TMA
    // bulk copies: the variant of the potrf trailing update, 0.91 of the DMMA peak) instead of 'N','N' (K-major B through
    // 16-byte cp.async; 0.80 measured here).  Same products in the same order: bitwise the same factor.
    bool use_bt = false;
    if constexpr (std::is_same<T, double>::value) {
        use_bt = switch_value(SW_GEMM_BT) != 0 && ! use_tc05;
    }
    DevBuf wsUt;
    const int64_t te_ = nb * nb;
    if (use_bt) SB_TRY(wsUt.alloc(size_t(2) * nt * te_ * sizeof(T)));
    auto ut = [&](int64_t k, int64_t j) -> T* { return wsUt.as<T>() + ((k & 1) * nt + j) * te_; };
    // per-step GEMM batches: lookahead column k+1 and trailing columns >= k+2; pack lists (tcgen05 path)
    struct Step {
        std::vector<const T*> ut_src_full, ut_src_last;     // U(k,j), j >= k+2 (use_bt)
        std::vector<T*> ut_dst_full, ut_dst_last;
        size_t ut_src_full_off = 0, ut_dst_full_off = 0, ut_src_last_off = 0, ut_dst_last_off = 0;
        std::vector<Batch> la, tr;
        std::vector<const void*> a_src, b_src;       // L(i,k), i > k (ragged last row at the end); U(k,j), j >= k+2 (ragged last col at the end)
        std::vector<void*> a_dst, b_dst;
        int a_full = 0, b_full = 0;
        size_t a_src_off = 0, a_dst_off = 0, b_src_off = 0, b_dst_off = 0;
    };
    std::vector<Step> steps(static_cast<size_t>(kt));
    PlanBuffer pb;
    for (int64_t k = 0; k < kt; ++k) {
        Step& s = steps[size_t(k)];
        const int kw = int(A.tile_nb(k));
        for (int64_t j = k + 1; j < nt; ++j)
            for (int64_t i = k + 1; i < mt; ++i) {
                const void* a = use_tc05 ? static_cast<const void*>(pkA(i, k)) : A.tile_as<T>(i, k);
                const void* b = use_tc05 ? static_cast<const void*>(pkB(j, k))
                              : (use_bt && j >= k + 2) ? static_cast<const void*>(ut(k, j)) : A.tile_as<T>(k, j);
                batch_add(j == k + 1 ? s.la : s.tr, int(A.tile_mb(i)), int(A.tile_nb(j)), kw, 0, a, b, A.tile_as<T>(i, j));
            }
        if (use_bt && k + 1 < mt)
            for (int64_t j = k + 2; j < nt; ++j) {
                if (A.tile_nb(j) == nb) { s.ut_src_full.push_back(A.tile_as<T>(k, j)); s.ut_dst_full.push_back(ut(k, j)); }
                else                    { s.ut_src_last.push_back(A.tile_as<T>(k, j)); s.ut_dst_last.push_back(ut(k, j)); }
            }
        s.ut_src_full_off = pb.push(s.ut_src_full); s.ut_dst_full_off = pb.push(s.ut_dst_full);
        s.ut_src_last_off = pb.push(s.ut_src_last); s.ut_dst_last_off = pb.push(s.ut_dst_last);
        if (use_tc05 && k + 1 < nt) {
            for (int64_t i = k + 1; i < mt; ++i) {
                s.a_src.push_back(A.tile_as<T>(i, k)); s.a_dst.push_back(pkA(i, k));
                if (A.tile_mb(i) == nb) ++s.a_full;
            }
            for (int64_t j = k + 2; j < nt; ++j) {
                s.b_src.push_back(A.tile_as<T>(k, j)); s.b_dst.push_back(pkB(j, k));
                if (A.tile_nb(j) == nb) ++s.b_full;
            }
        }
        pb.reserve(s.la);
        pb.reserve(s.tr);
        s.a_src_off = pb.push(s.a_src); s.a_dst_off = pb.push(s.a_dst);
        s.b_src_off = pb.push(s.b_src); s.b_dst_off = pb.push(s.b_dst);
    }
    const size_t tbl_off = pb.push(tbl), tblT_off = pb.push(tblT);

    DevBuf piv, dinfo, wt;
    SB_TRY(piv.alloc(size_t(2 * kt * nb) * sizeof(int64_t)));
    SB_TRY(dinfo.alloc(sizeof(int)));
    SB_TRY(wt.alloc(size_t(ceil_div(nb, 64)) * 64 * 64 * sizeof(T)));
    int64_t* dpiv_tile = piv.as<int64_t>();
    int64_t* dpiv_off = dpiv_tile + kt * nb;
    T* Wt = wt.as<T>();

    PanelScratch ps;
    SB_TRY(ps.init());
    Streams st;
    SB_TRY(st.init(size_t(2 * kt)));
    cudaStream_t P = st.panel, T_ = st.trail;
    TntScratch tnt;                                       // getrf_tntpiv only
    if (ps.tnt_ranks > 0) {
        if (use_tc05 || ! tnt_shape_supported(A)) return SB200_ENOTSUP;
        SB_TRY(tnt.init(mt, nb, A.m, int(sizeof(T)), ps.tnt_ranks, P));
    }
    auto P_done = [&](int64_t k) { return st.ev[size_t(k)]; };
    auto T_done = [&](int64_t k) { return st.ev[size_t(kt + k)]; };
    double trail_flops = 0; int64_t trail_launches = 0;
    PhaseTimer ph;
    SB_TRY(pb.upload(P));
    T* const* dtbl = pb.at<T>(tbl_off);
    T* const* dtblT = pb.at<T>(tblT_off);

    auto run_batches = [&](const std::vector<Batch>& bs, cudaStream_t s, bool b_transposed = false) -> int {
        if constexpr (is_float) {
            if (use_tc05) return launch_batches_tc05(bs, pb, -1.0f, 1.0f, ld, s);
        }
        return launch_batches<T>(bs, pb, 'N', b_transposed ? 'T' : 'N', MINUS_ONE, ONE, ld, 0, s);
    };
    // split-pack `cnt` tiles (full-size ones first, then the ragged one) for one operand role
    auto pack = [&](int role, size_t src_off, size_t dst_off, int cnt, int full, int rows_full, int rows_last,
                    int kw, cudaStream_t s) -> int {
        if constexpr (is_float) {
            for (int part = 0; part < 2; ++part) {
                const int c = part == 0 ? full : cnt - full;
                if (c <= 0) continue;
                const size_t o = part == 0 ? 0 : size_t(full);
                Tc05PackParams q{};
                q.X = reinterpret_cast<const float* const*>(pb.dev + src_off + o);
                q.P = reinterpret_cast<void* const*>(pb.dev + dst_off + o);
                q.rows = part == 0 ? rows_full : rows_last; q.k = kw;
                if (role == 'A') { q.rs = 1; q.ks = ld; q.ru = TC_BM; }     // L(i,k): operand row = tile row
                else             { q.rs = ld; q.ks = 1; q.ru = TC_BN; }     // U(k,j): operand row = tile column
                q.batch = c;
                SB_TRY(launch_tc05_pack(q, s));
            }
        }
        (void) role; (void) src_off; (void) dst_off; (void) cnt; (void) full; (void) rows_full; (void) rows_last; (void) kw; (void) s;
        return SB200_OK;
    };
    // row-k solve U(k, j0..j1) = L_kk^{-1} A(k, j0..j1) on stream s with workspace W
    auto row_trsm = [&](int64_t k, int64_t j0, int64_t j1, T* W, cudaStream_t s) -> int {
        if (j1 <= j0) return SB200_OK;
        const int kw = int(std::min(A.tile_mb(k), A.tile_nb(k)));
        const int64_t jfull_end = (A.tile_nb(nt - 1) == nb) ? j1 : std::min(j1, nt - 1);
        if (jfull_end > j0)
            SB_TRY(trsm_colmajor<T>(true, true, 'N', true, kw, int(nb), ONE, A.tile_as<T>(k, k), ld,
                                    dtblT + j0 + k * nt, 0, ld, int(jfull_end - j0), W, s));
        if (jfull_end < j1)
            SB_TRY(trsm_colmajor<T>(true, true, 'N', true, kw, int(A.tile_nb(nt - 1)), ONE, A.tile_as<T>(k, k), ld,
                                    dtblT + (nt - 1) + k * nt, 0, ld, 1, W, s));
        return SB200_OK;
    };

    CUDA_TRY(cudaMemsetAsync(dinfo.p, 0, sizeof(int), P));
    CUDA_TRY(cudaStreamSynchronize(P));
    CUDA_TRY(cudaEventRecord(st.t0, P));
    for (int64_t k = 0; k < kt; ++k) {
        Step& sk = steps[size_t(k)];
        const int kw = int(A.tile_nb(k));
        const int m_p = int(A.m - k * nb);
        const int diag_len = std::min(m_p, kw);
        int64_t* pt = dpiv_tile + k * nb;
        int64_t* po = dpiv_off + k * nb;
        T* const* stack_k = dtbl + k + k * mt;
        // ---- panel k (column k already carries every earlier update: lookahead below)
        SB_TRY(st.ptime(P));
        if (ps.tnt_ranks > 0) {
            std::vector<T*> htiles;
            for (int64_t i = k; i < mt; ++i) htiles.push_back(A.tile_as<T>(i, k));
            SB_TRY(getrf_panel_tnt<T>(stack_k, htiles, k, int(nb), m_p, kw, pt, po, dinfo.as<int>(), int(k * nb), ps, tnt, P,
                                      nullptr, &ph));
        }
        else
        SB_TRY(getrf_panel<T>(stack_k, A.tile_as<T>(k, k), int(mt - k), int(nb), m_p, kw, pt, po, dinfo.as<int>(),
                              int(k * nb), ps, P, nullptr, &ph));
        if (use_tc05 && ! sk.a_src.empty()) {
            ph.begin("pack_L", P);
            SB_TRY(pack('A', sk.a_src_off, sk.a_dst_off, int(sk.a_src.size()), sk.a_full, int(nb), int(A.tile_mb(mt - 1)), kw, P));
            ph.end(P);
        }
        SB_TRY(st.ptime(P));
        CUDA_TRY(cudaEventRecord(P_done(k), P));
        // ---- trailing update of columns >= k+2 and interchanges to the left, normal priority
        CUDA_TRY(cudaStreamWaitEvent(T_, P_done(k), 0));
        SB_TRY(launch_laswp<T>(dtbl + k, mt, int(nb), int(nb), ld, 1, pt, po, 0, diag_len, 1, 0, k * nb, T_));
        if (k + 2 < nt) {
            SB_TRY(launch_laswp<T>(dtbl + k, mt, int(nb), int(nb), ld, 1, pt, po, 0, diag_len, 1, (k + 2) * nb, A.n, T_));
            SB_TRY(row_trsm(k, k + 2, nt, Wt, T_));
            if (use_tc05 && ! sk.b_src.empty())
                SB_TRY(pack('B', sk.b_src_off, sk.b_dst_off, int(sk.b_src.size()), sk.b_full, int(nb), int(A.tile_nb(nt - 1)), kw, T_));
            if constexpr (std::is_same<T, double>::value) {
                if (use_bt) {
                    if (! sk.ut_src_full.empty())
                        SB_TRY(sb200_transpose_batched_d(0, kw, nb, pb.at<const double>(sk.ut_src_full_off), ld,
                                                         pb.at<double>(sk.ut_dst_full_off), ld, int64_t(sk.ut_src_full.size()), T_));
                    if (! sk.ut_src_last.empty())
                        SB_TRY(sb200_transpose_batched_d(0, kw, A.tile_nb(nt - 1), pb.at<const double>(sk.ut_src_last_off), ld,
                                                         pb.at<double>(sk.ut_dst_last_off), ld, int64_t(sk.ut_src_last.size()), T_));
                }
            }
            if (! sk.tr.empty()) {
                SB_TRY(st.time_begin(T_));
                SB_TRY(run_batches(sk.tr, T_, use_bt));
                SB_TRY(st.time_end(T_));
                trail_flops += batches_flops(sk.tr, IsComplex<T>::value);
                trail_launches += int64_t(sk.tr.size());
            }
        }
        CUDA_TRY(cudaEventRecord(T_done(k), T_));
        // ---- lookahead: bring column k+1 up to date on the panel stream
        if (k + 1 < nt) {
            if (k >= 1) CUDA_TRY(cudaStreamWaitEvent(P, T_done(k - 1), 0));
            ph.begin("la_swap_trsm", P);
            SB_TRY(launch_laswp<T>(dtbl + k, mt, int(nb), int(nb), ld, 1, pt, po, 0, diag_len, 1,
                                   (k + 1) * nb, std::min<int64_t>((k + 2) * nb, A.n), P));
            SB_TRY(row_trsm(k, k + 1, k + 2, reinterpret_cast<T*>(ps.W), P));
            if constexpr (is_float) {
                if (use_tc05 && ! sk.la.empty()) {
                    // U(k, k+1) packed on the panel stream into its own slot
                    Tc05PackParams q{};
                    q.X0 = A.tile_as<float>(k, k + 1); q.P0 = pkB(k + 1, k); q.strideX = 0; q.strideP = 0;
                    q.rows = int(A.tile_nb(k + 1)); q.k = kw; q.rs = ld; q.ks = 1; q.ru = TC_BN; q.batch = 1;
                    SB_TRY(launch_tc05_pack(q, P));
                }
            }
            ph.end(P);
            ph.begin("la_gemm", P);
            SB_TRY(run_batches(sk.la, P));
            ph.end(P);
        }
    }
    CUDA_TRY(cudaStreamWaitEvent(P, T_done(kt - 1), 0));
    CUDA_TRY(cudaEventRecord(st.t1, P));
    CUDA_TRY(cudaStreamSynchronize(P));
    CUDA_TRY(cudaStreamSynchronize(T_));
    ph.report("getrf", g.rank);
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, st.t0, st.t1));
    A.last_ms = ms;
    A.last_trail_ms = st.timed_ms(); A.last_trail_flops = trail_flops; A.last_trail_launches = trail_launches;
    A.last_panel_ms = st.panel_ms();
    int hinfo = 0;
    CUDA_TRY(cudaMemcpy(&hinfo, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (info_out) *info_out = hinfo;
    if (pivots_out) {
        std::vector<int64_t> ht(size_t(kt * nb)), ho(size_t(kt * nb));
        CUDA_TRY(cudaMemcpy(ht.data(), dpiv_tile, ht.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(ho.data(), dpiv_off, ho.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
        int64_t o = 0;
        for (int64_t k = 0; k < kt; ++k) {
            const int64_t dl = std::min(A.m - k * nb, A.tile_nb(k));
            for (int64_t j = 0; j < dl && o < mn; ++j, ++o) {
                pivots_out[2 * o] = ht[size_t(k * nb + j)];
                pivots_out[2 * o + 1] = ho[size_t(k * nb + j)];
            }
        }
    }
    return SB200_OK;
}

int getrf_driver(Matrix& A, int64_t* pivots_out, int64_t* info_out)
{
    // SB200_GETRF_DIST=1 runs the p x q algorithm on a single rank too (test hook: same code path
    // as the multi-GPU runs, minus the NCCL calls)
    const char* fd = getenv("SB200_GETRF_DIST");
    const bool force_dist = fd && atoi(fd) != 0;
    if (A.dtype != 'd') return SB200_EINVAL;
    if (A.g->size() > 1 || force_dist) return getrf_driver_dist(A, pivots_out, info_out);
    return getrf_driver_t<double>(A, pivots_out, info_out, false);
}

int getrf_driver_s(Matrix& A, int64_t* pivots_out, int64_t* info_out, bool use_tc05)
{
    const char* fd = getenv("SB200_GETRF_DIST");
    const bool force_dist = fd && atoi(fd) != 0;
    if (A.dtype != 's') return SB200_EINVAL;
    if (A.g->size() > 1 || force_dist) return getrf_driver_dist_s(A, pivots_out, info_out, use_tc05);
    return getrf_driver_t<float>(A, pivots_out, info_out, use_tc05);
}

// complex LU (1 x 1 grid): the same driver, base blocks by getrf_cplx.cu, trailing update on the complex GEMM kernels
int getrf_driver_cplx(Matrix& A, int64_t* pivots_out, int64_t* info_out)
{
    if (A.g->size() > 1) return SB200_ENOTSUP;
    if (A.dtype == 'z') return getrf_driver_t<cuDoubleComplex>(A, pivots_out, info_out, false);
    if (A.dtype == 'c') return getrf_driver_t<cuFloatComplex>(A, pivots_out, info_out, false);
    return SB200_EINVAL;
}

template int getrf_panel<double>(double* const*, double*, int, int, int, int, int64_t*, int64_t*, int*, int, PanelScratch&, cudaStream_t, int*, PhaseTimer*);
template int getrf_panel<float>(float* const*, float*, int, int, int, int, int64_t*, int64_t*, int*, int, PanelScratch&, cudaStream_t, int*, PhaseTimer*);
template int getrf_panel<cuFloatComplex>(cuFloatComplex* const*, cuFloatComplex*, int, int, int, int, int64_t*, int64_t*, int*, int, PanelScratch&, cudaStream_t, int*, PhaseTimer*);
template int getrf_panel<cuDoubleComplex>(cuDoubleComplex* const*, cuDoubleComplex*, int, int, int, int, int64_t*, int64_t*, int*, int, PanelScratch&, cudaStream_t, int*, PhaseTimer*);

template int launch_laswp<double>(double* const*, int64_t, int, int, int, int, const int64_t*, const int64_t*, int, int, int, int64_t, int64_t, cudaStream_t);
template int launch_laswp<float>(float* const*, int64_t, int, int, int, int, const int64_t*, const int64_t*, int, int, int, int64_t, int64_t, cudaStream_t);
template int launch_laswp<cuFloatComplex>(cuFloatComplex* const*, int64_t, int, int, int, int, const int64_t*, const int64_t*, int, int, int, int64_t, int64_t, cudaStream_t);
template int launch_laswp<cuDoubleComplex>(cuDoubleComplex* const*, int64_t, int, int, int, int, const int64_t*, const int64_t*, int, int, int, int64_t, int64_t, cudaStream_t);

} // namespace sb200

using namespace sb200;

extern "C" {

/* LU without pivoting (slate::getrf_nopiv, src/getrf_nopiv.cc): the getrf drivers with the pivot search switched off.
 * STATUS: validated on B200 in round 2 (1-, 2- and 8-GPU runs, profiles/r02*). */
static int getrf_nopiv_any(sb200_matrix_t h, int64_t* info, int dtype)
{
    if (! h || h->A.dtype != dtype) return SB200_EINVAL;
    struct Guard { Guard() { g_getrf_nopiv = true; } ~Guard() { g_getrf_nopiv = false; } } guard;
    std::vector<int64_t> piv(size_t(2 * std::max<int64_t>(std::min(h->A.m, h->A.n), 1)));
    if (dtype == 'z' || dtype == 'c') return getrf_driver_cplx(h->A, piv.data(), info);
    return dtype == 's' ? getrf_driver_s(h->A, piv.data(), info, false) : getrf_driver(h->A, piv.data(), info);
}
int sb200_getrf_nopiv_z(sb200_matrix_t h, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_nopiv_any(h, info, 'z'); }
int sb200_getrf_nopiv_c(sb200_matrix_t h, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_nopiv_any(h, info, 'c'); }
int sb200_getrf_nopiv_d(sb200_matrix_t h, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_nopiv_any(h, info, 'd'); }
int sb200_getrf_nopiv_s(sb200_matrix_t h, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_nopiv_any(h, info, 's'); }

/* LU with tournament pivoting (slate::getrf_tntpiv, src/getrf_tntpiv.cc; MethodLU::CALU of slate::lu_factor): the getrf
 * drivers with the panel of getrf_tnt.cu.  Participants per panel = process rows of the grid.
 * STATUS: validated on one B200 at the end of round 2 (profiles/r02r1_*, r02r2_*): 1 - 4 participants per panel through
 * SB200_TNT_RANKS, single-rank and grid driver, identical pivots against the oracle, which is pinned to the reference's own
 * runs on 2x1 ... 2x4 process grids (tests/golden/grid_getrf_tntpiv_d_*.npz); not yet run on a real p x q grid. */
static int getrf_tntpiv_any(sb200_matrix_t h, int64_t* pivots, int64_t* info, int dtype)
{
    if (! h || h->A.dtype != dtype) return SB200_EINVAL;
    if (! tnt_shape_supported(h->A)) return SB200_ENOTSUP;
    struct Guard { explicit Guard(int r) { g_getrf_tnt = r; } ~Guard() { g_getrf_tnt = 0; } } guard(tnt_ranks_for(*h->A.g));
    if (dtype == 'z' || dtype == 'c') return getrf_driver_cplx(h->A, pivots, info);           // 1 x 1 grid
    return dtype == 's' ? getrf_driver_s(h->A, pivots, info, false) : getrf_driver(h->A, pivots, info);
}
int sb200_getrf_tntpiv_d(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_tntpiv_any(h, pivots, info, 'd'); }
int sb200_getrf_tntpiv_s(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_tntpiv_any(h, pivots, info, 's'); }
int sb200_getrf_tntpiv_z(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_tntpiv_any(h, pivots, info, 'z'); }
int sb200_getrf_tntpiv_c(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info) { SB_TRY(options_status(opts)); return getrf_tntpiv_any(h, pivots, info, 'c'); }

/* complex LU with partial pivoting (cabs1 rule); 1 x 1 grid */
int sb200_getrf_z(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info)
{
    SB_TRY(options_status(opts));
    if (! h || h->A.dtype != 'z') return SB200_EINVAL;
    return getrf_driver_cplx(h->A, pivots, info);
}
int sb200_getrf_c(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info)
{
    SB_TRY(options_status(opts));
    if (! h || h->A.dtype != 'c') return SB200_EINVAL;
    return getrf_driver_cplx(h->A, pivots, info);
}

int sb200_getrf_d(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info)
{
    SB_TRY(options_status(opts));
    if (! h) return SB200_EINVAL;
    return getrf_driver(h->A, pivots, info);
}

/* FP32 LU (1 x 1 grid): the low-precision factorisation of gesv_mixed.  sb200_getrf_tc05_s runs the
 * trailing update on the tcgen05 FP32-emulated (3 x TF32) kernel. */
int sb200_getrf_s(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info)
{
    SB_TRY(options_status(opts));
    if (! h) return SB200_EINVAL;
    return getrf_driver_s(h->A, pivots, info, false);
}

int sb200_getrf_tc05_s(sb200_matrix_t h, int64_t* pivots, const sb200_options_t* opts, int64_t* info)
{
    SB_TRY(options_status(opts));
    if (! h) return SB200_EINVAL;
    return getrf_driver_s(h->A, pivots, info, true);
}

// one body for the four scalar types: pure data movement (sb200_permute_rows_{s,d,c,z})
#define SB200_DEF_PERMUTE_ROWS(X, T, CT) \
int sb200_permute_rows_##X(int layout, int forward, int64_t npiv, \
                           const int64_t* d_piv_tile, const int64_t* d_piv_off, \
                           T* const* dTiles, int64_t mt, int64_t ncolblocks, \
                           int64_t tile_mb, int64_t ncols, int64_t ld, sb200_stream_t stream) \
{ \
    if (! valid_layout(layout) || npiv < 0 || mt < 1 || ncolblocks < 0 || tile_mb < 1 || ncols < 0) return SB200_EINVAL; \
    if (npiv == 0 || ncolblocks == 0 || ncols == 0) return SB200_OK; \
    if (npiv > 0x7fffffff || tile_mb > 0x7fffffff || ld > 0x7fffffff) return SB200_EINVAL; \
    /* `ncols` columns per tile, block columns are `ncols` wide */ \
    return launch_laswp<CT>(reinterpret_cast<CT* const*>(dTiles), mt, int(tile_mb), int(ncols), int(ld), layout == 'C', \
                            d_piv_tile, d_piv_off, 0, int(npiv), forward != 0, 0, ncolblocks * ncols, cudaStream_t(stream)); \
}
SB200_DEF_PERMUTE_ROWS(s, float, float)
SB200_DEF_PERMUTE_ROWS(d, double, double)
SB200_DEF_PERMUTE_ROWS(c, sb200_c32, cuFloatComplex)
SB200_DEF_PERMUTE_ROWS(z, sb200_c64, cuDoubleComplex)

} // extern "C"
