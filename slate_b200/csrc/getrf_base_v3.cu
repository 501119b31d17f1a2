// getrf_base_v3.cu -- LU panel base block (<= 32 columns, partial pivoting) with ONE exchange round per column.
//
// Reference: the panel task of slate::getrf (src/getrf.cc:91-116 -> internal::getrf_panel ->
// src/internal/Tile_getrf.hh:160-447: per column a thread-team max search + MPI_Allreduce(MAXLOC) + row swap +
// scale + rank-1 update).  Same pivot rule (first maximum of |a| at or below the diagonal; the diagonal wins ties and
// NaNs), same arithmetic (reciprocal scaling above sfmin, FMA rank-1 update) as getrf_base_kernel in getrf.cu, whose
// factors this kernel reproduces bit for bit.
//
// Why a third kernel (profiles/r02a_perf_variants_phases_1gpu.txt): the cooperative kernel costs 10 us per column
// (327 ms of the 906 ms dgetrf at n = 32768 and the whole critical path of dgetrf on 8 GPUs), of which the arithmetic
// is < 0.5 us.  The rest is a chain of dependent L2 round trips: candidate stores + fence, grid barrier, candidate
// scan, diagonal value, winner's row, and -- on the critical path of EVERY column because every CTA waits for every
// other one -- CTA 0's row-map update and the interchange CTA's panel-wide swap (two dependent global accesses each).
// Here:
//   * every value that crosses CTAs travels as 8-byte words {32 data bits | 32-bit generation tag} (an aligned 8-byte
//     store is single-copy atomic: a reader that sees the tag sees the data; no fence, no barrier);
//   * slots are per COLUMN of the launch (32 x G records), never reused inside a launch, and tags grow over the launches
//     of a driver call: no flow control between CTAs is needed at all, a CTA never waits for a reader;
//   * a CTA publishes {|max|, row} + the candidate's 32 values + (CTA 0) the diagonal row in one go, polls the G
//     headers (one per thread) and the diagonal row together -- one round trip --, then the winner's row -- second one;
//   * the max search of column j+1 is folded into the rank-1 update of column j;
//   * the panel-wide interchanges and the row map are done by NWIDE extra CTAs that FOLLOW the pivot sequence
//     (tagged pivot words from CTA 0) and that nobody waits for.
// All CTAs must be co-resident (they wait for each other's publications): cooperative launch, G + NWIDE <= #SMs.
#include "runtime_internal.hh"
#include "getrf_internal.hh"
#include <cfloat>
#include <climits>
#include <cstdio>
#include <vector>

namespace sb200 {

namespace {

constexpr int V3_HDR = 8;                       // words: |max| (2), row (1), pad
constexpr int V3_REC = V3_HDR + 2 * PW;         // + candidate row (2 words per value)

__device__ __forceinline__ void v3_store(unsigned long long* p, unsigned data, unsigned gen)
{
    const unsigned long long w = (static_cast<unsigned long long>(gen) << 32) | data;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void v3_store_double(unsigned long long* p, double v, unsigned gen)   // p 16-byte aligned
{
    const unsigned long long g = static_cast<unsigned long long>(gen) << 32;
    const unsigned long long lo = g | unsigned(__double2loint(v)), hi = g | unsigned(__double2hiint(v));
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ void v3_load2(const unsigned long long* p, unsigned long long& a, unsigned long long& b)
{
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ double v3_wait_double(const unsigned long long* p, unsigned gen)
{
    unsigned long long lo, hi;
    const long long t0 = clock64();
    for (;;) {
        v3_load2(p, lo, hi);
        if (unsigned(lo >> 32) == gen && unsigned(hi >> 32) == gen) break;
        spin_watchdog(t0);
    }
    return __hiloint2double(int(unsigned(hi)), int(unsigned(lo)));
}
__device__ __forceinline__ unsigned v3_wait_word(const unsigned long long* p, unsigned gen)
{
    unsigned long long w;
    const long long t0 = clock64();
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
        if (unsigned(w >> 32) == gen) break;
        spin_watchdog(t0);
    }
    return unsigned(w);
}

template <typename T> __device__ __forceinline__ T v3_tiny();
template <> __device__ __forceinline__ float  v3_tiny<float>()  { return FLT_MIN; }
template <> __device__ __forceinline__ double v3_tiny<double>() { return DBL_MIN; }
__device__ __forceinline__ float  v3_fma(float a, float b, float c)    { return fmaf(a, b, c); }
__device__ __forceinline__ double v3_fma(double a, double b, double c) { return fma(a, b, c); }

// (value, row): larger value wins, equal values -> smaller row (the first maximum in row order)
__device__ __forceinline__ bool v3_better(double v, int r, double bv, int br) { return v > bv || (v == bv && r < br); }

template <typename T>
struct V3Args {
    T* const* tiles;
    int nb, m_p, c0, w, rows_per, G;          // G row CTAs (blockIdx < G), the rest are interchange CTAs
    int64_t* piv_tile; int64_t* piv_off;
    int* info; int info_base;
    int* rowmap; int kw_wide;
    unsigned long long* rec;                  // [PW][gmax][V3_REC]
    unsigned long long* drow;                 // [PW][2 PW]     diagonal row of column j (published by CTA 0)
    unsigned long long* pivrec;               // [PW]           pivot row of column j (published by CTA 0)
    unsigned gen_base; int gmax;
    int upd_c0;                               // >= 0 (v4 only): first fold in the update by the 32-column block [upd_c0, upd_c0 + 32)
    long long* trace;                         // debug (SB200_V3_TRACE=1): [G][8] cycles per phase, thread 0 of every row CTA
};
#define V3_MARK(k) do { if (a.trace && tid == 0) { const long long t_ = clock64(); tr_acc[k] += t_ - tr_last; tr_last = t_; } } while (0)

template <typename T>
__global__ void __launch_bounds__(PTHREADS)
getrf_base_v3_kernel(const V3Args<T> a)
{
    extern __shared__ __align__(16) unsigned char blk_raw[];
    T* blk = reinterpret_cast<T*>(blk_raw);          // [w][RP]
    __shared__ T s_prow[PW], s_drow[PW];
    __shared__ double s_val[PTHREADS / 32];
    __shared__ int    s_row[PTHREADS / 32];
    __shared__ double g_val[PTHREADS / 32];          // partials of the gather (separate from the candidate's: no reuse race)
    __shared__ int    g_row[PTHREADS / 32], g_cta[PTHREADS / 32];
    constexpr int NWARP = PTHREADS / 32;
    const int G = a.G, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = a.nb, w = a.w;

    if (b >= G) {
        // ---- interchange CTAs: follow the pivot sequence, swap rows d and p in every panel column outside the block
        const int wi = b - G, nw = int(gridDim.x) - G;
        const int outside = a.kw_wide - w;
        for (int j = 0; j < w; ++j) {
            const int d = a.c0 + j;
            const int p = int(v3_wait_word(a.pivrec + j, a.gen_base + unsigned(j) + 1u));
            if (p == d) continue;
            T* rd_ = a.tiles[d / nb] + (d % nb);
            T* rp_ = a.tiles[p / nb] + (p % nb);
            // thread t always handles the same columns: the interchanges of one column are applied in pivot order
            for (int ci = wi * PTHREADS + tid; ci < outside; ci += nw * PTHREADS) {
                const int c = ci < a.c0 ? ci : ci + w;
                const T t0 = rd_[int64_t(c) * nb], t1 = rp_[int64_t(c) * nb];
                rd_[int64_t(c) * nb] = t1;
                rp_[int64_t(c) * nb] = t0;
            }
            if (a.rowmap && wi == 0 && tid == 0) { const int t = a.rowmap[d]; a.rowmap[d] = a.rowmap[p]; a.rowmap[p] = t; }
        }
        return;
    }

    long long tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tr_last = clock64();
    const int RP = a.rows_per | 1;
    const int r_begin = a.c0 + b * a.rows_per;
    const int r_end = min(r_begin + a.rows_per, a.m_p);
    const int nr = max(r_end - r_begin, 0);

    for (int c = 0; c < w; ++c)
        for (int lr = tid; lr < nr; lr += PTHREADS) {
            const int r = r_begin + lr;
            blk[c * RP + lr] = a.tiles[r / nb][(r % nb) + int64_t(a.c0 + c) * nb];
        }
    __syncthreads();

    V3_MARK(0);                                     // slab load
    // candidate of the first column; later ones come out of the rank-1 update
    double best = -1.0;
    int brow = INT_MAX;
    for (int lr = tid; lr < nr; lr += PTHREADS) {
        const int r = r_begin + lr;
        if (r > a.c0) {
            const double v = double(fabs(blk[lr]));
            if (v > best) { best = v; brow = r; }          // rows ascend per thread: first max kept
        }
    }

    for (int j = 0; j < w; ++j) {
        const int d = a.c0 + j;                    // panel row of the diagonal entry; always one of CTA 0's rows
        const unsigned gen = a.gen_base + unsigned(j) + 1u;
        unsigned long long* slot = a.rec + size_t(j) * a.gmax * V3_REC;
        // ---- this CTA's candidate
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
            if (v3_better(ov, orow, best, brow)) { best = ov; brow = orow; }
        }
        if (lane == 0) { s_val[warp] = best; s_row[warp] = brow; }
        __syncthreads();                           // also: the update's writes to blk are visible below
        if (warp <= 1) {
            best = lane < NWARP ? s_val[lane] : -1.0;
            brow = lane < NWARP ? s_row[lane] : INT_MAX;
            #pragma unroll
            for (int o = NWARP / 2; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int orow = __shfl_xor_sync(0xffffffffu, brow, o);
                if (v3_better(ov, orow, best, brow)) { best = ov; brow = orow; }
            }
            best = __shfl_sync(0xffffffffu, best, 0);
            brow = __shfl_sync(0xffffffffu, brow, 0);
            unsigned long long* R = slot + size_t(b) * V3_REC;
            if (warp == 0) {
                if (brow != INT_MAX && lane < w)
                    v3_store_double(R + V3_HDR + 2 * lane, double(blk[lane * RP + (brow - r_begin)]), gen);
                if (lane == 0) { v3_store_double(R, best, gen); v3_store(R + 2, unsigned(brow), gen); }
            }
            else if (b == 0 && lane < w)
                v3_store_double(a.drow + (size_t(j) * PW + lane) * 2, double(blk[lane * RP + (d - r_begin)]), gen);
        }
        V3_MARK(1);                                 // candidate reduce + publish
        // ---- gather: thread c polls the header of CTA c; the last warp polls the diagonal row
        double bv = -1.0;
        int br = INT_MAX, bw = -1;
        for (int c = tid; c < G; c += PTHREADS) {
            const unsigned long long* R = slot + size_t(c) * V3_REC;
            unsigned long long w0, w1, w2, w3;
            const long long t0 = clock64();
            for (;;) {
                v3_load2(R, w0, w1); v3_load2(R + 2, w2, w3);
                if (unsigned(w0 >> 32) == gen && unsigned(w1 >> 32) == gen && unsigned(w2 >> 32) == gen) break;
                spin_watchdog(t0);
            }
            const double v = __hiloint2double(int(unsigned(w1)), int(unsigned(w0)));
            const int rr = int(unsigned(w2));
            if (v3_better(v, rr, bv, br)) { bv = v; br = rr; bw = c; }
        }
        if (warp == NWARP - 1 && lane < w)
            s_drow[lane] = T(v3_wait_double(a.drow + (size_t(j) * PW + lane) * 2, gen));
        V3_MARK(2);                                 // header poll (thread 0: the header of CTA 0)
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int orow = __shfl_xor_sync(0xffffffffu, br, o);
            const int ow = __shfl_xor_sync(0xffffffffu, bw, o);
            if (v3_better(ov, orow, bv, br)) { bv = ov; br = orow; bw = ow; }
        }
        if (lane == 0) { g_val[warp] = bv; g_row[warp] = br; g_cta[warp] = bw; }
        __syncthreads();
        bv = lane < NWARP ? g_val[lane] : -1.0;
        br = lane < NWARP ? g_row[lane] : INT_MAX;
        bw = lane < NWARP ? g_cta[lane] : -1;
        #pragma unroll
        for (int o = NWARP / 2; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int orow = __shfl_xor_sync(0xffffffffu, br, o);
            const int ow = __shfl_xor_sync(0xffffffffu, bw, o);
            if (v3_better(ov, orow, bv, br)) { bv = ov; br = orow; bw = ow; }
        }
        bv = __shfl_sync(0xffffffffu, bv, 0);
        br = __shfl_sync(0xffffffffu, br, 0);
        bw = __shfl_sync(0xffffffffu, bw, 0);
        V3_MARK(3);                                 // gather reduce (waits for the slowest poller of this CTA)
        // ---- every CTA picks the same winner: strictly larger than the diagonal, or the diagonal (ties, NaN)
        const double dv = double(fabs(s_drow[j]));
        const int p = (bv > dv) ? br : d;
        if (tid < w)
            s_prow[tid] = (p == d) ? s_drow[tid]
                                   : T(v3_wait_double(slot + size_t(bw) * V3_REC + V3_HDR + 2 * tid, gen));
        if (b == 0 && tid == 2 * 32) {
            a.piv_tile[d] = p / nb;
            a.piv_off[d] = p % nb;
            v3_store(a.pivrec + j, unsigned(p), gen);
        }
        __syncthreads();
        V3_MARK(4);                                 // winner's row
        if (p != d && tid < w) {
            if (p >= r_begin && p < r_end) blk[tid * RP + (p - r_begin)] = s_drow[tid];
            if (b == 0) blk[tid * RP + (d - r_begin)] = s_prow[tid];
        }
        __syncthreads();
        V3_MARK(5);                                 // swap
        // ---- scale + rank-1 update of rows below the diagonal; the next column's candidate falls out of it
        const T pv = s_prow[j];
        best = -1.0; brow = INT_MAX;
        if (pv == T(0)) {
            if (b == 0 && tid == 0 && *a.info == 0) *a.info = a.info_base + d + 1;
            if (j + 1 < w)
                for (int lr = tid; lr < nr; lr += PTHREADS) {
                    const int r = r_begin + lr;
                    if (r > d + 1) {
                        const double v = double(fabs(blk[(j + 1) * RP + lr]));
                        if (v > best) { best = v; brow = r; }
                    }
                }
        }
        else {
            const bool use_rcp = fabs(pv) >= v3_tiny<T>();
            const T rcp = T(1) / pv;
            for (int lr = tid; lr < nr; lr += PTHREADS) {
                const int r = r_begin + lr;
                if (r > d) {
                    T l = blk[j * RP + lr];
                    l = use_rcp ? l * rcp : l / pv;
                    blk[j * RP + lr] = l;
                    if (j + 1 < w) {
                        const T x = v3_fma(-l, s_prow[j + 1], blk[(j + 1) * RP + lr]);
                        blk[(j + 1) * RP + lr] = x;
                        if (r > d + 1) {
                            const double v = double(fabs(x));
                            if (v > best) { best = v; brow = r; }
                        }
                    }
                    for (int c = j + 2; c < w; ++c)
                        blk[c * RP + lr] = v3_fma(-l, s_prow[c], blk[c * RP + lr]);
                }
            }
        }
        V3_MARK(6);                                 // update
        // (the barrier at the top of the next column orders these writes before anybody reads them)
    }
    __syncthreads();
    for (int c = 0; c < w; ++c)
        for (int lr = tid; lr < nr; lr += PTHREADS) {
            const int r = r_begin + lr;
            a.tiles[r / nb][(r % nb) + int64_t(a.c0 + c) * nb] = blk[c * RP + lr];
        }
    V3_MARK(7);                                     // slab store
    if (a.trace && tid == 0)
        for (int k = 0; k < 8; ++k) a.trace[b * 8 + k] = tr_acc[k];
}


// ------------------------------------------------------------------------------------------
// Register-resident variant (the default for panels of <= (#SMs - NWIDE) x 512 rows): ONE row per thread, its <= 32
// block columns in registers, the column loop fully unrolled so that every register index is static.
// Why (v3 phase trace, profiles/r02f_v3_phase_trace.txt): with the block in shared memory the rank-1 update of one
// column re-reads and re-writes the whole remaining slab -- 3 700 of the 12 300 cycles per column, bound by shared-memory
// bandwidth -- and loading / storing the slab costs another 2 300 per column; the exchange itself has a floor of
// 1 500-1 900 cycles (profiles/r02g_allgather_floor.txt).  In registers the update is <= 31 DFMAs per thread.
// Same protocol, same pivot rule, same arithmetic as above.
// ------------------------------------------------------------------------------------------
constexpr int V4_THREADS = 512;

// argmax over a warp of (|value| as its 64 bits {hi, lo}, row): larger value first, then the smaller row.  Three
// redux.sync instead of five shuffle rounds of three registers; every lane gets the result.  (hi, lo) = 0 and
// row = INT_MAX stand for "no candidate": a candidate of value 0 and none at all lead to the same decision (the
// diagonal is kept unless a candidate is STRICTLY larger).
__device__ __forceinline__ void v4_argmax(unsigned& hi, unsigned& lo, int& row)
{
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const bool top = hi == mh && lo == ml;
    row = int(__reduce_min_sync(0xffffffffu, top ? unsigned(row) : unsigned(INT_MAX)));
    hi = mh; lo = ml;
}

template <typename T>
__global__ void __launch_bounds__(V4_THREADS, 1)
getrf_base_v4_kernel(const V3Args<T> a)
{
    __shared__ __align__(16) T s_prow[PW];
    __shared__ __align__(16) T s_drow[PW];
    constexpr int NWARP = V4_THREADS / 32;
    __shared__ unsigned s_hi[NWARP], s_lo[NWARP];
    __shared__ int      s_row[NWARP];
    __shared__ unsigned g_hi[NWARP], g_lo[NWARP];
    __shared__ int      g_row[NWARP];
    const int G = a.G, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nb = a.nb, w = a.w;

    if (b >= G) {
        // ---- interchange CTAs: follow the pivot sequence, swap rows d and p in every panel column outside the block
        const int wi = b - G, nw = int(gridDim.x) - G;
        const int outside = a.kw_wide - w;
        for (int j = 0; j < w; ++j) {
            const int d = a.c0 + j;
            const int p = int(v3_wait_word(a.pivrec + j, a.gen_base + unsigned(j) + 1u));
            if (p == d) continue;
            T* rd_ = a.tiles[d / nb] + (d % nb);
            T* rp_ = a.tiles[p / nb] + (p % nb);
            for (int ci = wi * V4_THREADS + tid; ci < outside; ci += nw * V4_THREADS) {
                const int c = ci < a.c0 ? ci : ci + w;
                const T t0 = rd_[int64_t(c) * nb], t1 = rp_[int64_t(c) * nb];
                rd_[int64_t(c) * nb] = t1;
                rp_[int64_t(c) * nb] = t0;
            }
            if (a.rowmap && wi == 0 && tid == 0) { const int t = a.rowmap[d]; a.rowmap[d] = a.rowmap[p]; a.rowmap[p] = t; }
        }
        return;
    }

    const int r = a.c0 + b * a.rows_per + tid;                 // this thread's panel row
    const bool have = tid < a.rows_per && r < a.m_p;
    T* rowp = have ? a.tiles[r / nb] + (r % nb) + int64_t(a.c0) * nb : nullptr;
    T x[PW];
    __shared__ T sL[PW][PW + 1];                               // fused update: L11 | A12 -> U12
    __shared__ __align__(16) T sU[PW][PW];

    if (a.upd_c0 >= 0) {
        // ---- fused update by the previous 32-column block [u0, u0 + 32) (the w1 = n2 = 32 updates of the recursive panel:
        //      8 of the 15 per nb = 512 panel, each a triangular-solve launch + a tile-GEMM launch before):
        //      U12 = L11^-1 A12 (every CTA, redundantly: 32 x 32), then this thread's row: x -= L21(r, :) U12.
        //      The interchange CTAs of THIS launch only touch L21 after the first pivot is known, i.e. after every
        //      row CTA has published its first candidate, i.e. after it has finished reading L21 here.
        const int u0 = a.upd_c0;
        for (int e = tid; e < PW * PW; e += V4_THREADS) {
            const int i = e % PW, k = e / PW;
            const int rr = u0 + i;
            const T* base = a.tiles[rr / nb] + (rr % nb);
            sL[i][k] = base[int64_t(u0 + k) * nb];
            sU[i][k] = k < w ? base[int64_t(a.c0 + k) * nb] : T(0);
        }
        __syncthreads();
        if (warp == 0) {
            // lane c owns column c of U12: forward substitution with the unit-lower L11, in place
            #pragma unroll
            for (int i = 1; i < PW; ++i) {
                T sacc = sU[i][lane];
                #pragma unroll
                for (int k = 0; k < i; ++k) sacc = v3_fma(-sL[i][k], sU[k][lane], sacc);
                sU[i][lane] = sacc;
            }
        }
        __syncthreads();
        #pragma unroll
        for (int c = 0; c < PW; ++c) x[c] = (have && c < w) ? rowp[int64_t(c) * nb] : T(0);
        if (have) {
            const T* lrow = a.tiles[r / nb] + (r % nb) + int64_t(u0) * nb;
            #pragma unroll 1
            for (int k0 = 0; k0 < PW; k0 += 8) {
                T l[8];
                #pragma unroll
                for (int kk = 0; kk < 8; ++kk) l[kk] = lrow[int64_t(k0 + kk) * nb];
                #pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    #pragma unroll
                    for (int c = 0; c < PW; ++c) x[c] = v3_fma(-l[kk], sU[k0 + kk][c], x[c]);
                }
            }
        }
    }
    else {
        #pragma unroll
        for (int c = 0; c < PW; ++c) x[c] = (have && c < w) ? rowp[int64_t(c) * nb] : T(0);
    }

    // candidate of the first column (NaN is never a candidate: the reference's `abs > max` is false for it)
    unsigned khi = 0, klo = 0;
    int krow = INT_MAX;
    {
        const double v = double(fabs(x[0]));
        if (have && r > a.c0 && v == v) { khi = unsigned(__double2hiint(v)); klo = unsigned(__double2loint(v)); krow = r; }
    }

    #pragma unroll
    for (int j = 0; j < PW; ++j) {
        if (j < w) {
            const int d = a.c0 + j;
            const unsigned gen = a.gen_base + unsigned(j) + 1u;
            unsigned long long* slot = a.rec + size_t(j) * a.gmax * V3_REC;
            // ---- this CTA's candidate: warp, then block
            v4_argmax(khi, klo, krow);
            if (lane == 0) { s_hi[warp] = khi; s_lo[warp] = klo; s_row[warp] = krow; }
            __syncthreads();
            khi = lane < NWARP ? s_hi[lane] : 0u;
            klo = lane < NWARP ? s_lo[lane] : 0u;
            krow = lane < NWARP ? s_row[lane] : INT_MAX;
            v4_argmax(khi, klo, krow);
            // ---- publish: the owner of the candidate row sends header + row, the owner of row d the diagonal row
            {
                unsigned long long* R = slot + size_t(b) * V3_REC;
                if (krow == INT_MAX ? tid == 0 : (have && r == krow)) {
                    if (krow != INT_MAX) {
                        #pragma unroll
                        for (int c = 0; c < PW; ++c) v3_store_double(R + V3_HDR + 2 * c, double(x[c]), gen);
                    }
                    const unsigned long long gg = static_cast<unsigned long long>(gen) << 32;
                    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(R), "l"(gg | klo), "l"(gg | khi) : "memory");
                    v3_store(R + 2, unsigned(krow), gen);
                }
                if (b == 0 && tid == j) {
                    #pragma unroll
                    for (int c = 0; c < PW; ++c) v3_store_double(a.drow + (size_t(j) * PW + c) * 2, double(x[c]), gen);
                }
            }
            // ---- gather: thread c polls the header of CTA c; the last warp polls the diagonal row
            unsigned bhi = 0, blo = 0;
            int br = INT_MAX;
            if (tid < G) {
                const unsigned long long* R = slot + size_t(tid) * V3_REC;
                unsigned long long w0, w1, w2, w3;
                const long long t0 = clock64();
                for (;;) {
                    v3_load2(R, w0, w1); v3_load2(R + 2, w2, w3);
                    if (unsigned(w0 >> 32) == gen && unsigned(w1 >> 32) == gen && unsigned(w2 >> 32) == gen) break;
                    spin_watchdog(t0);
                }
                blo = unsigned(w0); bhi = unsigned(w1);
                br = int(unsigned(w2));
            }
            if (warp == NWARP - 1)
                s_drow[lane] = T(v3_wait_double(a.drow + (size_t(j) * PW + lane) * 2, gen));
            if (warp * 32 < G) v4_argmax(bhi, blo, br);
            if (lane == 0) { g_hi[warp] = bhi; g_lo[warp] = blo; g_row[warp] = br; }
            __syncthreads();
            bhi = lane < NWARP ? g_hi[lane] : 0u;
            blo = lane < NWARP ? g_lo[lane] : 0u;
            br = lane < NWARP ? g_row[lane] : INT_MAX;
            v4_argmax(bhi, blo, br);
            // ---- every CTA picks the same winner: strictly larger than the diagonal, or the diagonal (ties, NaN)
            const double bv = __hiloint2double(int(bhi), int(blo));
            const double dv = double(fabs(s_drow[j]));
            const int p = (br != INT_MAX && bv > dv) ? br : d;
            if (warp == 0) {
                const int bw = (p - a.c0) / a.rows_per;                 // the CTA that holds row p
                s_prow[lane] = (p == d) ? s_drow[lane]
                                        : T(v3_wait_double(slot + size_t(bw) * V3_REC + V3_HDR + 2 * lane, gen));
            }
            if (b == 0 && tid == 2 * 32) {
                a.piv_tile[d] = p / nb;
                a.piv_off[d] = p % nb;
                v3_store(a.pivrec + j, unsigned(p), gen);
            }
            __syncthreads();
            // ---- interchange (registers of the two owning threads), scale, rank-1 update
            if (p != d) {
                if (have && r == p) {
                    #pragma unroll
                    for (int c = 0; c < PW; ++c) x[c] = s_drow[c];
                }
                if (b == 0 && tid == j) {
                    #pragma unroll
                    for (int c = 0; c < PW; ++c) x[c] = s_prow[c];
                }
            }
            const T pv = s_prow[j];
            if (pv == T(0)) {
                if (b == 0 && tid == 0 && *a.info == 0) *a.info = a.info_base + d + 1;
            }
            else if (have && r > d) {
                const bool use_rcp = fabs(pv) >= v3_tiny<T>();
                const T rcp = T(1) / pv;
                const T l = use_rcp ? x[j] * rcp : x[j] / pv;
                x[j] = l;
                #pragma unroll
                for (int c = j + 1; c < PW; ++c) x[c] = v3_fma(-l, s_prow[c], x[c]);
            }
            if (j + 1 < PW) {
                const double v = double(fabs(x[j + 1 < PW ? j + 1 : j]));
                const bool cand = have && r > d + 1 && v == v;
                khi = cand ? unsigned(__double2hiint(v)) : 0u;
                klo = cand ? unsigned(__double2loint(v)) : 0u;
                krow = cand ? r : INT_MAX;
            }
        }
    }
    if (have) {
        #pragma unroll
        for (int c = 0; c < PW; ++c)
            if (c < w) rowp[int64_t(c) * nb] = x[c];
    }
    // U12 goes back to the panel only now: every row CTA read A12 from that place before it published its first
    // candidate, and CTA 0 has seen all of those
    if (a.upd_c0 >= 0 && b == 0)
        for (int e = tid; e < PW * PW; e += V4_THREADS) {
            const int i = e % PW, k = e / PW;
            const int rr = a.upd_c0 + i;
            if (k < w) (a.tiles[rr / nb] + (rr % nb))[int64_t(a.c0 + k) * nb] = sU[i][k];
        }
}


// (A persistent variant -- up to 8 consecutive 32-column blocks in ONE cooperative launch, left-looking between the
// blocks: every CTA recomputes U12 = L11^-1 A12 and updates its rows with FP64 FMAs -- was written, validated (244 parity
// tests) and measured in round 2, and deleted: pnl_base + pnl_trsm + pnl_gemm = 273 / 280 / 250 / 268 ms at n = 32768 for
// 1 / 2 / 4 / 8 blocks per launch, dgetrf 822 / 831 / 830 / 847 ms.  Next to the trailing update a column costs 1.5 us
// more than on an idle device whatever the launch structure -- the exchange's L2 round trips compete with the GEMM's
// traffic -- and the launch itself only ~20 us; profiles/r02l_persistent_lu_kernel_negative_result.txt.)

} // namespace

// SB200_PANEL_V4=0: rows in shared memory (the fallback for panels taller than (#SMs - NWIDE) x 512) for every panel;
// read per call so that a test can switch it
static bool v4_enabled()
{
    const char* e = getenv("SB200_PANEL_V4");
    return ! e || atoi(e) != 0;
}

size_t base_v3_scratch_bytes(int max_ctas)
{
    return (size_t(PW) * max_ctas * V3_REC + size_t(PW) * 2 * PW + PW) * sizeof(unsigned long long);
}

int base_v3_init()
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    static thread_local bool done[64] = {};
    if (! done[dev & 63]) {
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_v3_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(PW * (PROWS_MAX | 1) * sizeof(double))));
        CUDA_TRY(cudaFuncSetAttribute(getrf_base_v3_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(PW * (PROWS_MAX | 1) * sizeof(float))));

        done[dev & 63] = true;
    }
    return SB200_OK;
}

// columns [c0, c0+w) of the panel over rows [c0, m_p); interchanges applied panel-wide (kw columns) by the extra CTAs
template <typename T>
int launch_base_v3(T* const* stack, int nb, int m_p, int c0, int w, int kw, int64_t* piv_tile, int64_t* piv_off,
                   int* dinfo, int info_base, int* rowmap, PanelScratch& ps, cudaStream_t s, int upd_c0)
{
    const int active = m_p - c0;
    const int ctas = ps.max_ctas - V3_NWIDE;
    const bool v4 = v4_enabled() && active <= ctas * V4_THREADS;    // register-resident rows: <= 512 rows per CTA
    int rows_per = v4 ? V4_THREADS : std::max(int(ceil_div(active, ctas)), std::min(active, PROWS_MAX));
    rows_per = std::max(rows_per, PW);
    if (rows_per > PROWS_MAX) return SB200_ENOTSUP;
    if (upd_c0 >= 0 && ! v4) return SB200_EINVAL;             // callers ask base_v3_can_fuse first
    const int G = int(ceil_div(active, rows_per));
    if (ps.v3_gen > 0xF0000000u) {                        // tags about to wrap: start over from clean slots
        CUDA_TRY(cudaMemsetAsync(ps.v3_buf, 0, base_v3_scratch_bytes(ps.max_ctas), s));
        ps.v3_gen = 0;
    }
    V3Args<T> a{};
    a.tiles = stack; a.nb = nb; a.m_p = m_p; a.c0 = c0; a.w = w; a.rows_per = rows_per; a.G = G;
    a.piv_tile = piv_tile; a.piv_off = piv_off; a.info = dinfo; a.info_base = info_base;
    a.rowmap = rowmap; a.kw_wide = kw; a.upd_c0 = upd_c0;
    a.rec = ps.v3_buf;
    a.drow = a.rec + size_t(PW) * ps.max_ctas * V3_REC;
    a.pivrec = a.drow + size_t(PW) * 2 * PW;
    a.gen_base = ps.v3_gen; a.gmax = ps.max_ctas;
    ps.v3_gen += unsigned(PW);
    const size_t smem = v4 ? 0 : size_t(w) * (rows_per | 1) * sizeof(T);
    static const bool trace_on = [] { const char* e = getenv("SB200_V3_TRACE"); return e && atoi(e) != 0; }();
    long long* dtrace = nullptr;
    if (trace_on && ! v4) { CUDA_TRY(cudaMalloc(&dtrace, size_t(G) * 8 * sizeof(long long))); a.trace = dtrace; }
    void* args[] = {&a};
    const cudaError_t e = v4
        ? cudaLaunchCooperativeKernel(reinterpret_cast<void*>(getrf_base_v4_kernel<T>), dim3(G + V3_NWIDE), dim3(V4_THREADS), args, 0, s)
        : cudaLaunchCooperativeKernel(reinterpret_cast<void*>(getrf_base_v3_kernel<T>), dim3(G + V3_NWIDE), dim3(PTHREADS), args, smem, s);
    if (e != cudaSuccess) return int(e);
    if (trace_on && ! v4) {
        // debug: cycles per phase (thread 0 of the first, a middle and the last row CTA), summed over the w columns
        std::vector<long long> h(size_t(G) * 8);
        CUDA_TRY(cudaStreamSynchronize(s));
        CUDA_TRY(cudaMemcpy(h.data(), dtrace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(dtrace);
        static int printed = 0;
        if (printed++ < 40)
            for (int c : {0, G / 2, G - 1})
                fprintf(stderr, "{\"v3_trace\": 1, \"m\": %d, \"c0\": %d, \"w\": %d, \"G\": %d, \"cta\": %d, \"load\": %lld, \"cand\": %lld, "
                        "\"poll\": %lld, \"greduce\": %lld, \"winrow\": %lld, \"swap\": %lld, \"update\": %lld, \"store\": %lld}\n",
                        m_p, c0, w, G, c, h[c * 8 + 0], h[c * 8 + 1], h[c * 8 + 2], h[c * 8 + 3], h[c * 8 + 4], h[c * 8 + 5], h[c * 8 + 6], h[c * 8 + 7]);
    }
    return launch_status();
}

template int launch_base_v3<double>(double* const*, int, int, int, int, int, int64_t*, int64_t*, int*, int, int*, PanelScratch&, cudaStream_t, int);
template int launch_base_v3<float>(float* const*, int, int, int, int, int, int64_t*, int64_t*, int*, int, int*, PanelScratch&, cudaStream_t, int);

// can the block [c0, c0 + w) of an m_p-row panel take its update by the previous 32-column block inside its own launch?
bool base_v3_can_fuse(const PanelScratch& ps, int m_p, int c0, int w1, int w)
{
    const char* e = getenv("SB200_PANEL_FUSE");
    const bool fuse = ! e || atoi(e) != 0;
    return ps.use_v3 && v4_enabled() && fuse && w1 == PW && w >= 1 && w <= PW
        && (m_p - c0) <= (ps.max_ctas - V3_NWIDE) * V4_THREADS && (m_p - c0) >= 1;
}


} // namespace sb200
