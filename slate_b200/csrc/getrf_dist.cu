// getrf_dist.cu -- LU with partial pivoting on a p x q process grid, one process per GPU, NCCL.
//
// Reference: src/getrf.cc:22-244 (panel task -> tileBcast of the panel along rows, permuteRows on
// every block column with MPI row exchanges (src/internal/internal_swap.cc:510-805), trsm on block
// row k, listBcast of U(k, j) down the columns, trailing gemm with lookahead).
//
// B200 / NVSwitch-first restatement (every GPU reaches every peer at full bandwidth, so fewer and
// larger exchanges win over the reference's per-row messages):
//   1. PANEL: the tiles of block column k are gathered on the owner of A(k,k) (<= 128 MiB at
//      n = 65536), factored there by the cooperative GPU panel kernel (getrf.cu), and broadcast --
//      tiles, pivots and the row map -- to ALL ranks in one NCCL group.  Every rank keeps the factored
//      panel in a double-buffered workspace (pws[k & 1]) and reads L(i,k) from it.
//   2. ROW INTERCHANGES are not applied as nb sequential swaps.  The panel kernel tracks where each
//      original row ends up (rowmap); the net effect of all nb swaps is
//          new top row j      <- original row src[j]            (anywhere in the panel)
//          lower row dst[j]   <- original TOP row tsrc[j]
//      so one all-gather inside the process column delivers the nb rows that move into the top block
//      (each rank contributes the rows it owns) and one broadcast delivers the displaced top rows.
//      All rows and columns move in parallel.
//   3. U(k, j) = L_kk^-1 A(k, j) is computed REDUNDANTLY by every rank of the process column from the
//      all-gathered top block (0.14 ms of DMMA work) instead of a second broadcast down the column.
//   4. Trailing update: one batched DMMA GEMM launch per shape class, A operand = pws, B operand = the
//      U workspace; lookahead column k+1 on the high-priority stream, the rest on the trailing stream.
// Collectives of the two streams use different communicators (col_comm / col_comm2).
//
// The driver is a template over float / double.  T = float is the low-precision factorisation of gesv_mixed on
// the grid; with use_tc05 its trailing / lookahead GEMMs run on the tcgen05 FP32-emulated kernel, the panel
// workspace tiles (A role) and the U slots (B role) being split-packed once per step.
// STATUS of T = float: validated on ONE GPU through the test hook SB200_GETRF_DIST=1 (the p x q algorithm minus the
// NCCL calls) and on 1x2, 2x1 and 2x4 grids (scratch/mgpu_check.py: gesv_mixed; profiles/r02m2_*, r02g8_*).
#include "runtime_internal.hh"
#include "getrf_internal.hh"
#include <algorithm>
#include <climits>
#include <cstdio>
#include <type_traits>
#include <vector>

namespace sb200 {

// geometry shared by the permutation kernels: local tile (il, jl) = pool + (jl*mt_loc + il)*te
template <typename T>
struct PermGeom {
    T* pool;
    int64_t te;
    int mt_loc, nt_loc, nb, p, q, prow, pcol;
    int64_t m, n;
    int k;              // panel step: panel row s lives in tile row k + s/nb, offset s % nb
    int ntop;           // rows of the top block (= pivots of this panel)
    int skip0, skip1;   // global block columns this call leaves alone (-1: none)
    const int* src;     // [ntop] original panel row that ends at top position j
    const int* dst;     // [ntop] panel row swapped with j (the pivot row)
    const int* tsrc;    // [ntop] original top row that ends at row dst[j] (when dst[j] >= ntop)
    __device__ int ncols(int jl) const { const int64_t j = pcol + int64_t(jl) * q; return int(min(int64_t(nb), n - j * nb)); }
    __device__ bool active(int jl) const { const int j = pcol + jl * q; return j != skip0 && j != skip1; }
    __device__ T* tile(int i, int jl) const { return pool + (int64_t(jl) * mt_loc + (i - prow) / p) * te; }
};

__global__ void iota_kernel(int* v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// perm = [src | dst | tsrc], ntop ints each
__global__ void perm_pack_kernel(const int* __restrict__ rowmap, const int64_t* __restrict__ piv_tile,
                                 const int64_t* __restrict__ piv_off, int nb, int ntop, int* __restrict__ perm)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ntop) return;
    const int d = int(piv_tile[j] * nb + piv_off[j]);
    perm[j] = rowmap[j];
    perm[ntop + j] = d;
    perm[2 * ntop + j] = rowmap[d];
}

// gather[slot jl - jl0][c*nb + pos] = A(row src[pos], column c of local block column jl), for the rows
// this rank owns.  grid = (column chunks of 16, block columns); threads run over pos (coalesced writes).
template <typename T>
__global__ void __launch_bounds__(256)
perm_gather_kernel(const PermGeom<T> g, int jl0, T* __restrict__ out)
{
    const int jl = jl0 + blockIdx.y;
    if (! g.active(jl)) return;
    const int nc = g.ncols(jl);
    const int c0 = blockIdx.x * 16, c1 = min(c0 + 16, nc);
    T* o = out + int64_t(blockIdx.y) * g.te;
    for (int pos = threadIdx.x; pos < g.ntop; pos += blockDim.x) {
        const int s = g.src[pos];
        const int i = g.k + s / g.nb;
        if (i % g.p != g.prow) continue;
        const T* a = g.tile(i, jl) + (s % g.nb);
        for (int c = c0; c < c1; ++c) o[int64_t(c) * g.nb + pos] = a[int64_t(c) * g.nb];
    }
}

// U workspace slot <- rows picked from the all-gathered buffers: row pos comes from the rank that owns
// original row src[pos].  all = [p][nslots][te].
template <typename T>
__global__ void __launch_bounds__(256)
perm_select_top_kernel(const PermGeom<T> g, int jl0, int nslots, const T* __restrict__ all, T* __restrict__ U)
{
    const int jl = jl0 + blockIdx.y;
    if (! g.active(jl)) return;
    const int nc = g.ncols(jl);
    const int c0 = blockIdx.x * 16, c1 = min(c0 + 16, nc);
    for (int pos = threadIdx.x; pos < g.ntop; pos += blockDim.x) {
        const int owner = (g.k + g.src[pos] / g.nb) % g.p;
        const T* a = all + (int64_t(owner) * nslots + blockIdx.y) * g.te;
        T* u = U + int64_t(jl) * g.te;
        for (int c = c0; c < c1; ++c) u[int64_t(c) * g.nb + pos] = a[int64_t(c) * g.nb + pos];
    }
}

// lower rows: row dst[pos] <- original top row tsrc[pos] (from the broadcast copy of the old top block)
template <typename T>
__global__ void __launch_bounds__(256)
perm_scatter_lower_kernel(const PermGeom<T> g, int jl0, const T* __restrict__ oldtop)
{
    const int jl = jl0 + blockIdx.y;
    if (! g.active(jl)) return;
    const int nc = g.ncols(jl);
    const int c0 = blockIdx.x * 16, c1 = min(c0 + 16, nc);
    const T* o = oldtop + int64_t(blockIdx.y) * g.te;
    for (int pos = threadIdx.x; pos < g.ntop; pos += blockDim.x) {
        const int d = g.dst[pos];
        if (d < g.ntop) continue;
        const int i = g.k + d / g.nb;
        if (i % g.p != g.prow) continue;
        T* a = g.tile(i, jl) + (d % g.nb);
        const int t = g.tsrc[pos];
        for (int c = c0; c < c1; ++c) a[int64_t(c) * g.nb] = o[int64_t(c) * g.nb + t];
    }
}

// rows [0, ntop) of tile (k, jl) <-> buffer slot; dir 0: tile -> buf (save old top), 1: buf -> tile
template <typename T>
__global__ void __launch_bounds__(256)
top_copy_kernel(const PermGeom<T> g, int jl0, T* __restrict__ buf, int buf_by_jl, int dir)
{
    const int jl = jl0 + blockIdx.y;
    if (! g.active(jl)) return;
    const int nc = g.ncols(jl);
    T* a = g.tile(g.k, jl);
    T* b = buf + int64_t(buf_by_jl ? jl : int(blockIdx.y)) * g.te;
    for (int c = blockIdx.x; c < nc; c += gridDim.x)
        for (int r = threadIdx.x; r < g.ntop; r += blockDim.x) {
            const int64_t e = int64_t(c) * g.nb + r;
            if (dir) a[e] = b[e]; else b[e] = a[e];
        }
}

template <typename T>
static int getrf_driver_dist_t(Matrix& A, int64_t* pivots_out, int64_t* info_out, bool use_tc05)
{
    Grid& g = *A.g;
    if (A.kind != 'G' || A.layout != 'C' || A.dtype != TypeChar<T>::value) return SB200_EINVAL;
    constexpr bool is_float = std::is_same<T, float>::value;
    if (use_tc05 && ! is_float) return SB200_EINVAL;
    using DBuf = DevBuf;
    const ncclDataType_t nccl_t = is_float ? ncclFloat : ncclDouble;
    HostTimes htm("getrf_dist");
    CUDA_TRY(cudaDeviceSynchronize());
    htm.mark("entry_sync");
    const int64_t mt = A.mt, nt = A.nt, nb = A.nb, te = A.tile_elems();
    const int p = g.p, q = g.q, prow = g.prow, pcol = g.pcol;
    const int ld = int(nb);
    const int64_t kt = std::min(mt, nt);
    const int64_t mn = std::min(A.m, A.n);
    if (mn == 0) { if (info_out) *info_out = 0; return SB200_OK; }
    const bool multi = g.size() > 1;
    const int mt_loc = int(A.mt_loc), nt_loc = int(A.nt_loc);

    // ---- panel workspace slots: tiles i >= k of block column k, grouped by owning process row
    auto first_row = [&](int r, int64_t k) { return k + ((r - k) % p + p) % p; };            // first i >= k, i % p == r
    auto count_rows = [&](int r, int64_t k) { const int64_t i0 = first_row(r, k); return i0 < mt ? (mt - 1 - i0) / p + 1 : 0; };
    auto region_off = [&](int r, int64_t k) { int64_t o = 0; for (int x = 0; x < r; ++x) o += count_rows(x, k); return o; };
    auto slot = [&](int64_t i, int64_t k) { const int r = int(i % p); return region_off(r, k) + (i - first_row(r, k)) / p; };

    DBuf pws, uws, ula, gmine, gall, oldtop, gmineP, gallP, oldtopP, permb, rowmapb, pivb, infob, wt, planb, stackb;
    SB_TRY(pws.alloc(size_t(2) * mt * te * sizeof(T)));
    SB_TRY(uws.alloc(size_t(std::max(nt_loc, 1)) * te * sizeof(T)));
    SB_TRY(ula.alloc(size_t(te) * sizeof(T)));
    // SB200_GEMM_BT (default 1; double, not the tcgen05 path): transposed copies of the U
    // slots, written once per step after the row solve, so that the trailing update runs as 'N','T' (see getrf.cu)
    bool use_bt = false;
    if constexpr (std::is_same<T, double>::value) {
        use_bt = switch_value(SW_GEMM_BT) != 0 && ! use_tc05;
    }
    DBuf uwsT;
    if (use_bt) SB_TRY(uwsT.alloc(size_t(std::max(nt_loc, 1)) * te * sizeof(T)));
    SB_TRY(gmine.alloc(size_t(std::max(nt_loc, 1)) * te * sizeof(T)));
    SB_TRY(oldtop.alloc(size_t(std::max(nt_loc, 1)) * te * sizeof(T)));
    SB_TRY(gmineP.alloc(size_t(te) * sizeof(T)));
    SB_TRY(oldtopP.alloc(size_t(te) * sizeof(T)));
    if (p > 1) {
        SB_TRY(gall.alloc(size_t(p) * std::max(nt_loc, 1) * te * sizeof(T)));
        SB_TRY(gallP.alloc(size_t(p) * te * sizeof(T)));
    }
    SB_TRY(permb.alloc(size_t(2) * 3 * nb * sizeof(int)));
    SB_TRY(rowmapb.alloc(size_t(A.m) * sizeof(int)));
    SB_TRY(pivb.alloc(size_t(2 * kt * nb) * sizeof(int64_t)));
    SB_TRY(infob.alloc(sizeof(int)));
    SB_TRY(wt.alloc(size_t(ceil_div(nb, 64)) * 64 * 64 * sizeof(T)));
    int64_t* dpiv_tile = pivb.as<int64_t>();
    int64_t* dpiv_off = dpiv_tile + kt * nb;
    auto pws_tile = [&](int64_t i, int64_t k) { return pws.as<T>() + ((k & 1) * mt + slot(i, k)) * te; };
    // tcgen05 path: split-packed copies of the panel workspace tiles (A role, [k & 1][slot]) and of the U slots (B role)
    const size_t pa_bytes = use_tc05 ? tc05_packed_bytes('A', nb, nb) : 0, pb_bytes = use_tc05 ? tc05_packed_bytes('B', nb, nb) : 0;
    DBuf pkA, pkB, pkBla;
    if (use_tc05) {
        SB_TRY(pkA.alloc(size_t(2) * mt * pa_bytes));
        SB_TRY(pkB.alloc(size_t(std::max(nt_loc, 1)) * pb_bytes));
        SB_TRY(pkBla.alloc(pb_bytes));
    }
    auto pkA_of = [&](int64_t i, int64_t k) { return pkA.as<unsigned char>() + ((k & 1) * mt + slot(i, k)) * pa_bytes; };

    // ---- plan: GEMM pointer batches per step, panel stacks (root's view), U-slot pointer arrays
    struct Step {
        std::vector<Batch> la, tr; size_t stack_off = 0;
        size_t a_src = 0, a_dst = 0, b_src = 0, b_dst = 0;       // pack lists (tcgen05 path): full-size operands first
        int a_cnt = 0, a_full = 0, b_cnt = 0, b_full = 0;
    };
    std::vector<Step> steps(static_cast<size_t>(kt));
    std::vector<const void*> hp;
    for (int64_t k = 0; k < kt; ++k) {
        const int kw = int(A.tile_nb(k));
        for (int64_t j = k + 1; j < nt; ++j) {
            if (int(j % q) != pcol) continue;
            const int64_t jl = (j - pcol) / q;
            const T* Bop = (j == k + 1) ? ula.as<T>() : (use_bt ? uwsT.as<T>() : uws.as<T>()) + jl * te;
            for (int64_t i = k + 1; i < mt; ++i) {
                if (int(i % p) != prow) continue;
                batch_add(j == k + 1 ? steps[k].la : steps[k].tr, int(A.tile_mb(i)), int(A.tile_nb(j)), kw, 0,
                          use_tc05 ? static_cast<const void*>(pkA_of(i, k)) : static_cast<const void*>(pws_tile(i, k)),
                          use_tc05 ? static_cast<const void*>(j == k + 1 ? pkBla.as<unsigned char>() : pkB.as<unsigned char>() + jl * pb_bytes)
                                   : static_cast<const void*>(Bop),
                          A.tile_as<T>(i, j));
            }
        }
        for (auto* lst : {&steps[k].la, &steps[k].tr})
            for (auto& b : *lst) {
                b.off = hp.size();
                hp.insert(hp.end(), b.A.begin(), b.A.end());
                hp.insert(hp.end(), b.B.begin(), b.B.end());
                hp.insert(hp.end(), b.C.begin(), b.C.end());
            }
        if (use_tc05 && k + 1 < nt) {
            std::vector<const void*> src; std::vector<void*> dst;
            for (int64_t i = k + 1; i < mt; ++i)
                if (int(i % p) == prow) {
                    src.push_back(pws_tile(i, k)); dst.push_back(pkA_of(i, k));
                    ++steps[k].a_cnt; if (A.tile_mb(i) == nb) ++steps[k].a_full;
                }
            steps[k].a_src = hp.size(); hp.insert(hp.end(), src.begin(), src.end());
            steps[k].a_dst = hp.size(); hp.insert(hp.end(), dst.begin(), dst.end());
            src.clear(); dst.clear();
            for (int64_t j = k + 2; j < nt; ++j)
                if (int(j % q) == pcol) {
                    const int64_t jl = (j - pcol) / q;
                    src.push_back(uws.as<T>() + jl * te); dst.push_back(pkB.as<unsigned char>() + jl * pb_bytes);
                    ++steps[k].b_cnt; if (A.tile_nb(j) == nb) ++steps[k].b_full;
                }
            steps[k].b_src = hp.size(); hp.insert(hp.end(), src.begin(), src.end());
            steps[k].b_dst = hp.size(); hp.insert(hp.end(), dst.begin(), dst.end());
        }
        if (use_bt && k + 1 < mt) {                       // U slots of the trailing columns -> their transposes (full-size first)
            std::vector<const void*> src; std::vector<void*> dst;
            for (int64_t j = k + 2; j < nt; ++j)
                if (int(j % q) == pcol) {
                    const int64_t jl = (j - pcol) / q;
                    src.push_back(uws.as<T>() + jl * te); dst.push_back(uwsT.as<T>() + jl * te);
                    ++steps[k].b_cnt; if (A.tile_nb(j) == nb) ++steps[k].b_full;
                }
            steps[k].b_src = hp.size(); hp.insert(hp.end(), src.begin(), src.end());
            steps[k].b_dst = hp.size(); hp.insert(hp.end(), dst.begin(), dst.end());
        }
        // panel stack as seen by the root of step k: own tiles in place, the others in pws
        if (g.rank == g.rank_of(k, k)) {
            steps[k].stack_off = hp.size();
            for (int64_t i = k; i < mt; ++i)
                hp.push_back(int(i % p) == prow ? A.tile_as<T>(i, k) : pws_tile(i, k));
        }
    }
    const size_t uptr_off = hp.size();                 // U-slot pointers: uws + jl*te
    for (int jl = 0; jl < nt_loc; ++jl) hp.push_back(uws.as<T>() + int64_t(jl) * te);
    const size_t ula_off = hp.size();
    hp.push_back(ula.as<T>());
    SB_TRY(planb.alloc(hp.size() * sizeof(void*)));
    void** dplan = planb.as<void*>();

    htm.mark("plan_and_buffers");
    PanelScratch ps;
    SB_TRY(ps.init());
    if (ps.tnt_ranks > 0 && (use_tc05 || ! tnt_shape_supported(A))) return SB200_ENOTSUP;       // getrf_tntpiv (getrf_tnt.cu)
    TntScratch tnt;
    cudaStream_t P = nullptr, T_ = nullptr;
    int lo, hi;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&P, cudaStreamNonBlocking, hi));
    CUDA_TRY(cudaStreamCreateWithPriority(&T_, cudaStreamNonBlocking, lo));
    std::vector<cudaEvent_t> ev(size_t(2 * kt)), tev, pev;
    auto ptime = [&](cudaStream_t st) -> int {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        pev.push_back(e);
        CUDA_TRY(cudaEventRecord(e, st));
        return SB200_OK;
    };
    for (auto& e : ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaEvent_t t0, t1;
    CUDA_TRY(cudaEventCreate(&t0)); CUDA_TRY(cudaEventCreate(&t1));
    auto P_done = [&](int64_t k) { return ev[size_t(k)]; };
    auto T_done = [&](int64_t k) { return ev[size_t(kt + k)]; };
    double trail_flops = 0; int64_t trail_launches = 0;
    PhaseTimer ph;

    PlanBuffer pbv;                      // non-owning view of the device plan for the shared batch launchers
    struct PbvGuard { PlanBuffer& v; ~PbvGuard() { v.dev = nullptr; } } pbv_guard{pbv};
    pbv.dev = dplan;
    auto run_batches = [&](const std::vector<Batch>& bs, cudaStream_t s, bool b_transposed = false) -> int {
        if constexpr (is_float) {
            if (use_tc05) return launch_batches_tc05(bs, pbv, -1.0f, 1.0f, ld, s);
        }
        return launch_batches<T>(bs, pbv, 'N', b_transposed ? 'T' : 'N', T(-1), T(1), ld, 0, s);
    };
    // split-pack `cnt` operands given by device pointer arrays (tcgen05 path)
    auto pack = [&](int role, const void* const* src, void* const* dst, int cnt, int rows, int kdim, cudaStream_t s) -> int {
        if constexpr (is_float) {
            if (cnt <= 0) return SB200_OK;
            Tc05PackParams qp{};
            qp.X = reinterpret_cast<const float* const*>(src);
            qp.P = dst;
            qp.rows = rows; qp.k = kdim;
            if (role == 'A') { qp.rs = 1; qp.ks = ld; qp.ru = TC_BM; }
            else             { qp.rs = ld; qp.ks = 1; qp.ru = TC_BN; }
            qp.batch = cnt;
            return launch_tc05_pack(qp, s);
        }
        (void) role; (void) src; (void) dst; (void) cnt; (void) rows; (void) kdim; (void) s;
        return SB200_OK;
    };

    // Row permutation of panel k applied to local block columns [jl0, jl1) (minus skip0/skip1), then
    // U(k, j) for the columns right of the panel: Uout slots <- L_kk^-1 * new top block.
    //   by_jl: U slot index = jl (trailing workspace) or slot 0 (lookahead, single column)
    auto permute_and_solve = [&](int64_t k, int jl0, int jl1, int skip0, int skip1, bool lookahead,
                                 cudaStream_t s, ncclComm_t comm, T* W) -> int {
        const int ns = jl1 - jl0;
        if (ns <= 0) return SB200_OK;
        const int kp = int(k % p);
        const int m_p = int(A.m - k * nb), kw = int(A.tile_nb(k));
        const int ntop = std::min(m_p, kw);
        const int* perm = permb.as<int>() + (k & 1) * 3 * nb;
        PermGeom<T> pg{reinterpret_cast<T*>(A.pool), te, mt_loc, nt_loc, int(nb), p, q, prow, pcol, A.m, A.n, int(k), ntop, skip0, skip1,
                    perm, perm + ntop, perm + 2 * ntop};
        T* mine = lookahead ? gmineP.as<T>() : gmine.as<T>();
        T* all  = p > 1 ? (lookahead ? gallP.as<T>() : gall.as<T>()) : mine;
        T* old  = lookahead ? oldtopP.as<T>() : oldtop.as<T>();
        T* U    = lookahead ? ula.as<T>() - int64_t(jl0) * te : uws.as<T>();   // U + jl*te
        const dim3 grid16(unsigned(ceil_div(nb, 16)), unsigned(ns));
        if (prow == kp) {
            top_copy_kernel<T><<<dim3(64, unsigned(ns)), 256, 0, s>>>(pg, jl0, old, 0, 0);
            SB_TRY(launch_status());
        }
        perm_gather_kernel<T><<<grid16, 256, 0, s>>>(pg, jl0, mine);
        SB_TRY(launch_status());
        if (p > 1) {
            NCCL_TRY(ncclGroupStart());
            NCCL_TRY(ncclAllGather(mine, all, size_t(ns) * te, nccl_t, comm, s));
            NCCL_TRY(ncclBroadcast(old, old, size_t(ns) * te, nccl_t, kp, comm, s));
            NCCL_TRY(ncclGroupEnd());
        }
        perm_select_top_kernel<T><<<grid16, 256, 0, s>>>(pg, jl0, ns, all, U);
        SB_TRY(launch_status());
        perm_scatter_lower_kernel<T><<<grid16, 256, 0, s>>>(pg, jl0, old);
        SB_TRY(launch_status());
        // U(k, j) = L_kk^-1 * top block for block columns right of the panel (contiguous tail of the range)
        int jr = jl0;
        while (jr < jl1 && (int64_t(pcol) + int64_t(jr) * q <= k || int64_t(pcol) + int64_t(jr) * q == skip1)) ++jr;
        if (jr < jl1) {
            const T* Lkk = pws_tile(k, k);
            T* const* uptr = lookahead ? reinterpret_cast<T* const*>(dplan + ula_off)
                                       : reinterpret_cast<T* const*>(dplan + uptr_off) + jr;
            const int64_t jlast = int64_t(pcol) + int64_t(jl1 - 1) * q;
            const int full = (A.tile_nb(jlast) == nb) ? jl1 - jr : jl1 - 1 - jr;
            if (full > 0)
                SB_TRY(trsm_colmajor<T>(true, true, 'N', true, ntop, int(nb), T(1), Lkk, ld, uptr, 0, ld, full, W, s));
            if (full < jl1 - jr)
                SB_TRY(trsm_colmajor<T>(true, true, 'N', true, ntop, int(A.tile_nb(jlast)), T(1), Lkk, ld,
                                       uptr + full, 0, ld, 1, W, s));
        }
        if (prow == kp) {
            // the owner row stores the new top block (U right of the panel, permuted L left of it)
            top_copy_kernel<T><<<dim3(64, unsigned(ns)), 256, 0, s>>>(pg, jl0, U + int64_t(jl0) * te, 0, 1);
            SB_TRY(launch_status());
        }
        return SB200_OK;
    };

    auto body = [&]() -> int {
        if (ps.tnt_ranks > 0) SB_TRY(tnt.init(mt, nb, A.m, int(sizeof(T)), ps.tnt_ranks, P));
        CUDA_TRY(cudaMemcpyAsync(dplan, hp.data(), hp.size() * sizeof(void*), cudaMemcpyHostToDevice, P));
        CUDA_TRY(cudaMemsetAsync(infob.p, 0, sizeof(int), P));
        CUDA_TRY(cudaStreamSynchronize(P));
        htm.mark("scratch_streams_upload");
        CUDA_TRY(cudaEventRecord(t0, P));
        for (int64_t k = 0; k < kt; ++k) {
            const int kw = int(A.tile_nb(k));
            const int m_p = int(A.m - k * nb);
            const int ntop = std::min(m_p, kw);
            const int kp = int(k % p), kq = int(k % q);
            const int root = g.rank_of(k, k);
            int64_t* pt = dpiv_tile + k * nb;
            int64_t* po = dpiv_off + k * nb;
            int* perm = permb.as<int>() + (k & 1) * 3 * nb;

            // ---- panel: gather on the root, factor, broadcast tiles + pivots + row map
            if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(P, T_done(k - 2), 0));       // pws[k & 1] is free again
            SB_TRY(ptime(P));
            ph.begin("gather", P);
            if (pcol == kq && p > 1) {
                NCCL_TRY(ncclGroupStart());
                if (g.rank == root) {
                    for (int r = 0; r < p; ++r)
                        if (r != kp && count_rows(r, k) > 0)
                            NCCL_TRY(ncclRecv(pws_tile(first_row(r, k), k), size_t(count_rows(r, k) * te), nccl_t, r, g.col_comm, P));
                }
                else if (count_rows(prow, k) > 0)
                    NCCL_TRY(ncclSend(A.tile_as<T>(first_row(prow, k), k), size_t(count_rows(prow, k) * te), nccl_t, kp, g.col_comm, P));
                NCCL_TRY(ncclGroupEnd());
            }
            ph.end(P);
            if (g.rank == root) {
                iota_kernel<<<unsigned(ceil_div(m_p, 256)), 256, 0, P>>>(rowmapb.as<int>(), m_p);
                SB_TRY(launch_status());
                T* const* stack_k = reinterpret_cast<T* const*>(dplan + steps[k].stack_off);
                if (ps.tnt_ranks > 0) {
                    // tournament over the process rows, entirely on this GPU: the panel is already gathered here
                    std::vector<T*> htiles;
                    for (int64_t i = k; i < mt; ++i) htiles.push_back(int(i % p) == prow ? A.tile_as<T>(i, k) : pws_tile(i, k));
                    SB_TRY(getrf_panel_tnt<T>(stack_k, htiles, k, int(nb), m_p, kw, pt, po, infob.as<int>(), int(k * nb), ps, tnt, P,
                                              rowmapb.as<int>(), &ph));
                }
                else
                SB_TRY(getrf_panel<T>(stack_k, A.tile_as<T>(k, k), int(mt - k), int(nb), m_p, kw, pt, po, infob.as<int>(),
                                     int(k * nb), ps, P, rowmapb.as<int>(), &ph));
                perm_pack_kernel<<<unsigned(ceil_div(ntop, 256)), 256, 0, P>>>(rowmapb.as<int>(), pt, po, int(nb), ntop, perm);
                SB_TRY(launch_status());
            }
            ph.begin("bcast", P);
            if (multi) {
                std::vector<BcastItem> items;
                for (int r = 0; r < p; ++r) {
                    if (count_rows(r, k) == 0) continue;
                    T* dstp = pws_tile(first_row(r, k), k);
                    const T* srcp = (g.rank == root && r == kp) ? A.tile_as<T>(first_row(r, k), k) : dstp;
                    items.push_back({srcp, dstp, size_t(count_rows(r, k) * te) * sizeof(T), root});
                }
                items.push_back({perm, perm, size_t(3 * ntop) * sizeof(int), root});
                items.push_back({pt, pt, size_t(ntop) * sizeof(int64_t), root});
                items.push_back({po, po, size_t(ntop) * sizeof(int64_t), root});
                SB_TRY(bcast_many(g, items, P));
                // the other owners of panel tiles take their factored tiles back from the workspace
                if (pcol == kq && g.rank != root && count_rows(prow, k) > 0)
                    CUDA_TRY(cudaMemcpyAsync(A.tile_as<T>(first_row(prow, k), k), pws_tile(first_row(prow, k), k),
                                             size_t(count_rows(prow, k) * te) * sizeof(T), cudaMemcpyDeviceToDevice, P));
            }
            else {
                // single rank: the workspace copy of the panel is what the GEMMs read
                CUDA_TRY(cudaMemcpyAsync(pws_tile(k, k), A.tile_as<T>(k, k), size_t((mt - k) * te) * sizeof(T),
                                         cudaMemcpyDeviceToDevice, P));
            }
            ph.end(P);
            if (use_tc05 && steps[k].a_cnt > 0) {
                const Step& sk = steps[k];
                const void* const* src = reinterpret_cast<const void* const*>(dplan + sk.a_src);
                void* const* dst = reinterpret_cast<void* const*>(dplan + sk.a_dst);
                SB_TRY(pack('A', src, dst, sk.a_full, int(nb), kw, P));
                SB_TRY(pack('A', src + sk.a_full, dst + sk.a_full, sk.a_cnt - sk.a_full, int(A.tile_mb(mt - 1)), kw, P));
            }
            SB_TRY(ptime(P));
            CUDA_TRY(cudaEventRecord(P_done(k), P));

            // ---- trailing stream: interchanges on every local column except k (and k+1: lookahead), U row, GEMM
            CUDA_TRY(cudaStreamWaitEvent(T_, P_done(k), 0));
            ph.begin("tr_permute_solve", T_);
            SB_TRY(permute_and_solve(k, 0, nt_loc, int(k), int(k + 1), false, T_, g.col_comm2, wt.as<T>()));
            ph.end(T_);
            if (use_tc05 && steps[k].b_cnt > 0) {
                const Step& sk = steps[k];
                const void* const* src = reinterpret_cast<const void* const*>(dplan + sk.b_src);
                void* const* dst = reinterpret_cast<void* const*>(dplan + sk.b_dst);
                SB_TRY(pack('B', src, dst, sk.b_full, int(nb), kw, T_));
                SB_TRY(pack('B', src + sk.b_full, dst + sk.b_full, sk.b_cnt - sk.b_full, int(A.tile_nb(nt - 1)), kw, T_));
            }
            if constexpr (std::is_same<T, double>::value) {
                if (use_bt && steps[k].b_cnt > 0) {
                    const Step& sk = steps[k];
                    const double* const* src = reinterpret_cast<const double* const*>(dplan + sk.b_src);
                    double* const* dst = reinterpret_cast<double* const*>(dplan + sk.b_dst);
                    if (sk.b_full > 0)
                        SB_TRY(sb200_transpose_batched_d(0, kw, nb, src, ld, dst, ld, sk.b_full, T_));
                    if (sk.b_cnt > sk.b_full)
                        SB_TRY(sb200_transpose_batched_d(0, kw, A.tile_nb(nt - 1), src + sk.b_full, ld, dst + sk.b_full, ld,
                                                         sk.b_cnt - sk.b_full, T_));
                }
            }
            if (! steps[k].tr.empty()) {
                cudaEvent_t a0, a1;
                CUDA_TRY(cudaEventCreate(&a0)); CUDA_TRY(cudaEventCreate(&a1));
                tev.push_back(a0); tev.push_back(a1);
                CUDA_TRY(cudaEventRecord(a0, T_));
                SB_TRY(run_batches(steps[k].tr, T_, use_bt));
                CUDA_TRY(cudaEventRecord(a1, T_));
                for (const auto& b : steps[k].tr) trail_flops += 2.0 * b.m * b.n * b.k * double(b.C.size());
                trail_launches += int64_t(steps[k].tr.size());
            }
            CUDA_TRY(cudaEventRecord(T_done(k), T_));

            // ---- lookahead: column k+1 brought up to date on the panel stream
            if (k + 1 < nt && int((k + 1) % q) == pcol) {
                if (k >= 1) CUDA_TRY(cudaStreamWaitEvent(P, T_done(k - 1), 0));
                const int jl = int((k + 1 - pcol) / q);
                ph.begin("la_permute_solve", P);
                SB_TRY(permute_and_solve(k, jl, jl + 1, -1, -1, true, P, g.col_comm, reinterpret_cast<T*>(ps.W)));
                if constexpr (is_float) {
                    if (use_tc05 && ! steps[k].la.empty()) {
                        Tc05PackParams qp{};
                        qp.X0 = ula.as<float>(); qp.P0 = pkBla.p;
                        qp.rows = int(A.tile_nb(k + 1)); qp.k = kw; qp.rs = ld; qp.ks = 1; qp.ru = TC_BN; qp.batch = 1;
                        SB_TRY(launch_tc05_pack(qp, P));
                    }
                }
                ph.end(P);
                ph.begin("la_gemm", P);
                SB_TRY(run_batches(steps[k].la, P));
                ph.end(P);
            }
        }
        htm.mark("enqueue");
        CUDA_TRY(cudaStreamWaitEvent(P, T_done(kt - 1), 0));
        CUDA_TRY(cudaEventRecord(t1, P));
        CUDA_TRY(cudaStreamSynchronize(P));
        CUDA_TRY(cudaStreamSynchronize(T_));
        htm.mark("sync");
        return SB200_OK;
    };
    int status = body();
    ph.report("getrf_dist", g.rank);
    if (status == SB200_OK) {
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        A.last_ms = ms;
        double tms = 0;
        for (size_t i = 0; i + 1 < tev.size(); i += 2) { float x = 0; if (cudaEventElapsedTime(&x, tev[i], tev[i + 1]) == cudaSuccess) tms += x; }
        A.last_trail_ms = tms; A.last_trail_flops = trail_flops; A.last_trail_launches = trail_launches;
        double pms = 0;
        for (size_t i = 0; i + 1 < pev.size(); i += 2) { float x = 0; if (cudaEventElapsedTime(&x, pev[i], pev[i + 1]) == cudaSuccess) pms += x; }
        A.last_panel_ms = pms;
        int hinfo = 0;
        cudaMemcpy(&hinfo, infob.p, sizeof(int), cudaMemcpyDeviceToHost);
        int64_t info = hinfo;
        if (multi) {
            // info is produced on the panel roots only: first failing column over all ranks
            // (reference: internal_reduce_info.cc:23-38, MPI_MIN over non-zero values)
            int64_t v = info ? info : INT64_MAX;
            int64_t* dv = reinterpret_cast<int64_t*>(rowmapb.p);
            cudaMemcpy(dv, &v, sizeof(v), cudaMemcpyHostToDevice);
            if (ncclAllReduce(dv, dv, 1, ncclInt64, ncclMin, g.world, P) != ncclSuccess) status = SB200_ENCCL;
            cudaStreamSynchronize(P);
            cudaMemcpy(&v, dv, sizeof(v), cudaMemcpyDeviceToHost);
            info = (v == INT64_MAX) ? 0 : v;
        }
        if (info_out) *info_out = info;
        if (pivots_out) {
            std::vector<int64_t> ht(size_t(kt * nb)), ho(size_t(kt * nb));
            cudaMemcpy(ht.data(), dpiv_tile, ht.size() * sizeof(int64_t), cudaMemcpyDeviceToHost);
            cudaMemcpy(ho.data(), dpiv_off, ho.size() * sizeof(int64_t), cudaMemcpyDeviceToHost);
            int64_t o = 0;
            for (int64_t k = 0; k < kt; ++k) {
                const int64_t dl = std::min(A.m - k * nb, A.tile_nb(k));
                for (int64_t j = 0; j < dl && o < mn; ++j, ++o) {
                    pivots_out[2 * o] = ht[size_t(k * nb + j)];
                    pivots_out[2 * o + 1] = ho[size_t(k * nb + j)];
                }
            }
        }
    }
    for (auto e : ev) cudaEventDestroy(e);
    for (auto e : tev) cudaEventDestroy(e);
    for (auto e : pev) cudaEventDestroy(e);
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    if (P) cudaStreamDestroy(P);
    if (T_) cudaStreamDestroy(T_);
    return status;
}

int getrf_driver_dist(Matrix& A, int64_t* pivots_out, int64_t* info_out)
{
    return getrf_driver_dist_t<double>(A, pivots_out, info_out, false);
}

int getrf_driver_dist_s(Matrix& A, int64_t* pivots_out, int64_t* info_out, bool use_tc05)
{
    return getrf_driver_dist_t<float>(A, pivots_out, info_out, use_tc05);
}

} // namespace sb200
