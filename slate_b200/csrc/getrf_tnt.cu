// getrf_tnt.cu -- panel of the LU with tournament pivoting (CALU), SURVEY section 8(f) item 2.
//
// Reference: src/getrf_tntpiv.cc:22-395 (driver), src/internal/internal_getrf_tntpiv.cc:357-640 (panel),
// :42-120 (permutation_to_sequential_pivot), src/internal/Tile_getrf_tntpiv.hh:69-300 (the LU of every tree node:
// the pivot rule of tile::getrf).
//
// What the reference does per panel: every MPI rank that owns tiles of the block column factors a COPY of its own
// rows with partial pivoting and keeps the ORIGINAL rows that ended in its first nb positions; a binary tree over the
// ranks (ordered by their first tile) stacks two candidate tiles, factors a copy, keeps the originals of the nb
// winners; the last LU gives the factored diagonal tile, the winners give the row interchanges; the rows below the
// diagonal tile are then solved against U_kk (trsm Right / Upper / NonUnit) instead of being eliminated in the panel.
//
// B200-first restatement: the p x q driver already gathers the panel on the owner of A(k, k) (getrf_dist.cu), so the
// whole tournament runs on ONE GPU with no message at all: a "rank" is the subset of panel tiles of one process row,
// i.e. a pointer sub-array of the workspace copy, and every tree node is one call of the partial-pivoting GPU panel
// (getrf.cu / getrf_base_v3.cu) whose row map says which original rows won.  Candidate rows are gathered from the
// ORIGINAL panel by index, so no candidate tile is ever copied between nodes.  Nothing goes through the host: the
// id lists, the winners -> sequential-interchange conversion and the final row map stay on the device.
//
// Shapes: as in the reference, every diagonal tile has to be square (its TriangularMatrix view throws otherwise,
// include/slate/TriangularMatrix.hh:459); the entry points return SB200_ENOTSUP for the others.
#include "runtime_internal.hh"
#include "getrf_internal.hh"
#include <algorithm>
#include <vector>

namespace sb200 {

namespace {

// dst tile t <- src tile t (whole nb x nb pool tiles)
template <typename T>
__global__ void __launch_bounds__(256)
tnt_copy_tiles_kernel(T* const* __restrict__ src, T* __restrict__ dst, int64_t te)
{
    const T* s = src[blockIdx.y];
    T* d = dst + int64_t(blockIdx.y) * te;
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < te; e += int64_t(gridDim.x) * blockDim.x) d[e] = s[e];
}

__global__ void tnt_iota_kernel(int* v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = i;
}

// stage 0: position j of a rank's sub-stack holds sub-stack row rowmap[j]; sub-stack tile ts is panel tile t0 + ts * stride
__global__ void tnt_ids_kernel(const int* __restrict__ rowmap, int cnt, int t0, int stride, int nb, int* __restrict__ ids)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cnt) return;
    const int r = rowmap[j];
    ids[j] = (t0 + (r / nb) * stride) * nb + (r % nb);
}

// dst(j, c) = original panel row ids[j], column c   (dst: tile with ld = nb; j < cnt, c < kw)
template <typename T>
__global__ void __launch_bounds__(256)
tnt_gather_rows_kernel(T* const* __restrict__ stack, int nb, int kw, const int* __restrict__ ids, int cnt, T* __restrict__ dst)
{
    const int c0 = blockIdx.x * 8, c1 = min(c0 + 8, kw);
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        const int r = ids[j];
        const T* a = stack[r / nb] + (r % nb);
        for (int c = c0; c < c1; ++c) dst[int64_t(c) * nb + j] = a[int64_t(c) * nb];
    }
}

// tree node: position j of the stacked pair holds stacked row rowmap[j]; stacked rows [0, na) are ids_a, the rest ids_b.
// out may alias ids_a: one CTA, every read before the first write.
__global__ void tnt_merge_ids_kernel(const int* __restrict__ rowmap, const int* ids_a, int na, const int* __restrict__ ids_b,
                                     int* out, int cnt)
{
    extern __shared__ int s_ids[];
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        const int r = rowmap[j];
        s_ids[j] = r < na ? ids_a[r] : ids_b[r - na];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) out[j] = s_ids[j];
}

// winners[j] = original panel row that has to sit at position j  ->  sequential interchanges j <-> piv[j] and the final
// arrangement row_at (row_at[x] = original row at position x).  One CTA; the chain of diag_len steps is serial.
__global__ void __launch_bounds__(256)
tnt_sequential_kernel(const int* __restrict__ winners, int diag_len, int m_p, int nb, int* row_at, int* pos_of,
                      int64_t* __restrict__ piv_tile, int64_t* __restrict__ piv_off)
{
    for (int i = threadIdx.x; i < m_p; i += blockDim.x) { row_at[i] = i; pos_of[i] = i; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    for (int j = 0; j < diag_len; ++j) {
        const int w = winners[j];
        const int x = pos_of[w];                 // where the winner sits now (>= j: positions < j hold earlier winners)
        piv_tile[j] = x / nb;
        piv_off[j] = x % nb;
        const int rj = row_at[j];
        row_at[j] = w; row_at[x] = rj;
        pos_of[w] = j; pos_of[rj] = x;
    }
}

} // namespace

TntScratch::~TntScratch() { if (raw) ws_cache_put(raw); }

// mt tile rows of nb x nb elements of `esize` bytes, at most m rows in a panel, `ranks` participants per panel
int TntScratch::init(int64_t mt, int64_t nb, int64_t m, int esize, int ranks_, cudaStream_t s)
{
    ranks = std::max(ranks_, 1);
    te = nb * nb;
    per_class = int((mt + ranks - 1) / ranks) + 1;
    auto up = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t wc_b = up(size_t(mt) * te * esize), tmp_b = up(size_t(2) * te * esize);
    const size_t ids_b = up(size_t(ranks) * nb * sizeof(int)), rm_b = up(size_t(std::max<int64_t>(m, 2 * nb)) * sizeof(int));
    const size_t piv_b = up(size_t(2) * nb * sizeof(int64_t));
    const size_t ptr_b = up((size_t(ranks) * per_class + 2) * sizeof(void*));
    const size_t bytes = wc_b + tmp_b + ids_b + 3 * rm_b + piv_b + ptr_b + 256;
    raw = ws_cache_get(bytes);
    if (! raw) return SB200_ENOMEM;
    char* p = static_cast<char*>(raw);
    wcopy = p; p += wc_b;
    tmp = p; p += tmp_b;
    ids = reinterpret_cast<int*>(p); p += ids_b;
    rm_sub = reinterpret_cast<int*>(p); p += rm_b;
    row_at = reinterpret_cast<int*>(p); p += rm_b;
    pos_of = reinterpret_cast<int*>(p); p += rm_b;
    spiv = reinterpret_cast<int64_t*>(p); p += piv_b;
    ptrs = reinterpret_cast<void**>(p); p += ptr_b;
    dummy_info = reinterpret_cast<int*>(p);
    // pointer sub-arrays of the workspace copy: class c lists the tiles c, c + ranks, c + 2 ranks, ...; then the pair
    std::vector<void*> h(size_t(ranks) * per_class + 2, nullptr);
    for (int c = 0; c < ranks; ++c)
        for (int i = 0; c + int64_t(i) * ranks < mt; ++i)
            h[size_t(c) * per_class + i] = static_cast<char*>(wcopy) + (c + int64_t(i) * ranks) * te * esize;
    h[size_t(ranks) * per_class] = tmp;
    h[size_t(ranks) * per_class + 1] = static_cast<char*>(tmp) + te * esize;
    CUDA_TRY(cudaMemcpyAsync(ptrs, h.data(), h.size() * sizeof(void*), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(dummy_info, 0, sizeof(int), s));
    CUDA_TRY(cudaStreamSynchronize(s));          // `h` is pageable and dies here
    return SB200_OK;
}

// Tournament panel.  `stack` (device) / `htiles` (the same pointers on the host): the ntile tiles of block column k
// from the diagonal tile down, m_p rows, kw columns, kw x kw diagonal tile.  On return the panel holds the factored
// diagonal tile and, below it, the interchanged original rows times U_kk^-1; piv_tile / piv_off the sequential
// interchanges; rowmap (optional, m_p ints) the final arrangement (rowmap[x] = original panel row now at position x).
template <typename T>
int getrf_panel_tnt(T* const* stack, const std::vector<T*>& htiles, int64_t k, int nb, int m_p, int kw,
                    int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                    PanelScratch& ps, TntScratch& ts, cudaStream_t s, int* rowmap, PhaseTimer* ph)
{
    const int ntile = int(htiles.size());
    const int ranks = ts.ranks;
    const int64_t te = ts.te;
    if (ntile < 1 || m_p < kw || (ntile > 1 && kw != nb)) return SB200_ENOTSUP;      // square diagonal tiles only
    const int diag_len = kw;
    PhaseTimer off_timer;
    off_timer.on = false;
    PhaseTimer& pt = ph ? *ph : off_timer;
    T* wc = static_cast<T*>(ts.wcopy);
    T* tmp0 = static_cast<T*>(ts.tmp);
    int64_t* spt = ts.spiv;
    int64_t* spo = ts.spiv + nb;
    auto rows_of_tile = [&](int t) { return std::min(nb, m_p - t * nb); };

    // participants in the order of their first tile (rank_rows, internal_getrf_tntpiv.cc:456-470)
    struct Part { int t0, ntile, rows, first_mb; };
    std::vector<Part> parts;
    for (int t = 0; t < ntile && int(parts.size()) < ranks; ++t) {
        // tile t belongs to process row (k + t) % ranks; its first tile in the panel is t < ranks
        Part q{t, 0, 0, rows_of_tile(t)};
        for (int u = t; u < ntile; u += ranks) { ++q.ntile; q.rows += rows_of_tile(u); }
        parts.push_back(q);
    }
    (void) k;          // the owner of tile t is (k + t) % ranks, but only the classes t mod ranks matter here
    const int nranks = int(parts.size());

    pt.begin("tnt_copy", s);
    tnt_copy_tiles_kernel<T><<<dim3(unsigned(std::min<int64_t>(ceil_div(te, 256 * 4), 64)), unsigned(ntile)), 256, 0, s>>>(stack, wc, te);
    SB_TRY(launch_status());
    pt.end(s);

    const T* top = nullptr;                       // factored diagonal tile
    if (nranks == 1) {
        // one participant: its local LU is the factorisation of the panel (internal_getrf_tntpiv.cc:602-619)
        if (rowmap) {
            tnt_iota_kernel<<<unsigned(ceil_div(m_p, 256)), 256, 0, s>>>(rowmap, m_p);
            SB_TRY(launch_status());
        }
        T* const* sub = reinterpret_cast<T* const*>(ts.ptrs);            // class 0 (ntile == 1, or ranks == 1: every tile)
        SB_TRY(getrf_panel<T>(sub, wc, ntile, nb, m_p, kw, piv_tile, piv_off, dinfo, info_base, ps, s, rowmap, ph));
        top = wc;
    }
    else {
        // ---- stage 0: every participant factors a copy of its own rows; ids[x] = original rows of its first tile
        for (int x = 0; x < nranks; ++x) {
            const Part& q = parts[size_t(x)];
            tnt_iota_kernel<<<unsigned(ceil_div(q.rows, 256)), 256, 0, s>>>(ts.rm_sub, q.rows);
            SB_TRY(launch_status());
            T* const* sub = reinterpret_cast<T* const*>(ts.ptrs) + size_t(q.t0) * ts.per_class;
            SB_TRY(getrf_panel<T>(sub, wc + int64_t(q.t0) * te, q.ntile, nb, q.rows, kw, spt, spo,
                                  x == 0 ? dinfo : ts.dummy_info, info_base, ps, s, ts.rm_sub, ph));
            tnt_ids_kernel<<<unsigned(ceil_div(q.first_mb, 256)), 256, 0, s>>>(ts.rm_sub, q.first_mb, q.t0, ranks, nb,
                                                                                ts.ids + int64_t(x) * nb);
            SB_TRY(launch_status());
        }
        // ---- tree: index x (x % 2 step == 0) against x + step (internal_getrf_tntpiv.cc:531-598)
        T* const* pair = reinterpret_cast<T* const*>(ts.ptrs) + size_t(ranks) * ts.per_class;
        T* tmp1 = tmp0 + te;
        for (int step = 1; step < nranks; step *= 2)
            for (int x = 0; x + step < nranks; x += 2 * step) {
                const int na = parts[size_t(x)].first_mb, nbm = parts[size_t(x + step)].first_mb;       // na == nb
                int* ia = ts.ids + int64_t(x) * nb;
                int* ib = ts.ids + int64_t(x + step) * nb;
                const unsigned gc = unsigned(ceil_div(kw, 8));
                tnt_gather_rows_kernel<T><<<gc, 256, 0, s>>>(stack, nb, kw, ia, na, tmp0);
                SB_TRY(launch_status());
                tnt_gather_rows_kernel<T><<<gc, 256, 0, s>>>(stack, nb, kw, ib, nbm, tmp1);
                SB_TRY(launch_status());
                tnt_iota_kernel<<<unsigned(ceil_div(na + nbm, 256)), 256, 0, s>>>(ts.rm_sub, na + nbm);
                SB_TRY(launch_status());
                SB_TRY(getrf_panel<T>(pair, tmp0, 2, nb, na + nbm, kw, spt, spo, x == 0 ? dinfo : ts.dummy_info, info_base,
                                      ps, s, ts.rm_sub, ph));
                tnt_merge_ids_kernel<<<1, 256, size_t(na) * sizeof(int), s>>>(ts.rm_sub, ia, na, ib, ia, na);
                SB_TRY(launch_status());
            }
        // the last node (index 0, largest step) left the factored diagonal tile in tmp0 and the winners in ids[0]
        int* rm = rowmap ? rowmap : ts.row_at;
        tnt_sequential_kernel<<<1, 256, 0, s>>>(ts.ids, diag_len, m_p, nb, rm, ts.pos_of, piv_tile, piv_off);
        SB_TRY(launch_status());
        top = tmp0;
    }

    // ---- the panel itself: interchanges on the original rows, factored diagonal tile, rows below times U_kk^-1
    pt.begin("tnt_apply", s);
    SB_TRY(launch_laswp<T>(stack, 0, nb, nb, nb, 1, piv_tile, piv_off, 0, diag_len, 1, 0, kw, s));
    CUDA_TRY(cudaMemcpyAsync(htiles[0], top, size_t(te) * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (ntile > 1) {
        const int last_rows = rows_of_tile(ntile - 1);
        const int full = last_rows == nb ? ntile - 1 : ntile - 2;
        if (full > 0)
            SB_TRY(trsm_colmajor<T>(false, false, 'N', false, nb, kw, from_real<T>(1), htiles[0], nb, stack + 1, 0, nb, full,
                                    reinterpret_cast<T*>(ps.W), s));
        if (full < ntile - 1)
            SB_TRY(trsm_colmajor<T>(false, false, 'N', false, last_rows, kw, from_real<T>(1), htiles[0], nb, stack + (ntile - 1), 0, nb, 1,
                                    reinterpret_cast<T*>(ps.W), s));
    }
    pt.end(s);
    return SB200_OK;
}

template int getrf_panel_tnt<double>(double* const*, const std::vector<double*>&, int64_t, int, int, int, int64_t*, int64_t*, int*, int,
                                     PanelScratch&, TntScratch&, cudaStream_t, int*, PhaseTimer*);
template int getrf_panel_tnt<float>(float* const*, const std::vector<float*>&, int64_t, int, int, int, int64_t*, int64_t*, int*, int,
                                    PanelScratch&, TntScratch&, cudaStream_t, int*, PhaseTimer*);
template int getrf_panel_tnt<cuFloatComplex>(cuFloatComplex* const*, const std::vector<cuFloatComplex*>&, int64_t, int, int, int, int64_t*,
                                             int64_t*, int*, int, PanelScratch&, TntScratch&, cudaStream_t, int*, PhaseTimer*);
template int getrf_panel_tnt<cuDoubleComplex>(cuDoubleComplex* const*, const std::vector<cuDoubleComplex*>&, int64_t, int, int, int, int64_t*,
                                              int64_t*, int*, int, PanelScratch&, TntScratch&, cudaStream_t, int*, PhaseTimer*);

// every diagonal tile square (the reference's TriangularMatrix view of A(k, k) requires it)
bool tnt_shape_supported(const Matrix& A)
{
    const int64_t kt = std::min(A.mt, A.nt);
    for (int64_t k = 0; k < kt; ++k)
        if (A.tile_mb(k) != A.tile_nb(k)) return false;
    return true;
}

} // namespace sb200
