// solve_dist.cu -- the solve path on a p x q grid: potrs / hemm / inf-norm / posv_mixed with the right-hand
// sides REPLICATED on every rank.
//
// Reference: work::trsm (src/work/work_trsm.cc:60-387) moves tiles of the triangular factor to the ranks that
// own the right-hand-side tiles (tileBcast / listBcast per step); potrs = two such sweeps (src/potrs.cc:54-77),
// hemm for the residual (src/hemmC.cc), posv_mixed on top (src/posv_mixed.cc:111-297).
//
// B200-first: with nrhs ~ 10 the right-hand sides are tiny (n x nrhs = 5 MB at n = 65536) while the factor is
// n^2 / 2 elements spread over the GPUs, so the FACTOR never moves.  Every rank keeps a replicated copy of the
// block vector X (mt blocks of nb x nrhs, ld = nb).  A sweep step i is left-looking:
//     partial_r = sum over the LOCAL tiles of block row i (NoTrans) / block column i (ConjTrans) of op(T) X_k
//     (one batched skinny GEMM into per-tile partials + a fixed-order reduction: deterministic),
//     all-reduce of the nb x nrhs partial inside the process row / column that owns those tiles,
//     the owner of T(i,i) solves its block (trsm_small) and broadcasts the nb x nrhs result to everyone.
// Per step: 2 small NCCL collectives; every factor tile is read exactly once per sweep, where it lives.
// The residual R = B - A X works the same way (local products into a replicated accumulator, one all-reduce).
//
// STATUS: validated in round 2 on 1x2, 2x1 and 2x4 grids (scratch/mgpu_check.py: potrs, posv_mixed, gesv_mixed against
// the oracle; profiles/r02m2_mgpu_check_2gpu.log, r02g8_mgpu_check_2x4.log) and through the one-rank hook
// SB200_DIST_SOLVE=2 (tests/test_zzz_gpu_dist_solve.py); dposv_mixed / dgesv_mixed at n = 65536 on 8 GPUs:
// profiles/r02g8_bench_{posv,gesv}_mixed_8gpu.json.  SB200_DIST_SOLVE=0 restores SB200_ENOTSUP on grids.
#include "runtime_internal.hh"
#include "getrf_internal.hh"
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>

namespace sb200 {

namespace {

template <typename T> struct NcclType;
template <> struct NcclType<float>  { static constexpr ncclDataType_t value = ncclFloat; };
template <> struct NcclType<double> { static constexpr ncclDataType_t value = ncclDouble; };

// replicated block vector: mt blocks of nb x nrhs (ld = nb), block i at base + i * nb * nrhs
template <typename T>
struct RepVec {
    T* base = nullptr;
    int64_t m = 0, nb = 0, mt = 0;
    int nrhs = 0;
    T* blk(int64_t i) const { return base + i * nb * nrhs; }
    int64_t rows(int64_t i) const { return i == mt - 1 ? m - i * nb : nb; }
    size_t elems() const { return size_t(mt) * nb * nrhs; }
    int alloc(int64_t m_, int64_t nb_, int nrhs_)
    {
        m = m_; nb = nb_; nrhs = nrhs_; mt = ceil_div(m, nb);
        CUDA_TRY(cudaMalloc(&base, std::max<size_t>(elems(), 1) * sizeof(T)));
        return SB200_OK;
    }
    ~RepVec() { if (base) cudaFree(base); }
};

template <typename S, typename D>
__global__ void __launch_bounds__(256) rv_convert_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t count)
{
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < count; e += int64_t(gridDim.x) * blockDim.x)
        dst[e] = D(src[e]);
}

// y = a + s * b  (element-wise)
template <typename T>
__global__ void __launch_bounds__(256) rv_axpby_kernel(const T* a, T s, const T* b, T* y, int64_t count)      // y may alias a or b
{
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < count; e += int64_t(gridDim.x) * blockDim.x)
        y[e] = a[e] + s * b[e];
}

// out = sum_{t < cnt} P[t] (each nb x nrhs block, fixed order), out may be then subtracted from x by the caller
template <typename T>
__global__ void __launch_bounds__(256) rv_reduce_partials_kernel(const T* __restrict__ P, int cnt, int64_t blk_elems, T* __restrict__ out)
{
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < blk_elems; e += int64_t(gridDim.x) * blockDim.x) {
        T s = T(0);
        for (int t = 0; t < cnt; ++t) s += P[int64_t(t) * blk_elems + e];
        out[e] = s;
    }
}

// out[c] = max_r |X(r, c)| over the replicated block vector (rows beyond m in the last block are skipped)
template <typename T>
__global__ void __launch_bounds__(256) rv_colmax_kernel(const T* __restrict__ X, double* __restrict__ out, int64_t m, int nb, int nrhs)
{
    __shared__ double red[256];
    const int c = blockIdx.x;
    double best = 0.0;
    for (int64_t r = threadIdx.x; r < m; r += blockDim.x) {
        const double v = fabs(double(X[(r / nb) * int64_t(nb) * nrhs + (r % nb) + int64_t(c) * nb]));
        best = (v > best || v != v) ? v : best;
    }
    red[threadIdx.x] = best;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (int(threadIdx.x) < o) {
            const double a = red[threadIdx.x + o], b = red[threadIdx.x];
            red[threadIdx.x] = (a > b || a != a) ? a : b;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = red[0];
}

// full Hermitian copy of local diagonal tiles (see he_fill_kernel in solve.cu)
template <typename T>
__global__ void __launch_bounds__(256) rv_he_fill_kernel(const T* const* __restrict__ diag, const int* __restrict__ dims,
                                                         T* __restrict__ out, int ld, int64_t te)
{
    const int k = blockIdx.y, n = dims[k];
    const T* __restrict__ a = diag[k];
    T* __restrict__ o = out + int64_t(k) * te;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
        const int r = e % n, c = e / n;
        o[r + int64_t(c) * ld] = (r >= c) ? a[r + int64_t(c) * ld] : a[c + int64_t(r) * ld];
    }
}

inline unsigned rv_grid(int64_t count) { return unsigned(std::min<int64_t>(ceil_div(std::max<int64_t>(count, 1), 256), 148 * 8)); }

struct WallClock {
    std::chrono::steady_clock::time_point t;
    void start() { cudaDeviceSynchronize(); t = std::chrono::steady_clock::now(); }
    double stop() { cudaDeviceSynchronize(); return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); }
};

bool dist_solve_enabled()
{
    const char* e = getenv("SB200_DIST_SOLVE");
    return ! (e && atoi(e) == 0);
}

// distributed tiles of B (n x nrhs, one tile column) -> replicated block vector (sum of the owners' copies)
template <typename T>
int gather_rep(Matrix& B, RepVec<T>& X, cudaStream_t s)
{
    Grid& g = *B.g;
    CUDA_TRY(cudaMemsetAsync(X.base, 0, X.elems() * sizeof(T), s));
    for (int64_t i = g.prow; i < B.mt; i += g.p)
        if (B.is_local(i, 0))
            CUDA_TRY(cudaMemcpyAsync(X.blk(i), B.tile_as<T>(i, 0), size_t(X.nb) * X.nrhs * sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (g.size() > 1) NCCL_TRY(ncclAllReduce(X.base, X.base, X.elems(), NcclType<T>::value, ncclSum, g.world, s));
    return SB200_OK;
}

template <typename T>
int scatter_rep(const RepVec<T>& X, Matrix& B, cudaStream_t s)
{
    Grid& g = *B.g;
    for (int64_t i = g.prow; i < B.mt; i += g.p)
        if (B.is_local(i, 0))
            CUDA_TRY(cudaMemcpyAsync(B.tile_as<T>(i, 0), X.blk(i), size_t(X.nb) * X.nrhs * sizeof(T), cudaMemcpyDeviceToDevice, s));
    return SB200_OK;
}

// ------------------------------------------------------------------------------------------
// X <- op(T)^{-1} X on the replicated block vector; T = lower (or upper) triangle of the tile matrix A.
// ------------------------------------------------------------------------------------------
template <typename T>
int sweep_dist(Matrix& A, bool lower, int op, bool unit, RepVec<T>& X, cudaStream_t s)
{
    Grid& g = *A.g;
    const int64_t kt = A.nt, nb = A.nb;
    const int ld = int(nb), nrhs = X.nrhs;
    const bool trans = (op != 'N');
    const bool forward = (lower != trans);
    if (! lower && trans) return SB200_ENOTSUP;
    if (A.kind == 'H' && ! lower) return SB200_EINVAL;
    if (kt == 0 || nrhs == 0) return SB200_OK;
    const int64_t blk_elems = nb * nrhs;
    const int nblk = int(ceil_div(nb, FACTOR_IB));

    // ---- plan: per step the local contributing tiles (A operand), the X blocks they multiply and partial slots
    struct Step { std::vector<Batch> prod; int cnt = 0; };
    std::vector<Step> steps(static_cast<size_t>(kt));
    int64_t max_cnt = 1;
    std::vector<const T*> my_diag;            // diagonal tiles this rank owns, ascending
    std::vector<int64_t> my_diag_idx(size_t(kt), -1);
    for (int64_t i = 0; i < kt; ++i)
        if (A.is_local(i, i)) { my_diag_idx[size_t(i)] = int64_t(my_diag.size()); my_diag.push_back(A.tile_as<T>(i, i)); }
    DevBuf partials;                          // [max_cnt] blocks of nb x nrhs
    // two passes: count first (the partial slots live in one buffer whose size must be known)
    for (int64_t i = 0; i < kt; ++i) {
        int64_t cnt = 0;
        const int64_t k0 = forward ? 0 : i + 1, k1 = forward ? i : kt;
        for (int64_t k = k0; k < k1; ++k)
            if (trans ? A.is_local(k, i) : A.is_local(i, k)) ++cnt;
        max_cnt = std::max(max_cnt, cnt);
    }
    SB_TRY(partials.alloc(size_t(max_cnt) * blk_elems * sizeof(T)));
    PlanBuffer pb;
    for (int64_t i = 0; i < kt; ++i) {
        Step& st = steps[size_t(i)];
        const int64_t k0 = forward ? 0 : i + 1, k1 = forward ? i : kt;
        for (int64_t k = k0; k < k1; ++k) {
            const bool mine = trans ? A.is_local(k, i) : A.is_local(i, k);
            if (! mine) continue;
            const T* tile = trans ? A.tile_as<T>(k, i) : A.tile_as<T>(i, k);
            // NoTrans: (rows(i) x rows(k)) * X_k ; Trans: op(tile (rows(k) x rows(i))) * X_k
            batch_add(st.prod, int(X.rows(i)), nrhs, int(X.rows(k)), 0, tile, X.blk(k),
                      partials.as<T>() + int64_t(st.cnt) * blk_elems);
            ++st.cnt;
        }
        pb.reserve(st.prod);
    }
    std::vector<T*> xblk;
    for (int64_t i = 0; i < kt; ++i) xblk.push_back(X.blk(i));
    const size_t xblk_off = pb.push(xblk);
    const size_t diag_off = pb.push(my_diag);
    DevBuf W, part;
    SB_TRY(W.alloc(std::max<size_t>(my_diag.size(), 1) * nblk * FACTOR_IB * FACTOR_IB * sizeof(T)));
    SB_TRY(part.alloc(size_t(blk_elems) * sizeof(T)));
    SB_TRY(pb.upload(s));
    if (! my_diag.empty()) {
        const bool own_last = A.is_local(kt - 1, kt - 1);
        SB_TRY(trtri_diag_all<T>(int(my_diag.size()), pb.at<const T>(diag_off), ld, int(nb),
                                 own_last ? int(A.tile_mb(kt - 1)) : int(nb), lower, unit, W.as<T>(), s));
    }
    for (int64_t sidx = 0; sidx < kt; ++sidx) {
        const int64_t i = forward ? sidx : kt - 1 - sidx;
        const Step& st = steps[size_t(i)];
        const bool in_set = trans ? (g.pcol == int(i % g.q)) : (g.prow == int(i % g.p));
        const int owner = g.rank_of(i, i);
        const bool have_terms = forward ? (i > 0) : (i < kt - 1);
        if (in_set && have_terms) {
            if (st.cnt > 0)
                SB_TRY(launch_batches<T>(st.prod, pb, trans ? op : 'N', 'N', T(1), T(0), ld, 0, s));
            rv_reduce_partials_kernel<T><<<rv_grid(blk_elems), 256, 0, s>>>(partials.as<T>(), st.cnt, blk_elems, part.as<T>());
            SB_TRY(launch_status());
            ncclComm_t comm = trans ? g.col_comm : g.row_comm;
            const int csize = trans ? g.p : g.q;
            if (csize > 1)
                NCCL_TRY(ncclAllReduce(part.p, part.p, size_t(blk_elems), NcclType<T>::value, ncclSum, comm, s));
        }
        if (g.rank == owner) {
            if (have_terms) {
                rv_axpby_kernel<T><<<rv_grid(blk_elems), 256, 0, s>>>(X.blk(i), T(-1), part.as<T>(), X.blk(i), blk_elems);
                SB_TRY(launch_status());
            }
            const T* Wi = W.as<T>() + my_diag_idx[size_t(i)] * int64_t(nblk) * FACTOR_IB * FACTOR_IB;
            const int st_small = trsm_small<T>(lower, op, int(X.rows(i)), nrhs, A.tile_as<T>(i, i), ld, Wi,
                                               pb.at<T>(xblk_off) + i, 0, ld, 1, s);
            if (st_small == SB200_ENOTSUP) {
                DevBuf w2;
                SB_TRY(w2.alloc(size_t(nblk) * FACTOR_IB * FACTOR_IB * sizeof(T)));
                SB_TRY(trsm_colmajor<T>(true, lower, op, unit, int(X.rows(i)), nrhs, T(1), A.tile_as<T>(i, i), ld,
                                        pb.at<T>(xblk_off) + i, 0, ld, 1, w2.as<T>(), s));
                CUDA_TRY(cudaStreamSynchronize(s));
            }
            else SB_TRY(st_small);
        }
        if (g.size() > 1) NCCL_TRY(ncclBroadcast(X.blk(i), X.blk(i), size_t(blk_elems), NcclType<T>::value, owner, g.world, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return SB200_OK;
}

// R = B - A X on replicated block vectors; A Hermitian (lower tiles) or general.  Local products into a zeroed
// accumulator, one all-reduce, then R = B + acc.
template <typename T>
int residual_dist(Matrix& A, const RepVec<T>& Bv, const RepVec<T>& X, RepVec<T>& R, cudaStream_t s)
{
    Grid& g = *A.g;
    const int64_t nt = A.nt, nb = A.nb, te = A.tile_elems();
    const int ld = int(nb), nrhs = X.nrhs;
    const bool herm = (A.kind == 'H');
    CUDA_TRY(cudaMemsetAsync(R.base, 0, R.elems() * sizeof(T), s));
    struct Step { std::vector<Batch> below, above, diag; };
    std::vector<Step> steps(static_cast<size_t>(nt));
    std::vector<const T*> dptr;
    std::vector<int> ddim;
    std::vector<int64_t> didx(size_t(nt), -1);
    if (herm)
        for (int64_t k = 0; k < nt; ++k)
            if (A.is_local(k, k)) { didx[size_t(k)] = int64_t(dptr.size()); dptr.push_back(A.tile_as<T>(k, k)); ddim.push_back(int(A.tile_mb(k))); }
    DevBuf dfull, ddims;
    SB_TRY(dfull.alloc(std::max<size_t>(dptr.size(), 1) * te * sizeof(T)));
    SB_TRY(ddims.alloc(std::max<size_t>(ddim.size(), 1) * sizeof(int)));
    PlanBuffer pb;
    for (int64_t k = 0; k < nt; ++k) {
        Step& st = steps[size_t(k)];
        if (herm) {
            for (int64_t i = k + 1; i < nt; ++i)
                if (A.is_local(i, k))
                    batch_add(st.below, int(X.rows(i)), nrhs, int(X.rows(k)), 0, A.tile_as<T>(i, k), X.blk(k), R.blk(i));
            for (int64_t i = 0; i < k; ++i)
                if (A.is_local(k, i))
                    batch_add(st.above, int(X.rows(i)), nrhs, int(X.rows(k)), 0, A.tile_as<T>(k, i), X.blk(k), R.blk(i));
            if (didx[size_t(k)] >= 0)
                batch_add(st.diag, int(X.rows(k)), nrhs, int(X.rows(k)), 0, dfull.as<T>() + didx[size_t(k)] * te, X.blk(k), R.blk(k));
        }
        else {
            for (int64_t i = 0; i < A.mt; ++i)
                if (A.is_local(i, k))
                    batch_add(st.below, int(X.rows(i)), nrhs, int(X.rows(k)), 0, A.tile_as<T>(i, k), X.blk(k), R.blk(i));
        }
        pb.reserve(st.below); pb.reserve(st.above); pb.reserve(st.diag);
    }
    const size_t dptr_off = pb.push(dptr);
    SB_TRY(pb.upload(s));
    if (! dptr.empty()) {
        CUDA_TRY(cudaMemcpyAsync(ddims.p, ddim.data(), ddim.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        rv_he_fill_kernel<T><<<dim3(64, unsigned(dptr.size())), 256, 0, s>>>(pb.at<const T>(dptr_off), ddims.as<int>(), dfull.as<T>(), ld, te);
        SB_TRY(launch_status());
    }
    for (int64_t k = 0; k < nt; ++k) {
        const Step& st = steps[size_t(k)];
        SB_TRY(launch_batches<T>(st.below, pb, 'N', 'N', T(-1), T(1), ld, 0, s));
        SB_TRY(launch_batches<T>(st.above, pb, 'T', 'N', T(-1), T(1), ld, 0, s));
        SB_TRY(launch_batches<T>(st.diag, pb, 'N', 'N', T(-1), T(1), ld, 0, s));
    }
    if (g.size() > 1) NCCL_TRY(ncclAllReduce(R.base, R.base, R.elems(), NcclType<T>::value, ncclSum, g.world, s));
    rv_axpby_kernel<T><<<rv_grid(int64_t(R.elems())), 256, 0, s>>>(Bv.base, T(1), R.base, R.base, int64_t(R.elems()));
    SB_TRY(launch_status());
    CUDA_TRY(cudaStreamSynchronize(s));
    return SB200_OK;
}

// norm(Norm::Inf, A) on the grid: per-tile row (and, Hermitian off-diagonal, column) sums from the batched tile-norm
// kernels (norms.cu), accumulated per global row on the host in a fixed order, summed over ranks, maximum taken.
int norm_inf_dist_d(Matrix& A, double* out, cudaStream_t s)
{
    Grid& g = *A.g;
    const int64_t nb = A.nb;
    const bool herm = (A.kind == 'H');
    std::vector<double> rowsum(size_t(std::max<int64_t>(A.m, 1)), 0.0);
    struct Item { int64_t i, j; };
    // group local tiles by (rows, cols, diagonal-of-Hermitian)
    std::vector<std::vector<Item>> groups;
    std::vector<std::array<int, 3>> keys;
    for (int64_t j = g.pcol; j < A.nt; j += g.q)
        for (int64_t i = g.prow; i < A.mt; i += g.p) {
            if (! A.stored(i, j)) continue;
            const std::array<int, 3> key{int(A.tile_mb(i)), int(A.tile_nb(j)), (herm && i == j) ? 1 : 0};
            size_t gi = 0;
            for (; gi < keys.size(); ++gi) if (keys[gi] == key) break;
            if (gi == keys.size()) { keys.push_back(key); groups.emplace_back(); }
            groups[gi].push_back({i, j});
        }
    for (size_t gi = 0; gi < groups.size(); ++gi) {
        const auto& items = groups[gi];
        const int mb = keys[gi][0], nbc = keys[gi][1], diag = keys[gi][2];
        std::vector<const double*> ptrs;
        for (const auto& it : items) ptrs.push_back(A.tile(it.i, it.j));
        DevBuf dp, dv;
        const int64_t ldv = diag ? mb : (herm ? mb + nbc : mb);
        SB_TRY(dp.alloc(ptrs.size() * sizeof(void*)));
        SB_TRY(dv.alloc(ptrs.size() * size_t(ldv) * sizeof(double)));
        CUDA_TRY(cudaMemcpyAsync(dp.p, ptrs.data(), ptrs.size() * sizeof(void*), cudaMemcpyHostToDevice, s));
        int st;
        if (diag)      st = sb200_henorm_batched_d('I', 'L', mb, dp.as<const double*>(), nb, dv.as<double>(), ldv, int64_t(ptrs.size()), s);
        else if (herm) st = sb200_synorm_offdiag_batched_d('I', mb, nbc, dp.as<const double*>(), nb, dv.as<double>(), ldv, int64_t(ptrs.size()), s);
        else           st = sb200_genorm_batched_d('I', 'M', mb, nbc, dp.as<const double*>(), nb, dv.as<double>(), ldv, int64_t(ptrs.size()), s);
        SB_TRY(st);
        std::vector<double> hv(ptrs.size() * size_t(ldv));
        CUDA_TRY(cudaMemcpyAsync(hv.data(), dv.p, hv.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        for (size_t t = 0; t < items.size(); ++t) {
            const double* v = hv.data() + t * size_t(ldv);
            const int64_t r0 = items[t].i * nb, c0 = items[t].j * nb;
            if (diag || ! herm) { for (int r = 0; r < mb; ++r) rowsum[size_t(r0 + r)] += v[r]; }
            else {
                // synorm_offdiag: column sums in v[0 .. nbc), row sums in v[nbc .. nbc + mb)
                for (int c = 0; c < nbc; ++c) rowsum[size_t(c0 + c)] += v[c];            // mirrored part: rows of block j
                for (int r = 0; r < mb; ++r)  rowsum[size_t(r0 + r)] += v[nbc + r];
            }
        }
    }
    DevBuf red;
    SB_TRY(red.alloc(rowsum.size() * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(red.p, rowsum.data(), rowsum.size() * sizeof(double), cudaMemcpyHostToDevice, s));
    if (g.size() > 1) NCCL_TRY(ncclAllReduce(red.p, red.p, rowsum.size(), ncclDouble, ncclSum, g.world, s));
    CUDA_TRY(cudaMemcpyAsync(rowsum.data(), red.p, rowsum.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    double mx = 0.0;
    for (int64_t r = 0; r < A.m; ++r) { const double v = rowsum[size_t(r)]; mx = (v > mx || v != v) ? v : mx; }
    *out = mx;
    return SB200_OK;
}

template <typename T>
int col_max_rep(const RepVec<T>& X, std::vector<double>& out, double* dscratch, cudaStream_t s)
{
    out.assign(size_t(X.nrhs), 0.0);
    if (X.nrhs == 0 || X.m == 0) return SB200_OK;
    rv_colmax_kernel<T><<<unsigned(X.nrhs), 256, 0, s>>>(X.base, dscratch, X.m, int(X.nb), X.nrhs);
    SB_TRY(launch_status());
    CUDA_TRY(cudaMemcpyAsync(out.data(), dscratch, out.size() * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SB200_OK;
}

template <typename S, typename D>
int convert_rep(const RepVec<S>& a, RepVec<D>& b, cudaStream_t s)
{
    const int64_t count = int64_t(a.elems());
    if (count == 0) return SB200_OK;
    rv_convert_kernel<S, D><<<rv_grid(count), 256, 0, s>>>(a.base, b.base, count);
    return launch_status();
}

} // namespace

// potrs on a p x q grid (B: one tile column, nrhs <= nb)
template <typename T>
int potrs_dist(Matrix& A, Matrix& B, cudaStream_t s)
{
    if (! dist_solve_enabled()) return SB200_ENOTSUP;
    if (A.kind != 'H' || B.kind != 'G' || B.m != A.n || B.nb != A.nb || B.g != A.g) return SB200_EINVAL;
    if (B.nt > 1) return SB200_ENOTSUP;              // nrhs > nb: not served on a grid yet
    if (B.nt == 0 || A.nt == 0) return SB200_OK;
    RepVec<T> X;
    SB_TRY(X.alloc(B.m, B.nb, int(B.n)));
    SB_TRY(gather_rep<T>(B, X, s));
    SB_TRY(sweep_dist<T>(A, true, 'N', false, X, s));
    SB_TRY(sweep_dist<T>(A, true, 'T', false, X, s));
    SB_TRY(scatter_rep<T>(X, B, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SB200_OK;
}
template int potrs_dist<float>(Matrix&, Matrix&, cudaStream_t);
template int potrs_dist<double>(Matrix&, Matrix&, cudaStream_t);

// getrs on a p x q grid (B: one tile column); pivots as returned by getrf (host)
template <typename T>
int getrs_dist(Matrix& A, const int64_t* pivots, Matrix& B, cudaStream_t s);

// row gather on a replicated block vector: out(x, :) = in(perm[x], :)   (permuteRows Forward of getrs, src/getrs.cc:45-46)
template <typename T>
__global__ void __launch_bounds__(256) rv_gather_rows_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                             const int* __restrict__ perm, int64_t m, int nb, int nrhs)
{
    for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < m * nrhs; e += int64_t(gridDim.x) * blockDim.x) {
        const int64_t x = e % m, c = e / m, y = perm[x];
        out[(x / nb) * int64_t(nb) * nrhs + (x % nb) + c * nb] = in[(y / nb) * int64_t(nb) * nrhs + (y % nb) + c * nb];
    }
}

static void pivots_to_perm_h(const int64_t* piv, int64_t m, int64_t n, int64_t nb, std::vector<int>& perm)
{
    perm.resize(size_t(m));
    for (int64_t i = 0; i < m; ++i) perm[size_t(i)] = int(i);
    const int64_t mn = std::min(m, n);
    for (int64_t o = 0; o < mn; ++o) {
        const int64_t r2 = (o / nb) * nb + piv[2 * o] * nb + piv[2 * o + 1];
        if (r2 != o && r2 >= 0 && r2 < m) std::swap(perm[size_t(o)], perm[size_t(r2)]);
    }
}

// X <- A^{-1} X from the LU factors: row gather with the composed permutation, unit-L sweep, U sweep
template <typename T>
static int getrs_rep(Matrix& A, const int* dperm, RepVec<T>& X, RepVec<T>& tmp, cudaStream_t s)
{
    const int64_t cnt = X.m * X.nrhs;
    if (cnt > 0) {
        CUDA_TRY(cudaMemcpyAsync(tmp.base, X.base, X.elems() * sizeof(T), cudaMemcpyDeviceToDevice, s));
        rv_gather_rows_kernel<T><<<rv_grid(cnt), 256, 0, s>>>(tmp.base, X.base, dperm, X.m, int(X.nb), X.nrhs);
        SB_TRY(launch_status());
    }
    SB_TRY(sweep_dist<T>(A, true, 'N', true, X, s));
    return sweep_dist<T>(A, false, 'N', false, X, s);
}

template <typename T>
int getrs_dist(Matrix& A, const int64_t* pivots, Matrix& B, cudaStream_t s)
{
    if (! dist_solve_enabled()) return SB200_ENOTSUP;
    if (A.kind != 'G' || A.m != A.n || B.kind != 'G' || B.m != A.n || B.nb != A.nb || B.g != A.g) return SB200_EINVAL;
    if (B.nt > 1) return SB200_ENOTSUP;
    if (B.nt == 0 || A.nt == 0) return SB200_OK;
    std::vector<int> perm;
    pivots_to_perm_h(pivots, A.m, A.n, A.nb, perm);
    DevBuf dp;
    SB_TRY(dp.alloc(perm.size() * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(dp.p, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    RepVec<T> X, tmp;
    SB_TRY(X.alloc(B.m, B.nb, int(B.n))); SB_TRY(tmp.alloc(B.m, B.nb, int(B.n)));
    SB_TRY(gather_rep<T>(B, X, s));
    SB_TRY(getrs_rep<T>(A, dp.as<int>(), X, tmp, s));
    SB_TRY(scatter_rep<T>(X, B, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return SB200_OK;
}
template int getrs_dist<float>(Matrix&, const int64_t*, Matrix&, cudaStream_t);
template int getrs_dist<double>(Matrix&, const int64_t*, Matrix&, cudaStream_t);

// posv_mixed / gesv_mixed <double, float> on a p x q grid (same control flow as solve_mixed_d in solve.cu /
// src/posv_mixed.cc:111-297, src/gesv_mixed.cc:106-300)
int solve_mixed_dist_d(bool hermitian, Matrix& A, int64_t* pivots_out, Matrix& B, Matrix& Xm, int64_t itermax, double tol,
                       bool use_fallback, int* iter_out, int64_t* info_out, double* timers_ms)
{
    if (! dist_solve_enabled()) return SB200_ENOTSUP;
    if (A.dtype != 'd' || B.dtype != 'd' || Xm.dtype != 'd' || A.kind != (hermitian ? 'H' : 'G') || A.m != A.n) return SB200_EINVAL;
    if (B.m != A.n || Xm.m != A.n || Xm.n != B.n || B.nb != A.nb || Xm.nb != A.nb || B.g != A.g || Xm.g != A.g) return SB200_EINVAL;
    if (B.nt > 1) return SB200_ENOTSUP;
    const double eps = std::numeric_limits<double>::epsilon();
    if (tol <= 0) tol = eps * std::sqrt(double(A.m));
    if (itermax < 0) itermax = 30;
    double tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    WallClock total, c;
    total.start();
    cudaStream_t s = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct SG { cudaStream_t s; ~SG() { cudaStreamDestroy(s); } } sguard{s};
    const int nrhs = int(B.n);

    Matrix A_lo;
    struct MG { Matrix& M; ~MG() { if (M.pool) cudaFree(M.pool); } } mguard{A_lo};
    SB_TRY(matrix_alloc(*A.g, 's', A.kind, A.m, A.n, A.nb, A_lo));
    RepVec<double> Bv, Xv, Rv, Tv;
    RepVec<float> Xlo, Tlo;
    SB_TRY(Bv.alloc(B.m, B.nb, nrhs)); SB_TRY(Xv.alloc(B.m, B.nb, nrhs)); SB_TRY(Rv.alloc(B.m, B.nb, nrhs));
    SB_TRY(Xlo.alloc(B.m, B.nb, nrhs));
    if (! hermitian) { SB_TRY(Tv.alloc(B.m, B.nb, nrhs)); SB_TRY(Tlo.alloc(B.m, B.nb, nrhs)); }
    std::vector<int64_t> piv_lo(size_t(2 * std::max<int64_t>(A.m, 1)));
    std::vector<int> perm;
    DevBuf dperm;
    DevBuf dnorm;
    SB_TRY(dnorm.alloc(size_t(std::max(nrhs, 1)) * sizeof(double)));
    std::vector<double> cn_x, cn_r;

    c.start();
    double Anorm = 0;
    SB_TRY(norm_inf_dist_d(A, &Anorm, s));
    const double cte = Anorm * tol;
    SB_TRY(gather_rep<double>(B, Bv, s));
    {
        const int64_t count = A.ntiles_loc * A.tile_elems();
        if (count > 0) {
            rv_convert_kernel<double, float><<<rv_grid(count), 256, 0, s>>>(A.pool, reinterpret_cast<float*>(A_lo.pool), count);
            SB_TRY(launch_status());
        }
    }
    tm[7] = c.stop();

    bool converged = false;
    int iter = 0;
    int64_t info = 0;
    const char* e = getenv("SB200_MIXED_TC05");
    const bool tc = ! (e && atoi(e) == 0);
    c.start();
    if (hermitian) SB_TRY(potrf_driver<float>(A_lo, &info, tc));
    else           SB_TRY(getrf_driver_dist_s(A_lo, piv_lo.data(), &info, tc));
    tm[1] = c.stop();
    auto upload_perm = [&](const int64_t* pv) -> int {
        pivots_to_perm_h(pv, A.m, A.n, A.nb, perm);
        if (! dperm.p) SB_TRY(dperm.alloc(perm.size() * sizeof(int)));
        CUDA_TRY(cudaMemcpyAsync(dperm.p, perm.data(), perm.size() * sizeof(int), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        return SB200_OK;
    };
    if (! hermitian && info == 0) SB_TRY(upload_perm(piv_lo.data()));
    Xm.last_trail_ms = A_lo.last_trail_ms; Xm.last_trail_flops = A_lo.last_trail_flops;
    Xm.last_trail_launches = A_lo.last_trail_launches; Xm.last_panel_ms = A_lo.last_panel_ms;

    auto solve_lo = [&]() -> int {            // Xlo <- A_lo^{-1} Xlo
        c.start();
        if (hermitian) {
            SB_TRY(sweep_dist<float>(A_lo, true, 'N', false, Xlo, s));
            SB_TRY(sweep_dist<float>(A_lo, true, 'T', false, Xlo, s));
        }
        else SB_TRY(getrs_rep<float>(A_lo, dperm.as<int>(), Xlo, Tlo, s));
        tm[2] += c.stop();
        return SB200_OK;
    };
    auto residual = [&]() -> int {
        c.start();
        SB_TRY(residual_dist<double>(A, Bv, Xv, Rv, s));
        tm[3] += c.stop();
        SB_TRY(col_max_rep<double>(Xv, cn_x, dnorm.as<double>(), s));
        SB_TRY(col_max_rep<double>(Rv, cn_r, dnorm.as<double>(), s));
        return SB200_OK;
    };
    auto conv = [&]() { for (size_t i = 0; i < cn_x.size(); ++i) if (cn_r[i] > cn_x[i] * cte) return false; return true; };

    if (info != 0) iter = -3;
    else {
        SB_TRY((convert_rep<double, float>(Bv, Xlo, s)));
        SB_TRY(solve_lo());
        SB_TRY((convert_rep<float, double>(Xlo, Xv, s)));
        SB_TRY(residual());
        if (conv()) { iter = 0; converged = true; }
        for (int64_t iiter = 0; iiter < itermax && ! converged; ++iiter) {
            SB_TRY((convert_rep<double, float>(Rv, Xlo, s)));
            SB_TRY(solve_lo());
            c.start();
            SB_TRY((convert_rep<float, double>(Xlo, Rv, s)));
            rv_axpby_kernel<double><<<rv_grid(int64_t(Xv.elems())), 256, 0, s>>>(Xv.base, 1.0, Rv.base, Xv.base, int64_t(Xv.elems()));
            SB_TRY(launch_status());
            tm[4] += c.stop();
            SB_TRY(residual());
            if (conv()) { iter = int(iiter) + 1; converged = true; }
        }
    }
    if (! converged) {
        if (info == 0) iter = -int(itermax) - 1;
        if (use_fallback) {
            c.start();
            int64_t* pv = pivots_out ? pivots_out : piv_lo.data();
            if (hermitian) SB_TRY(potrf_driver<double>(A, &info, false));
            else           SB_TRY(getrf_driver(A, pv, &info));
            tm[5] = c.stop();
            c.start();
            if (info == 0) {
                CUDA_TRY(cudaMemcpyAsync(Xv.base, Bv.base, Xv.elems() * sizeof(double), cudaMemcpyDeviceToDevice, s));
                if (hermitian) {
                    SB_TRY(sweep_dist<double>(A, true, 'N', false, Xv, s));
                    SB_TRY(sweep_dist<double>(A, true, 'T', false, Xv, s));
                }
                else {
                    SB_TRY(upload_perm(pv));
                    SB_TRY(getrs_rep<double>(A, dperm.as<int>(), Xv, Tv, s));
                }
            }
            tm[6] = c.stop();
        }
    }
    if (converged && pivots_out && ! hermitian)
        memcpy(pivots_out, piv_lo.data(), size_t(2 * std::min(A.m, A.n)) * sizeof(int64_t));
    SB_TRY(scatter_rep<double>(Xv, Xm, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    tm[0] = total.stop();
    Xm.last_ms = tm[0];
    if (timers_ms) memcpy(timers_ms, tm, sizeof(tm));
    if (iter_out) *iter_out = iter;
    if (info_out) *info_out = info;
    return SB200_OK;
}

} // namespace sb200
