// runtime.cu -- grid, tile matrix, generators, host<->device movement, potrf and gemm drivers.
// See runtime.hh for the mapping to the reference.
#include "runtime.hh"
#include "gemm_dmma.cuh"
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <map>
#include <tuple>

namespace sb200 {

// from factor_small.cu
int trsm_colmajor_d(bool left, bool lower, int op, bool unit, int m, int n, double alpha,
                    const double* T, int ldt, double* const* dB, int64_t offB, int ldb, int batch,
                    double* W, cudaStream_t stream);
int potrf_tile_lower_d(int n, double* A, int lda, int* dinfo, int info_base, double* W, cudaStream_t stream);

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return int(e_); } while (0)
#define NCCL_TRY(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { \
    fprintf(stderr, "slate_b200: NCCL error %s at %s:%d\n", ncclGetErrorString(r_), __FILE__, __LINE__); \
    return SB200_ENCCL; } } while (0)
#define SB_TRY(x) do { int s_ = (x); if (s_ != SB200_OK) return s_; } while (0)

// ------------------------------------------------------------------------------------------
// Philox-2x64 test-matrix generator on the device, bit-identical to the reference's matgen
// (matgen/random.cc:54-77 philox_2x64, :83-91 rand_to_real, generate_type_rand.hh:28-79).
// One thread per element of one tile; grid.y = local tile.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_2x64(uint64_t& s0, uint64_t& s1, uint64_t key)
{
    const uint64_t mult = 0x9E3779B97F4A7C15ull, inc = 0xD2B74407B1CE6E93ull;
    #pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r != 0) key += inc;
        const uint64_t lo = s1 * mult, hi = __umul64hi(s1, mult);
        const uint64_t L = s0;
        s0 = lo;
        s1 = hi ^ key ^ L;
    }
}

struct TileDesc { double* ptr; int64_t i0, j0; int mb, nbc; };

__global__ void generate_kernel(const TileDesc* __restrict__ tiles, int ld, int64_t seed,
                                int dominant, double diag_add)
{
    const TileDesc t = tiles[blockIdx.y];
    const int total = t.mb * t.nbc;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int i = e % t.mb, j = e / t.mb;
        uint64_t s0 = uint64_t(t.i0 + i), s1 = uint64_t(t.j0 + j);
        philox_2x64(s0, s1, uint64_t(seed));
        double v = double(s0 >> 11) * (1.0 / 9007199254740992.0);     // 53 bits -> [0, 1)
        if (dominant && (t.i0 + i == t.j0 + j)) v += diag_add;
        t.ptr[i + int64_t(j) * ld] = v;
    }
}

// ------------------------------------------------------------------------------------------
// batches: tiles of one step grouped by (m, n, k, tri) -- the reference's "regions"
// (src/internal/internal_batch.hh:169-347), built once per driver call for all steps.
// ------------------------------------------------------------------------------------------
struct Batch {
    int m, n, k, tri;
    std::vector<const double*> A, B;
    std::vector<double*> C;
    size_t off = 0;                 // offset (in pointers) of A | B | C blocks in the device plan
};

static void batch_add(std::vector<Batch>& v, int m, int n, int k, int tri,
                      const double* A, const double* B, double* C)
{
    for (auto& b : v)
        if (b.m == m && b.n == n && b.k == k && b.tri == tri) {
            b.A.push_back(A); b.B.push_back(B); b.C.push_back(C);
            return;
        }
    Batch b{m, n, k, tri, {A}, {B}, {C}, 0};
    v.push_back(std::move(b));
}

struct PlanBuffer {
    std::vector<const void*> host;
    void** dev = nullptr;
    size_t reserve(std::vector<Batch>& bs)
    {
        size_t first = host.size();
        for (auto& b : bs) {
            b.off = host.size();
            host.insert(host.end(), b.A.begin(), b.A.end());
            host.insert(host.end(), b.B.begin(), b.B.end());
            host.insert(host.end(), b.C.begin(), b.C.end());
        }
        return first;
    }
    size_t push(const std::vector<double*>& v)
    {
        size_t o = host.size();
        host.insert(host.end(), v.begin(), v.end());
        return o;
    }
    int upload(cudaStream_t s)
    {
        if (host.empty()) return SB200_OK;
        CUDA_TRY(cudaMalloc(&dev, host.size() * sizeof(void*)));
        CUDA_TRY(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(void*), cudaMemcpyHostToDevice, s));
        return SB200_OK;
    }
    ~PlanBuffer() { if (dev) cudaFree(dev); }
};

static int launch_batches(const std::vector<Batch>& bs, const PlanBuffer& pb, int opA, int opB,
                          double alpha, double beta, int ld, cudaStream_t s)
{
    for (const auto& b : bs) {
        GemmParamsD p{};
        const size_t cnt = b.C.size();
        p.A = reinterpret_cast<const double* const*>(pb.dev + b.off);
        p.B = reinterpret_cast<const double* const*>(pb.dev + b.off + cnt);
        p.C = reinterpret_cast<double* const*>(pb.dev + b.off + 2 * cnt);
        p.m = b.m; p.n = b.n; p.k = b.k; p.lda = ld; p.ldb = ld; p.ldc = ld;
        p.alpha = alpha; p.beta = beta; p.batch = int(cnt); p.tri = b.tri;
        SB_TRY(launch_gemm_d(opA, opB, p, s));
    }
    return SB200_OK;
}

static double batches_flops(const std::vector<Batch>& bs)
{
    double f = 0;
    for (const auto& b : bs) {
        // ALGORITHMIC flops: a triangle-masked (herk/syrk diagonal) tile counts n(n+1)k
        // (blaspp/include/blas/flops.hh syrk), whatever the kernel computes above the diagonal
        const double per = b.tri ? double(b.n) * (b.n + 1.0) * b.k : 2.0 * b.m * b.n * b.k;
        f += per * double(b.C.size());
    }
    return f;
}

static int64_t batches_launches(const std::vector<Batch>& bs) { return int64_t(bs.size()); }

struct Streams {
    cudaStream_t panel = nullptr, trail = nullptr;
    std::vector<cudaEvent_t> ev;
    std::vector<cudaEvent_t> tev;          // timing event pairs around the trailing-update launches
    std::vector<cudaEvent_t> pev;          // timing event pairs around the panel work of every step
    int ptime(cudaStream_t s)
    {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        pev.push_back(e);
        CUDA_TRY(cudaEventRecord(e, s));
        return SB200_OK;
    }
    double panel_ms()
    {
        double tot = 0;
        for (size_t i = 0; i + 1 < pev.size(); i += 2) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, pev[i], pev[i + 1]) == cudaSuccess) tot += ms;
        }
        return tot;
    }
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int time_begin(cudaStream_t s)
    {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        tev.push_back(e);
        CUDA_TRY(cudaEventRecord(e, s));
        return SB200_OK;
    }
    int time_end(cudaStream_t s) { return time_begin(s); }
    double timed_ms()
    {
        double tot = 0;
        for (size_t i = 0; i + 1 < tev.size(); i += 2) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, tev[i], tev[i + 1]) == cudaSuccess) tot += ms;
        }
        return tot;
    }
    int init(size_t nevents)
    {
        int lo, hi;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&panel, cudaStreamNonBlocking, hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&trail, cudaStreamNonBlocking, lo));
        ev.resize(nevents);
        for (auto& e : ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreate(&t0));
        CUDA_TRY(cudaEventCreate(&t1));
        return SB200_OK;
    }
    ~Streams()
    {
        for (auto e : ev) if (e) cudaEventDestroy(e);
        for (auto e : tev) if (e) cudaEventDestroy(e);
        for (auto e : pev) if (e) cudaEventDestroy(e);
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        if (panel) cudaStreamDestroy(panel);
        if (trail) cudaStreamDestroy(trail);
    }
};

struct DevBuf {
    void* p = nullptr;
    int alloc(size_t bytes) { CUDA_TRY(cudaMalloc(&p, bytes ? bytes : 16)); return SB200_OK; }
    ~DevBuf() { if (p) cudaFree(p); }
};

// ------------------------------------------------------------------------------------------
// potrf: right-looking tile Cholesky with lookahead 1, lower.
// reference schedule: src/potrf.cc:84-195 (panel task = potrf + tileBcast + trsm + listBcastMT,
// lookahead task = herk/gemm on column k+1, trailing task = herk on the rest).
// ------------------------------------------------------------------------------------------
int potrf_driver(Matrix& A, int64_t* info_out)
{
    Grid& g = *A.g;
    if (A.kind != 'H' || A.layout != 'C' || A.m != A.n) return SB200_EINVAL;
    CUDA_TRY(cudaDeviceSynchronize());       // inputs may have been produced on any stream
    const int64_t nt = A.nt, nb = A.nb;
    const int ld = int(nb);
    const bool multi = g.size() > 1;
    const int64_t te = A.tile_elems();
    const int64_t rows_max = (A.mt + g.p - 1) / g.p;        // panel workspace slots per process row

    DevBuf ws, dbuf, work, dinfo;
    if (multi) {
        SB_TRY(ws.alloc(size_t(2) * g.p * rows_max * te * sizeof(double)));
        SB_TRY(dbuf.alloc(size_t(2) * te * sizeof(double)));
    }
    SB_TRY(work.alloc(size_t(8) * 64 * 64 * sizeof(double) * 2));
    SB_TRY(dinfo.alloc(sizeof(int)));
    double* W_potrf = static_cast<double*>(work.p);
    double* W_trsm  = W_potrf + 64 * 64;

    auto pbuf = [&](int64_t i, int64_t k) -> double* {     // where step k's factored tile (i,k) is read from
        if (! multi) return A.tile(i, k);
        return static_cast<double*>(ws.p) + ((k & 1) * g.p * rows_max + (i % g.p) * rows_max + i / g.p) * te;
    };

    // ---- plan: every pointer batch of every step
    struct Step {
        std::vector<Batch> la, tr;           // lookahead column k+1 / trailing columns >= k+2
        std::vector<double*> panel;          // local tiles (i,k), i > k, full height
        std::vector<double*> panel_last;     // ragged last block row
        size_t panel_off = 0, panel_last_off = 0;
    };
    std::vector<Step> steps(nt);
    PlanBuffer pb;
    for (int64_t k = 0; k < nt; ++k) {
        Step& s = steps[k];
        const int kw = int(A.tile_nb(k));
        for (int64_t j = k + 1; j < nt; ++j)
            for (int64_t i = j; i < nt; ++i) {
                if (! A.is_local(i, j)) continue;
                auto& dst = (j == k + 1) ? s.la : s.tr;
                batch_add(dst, int(A.tile_mb(i)), int(A.tile_nb(j)), kw, i == j ? 1 : 0,
                          pbuf(i, k), pbuf(j, k), A.tile(i, j));
            }
        for (int64_t i = k + 1; i < nt; ++i)
            if (A.is_local(i, k)) {
                if (A.tile_mb(i) == nb) s.panel.push_back(A.tile(i, k));
                else                    s.panel_last.push_back(A.tile(i, k));
            }
        pb.reserve(s.la);
        pb.reserve(s.tr);
        s.panel_off = pb.push(s.panel);
        s.panel_last_off = pb.push(s.panel_last);
    }

    Streams st;
    PhaseTimer ph;
    SB_TRY(st.init(size_t(2 * nt)));
    double trail_flops = 0;
    int64_t trail_launches = 0;
    auto P_done = [&](int64_t k) { return st.ev[k]; };
    auto T_done = [&](int64_t k) { return st.ev[nt + k]; };
    SB_TRY(pb.upload(st.panel));
    CUDA_TRY(cudaMemsetAsync(dinfo.p, 0, sizeof(int), st.panel));
    CUDA_TRY(cudaStreamSynchronize(st.panel));
    CUDA_TRY(cudaEventRecord(st.t0, st.panel));

    for (int64_t k = 0; k < nt; ++k) {
        Step& s = steps[k];
        const int kw = int(A.tile_nb(k));
        const int owner = g.rank_of(k, k);
        const bool in_col = (g.pcol == int(k % g.q));
        cudaStream_t P = st.panel, T = st.trail;

        // -- lookahead update of column k by panel k-1 (after every older trailing update)
        if (k >= 1) {
            if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(P, T_done(k - 2), 0));
            ph.begin("la_update", P);
            SB_TRY(launch_batches(steps[k - 1].la, pb, 'N', 'T', -1.0, 1.0, ld, P));
            ph.end(P);
        }
        // -- diagonal tile
        SB_TRY(st.ptime(P));
        const double* Lkk = nullptr;
        ph.begin("potrf_tile", P);
        if (g.rank == owner)
            SB_TRY(potrf_tile_lower_d(kw, A.tile(k, k), ld, static_cast<int*>(dinfo.p), int(k * nb), W_potrf, P));
        ph.end(P);
        if (k + 1 < nt) {
            if (multi) {
                double* db = static_cast<double*>(dbuf.p) + (k & 1) * te;
                if (in_col && g.p > 1) {
                    const double* src = (g.rank == owner) ? A.tile(k, k) : db;
                    NCCL_TRY(ncclBroadcast(src, db, size_t(te), ncclDouble, int(k % g.p), g.col_comm, P));
                    Lkk = db;
                }
                else if (in_col) Lkk = A.tile(k, k);
            }
            else Lkk = A.tile(k, k);
            // -- panel solve A(i,k) <- A(i,k) L_kk^{-T}
            ph.begin("panel_trsm", P);
            if (in_col) {
                if (! s.panel.empty())
                    SB_TRY(trsm_colmajor_d(false, true, 'T', false, int(nb), kw, 1.0, Lkk, ld,
                                           reinterpret_cast<double* const*>(pb.dev + s.panel_off), 0, ld,
                                           int(s.panel.size()), W_trsm, P));
                if (! s.panel_last.empty())
                    SB_TRY(trsm_colmajor_d(false, true, 'T', false, int(A.tile_mb(nt - 1)), kw, 1.0, Lkk, ld,
                                           reinterpret_cast<double* const*>(pb.dev + s.panel_last_off), 0, ld,
                                           int(s.panel_last.size()), W_trsm, P));
            }
            ph.end(P);
            // -- panel broadcast: every rank receives the whole factored block column
            ph.begin("panel_bcast", P);
            if (multi) {
                if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(P, T_done(k - 2), 0));   // ws[k&1] is free again
                NCCL_TRY(ncclGroupStart());
                for (int r = 0; r < g.p; ++r) {
                    // tiles (i,k), i > k, i % p == r: contiguous in the root's pool
                    int64_t i0 = k + 1 + ((r - (k + 1)) % g.p + g.p) % g.p;
                    if (i0 >= nt) continue;
                    const int64_t cnt = (nt - 1 - i0) / g.p + 1;
                    const int root = g.rank_of(i0, k);
                    const double* src = (g.rank == root) ? A.tile(i0, k) : pbuf(i0, k);
                    NCCL_TRY(ncclBroadcast(src, pbuf(i0, k), size_t(cnt * te), ncclDouble, root, g.world, P));
                }
                NCCL_TRY(ncclGroupEnd());
            }
            ph.end(P);
        }
        SB_TRY(st.ptime(P));
        CUDA_TRY(cudaEventRecord(P_done(k), P));
        // -- trailing update of columns >= k+2
        CUDA_TRY(cudaStreamWaitEvent(T, P_done(k), 0));
        if (! s.tr.empty()) {
            SB_TRY(st.time_begin(T));
            SB_TRY(launch_batches(s.tr, pb, 'N', 'T', -1.0, 1.0, ld, T));
            SB_TRY(st.time_end(T));
            trail_flops += batches_flops(s.tr);
            trail_launches += batches_launches(s.tr);
        }
        CUDA_TRY(cudaEventRecord(T_done(k), T));
    }
    CUDA_TRY(cudaStreamWaitEvent(st.panel, T_done(nt - 1), 0));
    CUDA_TRY(cudaEventRecord(st.t1, st.panel));
    int hinfo = 0;
    CUDA_TRY(cudaMemcpyAsync(&hinfo, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, st.panel));
    CUDA_TRY(cudaStreamSynchronize(st.panel));
    CUDA_TRY(cudaStreamSynchronize(st.trail));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, st.t0, st.t1));
    A.last_ms = ms;
    A.last_trail_ms = st.timed_ms();
    A.last_trail_flops = trail_flops;
    A.last_trail_launches = trail_launches;
    A.last_panel_ms = st.panel_ms();
    ph.report("potrf", g.rank);
    int64_t info = hinfo;
    if (multi) {
        // first failing minor over all ranks (reference: internal_reduce_info.cc:23-38, MPI_MIN)
        DevBuf red;
        SB_TRY(red.alloc(sizeof(int64_t)));
        int64_t v = info ? info : INT64_MAX;
        CUDA_TRY(cudaMemcpyAsync(red.p, &v, sizeof(v), cudaMemcpyHostToDevice, st.panel));
        NCCL_TRY(ncclAllReduce(red.p, red.p, 1, ncclInt64, ncclMin, g.world, st.panel));
        CUDA_TRY(cudaMemcpyAsync(&v, red.p, sizeof(v), cudaMemcpyDeviceToHost, st.panel));
        CUDA_TRY(cudaStreamSynchronize(st.panel));
        info = (v == INT64_MAX) ? 0 : v;
    }
    if (info_out) *info_out = info;
    return SB200_OK;
}

// ------------------------------------------------------------------------------------------
// gemm: C = alpha A B + beta C, SUMMA over block columns of A (reference: src/gemmC.cc:39-202).
// Step k: A(:,k) goes along process rows, B(k,:) down process columns (one step ahead of the
// multiply, on the panel stream), then ONE batched launch updates every local C tile.
// ------------------------------------------------------------------------------------------
int gemm_driver(double alpha, Matrix& A, Matrix& B, double beta, Matrix& C)
{
    Grid& g = *C.g;
    if (A.g != &g || B.g != &g) return SB200_EINVAL;
    if (A.kind != 'G' || B.kind != 'G' || C.kind != 'G') return SB200_EINVAL;
    if (A.m != C.m || B.n != C.n || A.n != B.m || A.nb != C.nb || B.nb != C.nb) return SB200_EINVAL;
    const int64_t kt = A.nt, nb = C.nb, te = C.tile_elems();
    const int ld = int(nb);
    const bool multi = g.size() > 1;
    CUDA_TRY(cudaDeviceSynchronize());       // inputs may have been produced on any stream

    DevBuf wsA, wsB;
    if (multi) {
        SB_TRY(wsA.alloc(size_t(2) * C.mt_loc * te * sizeof(double)));
        SB_TRY(wsB.alloc(size_t(2) * C.nt_loc * te * sizeof(double)));
    }
    auto a_src = [&](int64_t i, int64_t k) -> double* {
        if (! multi) return A.tile(i, k);
        return static_cast<double*>(wsA.p) + ((k & 1) * C.mt_loc + (i - g.prow) / g.p) * te;
    };
    auto b_src = [&](int64_t k, int64_t j) -> double* {
        if (! multi) return B.tile(k, j);
        return static_cast<double*>(wsB.p) + ((k & 1) * C.nt_loc + (j - g.pcol) / g.q) * te;
    };

    std::vector<std::vector<Batch>> plan(kt);
    PlanBuffer pb;
    for (int64_t k = 0; k < kt; ++k) {
        for (int64_t j = g.pcol; j < C.nt; j += g.q)
            for (int64_t i = g.prow; i < C.mt; i += g.p)
                batch_add(plan[k], int(C.tile_mb(i)), int(C.tile_nb(j)), int(A.tile_nb(k)), 0,
                          a_src(i, k), b_src(k, j), C.tile(i, j));
        pb.reserve(plan[k]);
    }
    Streams st;
    double trail_flops = 0;
    int64_t trail_launches = 0;
    SB_TRY(st.init(size_t(2 * kt)));
    auto P_done = [&](int64_t k) { return st.ev[k]; };
    auto T_done = [&](int64_t k) { return st.ev[kt + k]; };
    SB_TRY(pb.upload(st.panel));
    CUDA_TRY(cudaStreamSynchronize(st.panel));
    CUDA_TRY(cudaEventRecord(st.t0, st.panel));
    CUDA_TRY(cudaStreamWaitEvent(st.trail, st.t0, 0));

    for (int64_t k = 0; k < kt; ++k) {
        cudaStream_t P = st.panel, T = st.trail;
        if (multi) {
            if (k >= 2) CUDA_TRY(cudaStreamWaitEvent(P, T_done(k - 2), 0));
            NCCL_TRY(ncclGroupStart());
            if (C.mt_loc > 0 && g.q > 1) {
                const int root = int(k % g.q);
                const double* src = (g.pcol == root) ? A.tile(g.prow, k) : a_src(g.prow, k);
                NCCL_TRY(ncclBroadcast(src, a_src(g.prow, k), size_t(C.mt_loc * te), ncclDouble, root, g.row_comm, P));
            }
            if (g.p > 1) {
                const int root = int(k % g.p);
                for (int64_t j = g.pcol; j < C.nt; j += g.q) {
                    const double* src = (g.prow == root) ? B.tile(k, j) : b_src(k, j);
                    NCCL_TRY(ncclBroadcast(src, b_src(k, j), size_t(te), ncclDouble, root, g.col_comm, P));
                }
            }
            NCCL_TRY(ncclGroupEnd());
            // operands that did not need a broadcast are copied so that every tile is read from ws
            if (g.q == 1 && C.mt_loc > 0)
                CUDA_TRY(cudaMemcpyAsync(a_src(g.prow, k), A.tile(g.prow, k), size_t(C.mt_loc * te) * sizeof(double),
                                         cudaMemcpyDeviceToDevice, P));
            if (g.p == 1)
                for (int64_t j = g.pcol; j < C.nt; j += g.q)
                    CUDA_TRY(cudaMemcpyAsync(b_src(k, j), B.tile(k, j), size_t(te) * sizeof(double),
                                             cudaMemcpyDeviceToDevice, P));
            CUDA_TRY(cudaEventRecord(P_done(k), P));
            CUDA_TRY(cudaStreamWaitEvent(T, P_done(k), 0));
        }
        SB_TRY(st.time_begin(T));
        SB_TRY(launch_batches(plan[k], pb, 'N', 'N', alpha, k == 0 ? beta : 1.0, ld, T));
        SB_TRY(st.time_end(T));
        trail_flops += batches_flops(plan[k]);
        trail_launches += batches_launches(plan[k]);
        CUDA_TRY(cudaEventRecord(T_done(k), T));
    }
    CUDA_TRY(cudaEventRecord(st.t1, st.trail));
    CUDA_TRY(cudaStreamSynchronize(st.trail));
    CUDA_TRY(cudaStreamSynchronize(st.panel));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, st.t0, st.t1));
    C.last_ms = ms;
    C.last_trail_ms = st.timed_ms();
    C.last_trail_flops = trail_flops;
    C.last_trail_launches = trail_launches;
    return SB200_OK;
}

} // namespace sb200

using namespace sb200;

struct sb200_grid_s   { Grid g; };
struct sb200_matrix_s { Matrix A; };

extern "C" {

int sb200_grid_unique_id(void* out_id_128)
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return SB200_ENCCL;
    memcpy(out_id_128, &id, sizeof(id));
    return SB200_OK;
}

int sb200_grid_create(int p, int q, int rank, const void* nccl_unique_id, sb200_grid_t* out)
{
    if (p < 1 || q < 1 || rank < 0 || rank >= p * q || ! out) return SB200_EINVAL;
    if (sb200_device_count() < 1) return SB200_ENODEV;
    auto* h = new sb200_grid_s();
    Grid& g = h->g;
    g.p = p; g.q = q; g.rank = rank; g.prow = rank % p; g.pcol = rank / p;
    if (p * q > 1) {
        if (! nccl_unique_id) { delete h; return SB200_EINVAL; }
        ncclUniqueId id;
        memcpy(&id, nccl_unique_id, sizeof(id));
        if (ncclCommInitRank(&g.world, p * q, id, rank) != ncclSuccess) { delete h; return SB200_ENCCL; }
        if (ncclCommSplit(g.world, g.prow, g.pcol, &g.row_comm, nullptr) != ncclSuccess
            || ncclCommSplit(g.world, p + g.pcol, g.prow, &g.col_comm, nullptr) != ncclSuccess
            || ncclCommSplit(g.world, p + g.pcol, g.prow, &g.col_comm2, nullptr) != ncclSuccess) {
            delete h; return SB200_ENCCL;
        }
    }
    *out = h;
    return SB200_OK;
}

int sb200_grid_destroy(sb200_grid_t h)
{
    if (! h) return SB200_OK;
    if (h->g.row_comm) ncclCommDestroy(h->g.row_comm);
    if (h->g.col_comm) ncclCommDestroy(h->g.col_comm);
    if (h->g.col_comm2) ncclCommDestroy(h->g.col_comm2);
    if (h->g.world) ncclCommDestroy(h->g.world);
    delete h;
    return SB200_OK;
}

int sb200_matrix_create_d(sb200_grid_t gh, int kind, int layout, int64_t m, int64_t n, int64_t nb,
                          sb200_matrix_t* out)
{
    if (! gh || ! out || (kind != 'G' && kind != 'H') || layout != 'C') return layout == 'R' ? SB200_ENOTSUP : SB200_EINVAL;
    if (m < 0 || n < 0 || nb < 1 || (kind == 'H' && m != n)) return SB200_EINVAL;
    auto* h = new sb200_matrix_s();
    Matrix& A = h->A;
    Grid& g = gh->g;
    A.g = &g; A.kind = kind; A.layout = layout; A.m = m; A.n = n; A.nb = nb;
    A.mt = ceil_div(m, nb); A.nt = ceil_div(n, nb);
    A.mt_loc = A.mt > g.prow ? (A.mt - g.prow + g.p - 1) / g.p : 0;
    A.nt_loc = A.nt > g.pcol ? (A.nt - g.pcol + g.q - 1) / g.q : 0;
    A.col_start.resize(A.nt_loc + 1);
    int64_t cnt = 0;
    for (int64_t jl = 0; jl < A.nt_loc; ++jl) {
        A.col_start[jl] = cnt;
        if (kind == 'G') cnt += A.mt_loc;
        else {
            const int64_t j = g.pcol + jl * g.q;
            const int64_t il0 = A.first_local_row(j);
            cnt += A.mt_loc > il0 ? A.mt_loc - il0 : 0;
        }
    }
    A.col_start[A.nt_loc] = cnt;
    A.ntiles_loc = cnt;
    const size_t bytes = size_t(cnt) * A.tile_elems() * sizeof(double);
    if (cudaMalloc(reinterpret_cast<void**>(&A.pool), bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError(); delete h; return SB200_ENOMEM;
    }
    *out = h;
    return SB200_OK;
}

int sb200_matrix_destroy(sb200_matrix_t h)
{
    if (! h) return SB200_OK;
    if (h->A.pool) cudaFree(h->A.pool);
    delete h;
    return SB200_OK;
}

int64_t sb200_matrix_local_tiles(sb200_matrix_t h) { return h ? h->A.ntiles_loc : 0; }
double  sb200_last_driver_ms(sb200_matrix_t h) { return h ? h->A.last_ms : 0.0; }
double  sb200_last_driver_panel_ms(sb200_matrix_t h) { return h ? h->A.last_panel_ms : 0.0; }

int sb200_last_driver_stats(sb200_matrix_t h, double* out4)
{
    if (! h || ! out4) return SB200_EINVAL;
    out4[0] = h->A.last_ms;
    out4[1] = h->A.last_trail_ms;
    out4[2] = h->A.last_trail_flops;
    out4[3] = double(h->A.last_trail_launches);
    return SB200_OK;
}

int sb200_matrix_generate_d(sb200_matrix_t h, int kind_code, int64_t seed, sb200_stream_t stream)
{
    if (! h || (kind_code != 0 && kind_code != 1)) return SB200_EINVAL;
    Matrix& A = h->A;
    std::vector<TileDesc> td;
    for (int64_t j = A.g->pcol; j < A.nt; j += A.g->q)
        for (int64_t i = A.g->prow; i < A.mt; i += A.g->p)
            if (A.stored(i, j))
                td.push_back({A.tile(i, j), i * A.nb, j * A.nb, int(A.tile_mb(i)), int(A.tile_nb(j))});
    if (td.empty()) return SB200_OK;
    TileDesc* dtd = nullptr;
    CUDA_TRY(cudaMalloc(&dtd, td.size() * sizeof(TileDesc)));
    cudaStream_t s = cudaStream_t(stream);
    CUDA_TRY(cudaMemcpyAsync(dtd, td.data(), td.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, s));
    int status = SB200_OK;
    for (size_t o = 0; o < td.size() && status == SB200_OK; o += 32768) {
        const unsigned cnt = unsigned(std::min<size_t>(32768, td.size() - o));
        generate_kernel<<<dim3(32, cnt), 256, 0, s>>>(dtd + o, int(A.nb), seed, kind_code, double(A.n));
        status = launch_status();
    }
    cudaStreamSynchronize(s);
    cudaFree(dtd);
    return status;
}

static int matrix_host_copy(Matrix& A, double* hA, int64_t lda, bool to_host, cudaStream_t s)
{
    if (lda < (A.m > 1 ? A.m : 1)) return SB200_EINVAL;
    for (int64_t j = A.g->pcol; j < A.nt; j += A.g->q)
        for (int64_t i = A.g->prow; i < A.mt; i += A.g->p) {
            if (! A.stored(i, j)) continue;
            double* d = A.tile(i, j);
            double* hp = hA + i * A.nb + j * A.nb * lda;
            const size_t w = size_t(A.tile_mb(i)) * sizeof(double), hgt = size_t(A.tile_nb(j));
            if (to_host) CUDA_TRY(cudaMemcpy2DAsync(hp, size_t(lda) * 8, d, size_t(A.nb) * 8, w, hgt, cudaMemcpyDeviceToHost, s));
            else         CUDA_TRY(cudaMemcpy2DAsync(d, size_t(A.nb) * 8, hp, size_t(lda) * 8, w, hgt, cudaMemcpyHostToDevice, s));
        }
    return SB200_OK;
}

int sb200_matrix_from_host_d(sb200_matrix_t h, const double* hA, int64_t lda, sb200_stream_t stream)
{
    if (! h || ! hA) return SB200_EINVAL;
    return matrix_host_copy(h->A, const_cast<double*>(hA), lda, false, cudaStream_t(stream));
}

int sb200_matrix_to_host_d(sb200_matrix_t h, double* hA, int64_t lda, sb200_stream_t stream)
{
    if (! h || ! hA) return SB200_EINVAL;
    return matrix_host_copy(h->A, hA, lda, true, cudaStream_t(stream));
}

// local tiles <-> a packed host buffer in pool order (local block column, then local block row;
// every tile nb*nb, ld = nb): the host-side layout a caller gets from Matrix::insertLocalTiles with
// its own contiguous storage (include/slate/Matrix.hh:631-662).  One contiguous copy.
int sb200_matrix_from_host_local_d(sb200_matrix_t h, const double* htiles, sb200_stream_t stream)
{
    if (! h || ! htiles) return SB200_EINVAL;
    Matrix& A = h->A;
    CUDA_TRY(cudaMemcpyAsync(A.pool, htiles, size_t(A.ntiles_loc) * A.tile_elems() * sizeof(double),
                             cudaMemcpyHostToDevice, cudaStream_t(stream)));
    return SB200_OK;
}

int sb200_matrix_to_host_local_d(sb200_matrix_t h, double* htiles, sb200_stream_t stream)
{
    if (! h || ! htiles) return SB200_EINVAL;
    Matrix& A = h->A;
    CUDA_TRY(cudaMemcpyAsync(htiles, A.pool, size_t(A.ntiles_loc) * A.tile_elems() * sizeof(double),
                             cudaMemcpyDeviceToHost, cudaStream_t(stream)));
    return SB200_OK;
}

int sb200_matrix_copy_d(sb200_matrix_t dst, sb200_matrix_t src, sb200_stream_t stream)
{
    if (! dst || ! src) return SB200_EINVAL;
    Matrix& D = dst->A; Matrix& S = src->A;
    if (D.g != S.g || D.kind != S.kind || D.m != S.m || D.n != S.n || D.nb != S.nb) return SB200_EINVAL;
    CUDA_TRY(cudaMemcpyAsync(D.pool, S.pool, size_t(S.ntiles_loc) * S.tile_elems() * sizeof(double),
                             cudaMemcpyDeviceToDevice, cudaStream_t(stream)));
    return SB200_OK;
}

int sb200_potrf_d(sb200_matrix_t h, const sb200_options_t* opts, int64_t* info)
{
    (void) opts;                       // lookahead is fixed at 1 (the reference default)
    if (! h) return SB200_EINVAL;
    return potrf_driver(h->A, info);
}

int sb200_gemm_d(double alpha, sb200_matrix_t A, sb200_matrix_t B, double beta, sb200_matrix_t C,
                 const sb200_options_t* opts)
{
    (void) opts;
    if (! A || ! B || ! C) return SB200_EINVAL;
    return gemm_driver(alpha, A->A, B->A, beta, C->A);
}

} // extern "C"
