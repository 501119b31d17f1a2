// runtime_internal.hh -- helpers shared by the drivers (runtime.cu, solve.cu, mixed.cu, getrf*.cu):
// status macros, per-step pointer batches ("device regions"), the once-per-call plan buffer,
// the two-stream/event scaffold, and the type-generic tile factor/solve entry points.
#pragma once
#include <cuda_profiler_api.h>
#include "runtime.hh"
#include "gemm_dmma.cuh"
#include "scalar_ops.cuh"
#include "tc05.hh"
#include <vector>
#include <cstdio>

namespace sb200 {

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return int(e_); } while (0)
#define NCCL_TRY(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) { \
    fprintf(stderr, "slate_b200: NCCL error %s at %s:%d\n", ncclGetErrorString(r_), __FILE__, __LINE__); \
    return SB200_ENCCL; } } while (0)
#define SB_TRY(x) do { int s_ = (x); if (s_ != SB200_OK) return s_; } while (0)

// from factor_small.cu (explicitly instantiated there for float, double, cuFloatComplex, cuDoubleComplex)
template <typename T>
int trsm_colmajor(bool left, bool lower, int op, bool unit, int m, int n, T alpha,
                  const T* Tm, int ldt, T* const* dB, int64_t offB, int ldb, int batch,
                  T* W, cudaStream_t stream);
template <typename T>
int potrf_tile_lower(int n, T* A, int lda, int* dinfo, int info_base, T* W, cudaStream_t stream, int fused_dflt = -1);
// small-nrhs solve path (factor_small.cu): inverted diagonal blocks of all diagonal tiles in one launch, and the
// one-CTA-per-8-columns tile solve; trsm_small returns SB200_ENOTSUP for cases it does not serve
template <typename T>
int trtri_diag_all(int ntiles, const T* const* Tarr, int ldt, int na, int na_last, bool lower, bool unit, T* W, cudaStream_t stream);
template <typename T>
int trsm_small(bool lower, int op, int na, int n, const T* Tm, int ldt, const T* Winv, T* const* dB, int64_t offB,
               int ldb, int batch, cudaStream_t stream);
constexpr int FACTOR_IB = 64;          // diagonal block of the tile factor / solve kernels
// opt-in one-launch tile Cholesky (potrf_tile_fused.cu; SB200_TILE_FUSED=1: divided form, 2: rsqrt form);
// returns FUSED_NOT_TAKEN when it does not apply
constexpr int FUSED_NOT_TAKEN = -1000001;
int potrf_tile_fused_d(int n, double* A, int lda, int* dinfo, int info_base, int variant, cudaStream_t stream);
int potrf_tile_fused_s(int n, float* A, int lda, int* dinfo, int info_base, int variant, cudaStream_t stream);
// skinny right operand (n <= 16): HBM-bound streaming kernel instead of a tensor-core tile kernel (gemm_skinny.cu)
template <typename T> bool gemm_skinny_applies(int opB, const GemmParamsT<T>& p);
template <typename T> int launch_gemm_skinny(int opA, const GemmParamsT<T>& p, cudaStream_t stream);

template <typename T> struct IsComplex { static constexpr bool value = false; };
template <> struct IsComplex<cuFloatComplex>  { static constexpr bool value = true; };
template <> struct IsComplex<cuDoubleComplex> { static constexpr bool value = true; };

// ------------------------------------------------------------------------------------------
// batches: tiles of one step grouped by (m, n, k, tri) -- the reference's "regions"
// (src/internal/internal_batch.hh:169-347), built once per driver call for all steps.
// ------------------------------------------------------------------------------------------
struct Batch {
    int m, n, k, tri;
    std::vector<const void*> A, B;
    std::vector<void*> C;
    size_t off = 0;                 // offset (in pointers) of A | B | C blocks in the device plan
};

inline void batch_add(std::vector<Batch>& v, int m, int n, int k, int tri,
                      const void* A, const void* B, void* C)
{
    for (auto& b : v)
        if (b.m == m && b.n == n && b.k == k && b.tri == tri) {
            b.A.push_back(A); b.B.push_back(B); b.C.push_back(C);
            return;
        }
    Batch b{m, n, k, tri, {A}, {B}, {C}, 0};
    v.push_back(std::move(b));
}

// Device workspaces of the drivers (panel rings, plan buffers, scratch) come from a grow-only per-device cache: a
// driver call at n = 65536 spent 40-130 ms of host time in cudaMalloc / cudaFree of buffers it allocates again,
// identically, at the next call (profiles/r02n_host_times_potrf_n65536.txt).  Every driver synchronises its streams
// before it returns, so a block handed back is idle.  sb200_release_workspaces() frees the cache.  (sm_partition.cu)
void* ws_cache_get(size_t bytes);
void ws_cache_put(void* p);

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
    template <typename P> size_t push(const std::vector<P*>& v)
    {
        size_t o = host.size();
        host.insert(host.end(), v.begin(), v.end());
        return o;
    }
    int upload(cudaStream_t s)
    {
        if (host.empty()) return SB200_OK;
        dev = static_cast<void**>(ws_cache_get(host.size() * sizeof(void*)));
        if (! dev) return SB200_ENOMEM;
        CUDA_TRY(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(void*), cudaMemcpyHostToDevice, s));
        return SB200_OK;
    }
    template <typename T> T* const* at(size_t off) const { return reinterpret_cast<T* const*>(dev + off); }
    ~PlanBuffer() { if (dev) ws_cache_put(dev); }
};

// one batched launch per shape class; herk != 0 forces the diagonal of triangle-masked complex tiles real
template <typename T>
inline int launch_batches(const std::vector<Batch>& bs, const PlanBuffer& pb, int opA, int opB_wide,
                          T alpha, T beta, int ld, int herk, cudaStream_t s, int opB_skinny = 0)
{
    for (const auto& b : bs) {
        // batches with a skinny right operand (n <= SKINNY_MAX_N) may carry their B in another orientation (the
        // transposed-B-panel path keeps them untransposed so that they stay on the HBM-bound streaming kernel)
        const int opB = (opB_skinny && b.n <= SKINNY_MAX_N) ? opB_skinny : opB_wide;
        GemmParamsT<T> p{};
        const size_t cnt = b.C.size();
        p.A = reinterpret_cast<const T* const*>(pb.dev + b.off);
        p.B = reinterpret_cast<const T* const*>(pb.dev + b.off + cnt);
        p.C = reinterpret_cast<T* const*>(pb.dev + b.off + 2 * cnt);
        p.m = b.m; p.n = b.n; p.k = b.k; p.lda = ld; p.ldb = ld; p.ldc = ld;
        p.alpha = alpha; p.beta = beta; p.batch = int(cnt); p.tri = b.tri;
        p.herk = (herk && b.tri) ? 1 : 0;
        if (gemm_skinny_applies<T>(opB, p)) SB_TRY(launch_gemm_skinny<T>(opA, p, s));
        else                                SB_TRY(launch_gemm<T>(opA, opB, p, s));
    }
    return SB200_OK;
}

// same batches on the tcgen05 FP32-emulated kernel: A / B entries point at PACKED operands
inline int launch_batches_tc05(const std::vector<Batch>& bs, const PlanBuffer& pb,
                               float alpha, float beta, int ld, cudaStream_t s)
{
    for (const auto& b : bs) {
        Tc05Params p{};
        const size_t cnt = b.C.size();
        p.Ap = reinterpret_cast<const void* const*>(pb.dev + b.off);
        p.Bp = reinterpret_cast<const void* const*>(pb.dev + b.off + cnt);
        p.C = reinterpret_cast<float* const*>(pb.dev + b.off + 2 * cnt);
        p.m = b.m; p.n = b.n; p.k = b.k; p.ldc = ld;
        p.alpha = alpha; p.beta = beta; p.batch = int(cnt); p.tri = b.tri;
        SB_TRY(launch_tc05_gemm(p, s));
    }
    return SB200_OK;
}

inline double batches_flops(const std::vector<Batch>& bs, bool complex_)
{
    double f = 0;
    for (const auto& b : bs) {
        // ALGORITHMIC flops: a triangle-masked (herk/syrk diagonal) tile counts n(n+1)k
        // (blaspp/include/blas/flops.hh syrk), whatever the kernel computes above the diagonal
        const double per = b.tri ? double(b.n) * (b.n + 1.0) * b.k : 2.0 * b.m * b.n * b.k;
        f += per * double(b.C.size());
    }
    return complex_ ? 4.0 * f : f;      // complex: 6 mul-flops + 2 add-flops per multiply-add (flops.hh:100-104)
}

struct Streams {
    cudaStream_t panel = nullptr, look = nullptr, trail = nullptr;     // chain (highest priority) | lookahead columns | trailing update
    cudaStream_t chain = nullptr;          // == panel, or a stream on its own SM partition (init with chain_sms > 0)
    bool own_chain = false;
    int chain_sm_count = 0;
    cudaEvent_t hop_ev[2] = {nullptr, nullptr};
    int hop_next = 0;
    std::vector<cudaEvent_t> ev;
    std::vector<cudaEvent_t> tev;          // timing event pairs around the trailing-update launches
    std::vector<cudaEvent_t> pev;          // timing event pairs around the panel work of every step
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    static double sum_pairs(const std::vector<cudaEvent_t>& v)
    {
        double tot = 0;
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, v[i], v[i + 1]) == cudaSuccess) tot += ms;
        }
        return tot;
    }
    static int stamp(std::vector<cudaEvent_t>& v, cudaStream_t s)
    {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        v.push_back(e);
        CUDA_TRY(cudaEventRecord(e, s));
        return SB200_OK;
    }
    int ptime(cudaStream_t s) { return stamp(pev, s); }
    // SB200_NCU_TIMED=i: the timed (trailing-update) launch groups i .. i+3 of a driver call sit inside a
    // cudaProfilerStart/Stop window, so that `ncu --profile-from-start off` captures exactly the kernels the
    // roofline is quoted on and not the small launches of the panel chain that share their name.
    int ncu_pick = -2, ncu_seen = 0;
    int time_begin(cudaStream_t s)
    {
        if (ncu_pick == -2) { const char* e = getenv("SB200_NCU_TIMED"); ncu_pick = e ? atoi(e) : -1; }
        if (ncu_pick >= 0 && ncu_seen == ncu_pick) cudaProfilerStart();
        return stamp(tev, s);
    }
    int time_end(cudaStream_t s)
    {
        if (ncu_pick >= 0 && ++ncu_seen == ncu_pick + 4) cudaProfilerStop();
        return stamp(tev, s);
    }
    double panel_ms() { return sum_pairs(pev); }
    double timed_ms() { return sum_pairs(tev); }
    int init(size_t nevents, int chain_sms = 0);          // sm_partition.cu
    int hop(cudaStream_t from, cudaStream_t to);          // event edge from -> to (no-op when they are the same stream)
    ~Streams()
    {
        for (auto e : ev) if (e) cudaEventDestroy(e);
        for (auto e : tev) if (e) cudaEventDestroy(e);
        for (auto e : pev) if (e) cudaEventDestroy(e);
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        for (auto e : hop_ev) if (e) cudaEventDestroy(e);
        if (own_chain && chain) cudaStreamDestroy(chain);
        if (panel) cudaStreamDestroy(panel);
        if (look) cudaStreamDestroy(look);
        if (trail) cudaStreamDestroy(trail);
    }
};

struct DevBuf {
    void* p = nullptr;
    int alloc(size_t bytes) { p = ws_cache_get(bytes ? bytes : 16); return p ? SB200_OK : SB200_ENOMEM; }
    template <typename T> T* as() const { return static_cast<T*>(p); }
    ~DevBuf() { if (p) ws_cache_put(p); }
};

// grouped root -> everybody broadcasts of contiguous ranges (runtime.cu); src is read on the root only
struct BcastItem { const void* src; void* dst; size_t bytes; int root; };
constexpr int SB200_BCAST_DEFAULT = 1;          // scatter + all-gather: measured on 8 GPUs (r2g8b): dpotrf 428 -> 408 ms, dgetrf 1043 -> 1008 ms
constexpr int SB200_NCCL_MAX_CTAS_DEFAULT = 8;
int bcast_many(Grid& g, const std::vector<BcastItem>& items, cudaStream_t s, ncclComm_t comm = nullptr);

// drivers (runtime.cu)
// lookahead <= 0: the library default (SB200_LOOKAHEAD or POTRF_DEFAULT_LOOKAHEAD)
constexpr int POTRF_DEFAULT_LOOKAHEAD = 2, MAX_LOOKAHEAD = 8;
template <typename T> int potrf_driver(Matrix& A, int64_t* info_out, bool use_tc05, void* host_out = nullptr,
                                        const void* host_in = nullptr, int lookahead = 0);
template <typename T> int gemm_driver(T alpha, Matrix& A, Matrix& B, T beta, Matrix& C);
template <typename T> int herk_driver(typename RealOf<T>::type alpha, Matrix& A, typename RealOf<T>::type beta, Matrix& C);
int matrix_alloc(Grid& g, int dtype, int kind, int64_t m, int64_t n, int64_t nb, Matrix& A);
// solve path on a p x q grid with replicated right-hand sides (solve_dist.cu; float / double)
template <typename T> int potrs_dist(Matrix& A, Matrix& B, cudaStream_t s);
template <typename T> int getrs_dist(Matrix& A, const int64_t* pivots, Matrix& B, cudaStream_t s);
int solve_mixed_dist_d(bool hermitian, Matrix& A, int64_t* pivots_out, Matrix& B, Matrix& Xm, int64_t itermax, double tol,
                       bool use_fallback, int* iter_out, int64_t* info_out, double* timers_ms);

} // namespace sb200

// Driver options (include/slate_b200.h sb200_options_t; reference: slate::Option::{Lookahead, InnerBlocking, PivotThreshold},
// src/potrf.cc:41-42, src/getrf.cc:38-43).  potrf honours Lookahead 1 .. MAX_LOOKAHEAD (depth of its lookahead stream); the
// other pipelines have a fixed depth of 1 and the LU panel always takes the largest candidate (threshold 1.0):
// anything else is REJECTED, never silently ignored.
// InnerBlocking only re-associates the panel's rank-ib updates in the reference (no effect on the pivot rule); the GPU
// panel has its own fixed blocking, so any positive value is accepted as the hint it is.
inline int options_status(const sb200_options_t* o, int max_lookahead = 1)
{
    if (! o) return SB200_OK;
    if (o->lookahead < 0 || o->inner_blocking < 1 || ! (o->pivot_threshold >= 0.0 && o->pivot_threshold <= 1.0)) return SB200_EINVAL;
    // lookahead 0 = the library's own depth (the result does not depend on the depth; only the overlap does)
    if (o->lookahead > max_lookahead || o->pivot_threshold != 1.0) return SB200_ENOTSUP;
    if (max_lookahead == 1 && o->lookahead > 1) return SB200_ENOTSUP;
    return SB200_OK;
}

struct sb200_grid_s   { sb200::Grid g; };
struct sb200_matrix_s { sb200::Matrix A; };
