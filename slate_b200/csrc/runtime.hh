// runtime.hh -- host runtime of libslate_b200: process grid, HBM-resident tile matrix,
// and the factorisation / multiply drivers that schedule the trailing-matrix update.
//
// What it mirrors in the reference (and how it differs, B200-first):
//   * slate::Matrix / HermitianMatrix (include/slate/Matrix.hh, BaseMatrix.hh): 2-D block-cyclic
//     tile map, tileRank(i,j) = (i % p) + (j % q) * p  (GridOrder::Col, include/slate/func.hh:96-104).
//     Here one process owns ONE B200 and all of its tiles live in one HBM pool for the whole
//     life of the matrix -- there is no host<->device MOSI traffic inside a driver; the MOSI state
//     machine collapses to "owner copy (Modified) + read-only panel workspace copies (Shared)"
//     that are overwritten two steps later.
//   * device_regions_build (src/internal/internal_batch.hh:227-347): pointer batches are built
//     ONCE per driver call on the host for every step (a "plan"), uploaded in one copy and reused;
//     the reference rebuilds and re-uploads them at every step.
//   * tileBcast / listBcast / listBcastMT (include/slate/BaseMatrix.hh:1889-2140): the panel
//     broadcast is p NCCL broadcasts of contiguous pool ranges per step on the panel stream.
//   * the OpenMP task DAG with lookahead (src/potrf.cc:84-195, src/getrf.cc:84-236,
//     src/gemmC.cc:89-196) becomes two CUDA streams (panel = high priority, trailing) ordered
//     by events; the host never blocks inside the step loop.
#pragma once
#include <ctime>
#include <utility>
#include "common.cuh"
#include <nccl.h>
#include <vector>
#include <string>
#include <cstdio>
#include <cstdlib>

namespace sb200 {

struct Grid {
    int p = 1, q = 1, rank = 0;
    int prow = 0, pcol = 0;
    ncclComm_t world = nullptr;      // null when p*q == 1
    ncclComm_t world_lo = nullptr;   // the same ranks, capped at a few CTAs: the Cholesky panel broadcast (may be null)
    ncclComm_t col_comm_lo = nullptr;   // col_comm capped the same way: the L_kk broadcast of the Cholesky chain
    ncclComm_t row_comm = nullptr;   // ranks with the same prow (size q), rank order = pcol
    ncclComm_t col_comm = nullptr;   // ranks with the same pcol (size p), rank order = prow
    ncclComm_t col_comm2 = nullptr;  // second communicator over the same ranks: collectives of the trailing stream
    int size() const { return p * q; }
    int rank_of(int64_t i, int64_t j) const { return int(i % p) + int(j % q) * p; }
};

struct Matrix {
    Grid*   g = nullptr;
    int     kind = 'G';            // 'G' general, 'H' Hermitian/symmetric, lower tiles stored
    int     layout = 'C';          // tile layout: 'C' column-major (all drivers here)
    int64_t m = 0, n = 0, nb = 0, mt = 0, nt = 0;
    int64_t mt_loc = 0, nt_loc = 0;          // local tile rows / cols
    std::vector<int64_t> col_start;          // tile index of the first stored tile of local col jl
    int64_t ntiles_loc = 0;
    int     dtype = 'd';                     // 's' float, 'd' double, 'c' complex<float>, 'z' complex<double>
    int     esize = 8;                       // bytes per element
    double* pool = nullptr;                  // ntiles_loc * nb*nb elements of `dtype` (typed double* for the FP64 drivers), every tile ld = nb
    double  last_ms = 0.0;                   // device time of the last driver call (CUDA events)
    double  last_trail_ms = 0.0;             // summed duration of its trailing-update GEMM launches
    double  last_trail_flops = 0.0;          // algorithmic flops of those launches
    int64_t last_trail_launches = 0;
    double  last_panel_ms = 0.0;             // summed duration of the panel-stream critical work (factor + solve + broadcast)

    int64_t tile_elems() const { return nb * nb; }
    int64_t tile_mb(int64_t i) const { return i == mt - 1 ? m - i * nb : nb; }
    int64_t tile_nb(int64_t j) const { return j == nt - 1 ? n - j * nb : nb; }
    bool    stored(int64_t i, int64_t j) const { return kind == 'G' || i >= j; }
    bool    is_local(int64_t i, int64_t j) const { return g->rank_of(i, j) == g->rank; }
    // first local tile row index >= i0 for process row prow
    int64_t first_local_row(int64_t i0) const
    {
        int64_t il = (i0 - g->prow + g->p - 1) / g->p;
        return il < 0 ? 0 : il;
    }
    // pool slot of local tile (i, j); caller guarantees is_local && stored
    int64_t tile_index(int64_t i, int64_t j) const
    {
        const int64_t il = (i - g->prow) / g->p, jl = (j - g->pcol) / g->q;
        if (kind == 'G') return col_start[jl] + il;
        return col_start[jl] + (il - first_local_row(j));
    }
    // device pointer of local tile (i, j) of an FP64 matrix
    double* tile(int64_t i, int64_t j) const { return pool + tile_index(i, j) * tile_elems(); }
    // ... of a matrix of element type T (sizeof(T) == esize)
    template <typename T> T* tile_as(int64_t i, int64_t j) const
    {
        return reinterpret_cast<T*>(pool) + tile_index(i, j) * tile_elems();
    }
    size_t pool_bytes() const { return size_t(ntiles_loc) * size_t(tile_elems()) * size_t(esize); }
};

// element type <-> ABI type character
template <typename T> struct TypeChar;
template <> struct TypeChar<float>           { static constexpr int value = 's'; };
template <> struct TypeChar<double>          { static constexpr int value = 'd'; };
template <> struct TypeChar<cuFloatComplex>  { static constexpr int value = 'c'; };
template <> struct TypeChar<cuDoubleComplex> { static constexpr int value = 'z'; };

// Optional per-phase device timing of a driver call (SB200_PHASES=1): event pairs around named
// phases of the panel / trailing streams, summed per phase and printed to stderr as one JSON line
// when the driver returns.  Used to find what the lookahead has to hide; off by default.
// SB200_HOST_TIMES=1: host wall clock between marks of a driver call, one JSON line on stderr when the object dies
// (i.e. after the driver's local buffers, streams and events have been released)
struct HostTimes {
    const char* what; bool on; double t0, last;
    std::vector<std::pair<const char*, double>> marks;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit HostTimes(const char* w) : what(w) { const char* e = getenv("SB200_HOST_TIMES"); on = e && atoi(e) != 0; t0 = last = on ? now() : 0; }
    void mark(const char* name) { if (! on) return; const double t = now(); marks.push_back({name, t - last}); last = t; }
    ~HostTimes()
    {
        if (! on) return;
        const double t = now();
        fprintf(stderr, "{\"sb200_host_ms\": \"%s\"", what);
        for (auto& m : marks) fprintf(stderr, ", \"%s\": %.3f", m.first, m.second);
        fprintf(stderr, ", \"teardown\": %.3f, \"total\": %.3f}\n", t - last, t - t0);
    }
};

struct PhaseTimer {
    struct Rec { const char* name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    bool on = false;
    PhaseTimer() { const char* e = getenv("SB200_PHASES"); on = e && atoi(e) != 0; }
    void begin(const char* name, cudaStream_t s)
    {
        if (! on) return;
        Rec r{name, nullptr, nullptr};
        cudaEventCreate(&r.a); cudaEventCreate(&r.b);
        cudaEventRecord(r.a, s);
        recs.push_back(r);
    }
    void end(cudaStream_t s) { if (on && ! recs.empty()) cudaEventRecord(recs.back().b, s); }
    void report(const char* what, int rank)
    {
        if (! on) return;
        cudaDeviceSynchronize();
        std::vector<std::pair<std::string, std::pair<double, int>>> sums;
        for (auto& r : recs) {
            float ms = 0;
            cudaEventElapsedTime(&ms, r.a, r.b);
            bool found = false;
            for (auto& x : sums) if (x.first == r.name) { x.second.first += ms; x.second.second++; found = true; }
            if (! found) sums.push_back({r.name, {ms, 1}});
            cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        }
        recs.clear();
        fprintf(stderr, "{\"sb200_phases\": \"%s\", \"rank\": %d", what, rank);
        for (auto& x : sums) fprintf(stderr, ", \"%s\": [%.3f, %d]", x.first.c_str(), x.second.first, x.second.second);
        fprintf(stderr, "}\n");
    }
};

} // namespace sb200
