// getrf_internal.hh -- pieces of the LU path shared by getrf.cu (single rank) and getrf_dist.cu (p x q grid)
#pragma once
#include "runtime.hh"

namespace sb200 {

constexpr int PW = 32;            // panel base-block width (the tester's ib 32)
constexpr int PROWS_MAX = 768;    // rows of the block one CTA keeps in shared memory
constexpr int PTHREADS = 256;
// complex<double> blocks keep half as many rows per CTA (16-byte elements: 32 x 385 x 16 B = 197 KB of shared memory)
template <typename T> constexpr int panel_rows_max() { return sizeof(T) > 8 ? PROWS_MAX / 2 : PROWS_MAX; }
constexpr int V3_NWIDE = 4;       // interchange CTAs of the one-round base kernel (getrf_base_v3.cu)

// scratch of the cooperative panel kernel (per driver call)
struct PanelScratch {
    double* gval = nullptr; int* grow = nullptr; double* gcand = nullptr; double* gdiag = nullptr;
    double* W = nullptr;            // trsm workspace of the panel stream
    unsigned* bar = nullptr;
    bool nopiv = false;                 // getrf_nopiv: base kernel without the pivot search
    int tnt_ranks = 0;                  // getrf_tntpiv: > 0 = participants of the tournament per panel (getrf_tnt.cu)
    unsigned long long* v3_buf = nullptr;   // per-column exchange records of getrf_base_v3_kernel (getrf_base_v3.cu)
    unsigned v3_gen = 0;
    bool use_v3 = false;
    int max_ctas = 0;
    void* raw = nullptr;
    int init();
    ~PanelScratch();
};

// arguments of the cooperative base-block kernels (getrf.cu: float / double; getrf_cplx.cu: complex)
template <typename T>
struct BaseArgs {
    T* const* tiles;
    int nb, m_p, c0, w, rows_per;
    int64_t* piv_tile; int64_t* piv_off;
    T* gval; int* grow; T* gcand; T* gdiag;     // [2][G], [2][G], [2][G][PW], [2][PW]
    int* info; int info_base;
    int* rowmap;       // optional: rowmap[x] = panel row whose ORIGINAL content now sits at position x
    int kw_wide;       // > 0: the LAST CTA of the grid owns no rows and applies every interchange of this
                       // block to the panel columns outside [c0, c0+w) (all kw_wide columns of the panel)
                       // while the other CTAs go on factoring -- no laswp launches between blocks
    unsigned* bar;     // unused (kept: ptxas's register allocation for this kernel depends on the size of the struct)
};
// (BaseArgs is left exactly as validated: ptxas's register allocation for getrf_base_kernel changes with the size of
// its parameter struct -- 64 registers as measured in round 1, 40 + a spill with two more fields.)


// complex base block (getrf_cplx.cu)
template <typename T> int launch_base_cplx(BaseArgs<T>& a, int grid, size_t smem, bool nopiv, cudaStream_t s);

// scratch of the tournament panel (getrf_tnt.cu), per driver call
struct TntScratch {
    void* wcopy = nullptr;          // workspace copy of the panel, mt tiles
    void* tmp = nullptr;            // the two stacked candidate tiles of a tree node
    int* ids = nullptr;             // [ranks][nb] original panel rows of every participant's candidates
    int* rm_sub = nullptr;          // row map of the node being factored
    int* row_at = nullptr; int* pos_of = nullptr;      // winners -> sequential interchanges
    int64_t* spiv = nullptr;        // [2][nb] pivots of the nodes (not the panel's)
    void** ptrs = nullptr;          // [ranks][per_class] tile pointers of the workspace copy by residue class, then the pair
    int* dummy_info = nullptr;      // info of the nodes the reference ignores (every rank but the first)
    int ranks = 1, per_class = 0;
    int64_t te = 0;
    void* raw = nullptr;
    int init(int64_t mt, int64_t nb, int64_t m, int esize, int ranks, cudaStream_t s);
    ~TntScratch();
};
// tournament panel: see getrf_tnt.cu.  htiles = the panel's tile pointers as the host knows them (stack is the same
// list on the device); k = block column (global tile row of the panel's first tile)
template <typename T>
int getrf_panel_tnt(T* const* stack, const std::vector<T*>& htiles, int64_t k, int nb, int m_p, int kw,
                    int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                    PanelScratch& ps, TntScratch& ts, cudaStream_t s, int* rowmap, PhaseTimer* ph);
bool tnt_shape_supported(const Matrix& A);
// participants per panel of a getrf_tntpiv call on grid g: its process rows, or SB200_TNT_RANKS (test hook: the
// tournament of a p-row grid on fewer ranks -- the tree never leaves the GPU that holds the gathered panel)
int tnt_ranks_for(const Grid& g);

// Factor the panel given as a stack of `ntile` tiles (device pointer array `stack`, nb x nb, ld = nb;
// last tile has m_p - (ntile-1)*nb rows), kw columns.  rowmap (optional, m_p ints, pre-set to the
// identity): on return rowmap[x] = panel row whose ORIGINAL content sits at position x.
int getrf_panel_d(double* const* stack, double* tile0, int ntile, int nb, int m_p, int kw,
                  int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                  PanelScratch& ps, cudaStream_t s, int* rowmap = nullptr, PhaseTimer* ph = nullptr);

int trsm_colmajor_d(bool left, bool lower, int op, bool unit, int m, int n, double alpha,
                    const double* T, int ldt, double* const* dB, int64_t offB, int ldb, int batch,
                    double* W, cudaStream_t stream);

int getrf_driver_dist(Matrix& A, int64_t* pivots_out, int64_t* info_out);
int getrf_driver_dist_s(Matrix& A, int64_t* pivots_out, int64_t* info_out, bool use_tc05);     // FP32, p x q grid
// type-generic panel (getrf.cu; float and double)
template <typename T>
int getrf_panel(T* const* stack, T* tile0, int ntile, int nb, int m_p, int kw,
                int64_t* piv_tile, int64_t* piv_off, int* dinfo, int info_base,
                PanelScratch& ps, cudaStream_t s, int* rowmap, PhaseTimer* ph);
int getrf_driver(Matrix& A, int64_t* pivots_out, int64_t* info_out);                      // FP64, any grid
int getrf_driver_s(Matrix& A, int64_t* pivots_out, int64_t* info_out, bool use_tc05);     // FP32, 1 x 1 grid
int getrf_driver_cplx(Matrix& A, int64_t* pivots_out, int64_t* info_out);                 // complex<float> / complex<double>, 1 x 1 grid

// one-round base block (getrf_base_v3.cu)
size_t base_v3_scratch_bytes(int max_ctas);
int base_v3_init();
template <typename T>
int launch_base_v3(T* const* stack, int nb, int m_p, int c0, int w, int kw, int64_t* piv_tile, int64_t* piv_off,
                   int* dinfo, int info_base, int* rowmap, PanelScratch& ps, cudaStream_t s, int upd_c0 = -1);
bool base_v3_can_fuse(const PanelScratch& ps, int m_p, int c0, int w1, int w);

// fused row interchanges of one panel over a block-column range (getrf.cu)
template <typename T>
int launch_laswp(T* const* tiles, int64_t ldt, int mb, int nb, int ld, int colmajor,
                 const int64_t* piv_tile, const int64_t* piv_off, int j0, int j1, int forward,
                 int64_t col_lo, int64_t col_hi, cudaStream_t s);

} // namespace sb200
