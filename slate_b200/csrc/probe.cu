// probe.cu -- probe-vector products with a distributed tile matrix: the residual checks of the reference tester
// without forming a second n x n matrix.
//
// The reference tester checks a factorisation by solving with it and forming ||B - A X|| with slate::gemm / hemm
// (test/test_posv.cc:336-342, test/test_gesv.cc:371-377).  At the bench sizes (n = 65536 on a p x q grid) the check
// here uses the identities  ||A x - L (L^H x)|| / (n ||A|| ||x||)  and  ||P A x - L (U x)|| / (n ||A|| ||x||)
// for a seeded probe vector x: three tile-matrix x vector products, each O(n^2) HBM-bound work on the tiles a rank
// stores.  Every rank accumulates its tiles' contributions into a replicated vector y (atomic adds; the caller zeroes
// y and sums it over the ranks).
//
//   part 'G': the whole tile matrix            'L': lower triangle (diag 'U': unit diagonal)      'U': upper triangle
//   part 'H': Hermitian matrix from its stored lower tiles (the strictly lower part also acts transposed-conjugated,
//             the diagonal is taken real; reference semantics of HermitianMatrix, include/slate/HermitianMatrix.hh)
//   op 'N': y += part(A) x,   op 'C': y += part(A)^H x          use_abs: |a_ij| instead of a_ij (x = ones gives the
//             row sums / column sums of |A|: the inf- and one-norm pieces)
#include "runtime_internal.hh"

namespace sb200 {

struct ProbeTile { const void* ptr; int64_t i0, j0; int mb, nbc; };

__device__ __forceinline__ void atomic_add_t(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_t(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_t(cuFloatComplex* p, cuFloatComplex v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }
__device__ __forceinline__ void atomic_add_t(cuDoubleComplex* p, cuDoubleComplex v) { atomicAdd(&p->x, v.x); atomicAdd(&p->y, v.y); }

__device__ __forceinline__ float  abs_t(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_t(double a) { return fabs(a); }
__device__ __forceinline__ cuFloatComplex  abs_t(cuFloatComplex a) { return make_cuFloatComplex(hypotf(a.x, a.y), 0.f); }
__device__ __forceinline__ cuDoubleComplex abs_t(cuDoubleComplex a) { return make_cuDoubleComplex(hypot(a.x, a.y), 0.0); }

// value of element (gi, gj) of part(A) given the stored value a; `keep` = false: the element is outside the part
template <typename T>
__device__ __forceinline__ T part_value(int part, int unit, int use_abs, int64_t gi, int64_t gj, T a, bool& keep)
{
    keep = true;
    if (part == 'L') {
        if (gi < gj) keep = false;
        else if (gi == gj && unit) a = from_real<T>(1);
    }
    else if (part == 'U') { if (gi > gj) keep = false; }
    else if (part == 'H') {
        if (gi < gj) keep = false;
        else if (gi == gj) a = real_part_only(a);
    }
    return use_abs ? abs_t(a) : a;
}

constexpr int PROBE_THREADS = 256;

// one CTA per stored local tile
template <typename T>
__global__ void __launch_bounds__(PROBE_THREADS)
probe_mv_kernel(const ProbeTile* __restrict__ tiles, int ld, int part, int op, int unit, int use_abs,
                const T* __restrict__ x, T* __restrict__ y)
{
    const ProbeTile t = tiles[blockIdx.x];
    const T* __restrict__ A = static_cast<const T*>(t.ptr);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool rows_pass = (op == 'N') || part == 'H';       // y[row] += sum_c v(r,c) x[col]
    const bool cols_pass = (op != 'N') || part == 'H';       // y[col] += sum_r conj(v(r,c)) x[row]
    if (rows_pass) {
        for (int r = tid; r < t.mb; r += PROBE_THREADS) {
            T acc = zero_of<T>();
            const int64_t gi = t.i0 + r;
            for (int c = 0; c < t.nbc; ++c) {
                bool keep;
                const T v = part_value<T>(part, unit, use_abs, gi, t.j0 + c, A[r + int64_t(c) * ld], keep);
                if (keep) fma_acc(acc, v, x[t.j0 + c]);
            }
            atomic_add_t(&y[gi], acc);
        }
    }
    if (cols_pass) {
        for (int c = warp; c < t.nbc; c += PROBE_THREADS / 32) {
            T acc = zero_of<T>();
            const int64_t gj = t.j0 + c;
            for (int r = lane; r < t.mb; r += 32) {
                const int64_t gi = t.i0 + r;
                bool keep;
                T v = part_value<T>(part, unit, use_abs, gi, gj, A[r + int64_t(c) * ld], keep);
                if (part == 'H' && gi == gj) keep = false;        // the diagonal was applied by the rows pass
                if (keep) fma_acc(acc, use_abs ? v : conj_(v), x[gi]);
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc = add(acc, shfl_xor_t(acc, o));
            if (lane == 0) atomic_add_t(&y[gj], acc);
        }
    }
}

template <typename T>
static int probe_mv_t(Matrix& A, int part, int op, int unit, int use_abs, const void* x, void* y, cudaStream_t s)
{
    std::vector<ProbeTile> td;
    for (int64_t j = A.g->pcol; j < A.nt; j += A.g->q)
        for (int64_t i = A.g->prow; i < A.mt; i += A.g->p) {
            if (! A.stored(i, j)) continue;
            if (part == 'L' && i < j) continue;
            if (part == 'U' && i > j) continue;
            td.push_back({A.tile_as<T>(i, j), i * A.nb, j * A.nb, int(A.tile_mb(i)), int(A.tile_nb(j))});
        }
    if (td.empty()) return SB200_OK;
    DevBuf d;
    SB_TRY(d.alloc(td.size() * sizeof(ProbeTile)));
    CUDA_TRY(cudaMemcpyAsync(d.p, td.data(), td.size() * sizeof(ProbeTile), cudaMemcpyHostToDevice, s));
    probe_mv_kernel<T><<<unsigned(td.size()), PROBE_THREADS, 0, s>>>(d.as<ProbeTile>(), int(A.nb), part, op, unit, use_abs,
                                                                      static_cast<const T*>(x), static_cast<T*>(y));
    const int st = launch_status();
    CUDA_TRY(cudaStreamSynchronize(s));          // td / d go out of scope
    return st;
}

} // namespace sb200

using namespace sb200;

extern "C" int sb200_matrix_probe_mv(sb200_matrix_t h, int part, int op, int diag, int use_abs,
                                     const void* x, void* y, sb200_stream_t stream)
{
    if (! h || ! x || ! y) return SB200_EINVAL;
    Matrix& A = h->A;
    if (part != 'G' && part != 'L' && part != 'U' && part != 'H') return SB200_EINVAL;
    if (op != 'N' && op != 'C') return SB200_EINVAL;
    if (diag != 'N' && diag != 'U') return SB200_EINVAL;
    if (part == 'H' && (A.kind != 'H' || A.m != A.n)) return SB200_EINVAL;
    if (A.kind == 'H' && part == 'U') return SB200_EINVAL;          // only the lower tiles exist
    if (A.kind == 'H' && part == 'G') return SB200_EINVAL;
    const int unit = diag == 'U';
    cudaStream_t s = cudaStream_t(stream);
    switch (A.dtype) {
        case 's': return probe_mv_t<float>(A, part, op, unit, use_abs, x, y, s);
        case 'd': return probe_mv_t<double>(A, part, op, unit, use_abs, x, y, s);
        case 'c': return probe_mv_t<cuFloatComplex>(A, part, op, unit, use_abs, x, y, s);
        case 'z': return probe_mv_t<cuDoubleComplex>(A, part, op, unit, use_abs, x, y, s);
    }
    return SB200_EINVAL;
}
