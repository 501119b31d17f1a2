// fake_cudart.cc -- TEST INFRASTRUCTURE (scratch/): a stand-in for libcudart that lets the HOST code of libslate_b200.so
// run in a container without a GPU, so that the drivers written after the GPU budget ended (solve.cu: trmm / hemm / symm
// variants, gemm with transposed views, transposed rank-k updates, the right-side triangular sweep, getrs with an op) can be
// executed for real -- their plan building, pointer arithmetic, operand roles, launch order -- with every KERNEL they launch
// replaced by a plain loop that computes what the kernel is specified to compute.
//
//   g++ -O1 -shared -fPIC -I/usr/local/cuda/include scratch/cpu_standin/fake_cudart.cc -o /tmp/libfakecudart.so
//   LD_PRELOAD=/tmp/libfakecudart.so python scratch/cpu_standin/run_real_host_code_on_cpu.py
//
// "Device" memory is host memory; streams and events are tokens; cudaLaunchKernel looks the kernel up by the name nvcc
// registered for it (__cudaRegisterFunction) and dispatches to an emulation below.  A kernel without an emulation makes the
// launch FAIL loudly (cudaErrorNotSupported, name printed), so nothing is silently skipped.
// What this checks: the C++ host code.  What it cannot check: the kernels themselves -- those are validated on B200s.
// The product never loads this file; nothing under slate_b200/ or tests/ refers to it.
#include <cuda_runtime_api.h>
#include <cxxabi.h>

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

std::map<const void*, std::string> g_kernels;          // host stub -> demangled kernel name
std::mutex g_mu;
thread_local cudaError_t g_last = cudaSuccess;
struct CallCfg { dim3 grid, block; size_t smem; cudaStream_t stream; };
thread_local std::vector<CallCfg> g_cfg;

std::string demangle(const char* n)
{
    int st = 0;
    char* d = abi::__cxa_demangle(n, nullptr, nullptr, &st);
    std::string s = (st == 0 && d) ? d : n;
    free(d);
    return s;
}

// ---- the parameter block of the tile GEMM kernels (slate_b200/csrc/gemm_dmma.cuh: GemmParamsT) ----------------------
struct alignas(8)  c32 { float x, y; };                // cuFloatComplex  = float2  (8-byte aligned)
struct alignas(16) c64 { double x, y; };               // cuDoubleComplex = double2 (16-byte aligned)
template <typename T>
struct GemmParamsT {
    const T* const* A; const T* const* B; T* const* C;
    int64_t offA, offB, offC;
    int64_t strideA, strideB, strideC;
    const T* A0; const T* B0; T* C0;
    int m, n, k;
    int lda, ldb, ldc;
    T alpha, beta;
    int batch;
    int tri;
    int herk;
};

template <typename T> struct Wide { using type = T; };
template <> struct Wide<c32> { using type = std::complex<float>; };
template <> struct Wide<c64> { using type = std::complex<double>; };
template <typename T> typename Wide<T>::type ld(const T& v) { return v; }
inline std::complex<float>  ld(const c32& v) { return {v.x, v.y}; }
inline std::complex<double> ld(const c64& v) { return {v.x, v.y}; }
inline void st(float& d, float v) { d = v; }
inline void st(double& d, double v) { d = v; }
inline void st(c32& d, std::complex<float> v)  { d.x = v.real(); d.y = v.imag(); }
inline void st(c64& d, std::complex<double> v) { d.x = v.real(); d.y = v.imag(); }
inline float  cj(float v) { return v; }
inline double cj(double v) { return v; }
template <typename R> std::complex<R> cj(std::complex<R> v) { return std::conj(v); }
inline float  re_only(float v) { return v; }
inline double re_only(double v) { return v; }
template <typename R> std::complex<R> re_only(std::complex<R> v) { return {v.real(), R(0)}; }

// C_t <- alpha op(A_t) op(B_t) + beta C_t over the batch; operands read completely before C is written (the kernels'
// in-place uses inside trsm_colmajor rely on one CTA owning the aliased block: same observable result)
template <typename T>
void emu_gemm(const GemmParamsT<T>& p, int opA, int opB)
{
    using W = typename Wide<T>::type;
    const W alpha = ld(p.alpha), beta = ld(p.beta);
    const bool use_beta = ! (beta == W(0));
    std::vector<W> a(size_t(p.m) * p.k), b(size_t(p.k) * p.n), c(size_t(p.m) * p.n);
    for (int t = 0; t < p.batch; ++t) {
        const T* A = (p.A ? p.A[t] : p.A0 + int64_t(t) * p.strideA) + p.offA;
        const T* B = (p.B ? p.B[t] : p.B0 + int64_t(t) * p.strideB) + p.offB;
        T* C = (p.C ? p.C[t] : p.C0 + int64_t(t) * p.strideC) + p.offC;
        for (int l = 0; l < p.k; ++l)
            for (int i = 0; i < p.m; ++i) {
                W v = opA == 'N' ? ld(A[i + int64_t(l) * p.lda]) : ld(A[l + int64_t(i) * p.lda]);
                a[size_t(i) + size_t(l) * p.m] = opA == 'C' ? cj(v) : v;
            }
        for (int j = 0; j < p.n; ++j)
            for (int l = 0; l < p.k; ++l) {
                W v = opB == 'N' ? ld(B[l + int64_t(j) * p.ldb]) : ld(B[j + int64_t(l) * p.ldb]);
                b[size_t(l) + size_t(j) * p.k] = opB == 'C' ? cj(v) : v;
            }
        for (int j = 0; j < p.n; ++j)
            for (int i = 0; i < p.m; ++i) {
                W s(0);
                for (int l = 0; l < p.k; ++l) s += a[size_t(i) + size_t(l) * p.m] * b[size_t(l) + size_t(j) * p.k];
                W v = alpha * s;
                if (use_beta) v += beta * ld(C[i + int64_t(j) * p.ldc]);
                c[size_t(i) + size_t(j) * p.m] = v;
            }
        for (int j = 0; j < p.n; ++j)
            for (int i = 0; i < p.m; ++i) {
                if (p.tri == 1 && i < j) continue;
                if (p.tri == 2 && i > j) continue;
                W v = c[size_t(i) + size_t(j) * p.m];
                if (p.herk && i == j) v = re_only(v);
                st(C[i + int64_t(j) * p.ldc], v);
            }
    }
}

template <typename T>
void emu_fill(int kind, const T* const* diag, T* out, int ldd, int64_t te, int nfull, int nlast, int ntiles, int unit)
{
    using W = typename Wide<T>::type;
    for (int k = 0; k < ntiles; ++k) {
        const int n = (k == ntiles - 1) ? nlast : nfull;
        const T* a = diag[k];
        T* o = out + int64_t(k) * te;
        for (int c = 0; c < n; ++c)
            for (int r = 0; r < n; ++r) {
                W v;
                if (kind == 'H')      v = r > c ? ld(a[r + int64_t(c) * ldd]) : r < c ? cj(ld(a[c + int64_t(r) * ldd])) : re_only(ld(a[r + int64_t(c) * ldd]));
                else if (kind == 'S') v = r >= c ? ld(a[r + int64_t(c) * ldd]) : ld(a[c + int64_t(r) * ldd]);
                else                  v = r > c ? ld(a[r + int64_t(c) * ldd]) : r == c ? (unit ? W(1) : ld(a[r + int64_t(c) * ldd])) : W(0);
                st(o[r + int64_t(c) * ldd], v);
            }
    }
}

// trtri_diag(_fast)_kernel (factor_small.cu): inverse of every IB x IB diagonal block of the triangular tile(s), block b of tile
// y at W + (y * nblk + b) * IB * IB (ld = IB), zero in the other triangle, identity padding for a ragged last block
constexpr int IB = 64;
template <typename T>
void emu_trtri(const T* Tm, int ldt, int na, int lower, int unit, T* Wout, const T* const* Tarr, int na_last, int nblk, int ntiles)
{
    using W = typename Wide<T>::type;
    for (int y = 0; y < ntiles; ++y) {
        const T* tm = Tarr ? Tarr[y] : Tm;
        const int n_this = (Tarr && y == ntiles - 1) ? na_last : na;
        for (int b = 0; b < nblk; ++b) {
            T* w = Wout + (int64_t(Tarr ? y : 0) * nblk + b) * IB * IB;
            const int jo = b * IB, nv = std::max(0, std::min(IB, n_this - jo));
            std::vector<W> L(size_t(IB) * IB, W(0)), X(size_t(IB) * IB, W(0));
            for (int i = 0; i < IB; ++i) L[size_t(i) + size_t(i) * IB] = W(1);
            for (int c = 0; c < nv; ++c)
                for (int r = 0; r < nv; ++r) {
                    const bool in = lower ? r >= c : r <= c;
                    if (! in) continue;
                    W v = ld(tm[(jo + r) + int64_t(jo + c) * ldt]);
                    if (r == c && unit) v = W(1);
                    L[size_t(r) + size_t(c) * IB] = v;
                }
            // invert the triangular IB x IB matrix (identity-padded) by substitution, column by column
            for (int c = 0; c < IB; ++c) {
                if (lower) {
                    for (int r = 0; r < IB; ++r) {
                        if (r < c) { X[size_t(r) + size_t(c) * IB] = W(0); continue; }
                        W s = (r == c) ? W(1) : W(0);
                        for (int l = c; l < r; ++l) s -= L[size_t(r) + size_t(l) * IB] * X[size_t(l) + size_t(c) * IB];
                        X[size_t(r) + size_t(c) * IB] = s / L[size_t(r) + size_t(r) * IB];
                    }
                }
                else {
                    for (int r = IB - 1; r >= 0; --r) {
                        if (r > c) { X[size_t(r) + size_t(c) * IB] = W(0); continue; }
                        W s = (r == c) ? W(1) : W(0);
                        for (int l = r + 1; l <= c; ++l) s -= L[size_t(r) + size_t(l) * IB] * X[size_t(l) + size_t(c) * IB];
                        X[size_t(r) + size_t(c) * IB] = s / L[size_t(r) + size_t(r) * IB];
                    }
                }
            }
            for (size_t e = 0; e < size_t(IB) * IB; ++e) st(w[e], X[e]);
        }
    }
}

template <typename T>
void emu_scale(T* x, T beta, int zero, int64_t count)
{
    using W = typename Wide<T>::type;
    for (int64_t e = 0; e < count; ++e) st(x[e], zero ? W(0) : ld(beta) * ld(x[e]));
}

template <typename T>
void emu_gather_rows(const T* in, T* out, const int* perm, int64_t m, int64_t n, int nb, int64_t mt)
{
    const int64_t te = int64_t(nb) * nb;
    for (int64_t c = 0; c < n; ++c)
        for (int64_t x = 0; x < m; ++x) {
            const int64_t y = perm[x], jt = c / nb, cc = c % nb;
            out[(jt * mt + x / nb) * te + (x % nb) + cc * nb] = in[(jt * mt + y / nb) * te + (y % nb) + cc * nb];
        }
}

// norm_kernel (norms.cu): per-tile partial results.  cfg.shape 0 general | 1 lower (i >= j) | 2 upper; cfg.sym: the stored
// triangle stands for the full symmetric / Hermitian tile (herm: |real| on the diagonal); cfg.unit: implicit unit diagonal.
// modes 'M' max, 'O' column sums (of the FULL tile when sym), 'I' row sums, 'F' (scale, sumsq).
struct NormCfg { int shape, sym, herm, unit; };
template <typename T, typename R>
bool emu_norm(int mode, NormCfg cfg, int m, int n, const T* const* A, int64_t lda, R* values, int64_t ldv, int batch)
{
    if (mode != 'M' && mode != 'O' && mode != 'I' && mode != 'F') return false;
    for (int t = 0; t < batch; ++t) {
        const T* a = A[t];
        R* out = values + int64_t(t) * ldv;
        auto in_shape = [&](int i, int j) { return cfg.shape == 0 || (cfg.shape == 1 ? i >= j : i <= j); };
        auto absv = [&](int i, int j) -> double {
            if (i == j && cfg.unit) return 1.0;
            auto v = ld(a[i + int64_t(j) * lda]);
            if (i == j && cfg.herm) return std::abs(double(std::real(v)));
            return double(std::abs(v));
        };
        if (mode == 'M') {
            double mx = 0.0; bool nan = false;
            for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) if (in_shape(i, j)) { double v = absv(i, j); nan |= v != v; if (v > mx) mx = v; }
            out[0] = R(nan ? NAN : mx);
        }
        else if (mode == 'F') {
            double sc = 0.0; bool nan = false;
            for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) if (in_shape(i, j)) { double v = absv(i, j); nan |= v != v; if (v > sc) sc = v; }
            double sq = 0.0;
            if (sc > 0.0)
                for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) if (in_shape(i, j)) {
                    const double r = absv(i, j) / sc; sq += ((cfg.sym && i != j) ? 2.0 : 1.0) * r * r;
                }
            out[0] = R(nan ? NAN : sc); out[1] = R(sc > 0.0 ? sq : 1.0);
        }
        else {
            const int len = mode == 'O' ? n : m;
            std::vector<double> acc(size_t(len), 0.0);
            for (int j = 0; j < n; ++j) for (int i = 0; i < m; ++i) if (in_shape(i, j)) {
                const double v = absv(i, j);
                if (mode == 'O') { acc[size_t(j)] += v; if (cfg.sym && i != j) acc[size_t(i)] += v; }
                else             acc[size_t(i)] += v;
            }
            for (int e = 0; e < len; ++e) out[e] = R(acc[size_t(e)]);
        }
    }
    return true;
}

template <typename V> V arg(void** args, int i) { return *static_cast<V*>(args[i]); }

// the scalar type of a kernel instantiation, from its demangled name
int type_of(const std::string& n)
{
    if (n.find("<double2") != std::string::npos || n.find("double2>") != std::string::npos) return 'z';
    if (n.find("<float2") != std::string::npos || n.find("float2>") != std::string::npos) return 'c';
    if (n.find("<double") != std::string::npos) return 'd';
    if (n.find("<float") != std::string::npos) return 's';
    return 0;
}

#define BY_TYPE(t, CALL) \
    switch (t) { case 's': { using T = float; CALL; break; } case 'd': { using T = double; CALL; break; } \
                 case 'c': { using T = c32; CALL; break; } case 'z': { using T = c64; CALL; break; } default: return false; }

bool emulate(const std::string& name, dim3 grid, void** args)
{
    const int t = type_of(name);
    if (name.find("gemm_dmma_kernel<") != std::string::npos) {
        // gemm_dmma_kernel<Cfg, AK, BKM>: AK = op(A) K-major (opA = T), BKM = op(B) K-major (opB = N)
        const size_t c2 = name.rfind(", "), c1 = name.rfind(", ", c2 - 1);
        const bool ak = name.compare(c1 + 2, 4, "true") == 0, bkm = name.compare(c2 + 2, 4, "true") == 0;
        emu_gemm<double>(arg<GemmParamsT<double>>(args, 0), ak ? 'T' : 'N', bkm ? 'N' : 'T');
        return true;
    }
    if (name.find("gemm_zdmma_kernel<") != std::string::npos) {
        const size_t lt = name.find('<'), cm = name.find(", ", lt);
        const bool ak = name.compare(lt + 1, 4, "true") == 0, bkm = name.compare(cm + 2, 4, "true") == 0;
        const int conjA = arg<int>(args, 1), conjB = arg<int>(args, 2);
        emu_gemm<c64>(arg<GemmParamsT<c64>>(args, 0), ak ? (conjA ? 'C' : 'T') : 'N', bkm ? 'N' : (conjB ? 'C' : 'T'));
        return true;
    }
    if (name.find("gemm_generic_kernel<") != std::string::npos) {
        BY_TYPE(t, emu_gemm<T>(arg<GemmParamsT<T>>(args, 0), arg<int>(args, 1), arg<int>(args, 2)));
        return true;
    }
    if (name.find("gemm_skinny_kernel<") != std::string::npos) {
        BY_TYPE(t, emu_gemm<T>(arg<GemmParamsT<T>>(args, 0), arg<int>(args, 1), 'N'));
        return true;
    }
    for (const char* k : {"he_fill_kernel<", "sy_fill_kernel<", "tr_fill_kernel<"})
        if (name.find(k) != std::string::npos && name.find("rv_") == std::string::npos) {
            const int kind = k[0] == 'h' ? 'H' : k[0] == 's' ? 'S' : 'T';
            BY_TYPE(t, emu_fill<T>(kind, arg<const T* const*>(args, 0), arg<T*>(args, 1), arg<int>(args, 2), arg<int64_t>(args, 3),
                                   arg<int>(args, 4), arg<int>(args, 5), arg<int>(args, 6), kind == 'T' ? arg<int>(args, 7) : 0));
            return true;
        }
    if (name.find("trtri_diag_kernel<") != std::string::npos || name.find("trtri_diag_fast_kernel<") != std::string::npos) {
        BY_TYPE(t, emu_trtri<T>(arg<const T*>(args, 0), arg<int>(args, 1), arg<int>(args, 2), arg<int>(args, 3), arg<int>(args, 4),
                                arg<T*>(args, 5), arg<const T* const*>(args, 6), arg<int>(args, 7), int(grid.x), int(grid.y)));
        return true;
    }
    if (name.find("::norm_kernel<") != std::string::npos) {
        bool ok = false;
        switch (t) {
            case 's': ok = emu_norm<float, float>(arg<int>(args, 0), arg<NormCfg>(args, 1), arg<int>(args, 2), arg<int>(args, 3), arg<const float* const*>(args, 4), arg<int64_t>(args, 5), arg<float*>(args, 6), arg<int64_t>(args, 7), int(grid.x)); break;
            case 'd': ok = emu_norm<double, double>(arg<int>(args, 0), arg<NormCfg>(args, 1), arg<int>(args, 2), arg<int>(args, 3), arg<const double* const*>(args, 4), arg<int64_t>(args, 5), arg<double*>(args, 6), arg<int64_t>(args, 7), int(grid.x)); break;
            case 'c': ok = emu_norm<c32, float>(arg<int>(args, 0), arg<NormCfg>(args, 1), arg<int>(args, 2), arg<int>(args, 3), arg<const c32* const*>(args, 4), arg<int64_t>(args, 5), arg<float*>(args, 6), arg<int64_t>(args, 7), int(grid.x)); break;
            case 'z': ok = emu_norm<c64, double>(arg<int>(args, 0), arg<NormCfg>(args, 1), arg<int>(args, 2), arg<int>(args, 3), arg<const c64* const*>(args, 4), arg<int64_t>(args, 5), arg<double*>(args, 6), arg<int64_t>(args, 7), int(grid.x)); break;
            default: break;
        }
        return ok;
    }
    if (name.find("::scale_kernel<") != std::string::npos) {
        BY_TYPE(t, emu_scale<T>(arg<T*>(args, 0), arg<T>(args, 1), arg<int>(args, 2), arg<int64_t>(args, 3)));
        return true;
    }
    if (name.find("::gather_rows_kernel<") != std::string::npos && name.find("tnt_") == std::string::npos) {
        BY_TYPE(t, emu_gather_rows<T>(arg<const T*>(args, 0), arg<T*>(args, 1), arg<const int*>(args, 2), arg<int64_t>(args, 3),
                                      arg<int64_t>(args, 4), arg<int>(args, 5), arg<int64_t>(args, 6)));
        return true;
    }
    return false;
}

} // namespace

extern "C" {

// ---- registration (called by the static initialisers nvcc generates) ------------------------------------------------
void** __cudaRegisterFatBinary(void*) { static void* h = nullptr; return &h; }
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char* hostFun, char*, const char* deviceName, int, uint3*, uint3*, dim3*, dim3*, int*)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_kernels[hostFun] = demangle(deviceName);
}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, struct CUstream_st* stream)
{
    g_cfg.push_back(CallCfg{grid, block, smem, stream});
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream)
{
    if (g_cfg.empty()) return cudaErrorInvalidConfiguration;
    CallCfg c = g_cfg.back(); g_cfg.pop_back();
    *grid = c.grid; *block = c.block; *smem = c.smem; *static_cast<cudaStream_t*>(stream) = c.stream;
    return cudaSuccess;
}

cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3, void** args, size_t, cudaStream_t)
{
    std::string name;
    { std::lock_guard<std::mutex> lk(g_mu); auto it = g_kernels.find(func); if (it != g_kernels.end()) name = it->second; }
    if (name.empty() || ! emulate(name, grid, args)) {
        std::fprintf(stderr, "[fake_cudart] NO EMULATION for kernel %s\n", name.empty() ? "<unregistered>" : name.c_str());
        g_last = cudaErrorNotSupported;
        return cudaErrorNotSupported;
    }
    if (getenv("FAKE_CUDART_TRACE")) std::fprintf(stderr, "[fake_cudart] %s\n", name.substr(0, 110).c_str());
    return cudaSuccess;
}
cudaError_t cudaLaunchCooperativeKernel(const void* f, dim3 g, dim3 b, void** a, size_t s, cudaStream_t st) { return cudaLaunchKernel(f, g, b, a, s, st); }

// ---- memory: the "device" is the host --------------------------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 16); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t)
{
    for (size_t r = 0; r < h; ++r) std::memmove(static_cast<char*>(d) + r * dp, static_cast<const char*>(s) + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* pool, int) { *pool = nullptr; return cudaErrorNotSupported; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }

// ---- device, streams, events: tokens ---------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int)
{
    *v = a == cudaDevAttrMultiProcessorCount ? 148 : a == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 232448 : 1;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = reinterpret_cast<cudaStream_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorNotSupported ? "fake_cudart: kernel without an emulation" : "fake_cudart error"; }
cudaError_t cudaGetDriverEntryPoint(const char*, void** f, unsigned long long, cudaDriverEntryPointQueryResult* r)
{
    *f = nullptr;
    if (r) *r = cudaDriverEntryPointSymbolNotFound;
    return cudaErrorNotSupported;
}
cudaError_t cudaProfilerStart(void) { return cudaSuccess; }
cudaError_t cudaProfilerStop(void) { return cudaSuccess; }

} // extern "C"
