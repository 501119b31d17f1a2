// gemm_dmma.cu -- launchers + C ABI for the FP64 batched tile GEMM / HERK / SYRK.
#include "gemm_dmma.cuh"
#include <mutex>
#include <cstdlib>

namespace sb200 {

std::atomic<int64_t> g_launch_count{0};

template <typename Cfg, bool AK, bool BKM>
static int launch_variant(const GemmParamsD& p, cudaStream_t stream)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, THREADS = Cfg::THREADS;
    constexpr size_t smem = Cfg::template smem_bytes<AK, BKM>();
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(once[dev & 63], [] {
        cudaFuncSetAttribute(gemm_dmma_kernel<Cfg, AK, BKM>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        cudaFuncSetAttribute(gemm_dmma_kernel<Cfg, AK, BKM>,
                             cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    });
    const int64_t tiles = ceil_div(p.m, BM) * ceil_div(p.n, BN);
    const int64_t grid = tiles * p.batch;
    if (grid <= 0) return SB200_OK;
    if (grid > 0x7fffffffLL) return SB200_EINVAL;
    gemm_dmma_kernel<Cfg, AK, BKM><<<unsigned(grid), THREADS, smem, stream>>>(p);
    return launch_status();
}

int launch_gemm_d(int opA, int opB, GemmParamsD p, cudaStream_t stream)
{
    if (p.m <= 0 || p.n <= 0 || p.batch <= 0) return SB200_OK;
    const bool ak = (opA != 'N');     // op(A) = A^T: k contiguous
    const bool bk = (opB == 'N');     // op(B) = B  : k contiguous
    static const int cfg_sel = [] { const char* e = getenv("SB200_GEMM_CFG"); return e ? atoi(e) : 0; }();
    if (cfg_sel == 1) {
        using Cfg = GemmCfg<128, 64, 64, 32, true>;    // 4 consumer warps/CTA (tuning experiment)
        if (ak) return bk ? launch_variant<Cfg, true, true>(p, stream)  : launch_variant<Cfg, true, false>(p, stream);
        else    return bk ? launch_variant<Cfg, false, true>(p, stream) : launch_variant<Cfg, false, false>(p, stream);
    }
    using Cfg = GemmCfgDefault;
    if (ak) return bk ? launch_variant<Cfg, true, true>(p, stream)  : launch_variant<Cfg, true, false>(p, stream);
    else    return bk ? launch_variant<Cfg, false, true>(p, stream) : launch_variant<Cfg, false, false>(p, stream);
}

static int gemm_batched_d_impl(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                               double alpha, const double* const* dA, int64_t offA, int64_t lda,
                               const double* const* dB, int64_t offB, int64_t ldb,
                               double beta, double* const* dC, int64_t offC, int64_t ldc,
                               int64_t batch, int tri, cudaStream_t stream)
{
    if (! valid_layout(layout) || ! valid_op(opA) || ! valid_op(opB)) return SB200_EINVAL;
    if (m < 0 || n < 0 || k < 0 || batch < 0) return SB200_EINVAL;
    if (m == 0 || n == 0 || batch == 0) return SB200_OK;
    if (m > 0x7fffffff || n > 0x7fffffff || k > 0x7fffffff
        || lda > 0x7fffffff || ldb > 0x7fffffff || ldc > 0x7fffffff || batch > 0x7fffffff)
        return SB200_EINVAL;
    if (layout == 'R') {
        // row-major C = op(A) op(B)  <=>  column-major C^T = op(B)^T op(A)^T: swap operands and m/n
        // (same trick as the reference: blaspp/src/device_batch_gemm.cc:114-121)
        std::swap(opA, opB); std::swap(dA, dB); std::swap(offA, offB); std::swap(lda, ldb); std::swap(m, n);
        if (tri == 1) tri = 2; else if (tri == 2) tri = 1;
    }
    const int64_t rowsA = (opA == 'N') ? m : k, rowsB = (opB == 'N') ? k : n;
    if (lda < (rowsA > 1 ? rowsA : 1) || ldb < (rowsB > 1 ? rowsB : 1) || ldc < m) return SB200_EINVAL;
    GemmParamsD p{};
    p.A = dA; p.B = dB; p.C = dC;
    p.offA = offA; p.offB = offB; p.offC = offC;
    p.m = int(m); p.n = int(n); p.k = int(k);
    p.lda = int(lda); p.ldb = int(ldb); p.ldc = int(ldc);
    p.alpha = alpha; p.beta = beta; p.batch = int(batch); p.tri = tri;
    return launch_gemm_d(opA, opB, p, stream);
}

} // namespace sb200

using namespace sb200;

extern "C" {

int sb200_version(void) { return SB200_VERSION; }

const char* sb200_strerror(int code)
{
    switch (code) {
        case SB200_OK:      return "success";
        case SB200_EINVAL:  return "invalid argument";
        case SB200_ENOTSUP: return "not supported";
        case SB200_ENOMEM:  return "out of memory";
        case SB200_ENODEV:  return "no CUDA device (slate_b200 has no CPU fallback)";
        case SB200_ENCCL:   return "NCCL error";
        default:            return code > 0 ? cudaGetErrorString(cudaError_t(code)) : "unknown error";
    }
}

int sb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return SB200_ENODEV; }
    return n;
}

int64_t sb200_launch_count(void) { return g_launch_count.load(); }

int sb200_gemm_batched_d(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                         double alpha, const double* const* dA, int64_t lda,
                         const double* const* dB, int64_t ldb,
                         double beta, double* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream)
{
    return gemm_batched_d_impl(layout, opA, opB, m, n, k, alpha, dA, 0, lda, dB, 0, ldb,
                               beta, dC, 0, ldc, batch, 0, cudaStream_t(stream));
}

int sb200_gemm_batched_off_d(int layout, int opA, int opB, int64_t m, int64_t n, int64_t k,
                             double alpha, const double* const* dA, int64_t offA, int64_t lda,
                             const double* const* dB, int64_t offB, int64_t ldb,
                             double beta, double* const* dC, int64_t offC, int64_t ldc,
                             int64_t batch, sb200_stream_t stream)
{
    return gemm_batched_d_impl(layout, opA, opB, m, n, k, alpha, dA, offA, lda, dB, offB, ldb,
                               beta, dC, offC, ldc, batch, 0, cudaStream_t(stream));
}

// herk == syrk for real types (the reference maps real blas::herk to syrk:
// blaspp/src/device_herk.cc:95-125)
int sb200_syrk_batched_d(int layout, int uplo, int op, int64_t n, int64_t k,
                         double alpha, const double* const* dA, int64_t lda,
                         double beta, double* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream)
{
    if (! valid_uplo(uplo) || ! valid_op(op)) return SB200_EINVAL;
    const int tri = (uplo == 'L') ? 1 : 2;
    const int opA = (op == 'N') ? 'N' : 'T';
    const int opB = (op == 'N') ? 'T' : 'N';
    return gemm_batched_d_impl(layout, opA, opB, n, n, k, alpha, dA, 0, lda, dA, 0, lda,
                               beta, dC, 0, ldc, batch, tri, cudaStream_t(stream));
}

int sb200_herk_batched_d(int layout, int uplo, int op, int64_t n, int64_t k,
                         double alpha, const double* const* dA, int64_t lda,
                         double beta, double* const* dC, int64_t ldc,
                         int64_t batch, sb200_stream_t stream)
{
    return sb200_syrk_batched_d(layout, uplo, op, n, k, alpha, dA, lda, beta, dC, ldc, batch, stream);
}

} // extern "C"
