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

} // extern "C"
