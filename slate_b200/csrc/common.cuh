// common.cuh -- shared device/host helpers for libslate_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <stdint.h>
#include <cstdlib>
#include <atomic>
#include "../../include/slate_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "slate_b200 is written for sm_100a (Blackwell B200) only"
#endif

namespace sb200 {

extern std::atomic<int64_t> g_launch_count;

// Every kernel launch in the library goes through this so that sb200_launch_count()
// is an honest count of OUR kernels (bench.py reports it as gpu_launches).
inline int launch_status()
{
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? SB200_OK : int(e);
}

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- opt-in switches
// Defaults of the per-call switches in ONE place (each was measured on a B200 before it was turned on; the losers
// of round 2 -- multi-warp / warp-synchronous diagonal kernels, tagged-word LU kernel of getrf.cu, hand-rolled grid
// barrier, skinny panel update, persistent trailing GEMM -- were deleted).  The environment variable of the same name overrides the default; every site reads it per call.
struct Switch { const char* env; int dflt; };
constexpr Switch SW_TILE_FUSED   {"SB200_TILE_FUSED", 0};     // one-launch tile Cholesky: 1 | 2 (rsqrt); the potrf driver turns it on when the chain has its own SM partition
constexpr Switch SW_TRSM_FUSED   {"SB200_TRSM_FUSED", 7};     // bit 0 panel solve, bit 1 row solve, bit 2 small-triangle (<= 32) solve: measured r2a, on
constexpr Switch SW_PANEL_V3     {"SB200_PANEL_V3", 1};       // LU base kernel with one exchange round per column (getrf_base_v3.cu)
constexpr Switch SW_GEMM_BT      {"SB200_GEMM_BT", 1};        // transposed B panel / U row: 'N','T' multiply (measured r2a: dgemm 30.5 -> 35.3 TF/s), on
inline int switch_value(const Switch& s)
{
    const char* e = getenv(s.env);
    return e ? atoi(e) : s.dflt;
}

inline bool valid_op(int op)       { return op == 'N' || op == 'T' || op == 'C'; }
inline bool valid_layout(int l)    { return l == 'C' || l == 'R'; }
inline bool valid_uplo(int u)      { return u == 'L' || u == 'U'; }
inline bool valid_diag(int d)      { return d == 'N' || d == 'U'; }
inline bool valid_side(int s)      { return s == 'L' || s == 'R'; }

// ----------------------------------------------------------------------------- PTX wrappers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// add expected transaction bytes without arriving
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// 16-byte Ampere-style async copy global -> shared (SASS: LDGSTS), L2-only caching
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// arrive on `bar` once all cp.async issued so far by this thread have landed
// (.noinc: the arrival counts against the barrier's initial expected count)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (! mbar_try_wait(bar, parity)) { }
}

// 1-D bulk async copy global -> shared through the TMA engine (SASS: UBLKCP), completion
// signalled as transaction bytes on an mbarrier.  src/dst 16-byte aligned, bytes % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// warp-specialised register re-allocation (whole warpgroup must execute it)
template <int REGS> __device__ __forceinline__ void setmaxnreg_inc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(REGS));
}
template <int REGS> __device__ __forceinline__ void setmaxnreg_dec()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(REGS));
}

// FP64 tensor-core MMA (SASS: DMMA.8x8x4): D(8x8) += A(8x4, row) * B(4x8, col).
// lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Watchdog for the spin-wait protocols between CTAs (flags of the fused tile Cholesky, tagged words and the arrival
// counter of the LU base kernels): a wait that lasts seconds is a protocol bug, and a kernel that spins forever takes
// the GPU down with it, so trap instead (the launch then fails with a CUDA error the host sees).
// 2^33 cycles = 4.4 s at 1.965 GHz; a legitimate wait is at most the time for a co-scheduled CTA to get an SM slot.
__device__ __forceinline__ void spin_watchdog(long long t0)
{
    if (clock64() - t0 > (1LL << 33)) __trap();
}
#endif // __CUDACC__

} // namespace sb200
