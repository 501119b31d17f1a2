// gemm_zdmma.cu -- complex128 batched tile GEMM / HERK on the FP64 tensor-core MMA.
//
// Serves zgemm / zherk trailing updates (BASELINE config 4; reference call sites
// src/internal/internal_gemm.cc:498-504, internal_herk.cc:510-516 -> cublasZgemmBatched and a
// per-tile cublasZherk loop, blaspp/src/device_batch_herk.cc:57-73).
//
// A complex MMA is four real DMMA.8x8x4 on split operands ("4M"):
//     Cr += Ar*Br + (-Ai)*Bi        Ci += Ar*Bi + Ai*Br
// Operands stay INTERLEAVED (re, im) in global and shared memory -- the layout SLATE tiles have --
// and are split in the fragment loads: one LDS.128 fetches (re, im) of one element, conjugation and
// the minus sign are sign-bit flips in registers.  Per k4-step a warp issues 8 LDS.128 for 64 DMMAs
// (the real kernel: 8 LDS.64 for 16), so the tensor pipe, not shared memory, is the limiter.
//
// CTA tile 64 x 64 x 8 (complex), 4 consumer warps (32 x 32 warp tiles: 128 accumulator registers)
// + 1 producer warp, 2 CTAs per SM; operands staged by TMA 1-D bulk copies (MN-major) or 16-byte
// cp.async (K-major) into a 4-stage mbarrier ring.  Shared strides are chosen so that every
// quarter-warp of an LDS.128 hits 8 distinct 16-byte bank groups.
#include "gemm_dmma.cuh"
#include "scalar_ops.cuh"
#include <mutex>

namespace sb200 {

using Z = cuDoubleComplex;

namespace zcfg {
constexpr int BM = 64, BN = 64, BK = 8, WM = 32, WN = 32, MI = WM / 8, NJ = WN / 8;
constexpr int CONSUMER_WARPS = 4, THREADS = CONSUMER_WARPS * 32 + 128, STAGES = 4;
constexpr int LDK = BK + 4;                 // K-major row stride (complex): 12 = 4 mod 8
constexpr int LDMN = BM + 2;                // MN-major k-slice stride (complex): 66 = 2 mod 8
constexpr int PRODUCER_REGS = 40, CONSUMER_REGS = 208;
template <bool KMAJ> constexpr int stage_elems() { return KMAJ ? BM * LDK : BK * LDMN; }
template <bool AK, bool BKM> constexpr size_t smem_bytes()
{
    return size_t(STAGES) * (stage_elems<AK>() + stage_elems<BKM>()) * sizeof(Z) + 2 * STAGES * sizeof(uint64_t);
}
static_assert(BM == BN, "one stride constant serves both operands");
}

__device__ __forceinline__ double flip(double x, bool f)
{
    return f ? __hiloint2double(__double2hiint(x) ^ int(0x80000000), __double2loint(x)) : x;
}

// A_KMAJ: op(A)(i,l) = A[l + i*lda] (opA = T or C);  B_KMAJ: op(B)(l,j) = B[l + j*ldb] (opB = N)
template <bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(zcfg::THREADS, 2)
gemm_zdmma_kernel(const GemmParamsT<Z> p, int conjA, int conjB)
{
    using namespace zcfg;
    constexpr int A_STAGE = stage_elems<A_KMAJ>(), B_STAGE = stage_elems<B_KMAJ>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Z* sA = reinterpret_cast<Z*>(smem_raw);
    Z* sB = sA + STAGES * A_STAGE;
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
    uint64_t* empty_bar = full_bar + STAGES;

    const int tiles_m = (p.m + BM - 1) / BM, tiles_n = (p.n + BN - 1) / BN;
    const int per_problem = tiles_m * tiles_n;
    const int t = blockIdx.x / per_problem, r = blockIdx.x - t * per_problem;
    const int m0 = (r % tiles_m) * BM, n0 = (r / tiles_m) * BN;
    if (p.tri == 1 && n0 >= m0 + BM) return;
    if (p.tri == 2 && m0 >= n0 + BN) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        #pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 32); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    const Z* __restrict__ A = (p.A ? p.A[t] : p.A0 + int64_t(t) * p.strideA) + p.offA;
    const Z* __restrict__ B = (p.B ? p.B[t] : p.B0 + int64_t(t) * p.strideB) + p.offB;
    const int mv = min(BM, p.m - m0), nv = min(BN, p.n - n0);
    const int num_kt = (p.k + BK - 1) / BK;

    if (warp >= CONSUMER_WARPS) {
        // ===================== producer warpgroup =====================
        setmaxnreg_dec<PRODUCER_REGS>();
        if (warp != CONSUMER_WARPS) return;
        for (int kt = 0; kt < num_kt; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int k0 = kt * BK, kv = min(BK, p.k - k0);
            Z* dA = sA + s * A_STAGE;
            Z* dB = sB + s * B_STAGE;
            if (kv & 3) {
                // ragged last stage: zero the k slots the consumers read beyond kv
                const int kz = (kv + 3) & ~3;
                const Z zero = make_cuDoubleComplex(0.0, 0.0);
                for (int e = lane; e < BM * (kz - kv); e += 32) {
                    const int i = e / (kz - kv), l = kv + e % (kz - kv);
                    if (A_KMAJ) dA[i * LDK + l] = zero; else dA[l * LDMN + i] = zero;
                    if (B_KMAJ) dB[i * LDK + l] = zero; else dB[l * LDMN + i] = zero;
                }
                __threadfence_block();
            }
            if (lane == 0) {
                const uint32_t bytes = (A_KMAJ ? 0u : uint32_t(mv) * uint32_t(kv) * 16u)
                                     + (B_KMAJ ? 0u : uint32_t(nv) * uint32_t(kv) * 16u);
                if (bytes) mbar_expect_tx(&full_bar[s], bytes);
            }
            __syncwarp();
            if (A_KMAJ) {
                const Z* src = A + k0 + int64_t(m0) * p.lda;
                for (int c = lane; c < mv * kv; c += 32) {
                    const int i = c / kv, q = c - i * kv;
                    cp_async16(dA + i * LDK + q, src + int64_t(i) * p.lda + q);
                }
            }
            else if (lane < kv)
                bulk_g2s(dA + lane * LDMN, A + m0 + int64_t(k0 + lane) * p.lda, uint32_t(mv) * 16u, &full_bar[s]);
            if (B_KMAJ) {
                const Z* src = B + k0 + int64_t(n0) * p.ldb;
                for (int c = lane; c < nv * kv; c += 32) {
                    const int j = c / kv, q = c - j * kv;
                    cp_async16(dB + j * LDK + q, src + int64_t(j) * p.ldb + q);
                }
            }
            else {
                const int l = lane - 16;          // lanes 16.. so that A and B copies issue in parallel
                if (l >= 0 && l < kv)
                    bulk_g2s(dB + l * LDMN, B + n0 + int64_t(k0 + l) * p.ldb, uint32_t(nv) * 16u, &full_bar[s]);
            }
            if (A_KMAJ || B_KMAJ) cp_async_mbar_arrive_noinc(&full_bar[s]);
            else                  mbar_arrive(&full_bar[s]);
        }
        return;
    }

    // ===================== consumer warps =====================
    setmaxnreg_inc<CONSUMER_REGS>();
    const int wm = (warp & 1) * WM, wn = (warp >> 1) * WN;
    const int lr = lane >> 2, lc = lane & 3;
    const bool cA = conjA != 0, cB = conjB != 0;

    double accr[MI][NJ][2], acci[MI][NJ][2];
    #pragma unroll
    for (int i = 0; i < MI; ++i)
        #pragma unroll
        for (int j = 0; j < NJ; ++j) { accr[i][j][0] = accr[i][j][1] = 0.0; acci[i][j][0] = acci[i][j][1] = 0.0; }

    const int a_base = A_KMAJ ? (wm + lr) * LDK + lc : lc * LDMN + wm + lr;
    const int b_base = B_KMAJ ? (wn + lr) * LDK + lc : lc * LDMN + wn + lr;
    constexpr int A_MI = A_KMAJ ? 8 * LDK : 8, A_K4 = A_KMAJ ? 4 : 4 * LDMN;
    constexpr int B_NJ = B_KMAJ ? 8 * LDK : 8, B_K4 = B_KMAJ ? 4 : 4 * LDMN;
    const int total_k4 = (p.k + 3) >> 2;

    for (int kt = 0; kt < num_kt; ++kt) {
        const int s = kt % STAGES;
        const uint32_t ph = (kt / STAGES) & 1;
        const int nk4 = min(BK / 4, total_k4 - kt * (BK / 4));
        mbar_wait(&full_bar[s], ph);
        #pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
            if (k4 < nk4) {
                const double2* __restrict__ pa = reinterpret_cast<const double2*>(sA + s * A_STAGE + a_base + k4 * A_K4);
                const double2* __restrict__ pb = reinterpret_cast<const double2*>(sB + s * B_STAGE + b_base + k4 * B_K4);
                double ar[MI], ai[MI], nai[MI], br[NJ], bi[NJ];
                #pragma unroll
                for (int i = 0; i < MI; ++i) {
                    const double2 v = pa[i * A_MI];
                    ar[i] = v.x; ai[i] = flip(v.y, cA); nai[i] = flip(v.y, !cA);
                }
                #pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const double2 v = pb[j * B_NJ];
                    br[j] = v.x; bi[j] = flip(v.y, cB);
                }
                #pragma unroll
                for (int i = 0; i < MI; ++i)
                    #pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        dmma884(accr[i][j][0], accr[i][j][1], ar[i], br[j]);
                        dmma884(acci[i][j][0], acci[i][j][1], ar[i], bi[j]);
                        dmma884(accr[i][j][0], accr[i][j][1], nai[i], bi[j]);
                        dmma884(acci[i][j][0], acci[i][j][1], ai[i], br[j]);
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // ===================== epilogue =====================
    Z* __restrict__ C = (p.C ? p.C[t] : p.C0 + int64_t(t) * p.strideC) + p.offC;
    const Z alpha = p.alpha, beta = p.beta;
    const bool use_beta = !(beta.x == 0.0 && beta.y == 0.0);
    const int tri = p.tri;
    #pragma unroll
    for (int j = 0; j < NJ; ++j) {
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = n0 + wn + j * 8 + 2 * lc + h;
            Z* Ccol = C + int64_t(col) * p.ldc;
            Z cv[MI];
            bool ok[MI];
            #pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int row = m0 + wm + i * 8 + lr;
                bool o = (row < p.m) && (col < p.n);
                if (tri == 1) o = o && (row >= col);
                if (tri == 2) o = o && (row <= col);
                ok[i] = o;
                cv[i] = (o && use_beta) ? Ccol[row] : make_cuDoubleComplex(0.0, 0.0);
            }
            #pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int row = m0 + wm + i * 8 + lr;
                if (ok[i]) {
                    const double xr = accr[i][j][h], xi = acci[i][j][h];
                    Z v;
                    v.x = alpha.x * xr - alpha.y * xi;
                    v.y = alpha.x * xi + alpha.y * xr;
                    if (use_beta) {
                        v.x += beta.x * cv[i].x - beta.y * cv[i].y;
                        v.y += beta.x * cv[i].y + beta.y * cv[i].x;
                    }
                    if (p.herk && row == col) v.y = 0.0;
                    Ccol[row] = v;
                }
            }
        }
    }
}

template <bool AK, bool BKM>
static int launch_zvariant(const GemmParamsT<Z>& p, int conjA, int conjB, cudaStream_t stream)
{
    constexpr size_t smem = zcfg::smem_bytes<AK, BKM>();
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(once[dev & 63], [] {
        cudaFuncSetAttribute(gemm_zdmma_kernel<AK, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        cudaFuncSetAttribute(gemm_zdmma_kernel<AK, BKM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    });
    const int64_t grid = ceil_div(p.m, zcfg::BM) * ceil_div(p.n, zcfg::BN) * int64_t(p.batch);
    if (grid <= 0) return SB200_OK;
    if (grid > 0x7fffffffLL) return SB200_EINVAL;
    gemm_zdmma_kernel<AK, BKM><<<unsigned(grid), zcfg::THREADS, smem, stream>>>(p, conjA, conjB);
    return launch_status();
}

int launch_gemm_z(int opA, int opB, GemmParamsT<Z> p, cudaStream_t stream)
{
    if (p.m <= 0 || p.n <= 0 || p.batch <= 0) return SB200_OK;
    const bool ak = (opA != 'N'), bk = (opB == 'N');
    const int cA = (opA == 'C'), cB = (opB == 'C');
    if (ak) return bk ? launch_zvariant<true, true>(p, cA, cB, stream)  : launch_zvariant<true, false>(p, cA, cB, stream);
    else    return bk ? launch_zvariant<false, true>(p, cA, cB, stream) : launch_zvariant<false, false>(p, cA, cB, stream);
}

} // namespace sb200
