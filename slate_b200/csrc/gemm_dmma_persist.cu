// gemm_dmma_persist.cu -- PERSISTENT variant of the FP64 DMMA batched tile GEMM (gemm_dmma.cuh), opt-in.
//
// Why: the panel kernels of potrf / getrf (high-priority stream) are tiny and strictly dependent; while a
// trailing update with thousands of CTAs is resident they each wait for a CTA slot to drain (one GEMM CTA
// runs ~60-100 us), which is what makes the panel chain the bound of the multi-GPU runs.
// Here the trailing update is launched with FEWER CTAs than there are slots (2 per SM minus a reserve) and
// every CTA loops over tiles, so that panel kernels always find a free slot.  The stage ring runs on across
// tiles: the producer warp is already loading the next tile while the consumer warps store the current one.
//
// Selected by SB200_GEMM_PERSIST=<reserved CTA slots> (e.g. 16); 0 / unset = the one-tile-per-CTA kernel.
// Same tile shape, fragment layout and arithmetic order as gemm_dmma_kernel: results are bit-identical.
#include "gemm_dmma.cuh"
#include <mutex>
#include <cstdlib>

namespace sb200 {

// A_KMAJ: op(A)(i,l) = A[l + i*lda] (k contiguous; op(A) = T in column-major terms)
// B_KMAJ: op(B)(l,j) = B[l + j*ldb] (k contiguous; op(B) = N)
template <typename Cfg, bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::CTAS_PER_SM)
gemm_dmma_persist_kernel(const GemmParamsD p, int total_work)
{
    constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, WM = Cfg::WM, WN = Cfg::WN;
    constexpr int MI = Cfg::MI, NJ = Cfg::NJ, CONSUMER_WARPS = Cfg::CONSUMER_WARPS;
    constexpr int LDK = Cfg::LDK, LDA_MN = Cfg::LDA_MN, LDB_MN = Cfg::LDB_MN;
    constexpr int STAGES = Cfg::template stages<A_KMAJ>();
    constexpr int A_STAGE = Cfg::template a_stage<A_KMAJ>();
    constexpr int B_STAGE = Cfg::template b_stage<B_KMAJ>();

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sA = reinterpret_cast<double*>(smem_raw);
    double* sB = sA + STAGES * A_STAGE;
    uint64_t* full_bar  = reinterpret_cast<uint64_t*>(sB + STAGES * B_STAGE);
    uint64_t* empty_bar = full_bar + STAGES;

    const int tiles_m = (p.m + BM - 1) / BM;
    const int tiles_n = (p.n + BN - 1) / BN;
    const int per_problem = tiles_m * tiles_n;
    (void) tiles_n;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        #pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 32);
            mbar_init(&empty_bar[s], CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int num_kt = (p.k + BK - 1) / BK;
    // work item -> (problem t, tile origin m0, n0); false for tiles wholly inside the discarded triangle
    auto decode = [&](int work, int& t, int& m0, int& n0) -> bool {
        t = work / per_problem;
        const int r = work - t * per_problem;
        m0 = (r % tiles_m) * BM;
        n0 = (r / tiles_m) * BN;
        if (p.tri == 1 && n0 >= m0 + BM) return false;
        if (p.tri == 2 && m0 >= n0 + BN) return false;
        return true;
    };

    if (warp >= CONSUMER_WARPS) {
        // ===================== producer warpgroup =====================
        setmaxnreg_dec<Cfg::PRODUCER_REGS>();
        if (warp != CONSUMER_WARPS) return;
        // MN-major operands: one bulk copy (UBLKCP) per k column.  K-major operands: 16-byte
        // cp.async (LDGSTS) chunks -- 128-byte bulk copies were measured 3x slower.  Every lane
        // arrives once per stage on full_bar (count 32); bulk bytes are added with expect_tx.
        uint32_t kt_base = 0;                   // k-steps issued so far: the stage ring runs on across tiles, so the
                                                // next tile's operands are in flight during this tile's epilogue
        for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        int t, m0, n0;
        if (! decode(work, t, m0, n0)) continue;
        const double* __restrict__ A = (p.A ? p.A[t] : p.A0 + int64_t(t) * p.strideA) + p.offA;
        const double* __restrict__ B = (p.B ? p.B[t] : p.B0 + int64_t(t) * p.strideB) + p.offB;
        const int mv = min(BM, p.m - m0);
        const int nv = min(BN, p.n - n0);
        const bool aligned =
            ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0
            && ((p.lda | p.ldb | mv | nv) & 1) == 0 && (p.k & 3) == 0;
        for (int kt = 0; kt < num_kt; ++kt) {
            const uint32_t g = kt_base + uint32_t(kt);
            const int s = int(g % STAGES);
            const uint32_t ph = (g / STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int k0 = kt * BK;
            const int kv = min(BK, p.k - k0);
            double* dA = sA + s * A_STAGE;
            double* dB = sB + s * B_STAGE;
            if (aligned) {
                if (lane == 0) {
                    const uint32_t bytes = (A_KMAJ ? 0u : uint32_t(mv) * uint32_t(kv) * 8u)
                                         + (B_KMAJ ? 0u : uint32_t(nv) * uint32_t(kv) * 8u);
                    if (bytes) mbar_expect_tx(&full_bar[s], bytes);
                }
                __syncwarp();
                const int cpr = kv >> 1;                     // 16-byte chunks per K-major row
                if (A_KMAJ) {
                    const double* src = A + k0 + int64_t(m0) * p.lda;
                    for (int c = lane; c < mv * cpr; c += 32) {
                        const int i = c / cpr, q = c - i * cpr;
                        cp_async16(dA + i * LDK + 2 * q, src + int64_t(i) * p.lda + 2 * q);
                    }
                }
                else {
                    if (lane < kv)
                        bulk_g2s(dA + lane * LDA_MN, A + m0 + int64_t(k0 + lane) * p.lda, mv * 8, &full_bar[s]);
                }
                if (B_KMAJ) {
                    const double* src = B + k0 + int64_t(n0) * p.ldb;
                    for (int c = lane; c < nv * cpr; c += 32) {
                        const int j = c / cpr, q = c - j * cpr;
                        cp_async16(dB + j * LDK + 2 * q, src + int64_t(j) * p.ldb + 2 * q);
                    }
                }
                else {
                    const int l = lane - 16;     // lanes 16..31 so that A and B issue in parallel (BK == 16)
                    if (l >= 0 && l < kv)
                        bulk_g2s(dB + l * LDB_MN, B + n0 + int64_t(k0 + l) * p.ldb, nv * 8, &full_bar[s]);
                }
                if (A_KMAJ || B_KMAJ) cp_async_mbar_arrive_noinc(&full_bar[s]);
                else                  mbar_arrive(&full_bar[s]);
            }
            else {
                // guarded fallback: element loads, zero fill to a multiple of 4 in k
                const int kz = (kv + 3) & ~3;
                if (A_KMAJ) {
                    for (int e = lane; e < BM * kz; e += 32) {
                        const int i = e / kz, l = e - i * kz;
                        dA[i * LDK + l] = (i < mv && l < kv) ? A[k0 + l + int64_t(m0 + i) * p.lda] : 0.0;
                    }
                }
                else {
                    for (int e = lane; e < BM * kz; e += 32) {
                        const int l = e / BM, i = e - l * BM;
                        dA[l * LDA_MN + i] = (i < mv && l < kv) ? A[m0 + i + int64_t(k0 + l) * p.lda] : 0.0;
                    }
                }
                if (B_KMAJ) {
                    for (int e = lane; e < BN * kz; e += 32) {
                        const int j = e / kz, l = e - j * kz;
                        dB[j * LDK + l] = (j < nv && l < kv) ? B[k0 + l + int64_t(n0 + j) * p.ldb] : 0.0;
                    }
                }
                else {
                    for (int e = lane; e < BN * kz; e += 32) {
                        const int l = e / BN, j = e - l * BN;
                        dB[l * LDB_MN + j] = (j < nv && l < kv) ? B[n0 + j + int64_t(k0 + l) * p.ldb] : 0.0;
                    }
                }
                mbar_arrive(&full_bar[s]);
            }
        }
        kt_base += uint32_t(num_kt);
        }
        return;
    }

    // ===================== consumer warps =====================
    setmaxnreg_inc<Cfg::CONSUMER_REGS>();
    const int wm = (warp % Cfg::WARPS_M) * WM;
    const int wn = (warp / Cfg::WARPS_M) * WN;
    const int lr = lane >> 2;      // 0..7  fragment row (A) / column (B)
    const int lc = lane & 3;       // 0..3  fragment k index

    double acc[MI][NJ][2];

    // per-lane base offsets into a stage
    const int a_base = A_KMAJ ? (wm + lr) * LDK + lc : lc * LDA_MN + wm + lr;
    const int b_base = B_KMAJ ? (wn + lr) * LDK + lc : lc * LDB_MN + wn + lr;
    constexpr int A_MI = A_KMAJ ? 8 * LDK : 8;          // step between 8-row blocks
    constexpr int A_K4 = A_KMAJ ? 4 : 4 * LDA_MN;       // step between k4 slices
    constexpr int B_NJ = B_KMAJ ? 8 * LDK : 8;
    constexpr int B_K4 = B_KMAJ ? 4 : 4 * LDB_MN;

    auto load_frag = [&](double (&a)[MI], double (&b)[NJ], int s, int k4) {
        const double* __restrict__ cA = sA + s * A_STAGE + a_base + k4 * A_K4;
        const double* __restrict__ cB = sB + s * B_STAGE + b_base + k4 * B_K4;
        #pragma unroll
        for (int i = 0; i < MI; ++i) a[i] = cA[i * A_MI];
        #pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] = cB[j * B_NJ];
    };
    auto mma_all = [&](const double (&a)[MI], const double (&b)[NJ]) {
        #pragma unroll
        for (int i = 0; i < MI; ++i)
            #pragma unroll
            for (int j = 0; j < NJ; ++j)
                dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    };

    static_assert(! Cfg::DBUF, "the persistent variant implements the 4-warps-per-sub-partition schedule");
    uint32_t kt_base = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
    int t, m0, n0;
    if (! decode(work, t, m0, n0)) continue;
    #pragma unroll
    for (int i = 0; i < MI; ++i)
        #pragma unroll
        for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    {
        // 4 warps per SM sub-partition hide each other's shared-memory latency: keep the
        // register footprint small instead of double-buffering.
        const int total_k4 = (p.k + 3) >> 2;
        for (int kt = 0; kt < num_kt; ++kt) {
            const uint32_t g = kt_base + uint32_t(kt);
            const int s = int(g % STAGES);
            const uint32_t ph = (g / STAGES) & 1;
            const int nk4 = min(BK / 4, total_k4 - kt * (BK / 4));
            mbar_wait(&full_bar[s], ph);
            if (nk4 == BK / 4) {
                #pragma unroll
                for (int k4 = 0; k4 < BK / 4; ++k4) {
                    double a[MI], b[NJ];
                    load_frag(a, b, s, k4);
                    mma_all(a, b);
                }
            }
            else {
                for (int k4 = 0; k4 < nk4; ++k4) {
                    double a[MI], b[NJ];
                    load_frag(a, b, s, k4);
                    mma_all(a, b);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        }
    }

    kt_base += uint32_t(num_kt);
    // ===================== epilogue =====================
    // Per 8-column block: issue ALL loads of C first (they are independent; interleaving them
    // with the stores to the same array serialises one DRAM round trip per element -- measured
    // 45 us per CTA), then scale and store.  A lane's 8 rows x 8 B form full 32-byte sectors.
    double* __restrict__ C = (p.C ? p.C[t] : p.C0 + int64_t(t) * p.strideC) + p.offC;
    const double alpha = p.alpha, beta = p.beta;
    const bool use_beta = (beta != 0.0);
    const int tri = p.tri;
    #pragma unroll
    for (int j = 0; j < NJ; ++j) {
        double cv[2][MI];
        bool ok[2][MI];
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = n0 + wn + j * 8 + 2 * lc + h;
            const double* Ccol = C + int64_t(col) * p.ldc;
            #pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int row = m0 + wm + i * 8 + lr;
                bool o = (row < p.m) && (col < p.n);
                if (tri == 1) o = o && (row >= col);
                if (tri == 2) o = o && (row <= col);
                ok[h][i] = o;
                cv[h][i] = (o && use_beta) ? __ldcg(Ccol + row) : 0.0;
            }
        }
        #pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = n0 + wn + j * 8 + 2 * lc + h;
            double* Ccol = C + int64_t(col) * p.ldc;
            #pragma unroll
            for (int i = 0; i < MI; ++i) {
                const int row = m0 + wm + i * 8 + lr;
                if (ok[h][i]) Ccol[row] = fma(alpha, acc[i][j][h], beta * cv[h][i]);
            }
        }
    }
    }   // work loop
}

template <typename Cfg, bool AK, bool BKM>
static int launch_persist_variant(const GemmParamsD& p, int64_t total, int ctas, cudaStream_t stream)
{
    constexpr size_t smem = Cfg::template smem_bytes<AK, BKM>();
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(once[dev & 63], [] {
        cudaFuncSetAttribute(gemm_dmma_persist_kernel<Cfg, AK, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        cudaFuncSetAttribute(gemm_dmma_persist_kernel<Cfg, AK, BKM>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    });
    gemm_dmma_persist_kernel<Cfg, AK, BKM><<<unsigned(ctas), Cfg::THREADS, smem, stream>>>(p, int(total));
    return launch_status();
}

constexpr int PERSIST_NOT_TAKEN = -1000000;     // distinct from every SB200_* / cudaError_t value
// Returns PERSIST_NOT_TAKEN when the persistent variant does not apply (switched off, or the launch is too small to fill the
// machine anyway); otherwise the launch status.
int launch_gemm_d_persist(int opA, int opB, const GemmParamsD& p, cudaStream_t stream)
{
    static const int reserve = [] { const char* e = getenv("SB200_GEMM_PERSIST"); return e ? atoi(e) : 0; }();
    if (reserve <= 0) return PERSIST_NOT_TAKEN;
    using Cfg = GemmCfgDefault;
    static thread_local int slots[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (slots[dev & 63] == 0) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots[dev & 63] = sms * Cfg::CTAS_PER_SM;
    }
    const int ctas = slots[dev & 63] - reserve;
    const int64_t total = ceil_div(p.m, Cfg::BM) * ceil_div(p.n, Cfg::BN) * int64_t(p.batch);
    if (ctas < 1 || total <= ctas || total > 0x7fffffffLL) return PERSIST_NOT_TAKEN;
    const bool ak = (opA != 'N'), bk = (opB == 'N');
    if (ak) return bk ? launch_persist_variant<Cfg, true, true>(p, total, ctas, stream)  : launch_persist_variant<Cfg, true, false>(p, total, ctas, stream);
    else    return bk ? launch_persist_variant<Cfg, false, true>(p, total, ctas, stream) : launch_persist_variant<Cfg, false, false>(p, total, ctas, stream);
}

} // namespace sb200
