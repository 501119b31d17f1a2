// gemm_dmma.cuh -- FP64 batched tile GEMM on the Blackwell FP64 tensor-core MMA.
//
// One kernel family serves every FP64 contraction of the hot path:
//   internal::gemm<Devices>  (src/internal/internal_gemm.cc:354-518)  -> cublasDgemmBatched
//   internal::herk<Devices>  (src/internal/internal_herk.cc:355-536)  -> cublasDgemmBatched + per-tile cublasDsyrk
//   the GEMM steps of the blocked trsm / potrf-tile / getrf-panel kernels in this library.
//
// Design (B200 / sm_100a):
//   * CTA tile 128 x 64 x 16, 4 consumer warps (64 x 32 warp tiles, 32 independent
//     DMMA.8x8x4 accumulator pairs per k-step) + 1 producer warp; 2 CTAs per SM so one CTA's
//     epilogue overlaps the other's main loop (FP64 tensor rate is only 128 flop/clk/SM, so
//     8 consumer warps per SM saturate it; operands need ~12 B/clk/SM from L2).
//   * Operands are staged global -> shared by the TMA engine with 1-D bulk copies
//     (cp.async.bulk, SASS UBLKCP) -- no tensor map is needed, so arbitrary per-tile
//     pointers from a cublas-style pointer array work -- into a 3/4-stage ring guarded by
//     full/empty mbarriers (transaction-count completion).
//   * Shared layouts are padded (+4 doubles) so that every DMMA fragment load
//     (lane -> (row l/4, k l%4)) is bank-conflict-free for both operand majors.
//   * Unaligned / odd-sized problems take the same kernel: the producer warp falls back to
//     guarded element loads with zero fill (correct, slower).
//   * Epilogue: alpha/beta in registers, optional triangle mask (herk/syrk diagonal tiles).
#pragma once
#include "common.cuh"

namespace sb200 {

template <typename T>
struct GemmParamsT {
    const T* const* A;           // device pointer arrays (batch entries)
    const T* const* B;
    T* const*       C;
    int64_t offA, offB, offC;    // element offsets added to every pointer
    int64_t strideA, strideB, strideC;  // used when the array pointer is null: base + t*stride
    const T* A0;
    const T* B0;
    T*       C0;
    int m, n, k;
    int lda, ldb, ldc;
    T alpha, beta;
    int batch;
    int tri;                     // 0 full, 1 keep lower (row >= col), 2 keep upper (row <= col)
    int herk;                    // complex herk: force the diagonal of C real (generic kernel only)
};
using GemmParamsD = GemmParamsT<double>;
constexpr int SKINNY_MAX_N = 16;       // widest right operand served by the HBM-bound streaming kernel (gemm_skinny.cu)

// Type-generic launcher: double -> DMMA kernel below; float / complex -> gemm_generic.cu.
// opA / opB in 'N','T','C' refer to the column-major problem C = alpha op(A) op(B) + beta C.
template <typename T> int launch_gemm(int opA, int opB, GemmParamsT<T> p, cudaStream_t stream);

// Tile configuration.  CTA tile BM x BN x 16; consumer warps own WM x WN warp tiles
// (WM/8 x WN/8 DMMA accumulator pairs); consumer warps fill whole warpgroups so that
// setmaxnreg can move registers from the producer warpgroup to them.
template <int BM_, int BN_, int WM_, int WN_, bool DBUF_>
struct GemmCfg {
    static constexpr int BM = BM_, BN = BN_, BK = 16, WM = WM_, WN = WN_;
    static constexpr bool DBUF = DBUF_;                 // double-buffer fragments in registers
    static constexpr int MI = WM / 8, NJ = WN / 8;
    static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
    static constexpr int CONSUMER_WARPS = WARPS_M * WARPS_N;
    static_assert(CONSUMER_WARPS % 4 == 0, "consumer warps must fill warpgroups");
    static constexpr int THREADS = CONSUMER_WARPS * 32 + 128;   // + producer warpgroup (1 active warp)
    static constexpr int CTAS_PER_SM = 2;
    static constexpr int PRODUCER_REGS = 40;
    // The CTA's register pool is what the launch allocates (THREADS x LAUNCH_REGS); consumers may
    // only grow into what the producer warpgroup gives back -- asking for more deadlocks in
    // setmaxnreg.inc.
    static constexpr int LAUNCH_REGS = (65536 / (CTAS_PER_SM * THREADS)) / 8 * 8;
    static constexpr int CONSUMER_REGS =
        ((THREADS * LAUNCH_REGS - 128 * PRODUCER_REGS) / (CONSUMER_WARPS * 32)) / 8 * 8;
    static_assert(CONSUMER_WARPS * 32 * CONSUMER_REGS + 128 * PRODUCER_REGS <= THREADS * LAUNCH_REGS,
                  "setmaxnreg budget exceeds the CTA register pool");
    static constexpr int PAD = 4;
    static constexpr int LDK = BK + PAD;                        // K-major stride (doubles)
    static constexpr int LDA_MN = BM + PAD, LDB_MN = BN + PAD;  // MN-major strides
    template <bool AK> static constexpr int a_stage() { return AK ? BM * LDK : BK * LDA_MN; }
    template <bool BKM> static constexpr int b_stage() { return BKM ? BN * LDK : BK * LDB_MN; }
    template <bool AK> static constexpr int stages() { return AK ? 3 : 4; }
    template <bool AK, bool BKM> static constexpr size_t smem_bytes()
    {
        return size_t(stages<AK>()) * (a_stage<AK>() + b_stage<BKM>()) * sizeof(double)
               + 2 * stages<AK>() * sizeof(uint64_t);
    }
};
using GemmCfgDefault = GemmCfg<128, 64, 32, 32, false>;   // 8 consumer warps/CTA, 4 per SM sub-partition

// A_KMAJ: op(A)(i,l) = A[l + i*lda] (k contiguous; op(A) = T in column-major terms)
// B_KMAJ: op(B)(l,j) = B[l + j*ldb] (k contiguous; op(B) = N)
template <typename Cfg, bool A_KMAJ, bool B_KMAJ>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::CTAS_PER_SM)
gemm_dmma_kernel(const GemmParamsD p)
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
    const int t  = blockIdx.x / per_problem;
    const int r  = blockIdx.x - t * per_problem;
    const int m0 = (r % tiles_m) * BM;
    const int n0 = (r / tiles_m) * BN;

    // triangle-masked problems: skip CTAs that lie entirely in the discarded triangle
    if (p.tri == 1 && n0 >= m0 + BM) return;
    if (p.tri == 2 && m0 >= n0 + BN) return;

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

    const double* __restrict__ A = (p.A ? p.A[t] : p.A0 + int64_t(t) * p.strideA) + p.offA;
    const double* __restrict__ B = (p.B ? p.B[t] : p.B0 + int64_t(t) * p.strideB) + p.offB;
    const int mv = min(BM, p.m - m0);
    const int nv = min(BN, p.n - n0);
    const int num_kt = (p.k + BK - 1) / BK;

    if (warp >= CONSUMER_WARPS) {
        // ===================== producer warpgroup =====================
        setmaxnreg_dec<Cfg::PRODUCER_REGS>();
        if (warp != CONSUMER_WARPS) return;
        // MN-major operands: one bulk copy (UBLKCP) per k column.  K-major operands: 16-byte
        // cp.async (LDGSTS) chunks -- 128-byte bulk copies were measured 3x slower.  Every lane
        // arrives once per stage on full_bar (count 32); bulk bytes are added with expect_tx.
        const bool aligned =
            ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0
            && ((p.lda | p.ldb | mv | nv) & 1) == 0 && (p.k & 3) == 0;
        for (int kt = 0; kt < num_kt; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
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
        return;
    }

    // ===================== consumer warps =====================
    setmaxnreg_inc<Cfg::CONSUMER_REGS>();
    const int wm = (warp % Cfg::WARPS_M) * WM;
    const int wn = (warp / Cfg::WARPS_M) * WN;
    const int lr = lane >> 2;      // 0..7  fragment row (A) / column (B)
    const int lc = lane & 3;       // 0..3  fragment k index

    double acc[MI][NJ][2];
    #pragma unroll
    for (int i = 0; i < MI; ++i)
        #pragma unroll
        for (int j = 0; j < NJ; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }

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

    if (Cfg::DBUF) {
        // Fragments are double-buffered in registers and the first slice of the NEXT stage is
        // fetched before the last DMMAs of the current one, so one warp alone never drains the
        // tensor pipe at a stage boundary (used when only 2 warps share an SM sub-partition).
        if (num_kt > 0) {
            double a0[MI], b0[NJ], a1[MI], b1[NJ];
            const int total_k4 = (p.k + 3) >> 2;
            mbar_wait(&full_bar[0], 0);
            load_frag(a0, b0, 0, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int kt = 0; kt < num_kt; ++kt) {
                const int nk4 = min(BK / 4, total_k4 - kt * (BK / 4));
                const int sn = (s + 1 == STAGES) ? 0 : s + 1;
                const uint32_t phn = (sn == 0) ? (ph ^ 1) : ph;
                if (nk4 == BK / 4) {
                    load_frag(a1, b1, s, 1);  mma_all(a0, b0);
                    load_frag(a0, b0, s, 2);  mma_all(a1, b1);
                    load_frag(a1, b1, s, 3);  mma_all(a0, b0);
                    if (kt + 1 < num_kt) {
                        mbar_wait(&full_bar[sn], phn);
                        load_frag(a0, b0, sn, 0);
                    }
                    mma_all(a1, b1);
                }
                else {
                    for (int k4 = 0; k4 < nk4; ++k4) {       // ragged last stage
                        if (k4 > 0) load_frag(a0, b0, s, k4);
                        mma_all(a0, b0);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                s = sn;
                ph = phn;
            }
        }
    }
    else {
        // 4 warps per SM sub-partition hide each other's shared-memory latency: keep the
        // register footprint small instead of double-buffering.
        const int total_k4 = (p.k + 3) >> 2;
        for (int kt = 0; kt < num_kt; ++kt) {
            const int s = kt % STAGES;
            const uint32_t ph = (kt / STAGES) & 1;
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
}

// Host-side launcher: column-major problem  C = alpha op(A) op(B) + beta C  (opA/opB in 'N','T','C';
// 'C' == 'T' for real).  Returns SB200 status.
int launch_gemm_d(int opA, int opB, GemmParamsD p, cudaStream_t stream);

} // namespace sb200
