// gemm_tc05.cu -- FP32-emulated batched tile GEMM on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM): the trailing-matrix update of the LOW-PRECISION
// factorisation inside gesv_mixed / posv_mixed.
//
// Reference: src/gesv_mixed.cc:106-300 factors a float copy of A with getrf<float>; on devices its
// trailing update is internal::gemm<Devices, float> -> cublasSgemmBatched (FP32 SIMT FMA,
// src/internal/internal_gemm.cc:354-518).
//
// B200-first restatement:
//   * FP32 operands are split ONCE per panel into two TF32 planes, x = hi + lo
//     (hi = round-to-nearest TF32 of x, lo = x - hi exactly representable in FP32; the tensor core
//     reads the upper 19 bits of lo), and each product is formed as
//         A*B ~= hi(A)*hi(B) + hi(A)*lo(B) + lo(A)*hi(B)          (the lo*lo term is < 2^-22 |a||b|)
//     i.e. three tcgen05.mma.kind::tf32 instructions accumulating into the same FP32 TMEM tile.
//   * The split operands are stored ("packed") in HBM directly in the tensor core's canonical
//     K-major shared-memory layout (8 x 16-byte core matrices, no swizzle), cut into units of
//     (128 | 256 rows) x 16 k with the hi plane followed by the lo plane.  The GEMM kernel therefore
//     stages a k-step with TWO 1-D bulk copies through the TMA engine (cp.async.bulk, SASS UBLKCP)
//     and never touches an operand with a thread: the panel is packed once (O(n nb) bytes) and read
//     by O(n^2 / nb^2) tile updates.
//   * CTA = 128 x 256 block of one C tile, 6 warps: warp 0 TMEM allocation + TMA producer (one
//     elected lane), warp 1 MMA issuer (one elected lane, tcgen05.commit onto the stage's "empty"
//     mbarrier), warps 2-5 epilogue (tcgen05.ld 32x32b, alpha/beta, coalesced C read-modify-write).
//     2 CTAs per SM (2 x 256 TMEM columns, 2 x 96 KiB shared memory) so that one CTA's epilogue
//     overlaps the other's main loop.
#include "gemm_dmma.cuh"
#include "tc05.hh"
#include <cstdio>

namespace sb200 {

constexpr int TC_THREADS = 192;
constexpr int TC_STAGES = 2;
constexpr uint32_t TC_A_PLANE = TC_BM * TC_KC * 4;          //  8 KiB: 128 rows x 16 k, TF32 in 4-byte containers
constexpr uint32_t TC_B_PLANE = TC_BN * TC_KC * 4;          // 16 KiB
constexpr uint32_t TC_A_UNIT = 2 * TC_A_PLANE;              // hi | lo
constexpr uint32_t TC_B_UNIT = 2 * TC_B_PLANE;
constexpr uint32_t TC_STAGE_BYTES = TC_A_UNIT + TC_B_UNIT;  // 48 KiB
constexpr uint32_t TC_TMEM_COLS = TC_BN;                    // 128 lanes x 256 FP32 columns
constexpr size_t   TC_SMEM = size_t(TC_STAGES) * TC_STAGE_BYTES + 256;

// canonical K-major / no-swizzle layout of one plane (rows x 16 k):
//   byte offset(r, kk) = (r / 8) * 512 + (kk / 4) * 128 + (r % 8) * 16 + (kk % 4) * 4
// -> leading (K-direction) core-matrix stride 128 B, stride (row-group) offset 512 B.
constexpr uint32_t TC_LBO = 128, TC_SBO = 512;

// ---------------------------------------------------------------------------- tcgen05 wrappers
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr)
{
    // cute::UMMA::SmemDescriptor: [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4,
    // [46,48) version = 1 (Blackwell), [61,64) layout type 0 = no swizzle
    return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(TC_LBO >> 4) << 16) | (uint64_t(TC_SBO >> 4) << 32)
         | (uint64_t(1) << 46);
}

// cute::UMMA::InstrDescriptor for kind::tf32, FP32 accumulate, both operands K-major
constexpr uint32_t umma_idesc_tf32(int M, int N)
{
    return (1u << 4)                 // c_format  = F32
         | (2u << 7) | (2u << 10)    // a_format, b_format = TF32
         | (0u << 15) | (0u << 16)   // a_major, b_major = K
         | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// mbarrier arrives once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------- the GEMM kernel
__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tf32x3_kernel(const Tc05Params p)
{
    extern __shared__ __align__(128) unsigned char tc_smem[];
    unsigned char* stage_base = tc_smem;
    uint64_t* full  = reinterpret_cast<uint64_t*>(tc_smem + size_t(TC_STAGES) * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* accbar = empty + TC_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int MB = (p.m + TC_BM - 1) / TC_BM, NB = (p.n + TC_BN - 1) / TC_BN;
    const int KC = (p.k + TC_KC - 1) / TC_KC;
    int b = blockIdx.x;
    const int mb = b % MB; b /= MB;
    const int nbk = b % NB;
    const int t = b / NB;
    // triangle-masked tile: a block that lies wholly above the diagonal has nothing to store
    if (p.tri == 1 && mb * TC_BM + TC_BM - 1 < nbk * TC_BN) return;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 32) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(accbar, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: two bulk copies per k-step
            const unsigned char* Ap = static_cast<const unsigned char*>(p.Ap ? p.Ap[t] : p.Ap0)
                                    + size_t(mb) * KC * TC_A_UNIT;
            const unsigned char* Bp = static_cast<const unsigned char*>(p.Bp ? p.Bp[t] : p.Bp0)
                                    + size_t(nbk) * KC * TC_B_UNIT;
            for (int kc = 0; kc < KC; ++kc) {
                const int s = kc % TC_STAGES;
                mbar_wait(&empty[s], ((kc / TC_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
                unsigned char* dst = stage_base + size_t(s) * TC_STAGE_BYTES;
                bulk_g2s(dst, Ap + size_t(kc) * TC_A_UNIT, TC_A_UNIT, &full[s]);
                bulk_g2s(dst + TC_A_UNIT, Bp + size_t(kc) * TC_B_UNIT, TC_B_UNIT, &full[s]);
            }
        }
        __syncwarp();
    }
    else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: per k-step 2 (K = 8 each) x 3 (hi*hi, hi*lo, lo*hi) instructions
            constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, TC_BN);
            uint32_t acc = 0;
            for (int kc = 0; kc < KC; ++kc) {
                const int s = kc % TC_STAGES;
                mbar_wait(&full[s], (kc / TC_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(stage_base + size_t(s) * TC_STAGE_BYTES);
                const uint32_t a_lo = a_hi + TC_A_PLANE;
                const uint32_t b_hi = a_hi + TC_A_UNIT;
                const uint32_t b_lo = b_hi + TC_B_PLANE;
                #pragma unroll
                for (int ks = 0; ks < TC_KC / 8; ++ks) {
                    const uint32_t ko = uint32_t(ks) * 2 * TC_LBO;       // two 16-byte K chunks per MMA
                    umma_tf32(tmem, umma_smem_desc(a_lo + ko), umma_smem_desc(b_hi + ko), idesc, acc);
                    acc = 1;
                    umma_tf32(tmem, umma_smem_desc(a_hi + ko), umma_smem_desc(b_lo + ko), idesc, 1);
                    umma_tf32(tmem, umma_smem_desc(a_hi + ko), umma_smem_desc(b_hi + ko), idesc, 1);
                }
                umma_commit(&empty[s]);          // frees the stage when these MMAs have read it
            }
            umma_commit(accbar);                 // accumulator complete
        }
        __syncwarp();
    }
    else {
        // ===== epilogue: warp q = warp % 4 owns TMEM lanes [32q, 32q + 32) = block rows
        const int q = warp & 3;
        const int row = mb * TC_BM + q * 32 + lane;
        const int n0 = nbk * TC_BN;
        float* __restrict__ C = (p.C ? p.C[t] : p.C0) + p.offC;
        const bool use_beta = (p.beta != 0.0f);
        mbar_wait(accbar, 0);
        tc_fence_after();
        #pragma unroll 1
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
            if (n0 + c0 >= p.n) break;                           // warp-uniform
            uint32_t v[32];
            tmem_ld32(tmem + (uint32_t(q * 32) << 16) + uint32_t(c0), v);
            float cin[32];
            #pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = n0 + c0 + j;
                cin[j] = (use_beta && row < p.m && col < p.n) ? C[row + int64_t(col) * p.ldc] : 0.0f;
            }
            #pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = n0 + c0 + j;
                if (row < p.m && col < p.n && (p.tri == 0 || row >= col)) {
                    float r = p.alpha * __uint_as_float(v[j]);
                    if (use_beta) r = fmaf(p.beta, cin[j], r);
                    C[row + int64_t(col) * p.ldc] = r;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TC_TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------- operand packing
// element (r, kk) of the operand = X[r * rs + kk * ks]; one thread packs 4 consecutive k of one row
__global__ void __launch_bounds__(256)
pack_tf32x3_kernel(const Tc05PackParams p)
{
    const int t = blockIdx.y;
    const float* __restrict__ X = (p.X ? p.X[t] : p.X0 + int64_t(t) * p.strideX) + p.offX;
    unsigned char* __restrict__ P = static_cast<unsigned char*>(p.P ? p.P[t] : static_cast<void*>(
                                        static_cast<unsigned char*>(p.P0) + int64_t(t) * p.strideP));
    const int ru = p.ru;
    const int RU = (p.rows + ru - 1) / ru, KC = (p.k + TC_KC - 1) / TC_KC;
    const int Rpad = RU * ru, cores = KC * (TC_KC / 4);
    const int64_t e = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (e >= int64_t(Rpad) * cores) return;
    int r, c;
    if (p.rs == 1) { r = int(e % Rpad); c = int(e / Rpad); }          // rows contiguous in memory
    else           { c = int(e % cores); r = int(e / cores); }       // k contiguous in memory
    float hi[4], lo[4];
    #pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int kk = 4 * c + i;
        const float x = (r < p.rows && kk < p.k) ? X[int64_t(r) * p.rs + int64_t(kk) * p.ks] : 0.0f;
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
        hi[i] = __uint_as_float(h);
        const float l = x - hi[i];
        lo[i] = (fabsf(x) <= 3.0e38f) ? l : 0.0f;                    // inf / nan stay in the hi plane only
    }
    const int u = r / ru, rr = r % ru, kc = c / (TC_KC / 4), cc = c % (TC_KC / 4);
    const size_t plane = size_t(ru) * TC_KC * 4;
    unsigned char* unit = P + (size_t(u) * KC + kc) * (2 * plane);
    const size_t off = size_t(rr / 8) * TC_SBO + size_t(cc) * TC_LBO + size_t(rr % 8) * 16;
    *reinterpret_cast<float4*>(unit + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(unit + plane + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

size_t tc05_packed_bytes(int side, int64_t rows, int64_t k)
{
    const int64_t ru = (side == 'A') ? TC_BM : TC_BN;
    return size_t(ceil_div(rows, ru) * ru) * size_t(ceil_div(k, TC_KC) * TC_KC) * 4 * 2;
}

int launch_tc05_pack(Tc05PackParams p, cudaStream_t stream)
{
    if (p.rows <= 0 || p.k <= 0 || p.batch <= 0) return SB200_OK;
    const int64_t Rpad = ceil_div(p.rows, p.ru) * p.ru, cores = ceil_div(p.k, TC_KC) * (TC_KC / 4);
    const int64_t blocks = ceil_div(Rpad * cores, 256);
    for (int b0 = 0; b0 < p.batch; b0 += 32768) {
        Tc05PackParams q = p;
        const int cnt = std::min(32768, p.batch - b0);
        if (q.X) q.X += b0; else q.X0 += int64_t(b0) * q.strideX;
        if (q.P) q.P += b0; else q.P0 = static_cast<unsigned char*>(q.P0) + int64_t(b0) * q.strideP;
        pack_tf32x3_kernel<<<dim3(unsigned(blocks), unsigned(cnt)), 256, 0, stream>>>(q);
        const int st = launch_status();
        if (st) return st;
    }
    return SB200_OK;
}

int launch_tc05_gemm(Tc05Params p, cudaStream_t stream)
{
    if (p.m <= 0 || p.n <= 0 || p.batch <= 0) return SB200_OK;
    if (p.k <= 0) return SB200_EINVAL;          // callers scale C themselves when k == 0
    static thread_local bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (! attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(TC_SMEM));
        if (e != cudaSuccess) return int(e);
        attr_done[dev & 63] = true;
    }
    const int64_t grid = ceil_div(p.m, TC_BM) * ceil_div(p.n, TC_BN) * int64_t(p.batch);
    if (grid > 0x7fffffffLL) return SB200_EINVAL;
    gemm_tf32x3_kernel<<<unsigned(grid), TC_THREADS, TC_SMEM, stream>>>(p);
    return launch_status();
}

} // namespace sb200

using namespace sb200;

extern "C" {

size_t sb200_tf32x3_packed_bytes(int side, int64_t rows, int64_t k)
{
    if ((side != 'A' && side != 'B') || rows < 0 || k < 0) return 0;
    return tc05_packed_bytes(side, rows, k);
}

int sb200_tf32x3_pack_batched_s(int side, int op, int64_t rows, int64_t k,
                                const float* const* dX, int64_t ldx, void* const* dPacked,
                                int64_t batch, sb200_stream_t stream)
{
    if ((side != 'A' && side != 'B') || ! valid_op(op) || rows < 0 || k < 0 || batch < 0) return SB200_EINVAL;
    if (rows == 0 || k == 0 || batch == 0) return SB200_OK;
    if (! dX || ! dPacked || ldx < 1 || rows > 0x7fffffff || k > 0x7fffffff || batch > 0x7fffffff) return SB200_EINVAL;
    // side A: operand rows = rows of op(X) (m index);  side B: operand rows = columns of op(X) (n index)
    const bool rows_contig = (side == 'A') == (op == 'N');
    Tc05PackParams p{};
    p.X = dX; p.P = dPacked; p.rows = int(rows); p.k = int(k);
    p.rs = rows_contig ? 1 : ldx; p.ks = rows_contig ? ldx : 1;
    p.ru = (side == 'A') ? TC_BM : TC_BN; p.batch = int(batch);
    return launch_tc05_pack(p, cudaStream_t(stream));
}

int sb200_gemm_tf32x3_packed_s(int64_t m, int64_t n, int64_t k, float alpha,
                               const void* const* dApacked, const void* const* dBpacked,
                               float beta, float* const* dC, int64_t ldc,
                               int64_t batch, sb200_stream_t stream)
{
    if (m < 0 || n < 0 || k < 1 || batch < 0) return SB200_EINVAL;
    if (m == 0 || n == 0 || batch == 0) return SB200_OK;
    if (! dApacked || ! dBpacked || ! dC || ldc < m || m > 0x7fffffff || n > 0x7fffffff || k > 0x7fffffff
        || batch > 0x7fffffff) return SB200_EINVAL;
    Tc05Params p{};
    p.Ap = dApacked; p.Bp = dBpacked; p.C = dC;
    p.m = int(m); p.n = int(n); p.k = int(k); p.ldc = int(ldc); p.alpha = alpha; p.beta = beta; p.batch = int(batch);
    return launch_tc05_gemm(p, cudaStream_t(stream));
}

} // extern "C"
