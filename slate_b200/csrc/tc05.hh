// tc05.hh -- interface of the tcgen05 FP32-emulated (3 x TF32) batched tile GEMM (gemm_tc05.cu).
#pragma once
#include "common.cuh"
#include <algorithm>

namespace sb200 {

constexpr int TC_BM = 128, TC_BN = 256, TC_KC = 16;     // CTA block of C and k per pipeline stage

// C_t = alpha * A_t * B_t + beta * C_t with A_t, B_t PACKED (see gemm_tc05.cu): A_t is the packed
// m x k operand (units of 128 rows), B_t the packed n x k operand (= (k x n)^T, units of 256 rows).
struct Tc05Params {
    const void* const* Ap;      // per-tile packed operands (device pointer arrays), or null -> Ap0 / Bp0 / C0
    const void* const* Bp;
    float* const*      C;
    const void* Ap0; const void* Bp0; float* C0;
    int64_t offC;
    int m, n, k, ldc;
    float alpha, beta;
    int batch;
    int tri;                    // 0 full, 1 keep lower (row >= col): herk/syrk diagonal tiles of the Cholesky update
};

// pack `batch` operands: element (r, kk) of operand t = X_t[r * rs + kk * ks], r < rows, kk < k
struct Tc05PackParams {
    const float* const* X; const float* X0; int64_t strideX, offX;
    void* const* P; void* P0; int64_t strideP;       // packed destinations (bytes)
    int rows, k;
    int64_t rs, ks;
    int ru;                     // rows per unit: TC_BM (A side) or TC_BN (B side)
    int batch;
};

size_t tc05_packed_bytes(int side, int64_t rows, int64_t k);
int launch_tc05_pack(Tc05PackParams p, cudaStream_t stream);
int launch_tc05_gemm(Tc05Params p, cudaStream_t stream);

} // namespace sb200
