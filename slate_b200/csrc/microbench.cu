// microbench.cu -- FP64 pipe peak probes (DMMA.8x8x4 and DFMA) used as the measured
// roofline denominator for the FP64 contractions (MEASURED_PEAKS.json has no FP64 figure).
#include "common.cuh"

namespace sb200 {

// each warp runs `iters` x 32 independent DMMA.8x8x4 (512 flop each)
__global__ void __launch_bounds__(256, 2) dmma_peak_kernel(double* out, int iters)
{
    double acc[32][2];
    #pragma unroll
    for (int i = 0; i < 32; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < 32; ++i) dmma884(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[0] = s;
}

// each thread runs `iters` x 16 independent DFMA (2 flop each)
__global__ void __launch_bounds__(256, 2) dfma_peak_kernel(double* out, int iters)
{
    double acc[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; ++it) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
    #pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

} // namespace sb200

extern "C" {
// Launch one peak probe on `stream`; returns the flop count of the launch in *flops.
// kind 0 = DMMA.8x8x4, 1 = DFMA.  ctas_per_sm x SM-count CTAs of 256 threads.
int sb200_fp64_peak_probe(int kind, int iters, int ctas_per_sm, double* d_scratch, double* flops,
                          sb200_stream_t stream)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return SB200_ENODEV;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * ctas_per_sm;
    if (kind == 0) {
        sb200::dmma_peak_kernel<<<grid, 256, 0, cudaStream_t(stream)>>>(d_scratch, iters);
        *flops = double(grid) * 8 * double(iters) * 32 * 512.0;
    }
    else {
        sb200::dfma_peak_kernel<<<grid, 256, 0, cudaStream_t(stream)>>>(d_scratch, iters);
        *flops = double(grid) * 256 * double(iters) * 16 * 2.0;
    }
    return sb200::launch_status();
}
}
