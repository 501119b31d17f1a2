// Floor of the per-column exchange of the LU base kernel: G co-resident CTAs (one per SM), every iteration each CTA
// publishes one tagged 16-byte header and polls the headers of all G CTAs (thread c polls CTA c), block barrier, repeat.
// Variants: POLLW = 0 thread c of the CTA polls record c (as getrf_base_v3_kernel), 1 = one warp polls everything
// (lanes stride over records, all loads in flight), 2 = as 0 but idle warps wait at the barrier instead of falling through.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o allgather_latency allgather_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
constexpr int ITERS = 512, THREADS = 256, REC = 72;
__device__ __forceinline__ void st2(unsigned long long* p, unsigned long long a, unsigned long long b)
{ asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ void ld2(const unsigned long long* p, unsigned long long& a, unsigned long long& b)
{ asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
template <int POLLW>
__global__ void __launch_bounds__(THREADS) allgather(unsigned long long* rec, int gmax, long long* out, unsigned base, int extra_work)
{
    extern __shared__ char pad[];
    __shared__ unsigned s_acc[8];
    const int G = gridDim.x, b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long t0 = clock64();
    unsigned acc = 0;
    for (int it = 0; it < ITERS; ++it) {
        const unsigned gen = base + it + 1;
        const unsigned long long g = (unsigned long long)gen << 32;
        unsigned long long* slot = rec + size_t(it % 32) * gmax * REC;
        if (extra_work) { for (int k = 0; k < extra_work; ++k) acc = acc * 1664525u + 1013904223u; }   // stand-in for the update
        __syncthreads();
        if (tid == 0) st2(slot + size_t(b) * REC, g | (acc & 0xffff), g | b);
        if (POLLW == 1) {
            if (warp == 0)
                for (int c = lane; c < G; c += 32) {
                    unsigned long long w0, w1;
                    do { ld2(slot + size_t(c) * REC, w0, w1); } while (unsigned(w0 >> 32) != gen || unsigned(w1 >> 32) != gen);
                    acc += unsigned(w0);
                }
        }
        else {
            for (int c = tid; c < G; c += THREADS) {
                unsigned long long w0, w1;
                do { ld2(slot + size_t(c) * REC, w0, w1); } while (unsigned(w0 >> 32) != gen || unsigned(w1 >> 32) != gen);
                acc += unsigned(w0);
            }
        }
        if (lane == 0) s_acc[warp] = acc;
        __syncthreads();
        acc += s_acc[(warp + 1) & 7];
    }
    if (b == 0 && tid == 0) { out[0] = (clock64() - t0) / ITERS; out[1] = acc; }
}
int main()
{
    unsigned long long* rec; long long* dout;
    const int gmax = 148;
    const size_t bytes = size_t(32) * gmax * REC * 8;
    cudaMalloc(&rec, bytes); cudaMemset(rec, 0, bytes); cudaMalloc(&dout, 16);
    cudaFuncSetAttribute(allgather<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024);
    cudaFuncSetAttribute(allgather<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024);
    unsigned base = 0;
    printf("cycles per all-gather iteration (publish 16 B, poll all G headers, block barrier); one CTA per SM\n");
    for (int extra : {0, 400})
    for (int pw = 0; pw < 2; ++pw)
        for (int G : {2, 8, 22, 44, 86, 144}) {
            int ew = extra;
            void* args[] = {&rec, (void*)&gmax, &dout, &base, &ew};
            cudaError_t e = cudaLaunchCooperativeKernel(pw == 0 ? (void*)allgather<0> : (void*)allgather<1>, dim3(G), dim3(THREADS), args, 190 * 1024, 0);
            base += ITERS + 8;
            long long h[2] = {-1, -1};
            if (e == cudaSuccess) e = cudaMemcpy(h, dout, 16, cudaMemcpyDeviceToHost);
            printf("extra_work=%3d  poll=%s  G=%3d  cycles/iter=%6lld %s\n", extra, pw ? "one warp       " : "thread per rec ", G, h[0], e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
