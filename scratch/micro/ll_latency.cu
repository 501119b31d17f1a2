// One-way latency of a tagged 8-byte word between two CTAs on different SMs of a B200, per store / load flavour.
// CTA 0 (thread 0) and CTA k ping-pong ITERS times; reported: cycles per one-way hop (round trip / 2), for several partners k.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ll_latency ll_latency.cu ; run: ./ll_latency
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2000;
template <int ST> __device__ __forceinline__ void put(unsigned long long* p, unsigned long long v)
{
    if (ST == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (ST == 1) asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (ST == 2) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); __threadfence(); }
    if (ST == 3) asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (ST == 4) { unsigned long long o; asm volatile("atom.relaxed.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(o) : "l"(p), "l"(v) : "memory"); }
    if (ST == 5) asm volatile("st.volatile.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (ST == 6) asm volatile("st.global.cg.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
    if (ST == 7) asm volatile("st.global.wt.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
template <int LD> __device__ __forceinline__ unsigned long long get(const unsigned long long* p)
{
    unsigned long long v;
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 1) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 2) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 3) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (LD == 4) { asm volatile("atom.relaxed.gpu.global.add.u64 %0, [%1], 0;" : "=l"(v) : "l"(p) : "memory"); }
    return v;
}
template <int ST, int LD>
__global__ void pingpong(unsigned long long* a, unsigned long long* b, int partner, long long* out, unsigned base)
{
    extern __shared__ char pad[];
    if (threadIdx.x != 0) return;
    if (blockIdx.x == 0) {
        const long long t0 = clock64();
        for (unsigned i = 1; i <= ITERS; ++i) {
            put<ST>(a, (unsigned long long)(base + i) << 32);
            while ((get<LD>(b) >> 32) != base + i) { }
        }
        *out = (clock64() - t0) / (2 * ITERS);
    }
    else if (int(blockIdx.x) == partner) {
        for (unsigned i = 1; i <= ITERS; ++i) {
            while ((get<LD>(a) >> 32) != base + i) { }
            put<ST>(b, (unsigned long long)(base + i) << 32);
        }
    }
}
// self visibility: store then poll own word (what thread 0 of CTA 0 sees in the LU kernel)
template <int ST, int LD>
__global__ void selfvis(unsigned long long* a, long long* out, unsigned base)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long t0 = clock64();
    for (unsigned i = 1; i <= ITERS; ++i) {
        put<ST>(a, (unsigned long long)(base + i) << 32);
        while ((get<LD>(a) >> 32) != base + i) { }
    }
    *out = (clock64() - t0) / ITERS;
}
unsigned g_base = 0;
template <int ST, int LD> void run(const char* name, unsigned long long* buf, long long* dout)
{
    cudaFuncSetAttribute(pingpong<ST, LD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("%-34s", name);
    for (int partner : {1, 2, 37, 73, 74, 110, 147}) {
        void* args[] = {nullptr};
        (void) args;
        pingpong<ST, LD><<<148, 32, 200 * 1024>>>(buf, buf + 64, partner, dout, g_base);
        g_base += ITERS + 8;
        long long h = -1;
        cudaError_t e = cudaMemcpy(&h, dout, sizeof(h), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf(" err %s", cudaGetErrorString(e)); break; }
        printf(" %6lld", h);
    }
    selfvis<ST, LD><<<1, 32>>>(buf + 128, dout, g_base); g_base += ITERS + 8;
    long long h = -1; cudaMemcpy(&h, dout, sizeof(h), cudaMemcpyDeviceToHost);
    printf("   self %6lld\n", h);
}
int main()
{
    unsigned long long* buf; long long* dout;
    cudaMalloc(&buf, 4096); cudaMemset(buf, 0, 4096); cudaMalloc(&dout, 8);
    printf("one-way hop, cycles (partner CTA = 1, 2, 37, 73, 74, 110, 147 of 148, one CTA per SM); self = store->own poll\n");
    run<0, 0>("st.relaxed.gpu / ld.relaxed.gpu", buf, dout);
    run<1, 1>("st.release.gpu / ld.acquire.gpu", buf, dout);
    run<2, 0>("st.relaxed+fence / ld.relaxed", buf, dout);
    run<3, 0>("red.max.u64 / ld.relaxed", buf, dout);
    run<4, 0>("atom.exch / ld.relaxed", buf, dout);
    run<5, 2>("st.volatile / ld.volatile", buf, dout);
    run<6, 3>("st.cg / ld.cv", buf, dout);
    run<7, 3>("st.wt / ld.cv", buf, dout);
    run<3, 4>("red.max / atom.add 0", buf, dout);
    run<0, 4>("st.relaxed / atom.add 0", buf, dout);
    return 0;
}
