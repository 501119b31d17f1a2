// sm_partition.cu -- whole SMs reserved for the panel chain of a factorisation (CUDA green contexts).
//
// Why (profiles/r02b_chain_kernels_contended_vs_idle.jsonl, profiles/r02d_greenctx_chain_partition.jsonl): the
// diagonal-tile Cholesky of a potrf step (reference: internal::potrf<Devices>, src/internal/internal_potrf.cc:57-81)
// takes 0.57 ms on an idle B200 and 2.48 ms next to the trailing update -- stream priorities only order CTAs that
// are still pending, they do not give a 128-thread chain CTA its issue slots or its L1 next to two resident
// trailing-update CTAs.  At 8 GPUs that chain (128 tiles at n = 65536) is longer than the trailing update.
// Splitting the 148 SMs into {chain: a few SMs} + {everything else} makes the tile 0.69 ms whatever the rest of the
// device does; the trailing update loses exactly the reserved SMs' share of the DMMA rate.
//
// The driver API symbols are fetched with cudaGetDriverEntryPoint, so that the library still links against
// -lcudart -lnccl only and loads on a box without a driver (the CPU test tier).
#include "runtime_internal.hh"
#include <cuda.h>
#include <mutex>

namespace sb200 {

namespace {

struct DriverApi {
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*,
                                          unsigned int, unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool ok = false;
};

template <typename F>
bool entry(const char* name, F& fn)
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || ! p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const DriverApi& driver_api()
{
    static DriverApi api = [] {
        DriverApi a;
        a.ok = entry("cuDeviceGet", a.DeviceGet)
            && entry("cuDeviceGetDevResource", a.DeviceGetDevResource)
            && entry("cuDevSmResourceSplitByCount", a.DevSmResourceSplitByCount)
            && entry("cuDevResourceGenerateDesc", a.DevResourceGenerateDesc)
            && entry("cuGreenCtxCreate", a.GreenCtxCreate)
            && entry("cuGreenCtxStreamCreate", a.GreenCtxStreamCreate);
        return a;
    }();
    return api;
}

struct Partition {
    int device = -1, want = 0;
    CUgreenCtx small = nullptr, big = nullptr;
    int sm_small = 0, sm_big = 0;
    bool ok = false;
};

// one partition per (device, size) for the life of the process: green contexts are expensive to create and the
// drivers are called in a loop
Partition* partition_for(int sms)
{
    static std::mutex mu;
    static std::vector<Partition*> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    for (Partition* p : cache)
        if (p->device == dev && p->want == sms) return p->ok ? p : nullptr;
    Partition* p = new Partition;
    p->device = dev; p->want = sms;
    cache.push_back(p);
    const DriverApi& d = driver_api();
    if (! d.ok) return nullptr;
    cudaFree(nullptr);                                   // the primary context must exist
    CUdevice cudev;
    CUdevResource all, grp, rest;
    if (d.DeviceGet(&cudev, dev) != CUDA_SUCCESS) return nullptr;
    if (d.DeviceGetDevResource(cudev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
    unsigned int ngroups = 1;
    // below the co-scheduling granularity (8 SMs on sm_100) the split needs the IGNORE flag; the chain kernels use
    // neither clusters nor cooperative launches, so they do not care
    const unsigned int flags = sms < 8 ? CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING : 0;
    if (d.DevSmResourceSplitByCount(&grp, &ngroups, &all, &rest, flags, unsigned(sms)) != CUDA_SUCCESS || ngroups < 1)
        return nullptr;
    CUdevResourceDesc ds, db;
    if (d.DevResourceGenerateDesc(&ds, &grp, 1) != CUDA_SUCCESS) return nullptr;
    if (d.DevResourceGenerateDesc(&db, &rest, 1) != CUDA_SUCCESS) return nullptr;
    if (d.GreenCtxCreate(&p->small, ds, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
    if (d.GreenCtxCreate(&p->big, db, cudev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return nullptr;
    p->sm_small = int(grp.sm.smCount);
    p->sm_big = int(rest.sm.smCount);
    p->ok = true;
    return p;
}

} // namespace

// Streams of one driver call.  chain_sms > 0: `chain` lives on a partition of that many SMs and panel / look / trail
// on the complementary partition (a stream of the primary context could be scheduled on the reserved SMs too);
// on any failure (old driver, MPS, ...) the call falls back to plain priority streams and chain == panel.
int Streams::init(size_t nevents, int chain_sms)
{
    int lo, hi;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    const int mid = hi < lo - 1 ? hi + 1 : hi;
    Partition* part = chain_sms > 0 ? partition_for(chain_sms) : nullptr;
    if (part) {
        const DriverApi& d = driver_api();
        CUstream c = nullptr, p = nullptr, l = nullptr, t = nullptr;
        const bool ok = d.GreenCtxStreamCreate(&c, part->small, CU_STREAM_NON_BLOCKING, hi) == CUDA_SUCCESS
                     && d.GreenCtxStreamCreate(&p, part->big, CU_STREAM_NON_BLOCKING, hi) == CUDA_SUCCESS
                     && d.GreenCtxStreamCreate(&l, part->big, CU_STREAM_NON_BLOCKING, mid) == CUDA_SUCCESS
                     && d.GreenCtxStreamCreate(&t, part->big, CU_STREAM_NON_BLOCKING, lo) == CUDA_SUCCESS;
        if (ok) {
            chain = c; panel = p; look = l; trail = t;
            own_chain = true;
            chain_sm_count = part->sm_small;
        }
        else {
            for (CUstream s : {c, p, l, t}) if (s) cudaStreamDestroy(s);
            part = nullptr;
        }
    }
    if (! part) {
        CUDA_TRY(cudaStreamCreateWithPriority(&panel, cudaStreamNonBlocking, hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&look, cudaStreamNonBlocking, mid));
        CUDA_TRY(cudaStreamCreateWithPriority(&trail, cudaStreamNonBlocking, lo));
        chain = panel;
    }
    ev.resize(nevents);
    for (auto& e : ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&hop_ev[0], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&hop_ev[1], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&t0));
    CUDA_TRY(cudaEventCreate(&t1));
    return SB200_OK;
}

// work submitted to `to` after this call starts after everything submitted to `from` so far
int Streams::hop(cudaStream_t from, cudaStream_t to)
{
    if (from == to) return SB200_OK;
    cudaEvent_t e = hop_ev[hop_next ^= 1];
    CUDA_TRY(cudaEventRecord(e, from));        // a wait captures the record it sees now: re-recording later is safe
    CUDA_TRY(cudaStreamWaitEvent(to, e, 0));
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// grow-only cache of device workspaces (see runtime_internal.hh)
namespace {
struct WsBlock { void* p; size_t bytes; int device; bool busy; };
std::mutex g_ws_mu;
std::vector<WsBlock> g_ws;
}

extern "C" int sb200_release_workspaces(void);

void* ws_cache_get(size_t bytes)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    bytes = (bytes + 255) / 256 * 256;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        WsBlock* best = nullptr;
        for (auto& b : g_ws)          // best fit among the idle blocks of this device that are not wastefully large
            if (! b.busy && b.device == dev && b.bytes >= bytes && b.bytes <= 2 * bytes + (size_t(1) << 20)
                && (! best || b.bytes < best->bytes)) best = &b;
        if (best) { best->busy = true; return best->p; }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        sb200_release_workspaces();                  // idle blocks of other sizes may be what is in the way
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    std::lock_guard<std::mutex> lk(g_ws_mu);
    g_ws.push_back({p, bytes, dev, true});
    return p;
}

void ws_cache_put(void* p)
{
    // what cudaFree did implicitly and some callers (asynchronous solve-path helpers) rely on: nothing on the device
    // still uses the block when it becomes available again.  A no-op after a driver (its streams are synchronised).
    cudaDeviceSynchronize();
    // a process that factors matrices of many different shapes must not accumulate idle blocks without bound:
    // above SB200_WS_CACHE_MB (default 16 GiB) of idle memory everything idle is given back to the driver
    static const size_t cap = [] { const char* e = getenv("SB200_WS_CACHE_MB"); return size_t(e ? atoll(e) : 16384) << 20; }();
    size_t idle = 0;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        for (auto& b : g_ws) {
            if (b.p == p) b.busy = false;
            if (! b.busy) idle += b.bytes;
        }
    }
    if (idle > cap) sb200_release_workspaces();
}

extern "C" int sb200_release_workspaces(void)
{
    std::vector<void*> gone;
    {
        std::lock_guard<std::mutex> lk(g_ws_mu);
        for (size_t i = 0; i < g_ws.size();)
            if (! g_ws[i].busy) { gone.push_back(g_ws[i].p); g_ws[i] = g_ws.back(); g_ws.pop_back(); }
            else ++i;
    }
    for (void* p : gone) cudaFree(p);
    return SB200_OK;
}

extern "C" int sb200_sm_partition_probe(int chain_sms, int* sm_chain, int* sm_rest)
{
    Partition* p = partition_for(chain_sms);
    if (! p) return SB200_ENOTSUP;
    if (sm_chain) *sm_chain = p->sm_small;
    if (sm_rest) *sm_rest = p->sm_big;
    return SB200_OK;
}

} // namespace sb200
