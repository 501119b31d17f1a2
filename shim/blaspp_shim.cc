// blaspp_shim.cc -- seam 1 of the drop-in boundary.
//
// Defines the exact C++ symbols SLATE's src/internal links against for the device BLAS:
//   blas::batch::gemm / herk / syrk / trsm (vector-of-args, blas::Queue&)   4 scalar types each
//   blas::herk / blas::syrk (single tile, blas::Queue&)                      4 scalar types each
// forwarding to libslate_b200.so through its C ABI (include/slate_b200.h).  Linked into the
// reference build IN PLACE OF blaspp/src/device_batch_{gemm,herk,syrk,trsm}.cc and
// blaspp/src/device_{herk,syrk}.cc (oracle/build_ref_gpu.sh shows the link line); nothing of
// the reference is modified.  Semantics mirrored from those files: argument checks when `info`
// is non-empty, size-1 vectors = fixed-size batch (the only form SLATE uses,
// src/internal/internal_gemm.cc:467-504), host pointer vectors uploaded into queue.work(),
// asynchronous on queue.stream(), RowMajor handled by operand swap.
#include "blas.hh"
#include "blas/device.hh"
#include "blas/batch_common.hh"
#include "sb200_abi.hh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

namespace {

using namespace sb200_shim;

inline int ch(blas::Layout v) { return int(blas::to_char(v)); }
inline int ch(blas::Op v)     { return int(blas::to_char(v)); }
inline int ch(blas::Uplo v)   { return int(blas::to_char(v)); }
inline int ch(blas::Side v)   { return int(blas::to_char(v)); }
inline int ch(blas::Diag v)   { return int(blas::to_char(v)); }

template <typename V> auto const& at(V const& v, size_t i) { return v.size() == 1 ? v[0] : v[i]; }

// Upload `count` host pointers starting at src into the queue workspace slot `slot` (each slot holds
// MaxBatchChunk pointers), as blaspp does (device_batch_gemm.cc:96-112).
template <typename T>
T** upload(std::vector<T*> const& src, size_t first, size_t count, int slot, blas::Queue& queue)
{
    T** dst = reinterpret_cast<T**>(queue.work()) + size_t(slot) * blas::MaxBatchChunk;
    blas::device_copy_vector(int64_t(count), &src[first], 1, dst, 1, queue);
    return dst;
}

// bytes of queue.work() reserved for pointer arrays; kernel scratch (trsm) lives behind them
constexpr size_t kPtrBytes = 3 * size_t(blas::MaxBatchChunk) * sizeof(void*);

// ---- 'N','N' -> 'N','T': the B tiles of a batch are few and shared (SUMMA step: nt distinct B(k,j) under mt x nt
// products; LU trailing update: the row U(k, :)), and our FP64 kernel stages an N-contiguous B operand with TMA bulk
// copies (0.95 of the DMMA peak) but a K-contiguous one with 16-byte cp.async (0.85; the incumbent cuBLAS: 0.86).
// So each DISTINCT B tile is transposed once into a per-stream scratch (<= 0.1 % of the step's traffic) and the
// batch runs as 'N','T'.  Same products in the same order: bitwise the same C.  SB200_SHIM_GEMM_BT=0 turns it off.
struct BtScratch { void* buf = nullptr; size_t bytes = 0; };
inline void* bt_scratch(cudaStream_t s, size_t bytes)
{
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, BtScratch> map;      // stream-ordered reuse: one buffer per queue
    std::lock_guard<std::mutex> lk(mu);
    BtScratch& e = map[s];
    if (e.bytes < bytes) {
        if (e.buf) { cudaStreamSynchronize(s); cudaFree(e.buf); }
        e.buf = nullptr; e.bytes = 0;
        if (cudaMalloc(&e.buf, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        e.bytes = bytes;
    }
    return e.buf;
}
inline bool bt_enabled()
{
    static const bool on = [] { const char* e = std::getenv("SB200_SHIM_GEMM_BT"); return ! e || std::atoi(e) != 0; }();
    return on;
}
// Barray[first .. first+cnt) -> pointers to transposed copies (n x k, ld n); false = not applicable, nothing done
template <typename T>
bool transpose_b_tiles(std::vector<T*> const& Barray, size_t first, size_t cnt, int64_t k, int64_t n, int64_t ldb,
                       std::vector<T*>& Bt, blas::Queue& queue)
{
    if constexpr (! (std::is_same<T, double>::value || std::is_same<T, std::complex<double>>::value)) return false;
    if (! bt_enabled() || k > 1024 || n > 1024 || k < 16 || n < 16 || cnt < 8) return false;
    std::unordered_map<T*, size_t> slot;
    std::vector<T*> uniq;
    for (size_t i = 0; i < cnt; ++i)
        if (slot.emplace(Barray[first + i], uniq.size()).second) uniq.push_back(Barray[first + i]);
    if (uniq.size() * 4 > cnt) return false;                     // not shared enough to pay for the copies
    const size_t te = size_t(n) * size_t(k);
    T* buf = static_cast<T*>(bt_scratch(queue.stream(), uniq.size() * te * sizeof(T)));
    if (! buf) return false;
    std::vector<T*> dstv(uniq.size());
    for (size_t u = 0; u < uniq.size(); ++u) dstv[u] = buf + u * te;
    T** dsrc = upload(uniq, 0, uniq.size(), 0, queue);           // slots 0 / 1 are re-filled with A / B afterwards
    T** ddst = upload(dstv, 0, dstv.size(), 1, queue);           // (same stream: ordered behind the transposes)
    check(transpose_batched(tag<T>(), 0, k, n, cpp(dsrc), ldb, pp(ddst), n, int64_t(uniq.size()), queue.stream()),
          "batch::gemm (B transposes)");
    Bt.resize(cnt);
    for (size_t i = 0; i < cnt; ++i) Bt[i] = dstv[slot[Barray[first + i]]];
    return true;
}

template <typename T>
void batch_gemm(blas::Layout layout,
                std::vector<blas::Op> const& transA, std::vector<blas::Op> const& transB,
                std::vector<int64_t> const& m, std::vector<int64_t> const& n, std::vector<int64_t> const& k,
                std::vector<T> const& alpha,
                std::vector<T*> const& Aarray, std::vector<int64_t> const& lda,
                std::vector<T*> const& Barray, std::vector<int64_t> const& ldb,
                std::vector<T> const& beta,
                std::vector<T*> const& Carray, std::vector<int64_t> const& ldc,
                size_t batch_size, std::vector<int64_t>& info, blas::Queue& queue)
{
    blas_error_if(layout != blas::Layout::ColMajor && layout != blas::Layout::RowMajor);
    blas_error_if(info.size() != 0 && info.size() != 1 && info.size() != batch_size);
    if (info.size() > 0)
        blas::batch::gemm_check(layout, transA, transB, m, n, k, alpha, Aarray, lda, Barray, ldb,
                                beta, Carray, ldc, batch_size, info);
    if (batch_size == 0) return;
    blas::internal_set_device(queue.device());
    const bool fixed = transA.size() == 1 && transB.size() == 1 && m.size() == 1 && n.size() == 1
        && k.size() == 1 && alpha.size() == 1 && lda.size() == 1 && ldb.size() == 1 && beta.size() == 1
        && ldc.size() == 1 && Aarray.size() == batch_size && Barray.size() == batch_size
        && Carray.size() == batch_size;
    if (fixed) {
        queue.work_ensure_size<char>(kPtrBytes);
        for (size_t i = 0; i < batch_size; i += blas::MaxBatchChunk) {
            const size_t cnt = std::min(size_t(blas::MaxBatchChunk), batch_size - i);
            std::vector<T*> Bt;
            const bool bt = layout == blas::Layout::ColMajor && transB[0] == blas::Op::NoTrans
                && transpose_b_tiles(Barray, i, cnt, k[0], n[0], ldb[0], Bt, queue);
            T** dA = upload(Aarray, i, cnt, 0, queue);
            T** dB = bt ? upload(Bt, 0, cnt, 1, queue) : upload(Barray, i, cnt, 1, queue);
            T** dC = upload(Carray, i, cnt, 2, queue);
            check(gemm_batched(tag<T>(), ch(layout), ch(transA[0]), bt ? int('T') : ch(transB[0]), m[0], n[0], k[0],
                               to_abi(alpha[0]), cpp(dA), lda[0], cpp(dB), bt ? n[0] : ldb[0],
                               to_abi(beta[0]), pp(dC), ldc[0], int64_t(cnt), queue.stream()),
                  "batch::gemm");
        }
    }
    else {
        // variable-size batch: one strided launch (batch 1) per problem, all on queue.stream()
        for (size_t i = 0; i < batch_size; ++i)
            check(gemm_strided(tag<T>(), ch(layout), ch(at(transA, i)), ch(at(transB, i)),
                               at(m, i), at(n, i), at(k, i), to_abi(at(alpha, i)),
                               p((const T*) at(Aarray, i)), at(lda, i), int64_t(0),
                               p((const T*) at(Barray, i)), at(ldb, i), int64_t(0),
                               to_abi(at(beta, i)), p(at(Carray, i)), at(ldc, i), int64_t(0),
                               int64_t(1), queue.stream()),
                  "batch::gemm (variable size)");
    }
}

// herk / syrk: the reference loops over per-tile cublas calls on forked streams
// (device_batch_herk.cc:57-73); here the fixed-size case is ONE launch.
template <typename T, typename S, bool Herk>
void batch_rank_k(blas::Layout layout, std::vector<blas::Uplo> const& uplo, std::vector<blas::Op> const& trans,
                  std::vector<int64_t> const& n, std::vector<int64_t> const& k,
                  std::vector<S> const& alpha, std::vector<T*> const& Aarray, std::vector<int64_t> const& lda,
                  std::vector<S> const& beta, std::vector<T*> const& Carray, std::vector<int64_t> const& ldc,
                  size_t batch_size, std::vector<int64_t>& info, blas::Queue& queue)
{
    blas_error_if(layout != blas::Layout::ColMajor && layout != blas::Layout::RowMajor);
    blas_error_if(info.size() != 0 && info.size() != 1 && info.size() != batch_size);
    if (info.size() > 0) {
        if constexpr (Herk)
            blas::batch::herk_check(layout, uplo, trans, n, k, alpha, Aarray, lda, beta, Carray, ldc, batch_size, info);
        else
            blas::batch::syrk_check(layout, uplo, trans, n, k, alpha, Aarray, lda, beta, Carray, ldc, batch_size, info);
    }
    if (batch_size == 0) return;
    blas::internal_set_device(queue.device());
    const bool fixed = uplo.size() == 1 && trans.size() == 1 && n.size() == 1 && k.size() == 1
        && alpha.size() == 1 && lda.size() == 1 && beta.size() == 1 && ldc.size() == 1
        && Aarray.size() == batch_size && Carray.size() == batch_size;
    if (fixed) {
        queue.work_ensure_size<char>(kPtrBytes);
        for (size_t i = 0; i < batch_size; i += blas::MaxBatchChunk) {
            const size_t cnt = std::min(size_t(blas::MaxBatchChunk), batch_size - i);
            T** dA = upload(Aarray, i, cnt, 0, queue);
            T** dC = upload(Carray, i, cnt, 2, queue);
            int st;
            if constexpr (Herk)
                st = herk_batched(tag<T>(), ch(layout), ch(uplo[0]), ch(trans[0]), n[0], k[0], alpha[0],
                                  cpp(dA), lda[0], beta[0], pp(dC), ldc[0], int64_t(cnt), queue.stream());
            else
                st = syrk_batched(tag<T>(), ch(layout), ch(uplo[0]), ch(trans[0]), n[0], k[0], to_abi(alpha[0]),
                                  cpp(dA), lda[0], to_abi(beta[0]), pp(dC), ldc[0], int64_t(cnt), queue.stream());
            check(st, Herk ? "batch::herk" : "batch::syrk");
        }
    }
    else {
        for (size_t i = 0; i < batch_size; ++i) {
            int st;
            if constexpr (Herk)
                st = herk(tag<T>(), ch(layout), ch(at(uplo, i)), ch(at(trans, i)), at(n, i), at(k, i), at(alpha, i),
                          p((const T*) at(Aarray, i)), at(lda, i), at(beta, i), p(at(Carray, i)), at(ldc, i), queue.stream());
            else
                st = syrk(tag<T>(), ch(layout), ch(at(uplo, i)), ch(at(trans, i)), at(n, i), at(k, i), to_abi(at(alpha, i)),
                          p((const T*) at(Aarray, i)), at(lda, i), to_abi(at(beta, i)), p(at(Carray, i)), at(ldc, i), queue.stream());
            check(st, Herk ? "batch::herk (variable size)" : "batch::syrk (variable size)");
        }
    }
}

template <typename T>
void batch_trsm(blas::Layout layout, std::vector<blas::Side> const& side, std::vector<blas::Uplo> const& uplo,
                std::vector<blas::Op> const& trans, std::vector<blas::Diag> const& diag,
                std::vector<int64_t> const& m, std::vector<int64_t> const& n, std::vector<T> const& alpha,
                std::vector<T*> const& Aarray, std::vector<int64_t> const& lda,
                std::vector<T*> const& Barray, std::vector<int64_t> const& ldb,
                size_t batch_size, std::vector<int64_t>& info, blas::Queue& queue)
{
    blas_error_if(layout != blas::Layout::ColMajor && layout != blas::Layout::RowMajor);
    blas_error_if(info.size() != 0 && info.size() != 1 && info.size() != batch_size);
    if (info.size() > 0)
        blas::batch::trsm_check(layout, side, uplo, trans, diag, m, n, alpha, Aarray, lda, Barray, ldb, batch_size, info);
    if (batch_size == 0) return;
    blas::internal_set_device(queue.device());
    constexpr int dtype = std::is_same<T, float>::value ? 's' : std::is_same<T, double>::value ? 'd'
                        : std::is_same<T, std::complex<float>>::value ? 'c' : 'z';
    // SLATE passes the SAME triangular tile for the whole batch (internal_trsm.cc:225-249); the
    // kernel inverts its diagonal blocks once.  Runs of equal A pointers become one launch each.
    size_t i = 0;
    while (i < batch_size) {
        size_t j = i + 1;
        const bool same_shape_tail = side.size() == 1 && uplo.size() == 1 && trans.size() == 1 && diag.size() == 1
            && m.size() == 1 && n.size() == 1 && alpha.size() == 1 && lda.size() == 1 && ldb.size() == 1;
        if (same_shape_tail)
            while (j < batch_size && j - i < size_t(blas::MaxBatchChunk) && Aarray[j] == Aarray[i]) ++j;
        const size_t cnt = j - i;
        const size_t wbytes = sb200_trsm_work_bytes(dtype, ch(at(side, i)), at(m, i), at(n, i));
        queue.work_ensure_size<char>(kPtrBytes + wbytes);
        T** dB = upload(Barray, i, cnt, 1, queue);
        void* work = static_cast<char*>(queue.work()) + kPtrBytes;
        check(trsm_batched(tag<T>(), ch(layout), ch(at(side, i)), ch(at(uplo, i)), ch(at(trans, i)), ch(at(diag, i)),
                           at(m, i), at(n, i), to_abi(at(alpha, i)), p((const T*) Aarray[i]), at(lda, i),
                           pp(dB), at(ldb, i), int64_t(cnt), work, queue.stream()),
              "batch::trsm");
        i = j;
    }
}

} // namespace

namespace blas {

// ------------------------------------------------------------------------------------------
// single-tile herk / syrk on a queue (blaspp/src/device_herk.cc, device_syrk.cc)
#define SB200_SHIM_HERK(T, S) \
void herk(blas::Layout layout, blas::Uplo uplo, blas::Op trans, int64_t n, int64_t k, \
          S alpha, T const* A, int64_t lda, S beta, T* C, int64_t ldc, blas::Queue& queue) \
{ \
    blas::internal_set_device(queue.device()); \
    check(sb200_shim::herk(tag<T>(), ch(layout), ch(uplo), ch(trans), n, k, alpha, p(A), lda, beta, p(C), ldc, \
                           queue.stream()), "herk"); \
}
SB200_SHIM_HERK(float, float)
SB200_SHIM_HERK(double, double)
SB200_SHIM_HERK(std::complex<float>, float)
SB200_SHIM_HERK(std::complex<double>, double)

#define SB200_SHIM_SYRK(T) \
void syrk(blas::Layout layout, blas::Uplo uplo, blas::Op trans, int64_t n, int64_t k, \
          T alpha, T const* A, int64_t lda, T beta, T* C, int64_t ldc, blas::Queue& queue) \
{ \
    blas::internal_set_device(queue.device()); \
    check(sb200_shim::syrk(tag<T>(), ch(layout), ch(uplo), ch(trans), n, k, to_abi(alpha), p(A), lda, to_abi(beta), \
                           p(C), ldc, queue.stream()), "syrk"); \
}
SB200_SHIM_SYRK(float)
SB200_SHIM_SYRK(double)
SB200_SHIM_SYRK(std::complex<float>)
SB200_SHIM_SYRK(std::complex<double>)

namespace batch {

#define SB200_SHIM_BATCH_GEMM(T) \
void gemm(blas::Layout layout, std::vector<blas::Op> const& transA, std::vector<blas::Op> const& transB, \
          std::vector<int64_t> const& m, std::vector<int64_t> const& n, std::vector<int64_t> const& k, \
          std::vector<T> const& alpha, std::vector<T*> const& Aarray, std::vector<int64_t> const& lda, \
          std::vector<T*> const& Barray, std::vector<int64_t> const& ldb, std::vector<T> const& beta, \
          std::vector<T*> const& Carray, std::vector<int64_t> const& ldc, \
          size_t batch_size, std::vector<int64_t>& info, blas::Queue& queue) \
{ batch_gemm<T>(layout, transA, transB, m, n, k, alpha, Aarray, lda, Barray, ldb, beta, Carray, ldc, batch_size, info, queue); }
SB200_SHIM_BATCH_GEMM(float)
SB200_SHIM_BATCH_GEMM(double)
SB200_SHIM_BATCH_GEMM(std::complex<float>)
SB200_SHIM_BATCH_GEMM(std::complex<double>)

#define SB200_SHIM_BATCH_RANKK(name, T, S, H) \
void name(blas::Layout layout, std::vector<blas::Uplo> const& uplo, std::vector<blas::Op> const& trans, \
          std::vector<int64_t> const& n, std::vector<int64_t> const& k, \
          std::vector<S> const& alpha, std::vector<T*> const& Aarray, std::vector<int64_t> const& lda, \
          std::vector<S> const& beta, std::vector<T*> const& Carray, std::vector<int64_t> const& ldc, \
          size_t batch_size, std::vector<int64_t>& info, blas::Queue& queue) \
{ batch_rank_k<T, S, H>(layout, uplo, trans, n, k, alpha, Aarray, lda, beta, Carray, ldc, batch_size, info, queue); }
SB200_SHIM_BATCH_RANKK(herk, float, float, true)
SB200_SHIM_BATCH_RANKK(herk, double, double, true)
SB200_SHIM_BATCH_RANKK(herk, std::complex<float>, float, true)
SB200_SHIM_BATCH_RANKK(herk, std::complex<double>, double, true)
SB200_SHIM_BATCH_RANKK(syrk, float, float, false)
SB200_SHIM_BATCH_RANKK(syrk, double, double, false)
SB200_SHIM_BATCH_RANKK(syrk, std::complex<float>, std::complex<float>, false)
SB200_SHIM_BATCH_RANKK(syrk, std::complex<double>, std::complex<double>, false)

#define SB200_SHIM_BATCH_TRSM(T) \
void trsm(blas::Layout layout, std::vector<blas::Side> const& side, std::vector<blas::Uplo> const& uplo, \
          std::vector<blas::Op> const& trans, std::vector<blas::Diag> const& diag, \
          std::vector<int64_t> const& m, std::vector<int64_t> const& n, std::vector<T> const& alpha, \
          std::vector<T*> const& Aarray, std::vector<int64_t> const& lda, \
          std::vector<T*> const& Barray, std::vector<int64_t> const& ldb, \
          size_t batch_size, std::vector<int64_t>& info, blas::Queue& queue) \
{ batch_trsm<T>(layout, side, uplo, trans, diag, m, n, alpha, Aarray, lda, Barray, ldb, batch_size, info, queue); }
SB200_SHIM_BATCH_TRSM(float)
SB200_SHIM_BATCH_TRSM(double)
SB200_SHIM_BATCH_TRSM(std::complex<float>)
SB200_SHIM_BATCH_TRSM(std::complex<double>)

} // namespace batch
} // namespace blas
