// slate_device_shim.cc -- seam 2 of the drop-in boundary.
//
// Defines every template of `namespace slate::device` declared in
// include/slate/internal/device.hh:92-281, explicitly specialised for the four scalar types the
// reference instantiates in src/cuda/device_*.cu, forwarding to libslate_b200.so through its
// C ABI (include/slate_b200.h).  Linked into the reference build IN PLACE OF src/cuda/*.cu
// (or src/omptarget/*.cc).  Conventions mirrored from the reference kernels' host wrappers:
// quick return on empty sizes (device_geadd.cu:127-129), cudaSetDevice(queue.device())
// (device_geadd.cu:131), asynchronous launch on queue.stream(), failure -> exception
// (slate_assert in the reference).  Pointer arrays are DEVICE arrays uploaded by the caller
// (src/internal/internal_geadd.cc:163-165).
#include "blas.hh"
#include "blas/device.hh"
#include "slate/Exception.hh"
#include "slate/internal/device.hh"
#include "sb200_abi.hh"

#include <complex>

namespace slate {
namespace device {

namespace {
using namespace sb200_shim;
inline int ch(Uplo v)      { return int(blas::to_char(v)); }
inline int ch(Diag v)      { return int(blas::to_char(v)); }
inline int ch(Norm v)      { return int(lapack::to_char(v)); }
inline int ch(NormScope v) { return int(char(v)); }
inline int ch(Equed v)     { return int(lapack::to_char(v)); }
inline int tz_uplo(Uplo u) { return u == Uplo::Lower ? 'L' : (u == Uplo::Upper ? 'U' : 'G'); }
inline void dev(blas::Queue& q) { blas::internal_set_device(q.device()); }
} // namespace

// ------------------------------------------------------------------------------ gecopy / tzcopy
#define SB200_SHIM_COPY(XY, S, D) \
template <> \
void gecopy(int64_t m, int64_t n, S const* const* Aarray, int64_t lda, D** Barray, int64_t ldb, \
            int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(sb200_gecopy_batched_##XY(m, n, cpp(Aarray), lda, pp(Barray), ldb, batch_count, queue.stream()), "device::gecopy"); \
}
#define SB200_SHIM_TZCOPY(XY, S, D) \
template <> \
void tzcopy(Uplo uplo, int64_t m, int64_t n, S const* const* Aarray, int64_t lda, D** Barray, int64_t ldb, \
            int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(sb200_tzcopy_batched_##XY(ch(uplo), m, n, cpp(Aarray), lda, pp(Barray), ldb, batch_count, queue.stream()), "device::tzcopy"); \
}
using cf = std::complex<float>;
using cd = std::complex<double>;
SB200_SHIM_COPY(ss, float, float)   SB200_SHIM_COPY(sd, float, double)
SB200_SHIM_COPY(dd, double, double) SB200_SHIM_COPY(ds, double, float)
SB200_SHIM_COPY(cc, cf, cf)         SB200_SHIM_COPY(cz, cf, cd)
SB200_SHIM_COPY(zz, cd, cd)         SB200_SHIM_COPY(zc, cd, cf)
SB200_SHIM_COPY(sc, float, cf)      SB200_SHIM_COPY(dz, double, cd)
SB200_SHIM_TZCOPY(ss, float, float)   SB200_SHIM_TZCOPY(sd, float, double)
SB200_SHIM_TZCOPY(dd, double, double) SB200_SHIM_TZCOPY(ds, double, float)
SB200_SHIM_TZCOPY(cc, cf, cf)         SB200_SHIM_TZCOPY(cz, cf, cd)
SB200_SHIM_TZCOPY(zz, cd, cd)         SB200_SHIM_TZCOPY(zc, cd, cf)

// ------------------------------------------------------------------------------ per-type families
#define SB200_SHIM_TYPE(T, R) \
template <> \
void geadd(int64_t m, int64_t n, T const& alpha, T* A, int64_t lda, T const& beta, T* B, int64_t ldb, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0) return; \
    dev(queue); \
    check(sb200_shim::geadd(tag<T>(), m, n, to_abi(alpha), p((const T*) A), lda, to_abi(beta), p(B), ldb, queue.stream()), "device::geadd"); \
} \
template <> \
void tzadd(Uplo uplo, int64_t m, int64_t n, T const& alpha, T** Aarray, int64_t lda, T const& beta, T** Barray, int64_t ldb, \
           int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(tzadd_batched(tag<T>(), ch(uplo), m, n, to_abi(alpha), cpp(Aarray), lda, to_abi(beta), pp(Barray), ldb, \
                        batch_count, queue.stream()), "device::tzadd"); \
} \
template <> \
void gescale(int64_t m, int64_t n, T numer, T denom, T* A, int64_t lda, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0) return; \
    dev(queue); \
    check(sb200_shim::gescale(tag<T>(), m, n, to_abi(numer), to_abi(denom), p(A), lda, queue.stream()), "device::gescale"); \
} \
template <> \
void gescale_row_col_batch(Equed equed, int64_t m, int64_t n, T const* const* Rarray, T const* const* Carray, \
                           T** Aarray, int64_t lda, int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(gescale_row_col_batched(tag<T>(), ch(equed), m, n, cpp(Rarray), cpp(Carray), pp(Aarray), lda, batch_count, \
                                  queue.stream()), "device::gescale_row_col_batch"); \
} \
template <> \
void geset(int64_t m, int64_t n, T const& offdiag_value, T const& diag_value, T* A, int64_t lda, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0) return; \
    dev(queue); \
    check(sb200_shim::geset(tag<T>(), int('G'), m, n, to_abi(offdiag_value), to_abi(diag_value), p(A), lda, queue.stream()), "device::geset"); \
} \
template <> \
void tzset(Uplo uplo, int64_t m, int64_t n, T const& offdiag_value, T const& diag_value, T* A, int64_t lda, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0) return; \
    dev(queue); \
    check(sb200_shim::geset(tag<T>(), tz_uplo(uplo), m, n, to_abi(offdiag_value), to_abi(diag_value), p(A), lda, queue.stream()), "device::tzset"); \
} \
namespace batch { \
template <> \
void gescale(int64_t m, int64_t n, T numer, T denom, T** Aarray, int64_t lda, int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(gescale_batched(tag<T>(), m, n, to_abi(numer), to_abi(denom), pp(Aarray), lda, batch_count, queue.stream()), "device::batch::gescale"); \
} \
template <> \
void tzscale(Uplo uplo, int64_t m, int64_t n, R numer, R denom, T** Aarray, int64_t lda, int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(tzscale_batched(tag<T>(), ch(uplo), m, n, numer, denom, pp(Aarray), lda, batch_count, queue.stream()), "device::batch::tzscale"); \
} \
template <> \
void geadd(int64_t m, int64_t n, T const& alpha, T** Aarray, int64_t lda, T const& beta, T** Barray, int64_t ldb, \
           int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(geadd_batched(tag<T>(), m, n, to_abi(alpha), cpp(Aarray), lda, to_abi(beta), pp(Barray), ldb, batch_count, \
                        queue.stream()), "device::batch::geadd"); \
} \
template <> \
void geset(int64_t m, int64_t n, T const& offdiag_value, T const& diag_value, T** Aarray, int64_t lda, \
           int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(geset_batched(tag<T>(), m, n, to_abi(offdiag_value), to_abi(diag_value), pp(Aarray), lda, batch_count, \
                        queue.stream()), "device::batch::geset"); \
} \
template <> \
void tzset(Uplo uplo, int64_t m, int64_t n, T const& offdiag_value, T const& diag_value, T** Aarray, int64_t lda, \
           int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(tzset_batched(tag<T>(), ch(uplo), m, n, to_abi(offdiag_value), to_abi(diag_value), pp(Aarray), lda, batch_count, \
                        queue.stream()), "device::batch::tzset"); \
} \
} /* namespace batch */ \
template <> \
void genorm(Norm norm, NormScope scope, int64_t m, int64_t n, T const* const* Aarray, int64_t lda, \
            R* values, int64_t ldv, int64_t batch_count, blas::Queue& queue) \
{ \
    if (batch_count == 0) return; \
    dev(queue); \
    check(genorm_batched(tag<T>(), ch(norm), ch(scope), m, n, cpp(Aarray), lda, values, ldv, batch_count, queue.stream()), "device::genorm"); \
} \
template <> \
void henorm(Norm norm, Uplo uplo, int64_t n, T const* const* Aarray, int64_t lda, \
            R* values, int64_t ldv, int64_t batch_count, blas::Queue& queue) \
{ \
    if (batch_count == 0) return; \
    dev(queue); \
    check(henorm_batched(tag<T>(), ch(norm), ch(uplo), n, cpp(Aarray), lda, values, ldv, batch_count, queue.stream()), "device::henorm"); \
} \
template <> \
void synorm(Norm norm, Uplo uplo, int64_t n, T const* const* Aarray, int64_t lda, \
            R* values, int64_t ldv, int64_t batch_count, blas::Queue& queue) \
{ \
    if (batch_count == 0) return; \
    dev(queue); \
    check(synorm_batched(tag<T>(), ch(norm), ch(uplo), n, cpp(Aarray), lda, values, ldv, batch_count, queue.stream()), "device::synorm"); \
} \
template <> \
void synormOffdiag(Norm norm, int64_t m, int64_t n, T const* const* Aarray, int64_t lda, \
                   R* values, int64_t ldv, int64_t batch_count, blas::Queue& queue) \
{ \
    if (batch_count == 0) return; \
    dev(queue); \
    check(synorm_offdiag_batched(tag<T>(), ch(norm), m, n, cpp(Aarray), lda, values, ldv, batch_count, queue.stream()), "device::synormOffdiag"); \
} \
template <> \
void trnorm(Norm norm, Uplo uplo, Diag diag, int64_t m, int64_t n, T const* const* Aarray, int64_t lda, \
            R* values, int64_t ldv, int64_t batch_count, blas::Queue& queue) \
{ \
    if (batch_count == 0) return; \
    dev(queue); \
    check(trnorm_batched(tag<T>(), ch(norm), ch(uplo), ch(diag), m, n, cpp(Aarray), lda, values, ldv, batch_count, queue.stream()), "device::trnorm"); \
} \
template <> \
void transpose(bool is_conj, int64_t n, T* A, int64_t lda, blas::Queue& queue) \
{ \
    if (n <= 1 && ! is_conj) return; \
    dev(queue); \
    check(transpose_inplace(tag<T>(), int(is_conj), n, p(A), lda, queue.stream()), "device::transpose"); \
} \
template <> \
void transpose_batch(bool is_conj, int64_t n, T** Aarray, int64_t lda, int64_t batch_count, blas::Queue& queue) \
{ \
    if (batch_count == 0 || (n <= 1 && ! is_conj)) return; \
    dev(queue); \
    check(transpose_inplace_batched(tag<T>(), int(is_conj), n, pp(Aarray), lda, batch_count, queue.stream()), "device::transpose_batch"); \
} \
template <> \
void transpose(bool is_conj, int64_t m, int64_t n, T* dA, int64_t lda, T* dAT, int64_t ldat, blas::Queue& queue) \
{ \
    if (m <= 0 || n <= 0) return; \
    dev(queue); \
    check(sb200_shim::transpose(tag<T>(), int(is_conj), m, n, p((const T*) dA), lda, p(dAT), ldat, queue.stream()), "device::transpose"); \
} \
template <> \
void transpose_batch(bool is_conj, int64_t m, int64_t n, T** dA_array, int64_t lda, T** dAT_array, int64_t ldat, \
                     int64_t batch_count, blas::Queue& queue) \
{ \
    if (m <= 0 || n <= 0 || batch_count == 0) return; \
    dev(queue); \
    check(transpose_batched(tag<T>(), int(is_conj), m, n, cpp(dA_array), lda, pp(dAT_array), ldat, batch_count, \
                            queue.stream()), "device::transpose_batch"); \
}

SB200_SHIM_TYPE(float, float)
SB200_SHIM_TYPE(double, double)
SB200_SHIM_TYPE(cf, float)
SB200_SHIM_TYPE(cd, double)

// mixed scalar / real-scale variants the reference also instantiates for complex tiles
// (device_gescale.cu:126-170, 258-305; device_gescale_row_col.cu:228-256)
#define SB200_SHIM_REALSCALE(T, R) \
template <> \
void gescale(int64_t m, int64_t n, R numer, R denom, T* A, int64_t lda, blas::Queue& queue) \
{ gescale(m, n, T(numer), T(denom), A, lda, queue); } \
namespace batch { \
template <> \
void gescale(int64_t m, int64_t n, R numer, R denom, T** Aarray, int64_t lda, int64_t batch_count, blas::Queue& queue) \
{ gescale(m, n, T(numer), T(denom), Aarray, lda, batch_count, queue); } \
} \
template <> \
void gescale_row_col_batch(Equed equed, int64_t m, int64_t n, R const* const* Rarray, R const* const* Carray, \
                           T** Aarray, int64_t lda, int64_t batch_count, blas::Queue& queue) \
{ \
    if (m == 0 || n == 0 || batch_count == 0) return; \
    dev(queue); \
    check(gescale_row_col_real_batched(tag<T>(), ch(equed), m, n, Rarray, Carray, pp(Aarray), lda, batch_count, \
                                       queue.stream()), "device::gescale_row_col_batch"); \
}
SB200_SHIM_REALSCALE(cf, float)
SB200_SHIM_REALSCALE(cd, double)

} // namespace device
} // namespace slate
