// lapackpp_shim.cc -- seam 1, LAPACK part: lapack::potrf(uplo, n, dA, ldda, dev_info, queue)
// (declared in lapackpp/include/lapack/device.hh:127-131; the reference implementation is
// lapackpp/src/cuda/cuda_potrf.cc -> cusolverDn?potrf).  Called by SLATE for every diagonal tile
// (src/internal/internal_potrf.cc:72-78).  Linked in place of lapackpp/src/cuda/cuda_potrf.cc.
// Asynchronous on queue.stream(); LAPACK info lands in *dev_info on the device.
#include "lapack.hh"
#include "lapack/device.hh"
#include "sb200_abi.hh"

namespace lapack {

namespace {
template <typename T>
void potrf_impl(lapack::Uplo uplo, int64_t n, T* dA, int64_t ldda, device_info_int* dev_info, lapack::Queue& queue)
{
    using namespace sb200_shim;
    static_assert(sizeof(device_info_int) == sizeof(int), "cuSOLVER-style int info expected");
    blas::internal_set_device(queue.device());
    // scratch for the inverted 64x64 diagonal block lives in the queue workspace (per stream)
    queue.work_ensure_size<char>(64 * 64 * sizeof(T));
    check(potrf_tile(tag<T>(), int(blas::to_char(uplo)), n, p(dA), ldda, reinterpret_cast<int*>(dev_info),
                     queue.work(), queue.stream()), "lapack::potrf");
}
} // namespace

// the reference declares a template and instantiates it explicitly (cuda_potrf.cc:218-246)
template <>
void potrf<float>(lapack::Uplo uplo, int64_t n, float* dA, int64_t ldda, device_info_int* dev_info, lapack::Queue& queue)
{ potrf_impl(uplo, n, dA, ldda, dev_info, queue); }
template <>
void potrf<double>(lapack::Uplo uplo, int64_t n, double* dA, int64_t ldda, device_info_int* dev_info, lapack::Queue& queue)
{ potrf_impl(uplo, n, dA, ldda, dev_info, queue); }
template <>
void potrf<std::complex<float>>(lapack::Uplo uplo, int64_t n, std::complex<float>* dA, int64_t ldda,
                                device_info_int* dev_info, lapack::Queue& queue)
{ potrf_impl(uplo, n, dA, ldda, dev_info, queue); }
template <>
void potrf<std::complex<double>>(lapack::Uplo uplo, int64_t n, std::complex<double>* dA, int64_t ldda,
                                 device_info_int* dev_info, lapack::Queue& queue)
{ potrf_impl(uplo, n, dA, ldda, dev_info, queue); }

} // namespace lapack
