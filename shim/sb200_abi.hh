// sb200_abi.hh -- C++ overload set over the C ABI (include/slate_b200.h) so the shim's templates
// can dispatch on the scalar type.  Part of the drop-in shim; includes no reference header.
#pragma once
#include "slate_b200.h"
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace sb200_shim {

inline float      to_abi(float v) { return v; }
inline double     to_abi(double v) { return v; }
inline sb200_c32  to_abi(std::complex<float> v) { return sb200_c32{v.real(), v.imag()}; }
inline sb200_c64  to_abi(std::complex<double> v) { return sb200_c64{v.real(), v.imag()}; }

template <typename T> struct Abi { using type = T; };
template <> struct Abi<std::complex<float>>  { using type = sb200_c32; };
template <> struct Abi<std::complex<double>> { using type = sb200_c64; };
template <typename T> using abi_t = typename Abi<T>::type;

template <typename T> inline abi_t<T>* p(T* x) { return reinterpret_cast<abi_t<T>*>(x); }
template <typename T> inline const abi_t<T>* p(const T* x) { return reinterpret_cast<const abi_t<T>*>(x); }
template <typename T> inline abi_t<T>* const* pp(T* const* x) { return reinterpret_cast<abi_t<T>* const*>(x); }
template <typename T> inline const abi_t<T>* const* cpp(const T* const* x) { return reinterpret_cast<const abi_t<T>* const*>(x); }
template <typename T> inline const abi_t<T>* const* cpp(T* const* x) { return reinterpret_cast<const abi_t<T>* const*>(x); }

// status -> exception (the reference's boundary throws blas::Error / slate::Exception;
// SURVEY.md section 8b "Errors")
inline void check(int status, const char* what)
{
    if (status != SB200_OK)
        throw std::runtime_error(std::string("slate_b200: ") + what + ": " + sb200_strerror(status));
}

// Overload sets: fn(float...) -> sb200_fn_s, fn(double...) -> _d, complex<float> -> _c, complex<double> -> _z
#define SB200_SHIM_DISPATCH(name) \
    template <typename... Args> inline int name(float*, Args... a)                { return sb200_##name##_s(a...); } \
    template <typename... Args> inline int name(double*, Args... a)               { return sb200_##name##_d(a...); } \
    template <typename... Args> inline int name(std::complex<float>*, Args... a)  { return sb200_##name##_c(a...); } \
    template <typename... Args> inline int name(std::complex<double>*, Args... a) { return sb200_##name##_z(a...); }
// usage: name((T*) nullptr, args...) -- the first argument only selects the type
SB200_SHIM_DISPATCH(gemm_batched)
SB200_SHIM_DISPATCH(gemm_strided)
SB200_SHIM_DISPATCH(herk_batched)
SB200_SHIM_DISPATCH(syrk_batched)
SB200_SHIM_DISPATCH(herk)
SB200_SHIM_DISPATCH(syrk)
SB200_SHIM_DISPATCH(trsm_batched)
SB200_SHIM_DISPATCH(potrf_tile)
SB200_SHIM_DISPATCH(geadd)
SB200_SHIM_DISPATCH(geadd_batched)
SB200_SHIM_DISPATCH(tzadd_batched)
SB200_SHIM_DISPATCH(gescale)
SB200_SHIM_DISPATCH(gescale_batched)
SB200_SHIM_DISPATCH(tzscale_batched)
SB200_SHIM_DISPATCH(gescale_row_col_batched)
SB200_SHIM_DISPATCH(gescale_row_col_real_batched)
SB200_SHIM_DISPATCH(geset)
SB200_SHIM_DISPATCH(geset_batched)
SB200_SHIM_DISPATCH(tzset_batched)
SB200_SHIM_DISPATCH(transpose_inplace)
SB200_SHIM_DISPATCH(transpose)
SB200_SHIM_DISPATCH(transpose_inplace_batched)
SB200_SHIM_DISPATCH(transpose_batched)
SB200_SHIM_DISPATCH(genorm_batched)
SB200_SHIM_DISPATCH(henorm_batched)
SB200_SHIM_DISPATCH(synorm_batched)
SB200_SHIM_DISPATCH(synorm_offdiag_batched)
SB200_SHIM_DISPATCH(trnorm_batched)
#undef SB200_SHIM_DISPATCH

template <typename T> constexpr T* tag() { return static_cast<T*>(nullptr); }

} // namespace sb200_shim
