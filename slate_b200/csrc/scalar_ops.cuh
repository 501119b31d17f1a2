// scalar_ops.cuh -- arithmetic helpers shared by the type-generic kernels
// (float, double, cuFloatComplex, cuDoubleComplex).
#pragma once
#include <cuComplex.h>

namespace sb200 {

template <typename T> struct RealOf { using type = T; };
template <> struct RealOf<cuDoubleComplex> { using type = double; };
template <> struct RealOf<cuFloatComplex>  { using type = float; };

template <typename T> __host__ __device__ inline T zero_of();
template <> __host__ __device__ inline float  zero_of<float>() { return 0.f; }
template <> __host__ __device__ inline double zero_of<double>() { return 0.0; }
template <> __host__ __device__ inline cuFloatComplex  zero_of<cuFloatComplex>() { return make_cuFloatComplex(0.f, 0.f); }
template <> __host__ __device__ inline cuDoubleComplex zero_of<cuDoubleComplex>() { return make_cuDoubleComplex(0.0, 0.0); }

template <typename T> __host__ __device__ inline T from_real(typename RealOf<T>::type r);
template <> __host__ __device__ inline float  from_real<float>(float r) { return r; }
template <> __host__ __device__ inline double from_real<double>(double r) { return r; }
template <> __host__ __device__ inline cuFloatComplex  from_real<cuFloatComplex>(float r) { return make_cuFloatComplex(r, 0.f); }
template <> __host__ __device__ inline cuDoubleComplex from_real<cuDoubleComplex>(double r) { return make_cuDoubleComplex(r, 0.0); }

__host__ __device__ inline float  mul(float a, float b) { return a * b; }
__host__ __device__ inline double mul(double a, double b) { return a * b; }
__host__ __device__ inline cuFloatComplex  mul(cuFloatComplex a, cuFloatComplex b) { return cuCmulf(a, b); }
__host__ __device__ inline cuDoubleComplex mul(cuDoubleComplex a, cuDoubleComplex b) { return cuCmul(a, b); }
__host__ __device__ inline float  add(float a, float b) { return a + b; }
__host__ __device__ inline double add(double a, double b) { return a + b; }
__host__ __device__ inline cuFloatComplex  add(cuFloatComplex a, cuFloatComplex b) { return cuCaddf(a, b); }
__host__ __device__ inline cuDoubleComplex add(cuDoubleComplex a, cuDoubleComplex b) { return cuCadd(a, b); }
__host__ __device__ inline float  divide(float a, float b) { return a / b; }
__host__ __device__ inline double divide(double a, double b) { return a / b; }
__host__ __device__ inline cuFloatComplex  divide(cuFloatComplex a, cuFloatComplex b) { return cuCdivf(a, b); }
__host__ __device__ inline cuDoubleComplex divide(cuDoubleComplex a, cuDoubleComplex b) { return cuCdiv(a, b); }
__host__ __device__ inline float  conj_(float a) { return a; }
__host__ __device__ inline double conj_(double a) { return a; }
__host__ __device__ inline cuFloatComplex  conj_(cuFloatComplex a) { return cuConjf(a); }
__host__ __device__ inline cuDoubleComplex conj_(cuDoubleComplex a) { return cuConj(a); }
__host__ __device__ inline bool is_zero(float a) { return a == 0.f; }
__host__ __device__ inline bool is_zero(double a) { return a == 0.0; }
__host__ __device__ inline bool is_zero(cuFloatComplex a) { return a.x == 0.f && a.y == 0.f; }
__host__ __device__ inline bool is_zero(cuDoubleComplex a) { return a.x == 0.0 && a.y == 0.0; }
__host__ __device__ inline float  real_part_only(float a) { return a; }
__host__ __device__ inline double real_part_only(double a) { return a; }
__host__ __device__ inline cuFloatComplex  real_part_only(cuFloatComplex a) { return make_cuFloatComplex(a.x, 0.f); }
__host__ __device__ inline cuDoubleComplex real_part_only(cuDoubleComplex a) { return make_cuDoubleComplex(a.x, 0.0); }


__host__ __device__ inline float  real_of(float a) { return a; }
__host__ __device__ inline double real_of(double a) { return a; }
__host__ __device__ inline float  real_of(cuFloatComplex a) { return a.x; }
__host__ __device__ inline double real_of(cuDoubleComplex a) { return a.x; }
__host__ __device__ inline float  sub(float a, float b) { return a - b; }
__host__ __device__ inline double sub(double a, double b) { return a - b; }
__host__ __device__ inline cuFloatComplex  sub(cuFloatComplex a, cuFloatComplex b) { return cuCsubf(a, b); }
__host__ __device__ inline cuDoubleComplex sub(cuDoubleComplex a, cuDoubleComplex b) { return cuCsub(a, b); }
__host__ __device__ inline float  neg(float a) { return -a; }
__host__ __device__ inline double neg(double a) { return -a; }
__host__ __device__ inline cuFloatComplex  neg(cuFloatComplex a) { return make_cuFloatComplex(-a.x, -a.y); }
__host__ __device__ inline cuDoubleComplex neg(cuDoubleComplex a) { return make_cuDoubleComplex(-a.x, -a.y); }
// a / r with r real
__host__ __device__ inline float  div_real(float a, float r) { return a / r; }
__host__ __device__ inline double div_real(double a, double r) { return a / r; }
__host__ __device__ inline cuFloatComplex  div_real(cuFloatComplex a, float r) { return make_cuFloatComplex(a.x / r, a.y / r); }
__host__ __device__ inline cuDoubleComplex div_real(cuDoubleComplex a, double r) { return make_cuDoubleComplex(a.x / r, a.y / r); }

#ifdef __CUDACC__
__device__ inline float  shfl_xor_t(float v, int m)  { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ inline double shfl_xor_t(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ inline cuFloatComplex shfl_xor_t(cuFloatComplex v, int m)
{
    return make_cuFloatComplex(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
__device__ inline cuDoubleComplex shfl_xor_t(cuDoubleComplex v, int m)
{
    return make_cuDoubleComplex(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
}
#endif

// acc += a * b
__device__ inline void fma_acc(float& acc, float a, float b) { acc = fmaf(a, b, acc); }
__device__ inline void fma_acc(double& acc, double a, double b) { acc = fma(a, b, acc); }
__device__ inline void fma_acc(cuFloatComplex& acc, cuFloatComplex a, cuFloatComplex b)
{
    acc.x = fmaf(a.x, b.x, acc.x); acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y); acc.y = fmaf(a.y, b.x, acc.y);
}
__device__ inline void fma_acc(cuDoubleComplex& acc, cuDoubleComplex a, cuDoubleComplex b)
{
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}

} // namespace sb200
