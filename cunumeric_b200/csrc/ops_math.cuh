// Scalar math helpers shared by the op functors.
//
// fp16 never computes natively: every fp16 op converts to fp32, computes, and rounds ONCE with
// round-to-nearest-even — the reference's CPU semantics (Legion's software half; e.g.
// binary_op_util.h:141-147 `lift`, unary_op_util.h:266-278).  For + - * / that is bit-identical to a
// correctly rounded fp16 op (24 >= 2*11+2).
#pragma once

#include "cnb_common.cuh"

#include <cuda/std/limits>

namespace cnb {

__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }
__device__ __forceinline__ __half d2h(double d) { return __double2half(d); }

// Type the arithmetic of T is carried out in
template <typename T>
struct ComputeT {
  using type = T;
};
template <>
struct ComputeT<__half> {
  using type = float;
};
template <typename T>
using compute_t = typename ComputeT<T>::type;

template <typename T>
__device__ __forceinline__ compute_t<T> up(T x)
{
  if constexpr (is_half_v<T>)
    return h2f(x);
  else
    return x;
}
template <typename T>
__device__ __forceinline__ T down(compute_t<T> x)
{
  if constexpr (is_half_v<T>)
    return f2h(x);
  else
    return x;
}

// truthiness: `static_cast<bool>(x)`; complex looks only at the real part
// (binary_op_util.h:656-666, unary_op_util.h:865-875)
template <typename T>
__device__ __forceinline__ bool truth(const T& x)
{
  if constexpr (is_complex_v<T>)
    return x.real() != 0;
  else if constexpr (is_half_v<T>)
    return h2f(x) != 0.0f;
  else
    return static_cast<bool>(x);
}

// NumPy's lexicographic complex ordering, plain < for everything else (fp16 through fp32)
template <typename T>
__device__ __forceinline__ bool lt(const T& a, const T& b)
{
  if constexpr (is_complex_v<T>)
    return a.real() < b.real() || (a.real() == b.real() && a.imag() < b.imag());
  else if constexpr (is_half_v<T>)
    return h2f(a) < h2f(b);
  else
    return a < b;
}
template <typename T>
__device__ __forceinline__ bool le(const T& a, const T& b)
{
  if constexpr (is_complex_v<T>)
    return a.real() < b.real() || (a.real() == b.real() && a.imag() <= b.imag());
  else if constexpr (is_half_v<T>)
    return h2f(a) <= h2f(b);
  else
    return a <= b;
}
template <typename T>
__device__ __forceinline__ bool eq(const T& a, const T& b)
{
  if constexpr (is_complex_v<T>)
    return a.real() == b.real() && a.imag() == b.imag();
  else if constexpr (is_half_v<T>)
    return h2f(a) == h2f(b);
  else
    return a == b;
}

template <typename T>
__device__ __forceinline__ bool isnan_any(const T& x)
{
  if constexpr (is_complex_v<T>)
    return isnan(x.real()) || isnan(x.imag());
  else if constexpr (is_half_v<T>)
    return __hisnan(x);
  else if constexpr (std::is_floating_point<T>::value)
    return isnan(x);
  else
    return false;
}

}  // namespace cnb
