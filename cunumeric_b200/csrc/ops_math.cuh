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

// binary64 -> binary32, round-to-nearest-even, in integer arithmetic.  Bit-identical to
// `static_cast<float>(double)` (cvt.rn.f32.f64), but that instruction (F2F.F32.F64) issues at about
// one result per 32 clocks per SM on B200 — measured 1.6 % of the HBM roofline for astype
// float64->float32 — whereas ~15 integer ops per element stay far below the memory time.
// (out-of-line: results that are not normal fp32 numbers — zero, subnormal, overflow, inf, nan)
static __device__ __noinline__ float d2f_rare(unsigned int hi, unsigned int lo)
{
  const unsigned int sign = hi & 0x80000000u;
  const unsigned int ex   = (hi >> 20) & 0x7ffu;
  const unsigned long long a =
    (static_cast<unsigned long long>(hi & 0x7fffffffu) << 32) | static_cast<unsigned long long>(lo);
  unsigned int r;
  if (a >= 0x7ff0000000000000ull) {  // inf / nan (nan stays quiet, payload truncated like the HW)
    r = (a > 0x7ff0000000000000ull)
          ? (0x7fc00000u | static_cast<unsigned int>((a >> 29) & 0x003fffffu))
          : 0x7f800000u;
  } else {
    const int e = static_cast<int>(ex) - 1023 + 127;  // biased fp32 exponent, <= 0 or >= 255 here
    if (e >= 255) {
      r = 0x7f800000u;  // overflow
    } else if (e < -24) {
      r = 0u;  // below half of the smallest subnormal (also zero and fp64 subnormals)
    } else {
      const unsigned long long full = (a & 0x000fffffffffffffull) | 0x0010000000000000ull;
      const int shift               = 29 + (1 - e);  // 30 .. 54
      const unsigned long long q    = full >> shift;
      const unsigned long long rem  = full & ((1ull << shift) - 1ull);
      const unsigned long long half = 1ull << (shift - 1);
      r = static_cast<unsigned int>(q) + ((rem > half || (rem == half && (q & 1ull))) ? 1u : 0u);
    }
  }
  return __uint_as_float(sign | r);
}

__device__ __forceinline__ float d2f(double d)
{
  const unsigned int hi = static_cast<unsigned int>(__double2hiint(d));
  const unsigned int lo = static_cast<unsigned int>(__double2loint(d));
  const unsigned int ex = (hi >> 20) & 0x7ffu;  // biased fp64 exponent
  // fast path (a dozen 32-bit ops, branch-free): the result is a normal fp32 number
  const unsigned int q   = ((hi & 0x000fffffu) << 3) | (lo >> 29);
  const unsigned int rem = lo & 0x1fffffffu;
  unsigned int r         = (hi & 0x80000000u) | (((ex - 896u) << 23) | q);
  // a carry out of the mantissa bumps the exponent (and rounds up to inf) correctly
  r += (rem > 0x10000000u || (rem == 0x10000000u && (q & 1u))) ? 1u : 0u;
  float out = __uint_as_float(r);
  if (__builtin_expect(ex - 897u >= 254u, 0)) out = d2f_rare(hi, lo);
  return out;
}

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
