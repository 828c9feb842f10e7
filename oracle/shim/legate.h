// Stand-in for legate.core 24.01's public header `legate.h` (external dependency of the
// reference, pinned at /root/reference/cmake/versions.json:3-9, NOT vendored in the tree).
//
// TEST INFRASTRUCTURE ONLY. This shim exists so that the reference's own functor headers
//   src/cunumeric/binary/binary_op_util.h, unary/unary_op_util.h, unary/unary_red_util.h,
//   unary/convert_util.h, arg.h, arg.inl, unary/isnan.h
// compile UNMODIFIED with g++ (included by path from /root/reference, never copied) into
// oracle/_ref/libcunumeric_ref.so.  Nothing in the product (cunumeric_b200/) includes this.
//
// What legate.core supplies and is re-stated here (all "external / unpinned" in DESIGN.md):
//   * __CUDA_HD__, LEGATE_MAX_DIM
//   * a host `__half` (Legion mathtypes/half.h gives CPU builds a software half: storage
//     uint16, arithmetic in float rounded once to nearest-even)
//   * global `complex<T>` with NumPy's lexicographic ordering
//   * legate::Type::Code, legate_type_of, is_integral/is_signed/is_floating_point/is_complex
//   * legate::Store::scalar<T>(), Point/Rect
//   * legate::{Sum,Prod,Max,Min}Reduction<T> (identity + fold) and the Legion:: aliases
#pragma once

#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <type_traits>
#include <vector>

#define __CUDA_HD__
#define LEGATE_MAX_DIM 4
#define LEGATE_ABORT abort()

// ----------------------------------------------------------------------------------------
// software half: IEEE binary16 storage, float arithmetic, one round-to-nearest-even.
// ----------------------------------------------------------------------------------------
namespace shim_detail {

inline uint16_t double_to_half_bits(double d)
{
  // Correctly rounded (single rounding) binary64 -> binary16, round-to-nearest-even.
  uint64_t u;
  std::memcpy(&u, &d, 8);
  const uint16_t sign = static_cast<uint16_t>((u >> 48) & 0x8000u);
  const int64_t exp   = static_cast<int64_t>((u >> 52) & 0x7ff);
  uint64_t man        = u & 0xfffffffffffffull;
  if (exp == 0x7ff) {  // inf / nan
    if (man == 0) return sign | 0x7c00u;
    return sign | 0x7e00u | static_cast<uint16_t>(man >> 42);
  }
  int64_t e = exp - 1023;  // unbiased
  if (exp == 0) return sign;  // double subnormals are far below half range -> +-0
  if (e > 15) return sign | 0x7c00u;  // overflow -> inf (values >= 65520 round to inf; see below)
  man |= (1ull << 52);  // implicit 1
  int shift;
  uint16_t hexp;
  if (e >= -14) {
    shift = 42;  // keep 10 fraction bits
    hexp  = static_cast<uint16_t>(e + 15);
  } else {
    // subnormal half: value = man * 2^(e-52); target unit 2^-24
    shift = static_cast<int>(42 + (-14 - e));
    hexp  = 0;
    if (shift > 63) return sign;  // underflow to zero (below half of min subnormal)
  }
  uint64_t q   = man >> shift;
  uint64_t rem = man & ((1ull << shift) - 1);
  uint64_t half = 1ull << (shift - 1);
  if (rem > half || (rem == half && (q & 1))) q++;
  // q holds implicit bit (if normal) + 10 fraction bits; adding with exponent handles carry
  uint32_t out;
  if (hexp == 0)
    out = static_cast<uint32_t>(q);  // subnormal (carry into exponent 1 is correct)
  else
    out = (static_cast<uint32_t>(hexp - 1) << 10) + static_cast<uint32_t>(q);
  if (out >= 0x7c00u) out = 0x7c00u;
  return sign | static_cast<uint16_t>(out);
}

inline float half_bits_to_float(uint16_t h)
{
  const uint32_t sign = (static_cast<uint32_t>(h) & 0x8000u) << 16;
  const uint32_t exp  = (h >> 10) & 0x1f;
  const uint32_t man  = h & 0x3ffu;
  uint32_t out;
  if (exp == 0) {
    if (man == 0) {
      out = sign;
    } else {
      float f = static_cast<float>(man) * 5.9604644775390625e-08f;  // 2^-24, exact
      uint32_t fb;
      std::memcpy(&fb, &f, 4);
      out = fb | sign;
    }
  } else if (exp == 0x1f) {
    out = sign | 0x7f800000u | (man << 13);
  } else {
    out = sign | ((exp + 112) << 23) | (man << 13);
  }
  float r;
  std::memcpy(&r, &out, 4);
  return r;
}

}  // namespace shim_detail

struct __half {
  uint16_t raw;
  __half() : raw(0) {}
  // float -> half is exact through double (float is a subset of double): single rounding.
  __half(float f) : raw(shim_detail::double_to_half_bits(static_cast<double>(f))) {}
  __half(double d) : raw(shim_detail::double_to_half_bits(d)) {}
  __half(int v) : raw(shim_detail::double_to_half_bits(static_cast<double>(v))) {}
  __half(long v) : raw(shim_detail::double_to_half_bits(static_cast<double>(v))) {}
  __half(long long v) : raw(shim_detail::double_to_half_bits(static_cast<double>(v))) {}
  __half(unsigned v) : raw(shim_detail::double_to_half_bits(static_cast<double>(v))) {}
  __half(unsigned long v) : raw(shim_detail::double_to_half_bits(static_cast<double>(v))) {}
  __half(unsigned long long v) : raw(shim_detail::double_to_half_bits(static_cast<double>(v))) {}
  __half(bool v) : raw(v ? 0x3c00 : 0) {}
  operator float() const { return shim_detail::half_bits_to_float(raw); }
};

inline __half operator+(const __half& a, const __half& b) { return __half(float(a) + float(b)); }
inline __half operator-(const __half& a, const __half& b) { return __half(float(a) - float(b)); }
inline __half operator*(const __half& a, const __half& b) { return __half(float(a) * float(b)); }
inline __half operator/(const __half& a, const __half& b) { return __half(float(a) / float(b)); }
inline __half operator-(const __half& a)
{
  __half r;
  r.raw = a.raw ^ 0x8000u;
  return r;
}
inline __half& operator+=(__half& a, const __half& b) { return a = a + b; }
inline __half& operator*=(__half& a, const __half& b) { return a = a * b; }
inline bool operator==(const __half& a, const __half& b) { return float(a) == float(b); }
inline bool operator!=(const __half& a, const __half& b) { return float(a) != float(b); }
inline bool operator<(const __half& a, const __half& b) { return float(a) < float(b); }
inline bool operator<=(const __half& a, const __half& b) { return float(a) <= float(b); }
inline bool operator>(const __half& a, const __half& b) { return float(a) > float(b); }
inline bool operator>=(const __half& a, const __half& b) { return float(a) >= float(b); }

#define SHIM_HALF_FN1(name) \
  inline __half name(const __half& a) { return __half(std::name(float(a))); }
SHIM_HALF_FN1(fabs)
SHIM_HALF_FN1(acos)
SHIM_HALF_FN1(asin)
SHIM_HALF_FN1(atan)
SHIM_HALF_FN1(ceil)
SHIM_HALF_FN1(cos)
SHIM_HALF_FN1(exp)
SHIM_HALF_FN1(floor)
SHIM_HALF_FN1(log)
SHIM_HALF_FN1(sin)
SHIM_HALF_FN1(sqrt)
SHIM_HALF_FN1(tan)
SHIM_HALF_FN1(tanh)
#undef SHIM_HALF_FN1
inline __half pow(const __half& a, const __half& b) { return __half(std::pow(float(a), float(b))); }
inline bool isnan(const __half& a) { return std::isnan(float(a)); }
inline bool isinf(const __half& a) { return std::isinf(float(a)); }
inline bool isfinite(const __half& a) { return std::isfinite(float(a)); }

// ----------------------------------------------------------------------------------------
// complex<T>: std::complex plus NumPy's lexicographic ordering (legate.core supplies the
// ordering the reference relies on at binary_op_util.h:431-442,571-581 and unary_op_util.h:418).
// ----------------------------------------------------------------------------------------
template <typename T>
struct complex : public std::complex<T> {
  using base = std::complex<T>;
  constexpr complex() : base() {}
  constexpr complex(const base& b) : base(b) {}
  constexpr complex(T re, T im) : base(re, im) {}
  template <typename U, std::enable_if_t<std::is_arithmetic<U>::value>* = nullptr>
  constexpr complex(U re) : base(static_cast<T>(re), T(0))
  {
  }
  template <typename U, std::enable_if_t<!std::is_same<U, T>::value>* = nullptr>
  explicit constexpr complex(const complex<U>& o)
    : base(static_cast<T>(o.real()), static_cast<T>(o.imag()))
  {
  }
};

template <typename T>
inline bool operator<(const complex<T>& a, const complex<T>& b)
{
  return a.real() < b.real() || (a.real() == b.real() && a.imag() < b.imag());
}
template <typename T>
inline bool operator>(const complex<T>& a, const complex<T>& b)
{
  return b < a;
}
template <typename T>
inline bool operator<=(const complex<T>& a, const complex<T>& b)
{
  return a.real() < b.real() || (a.real() == b.real() && a.imag() <= b.imag());
}
template <typename T>
inline bool operator>=(const complex<T>& a, const complex<T>& b)
{
  return b <= a;
}

// ----------------------------------------------------------------------------------------
namespace legate {

struct Type {
  // Same ordering as Legion's legion_type_id_t (LEGION_TYPE_BOOL = 0 ...).
  enum class Code : int32_t {
    BOOL = 0,
    INT8,
    INT16,
    INT32,
    INT64,
    UINT8,
    UINT16,
    UINT32,
    UINT64,
    FLOAT16,
    FLOAT32,
    FLOAT64,
    COMPLEX64,
    COMPLEX128,
  };
};

template <Type::Code CODE>
struct LegateTypeOf;
#define SHIM_TYPE(CODE, T)           \
  template <>                        \
  struct LegateTypeOf<Type::Code::CODE> { \
    using type = T;                  \
  };
SHIM_TYPE(BOOL, bool)
SHIM_TYPE(INT8, int8_t)
SHIM_TYPE(INT16, int16_t)
SHIM_TYPE(INT32, int32_t)
SHIM_TYPE(INT64, int64_t)
SHIM_TYPE(UINT8, uint8_t)
SHIM_TYPE(UINT16, uint16_t)
SHIM_TYPE(UINT32, uint32_t)
SHIM_TYPE(UINT64, uint64_t)
SHIM_TYPE(FLOAT16, __half)
SHIM_TYPE(FLOAT32, float)
SHIM_TYPE(FLOAT64, double)
SHIM_TYPE(COMPLEX64, ::complex<float>)
SHIM_TYPE(COMPLEX128, ::complex<double>)
#undef SHIM_TYPE

template <Type::Code CODE>
using legate_type_of = typename LegateTypeOf<CODE>::type;

template <Type::Code CODE>
struct is_integral {
  static constexpr bool value = std::is_integral<legate_type_of<CODE>>::value;
};
template <Type::Code CODE>
struct is_signed {
  static constexpr bool value = std::is_signed<legate_type_of<CODE>>::value;
};
template <>
struct is_signed<Type::Code::FLOAT16> {
  static constexpr bool value = true;
};
template <Type::Code CODE>
struct is_unsigned {
  static constexpr bool value = std::is_unsigned<legate_type_of<CODE>>::value;
};
template <Type::Code CODE>
struct is_floating_point {
  static constexpr bool value = std::is_floating_point<legate_type_of<CODE>>::value;
};
template <>
struct is_floating_point<Type::Code::FLOAT16> {
  static constexpr bool value = true;
};
template <Type::Code CODE>
struct is_complex : std::false_type {};
template <>
struct is_complex<Type::Code::COMPLEX64> : std::true_type {};
template <>
struct is_complex<Type::Code::COMPLEX128> : std::true_type {};
template <typename T>
struct is_complex_type : std::false_type {};
template <>
struct is_complex_type<::complex<float>> : std::true_type {};
template <>
struct is_complex_type<::complex<double>> : std::true_type {};

// Scalar-carrying store: the only Store API the functor headers touch is scalar<T>().
class Store {
 public:
  Store() : data_(nullptr) {}
  explicit Store(const void* p) : data_(p) {}
  template <typename T>
  T scalar() const
  {
    T v;
    std::memcpy(&v, data_, sizeof(T));
    return v;
  }

 private:
  const void* data_;
};

template <int DIM, typename T = int64_t>
struct Point {
  T x[DIM > 0 ? DIM : 1];
  T& operator[](int i) { return x[i]; }
  const T& operator[](int i) const { return x[i]; }
};
template <int DIM, typename T = int64_t>
struct Rect {
  Point<DIM, T> lo, hi;
};

// Legion reduction operators (external): identity + fold.
template <typename T>
struct SumReduction {
  using LHS = T;
  using RHS = T;
  static const T identity;
  template <bool EXCLUSIVE>
  static void fold(T& a, T b)
  {
    a = a + b;
  }
};
template <>
template <bool EXCLUSIVE>
inline void SumReduction<bool>::fold(bool& a, bool b)
{
  a = a || b;
}
template <typename T>
struct ProdReduction {
  using LHS = T;
  using RHS = T;
  static const T identity;
  template <bool EXCLUSIVE>
  static void fold(T& a, T b)
  {
    a = a * b;
  }
};
template <>
template <bool EXCLUSIVE>
inline void ProdReduction<bool>::fold(bool& a, bool b)
{
  a = a && b;
}
template <typename T>
struct MaxReduction {
  using LHS = T;
  using RHS = T;
  static const T identity;
  template <bool EXCLUSIVE>
  static void fold(T& a, T b)
  {
    if (b > a) a = b;
  }
};
template <typename T>
struct MinReduction {
  using LHS = T;
  using RHS = T;
  static const T identity;
  template <bool EXCLUSIVE>
  static void fold(T& a, T b)
  {
    if (b < a) a = b;
  }
};

namespace shim_ident {
template <typename T, typename = void>
struct Lim {
  static T lowest() { return std::numeric_limits<T>::lowest(); }
  static T highest() { return std::numeric_limits<T>::max(); }
};
template <typename T>
struct Lim<T, std::enable_if_t<std::is_floating_point<T>::value>> {
  static T lowest() { return -std::numeric_limits<T>::infinity(); }
  static T highest() { return std::numeric_limits<T>::infinity(); }
};
template <>
struct Lim<__half, void> {
  static __half lowest()
  {
    __half h;
    h.raw = 0xfc00;
    return h;
  }
  static __half highest()
  {
    __half h;
    h.raw = 0x7c00;
    return h;
  }
};
}  // namespace shim_ident

template <typename T>
const T SumReduction<T>::identity = T(0);
template <typename T>
const T ProdReduction<T>::identity = T(1);
template <typename T>
const T MaxReduction<T>::identity = shim_ident::Lim<T>::lowest();
template <typename T>
const T MinReduction<T>::identity = shim_ident::Lim<T>::highest();

class TaskContext {};
class TaskRegistrar {};
template <typename T>
struct LegateTask {};

}  // namespace legate

namespace Legion {
template <int DIM, typename T = int64_t>
using Point = legate::Point<DIM, T>;
template <typename T>
using SumReduction = legate::SumReduction<T>;
}  // namespace Legion
