// UNARY_OP functors for sm_100a — one struct per CuNumericUnaryOpCode.
//
// Semantics follow the reference's CPU variant (the parity oracle), cited as
// unary_op_util.h:<line>, including its deviations from NumPy: RECIPROCAL(0)=0,
// SIGN(NaN)=0, complex SIGN / LOGICAL_NOT rules, RAD2DEG in double, fp16 math in fp32 + one rounding.
// Integer inputs to EXP / SQRT / RINT produce float64, as the C++ functors do.
#pragma once

#include "ops_math.cuh"

namespace cnb {
namespace uop {

template <typename T>
struct Base {
  using Out = T;
  __host__ __device__ Base() {}
  __host__ Base(const void*) {}
};
template <typename T>
struct BoolBase {
  using Out = bool;
  __host__ __device__ BoolBase() {}
  __host__ BoolBase(const void*) {}
};

// ---- float-or-complex math: e f d F D (fp16 in fp32) -------------------------------------------
// F32 / F64 / CPLX are expressions over `x`
#define CNB_FC_UNOP(NAME, F32, F64, CPLX)                                   \
  template <typename T>                                                     \
  struct NAME : Base<T> {                                                   \
    static constexpr bool valid = is_float_v<T> || is_complex_v<T>;         \
    using Base<T>::Base;                                                    \
    __device__ __forceinline__ T operator()(const T& x_) const              \
    {                                                                       \
      if constexpr (is_half_v<T>) {                                         \
        float x = h2f(x_);                                                  \
        return f2h(F32);                                                    \
      } else if constexpr (std::is_same<T, float>::value) {                 \
        float x = x_;                                                       \
        return F32;                                                         \
      } else if constexpr (std::is_same<T, double>::value) {                \
        double x = x_;                                                      \
        return F64;                                                         \
      } else if constexpr (is_complex_v<T>) {                               \
        const T& x = x_;                                                    \
        return CPLX;                                                        \
      } else                                                                \
        return x_;                                                          \
    }                                                                       \
  };
CNB_FC_UNOP(Arccos, acosf(x), acos(x), cuda::std::acos(x))
CNB_FC_UNOP(Arccosh, acoshf(x), acosh(x), cuda::std::acosh(x))
CNB_FC_UNOP(Arcsin, asinf(x), asin(x), cuda::std::asin(x))
CNB_FC_UNOP(Arcsinh, asinhf(x), asinh(x), cuda::std::asinh(x))
CNB_FC_UNOP(Arctan, atanf(x), atan(x), cuda::std::atan(x))
CNB_FC_UNOP(Arctanh, atanhf(x), atanh(x), cuda::std::atanh(x))
CNB_FC_UNOP(Cos, cosf(x), cos(x), cuda::std::cos(x))
CNB_FC_UNOP(Cosh, coshf(x), cosh(x), cuda::std::cosh(x))
CNB_FC_UNOP(Sin, sinf(x), sin(x), cuda::std::sin(x))
// sinhf/tanhf are 3-ulp functions on the device; evaluate in double and round once so fp32
// results stay within the 2-ulp parity budget against the CPU libm
CNB_FC_UNOP(Sinh, d2f(sinh(static_cast<double>(x))), sinh(x), cuda::std::sinh(x))
CNB_FC_UNOP(Tan, tanf(x), tan(x), cuda::std::tan(x))
CNB_FC_UNOP(Tanh, d2f(tanh(static_cast<double>(x))), tanh(x), cuda::std::tanh(x))
CNB_FC_UNOP(Log, logf(x), log(x), cuda::std::log(x))
CNB_FC_UNOP(Log10, log10f(x), log10(x), cuda::std::log10(x))
// complex variants restate :546-557, :588-593, :799-804, :836-841
CNB_FC_UNOP(Exp2, exp2f(x), exp2(x),
            cuda::std::exp(T(static_cast<typename T::value_type>(0.69314718055994530942), 0) * x))
CNB_FC_UNOP(Expm1, expm1f(x), expm1(x), cuda::std::exp(x) - T(1))
CNB_FC_UNOP(Log1p, log1pf(x), log1p(x), cuda::std::log(T(1) + x))
CNB_FC_UNOP(Log2, log2f(x), log2(x), cuda::std::log(x) / cuda::std::log(T(2)))
#undef CNB_FC_UNOP

// EXP (:519-531), SQRT (:1115-1127): valid for every dtype; integers (incl. bool) compute in double
#define CNB_ALL_MATH_UNOP(NAME, F32, F64, CPLX)                                         \
  template <typename T>                                                                 \
  struct NAME {                                                                         \
    static constexpr bool valid = true;                                                 \
    using Out = std::conditional_t<std::is_integral<T>::value, double, T>;              \
    __host__ __device__ NAME(const void* = nullptr) {}                                                       \
    __device__ __forceinline__ Out operator()(const T& x_) const                        \
    {                                                                                   \
      if constexpr (is_half_v<T>) {                                                     \
        float x = h2f(x_);                                                              \
        return f2h(F32);                                                                \
      } else if constexpr (std::is_same<T, float>::value) {                             \
        float x = x_;                                                                   \
        return F32;                                                                     \
      } else if constexpr (is_complex_v<T>) {                                           \
        const T& x = x_;                                                                \
        return CPLX;                                                                    \
      } else {                                                                          \
        double x = static_cast<double>(x_);                                             \
        return F64;                                                                     \
      }                                                                                 \
    }                                                                                   \
  };
CNB_ALL_MATH_UNOP(Exp, expf(x), exp(x), cuda::std::exp(x))
CNB_ALL_MATH_UNOP(Sqrt, sqrtf(x), sqrt(x), cuda::std::sqrt(x))
// RINT (:948-980): complex rounds component-wise
CNB_ALL_MATH_UNOP(Rint, rintf(x), rint(x), T(rint(x.real()), rint(x.imag())))
#undef CNB_ALL_MATH_UNOP

// ---- real-float only: e f d --------------------------------------------------------------------
#define CNB_F_UNOP(NAME, F32, F64)                                          \
  template <typename T>                                                     \
  struct NAME : Base<T> {                                                   \
    static constexpr bool valid = is_float_v<T>;                            \
    using Base<T>::Base;                                                    \
    __device__ __forceinline__ T operator()(const T& x_) const              \
    {                                                                       \
      if constexpr (is_half_v<T>) {                                         \
        float x = h2f(x_);                                                  \
        return f2h(F32);                                                    \
      } else if constexpr (std::is_same<T, float>::value) {                 \
        float x = x_;                                                       \
        return F32;                                                         \
      } else if constexpr (std::is_same<T, double>::value) {                \
        double x = x_;                                                      \
        return F64;                                                         \
      } else                                                                \
        return x_;                                                          \
    }                                                                       \
  };
CNB_F_UNOP(Cbrt, cbrtf(x), cbrt(x))
CNB_F_UNOP(Ceil, ceilf(x), ceil(x))
CNB_F_UNOP(Floor, floorf(x), floor(x))
CNB_F_UNOP(Trunc, truncf(x), trunc(x))
// DEG2RAD (:496-517): multiply by a T-precision constant
CNB_F_UNOP(Deg2rad, x * static_cast<float>(3.14159265358979323846 / 180.0),
           x * (3.14159265358979323846 / 180.0))
// RAD2DEG (:888-909): fp32/fp64 evaluate x*180.0/M_PI in double and narrow; fp16 uses one fp32
// multiply by float(180/pi)
template <typename T>
struct Rad2deg : Base<T> {
  static constexpr bool valid = is_float_v<T>;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const
  {
    if constexpr (is_half_v<T>)
      return f2h(h2f(x) * static_cast<float>(180.0 / 3.14159265358979323846));
    else if constexpr (std::is_same<T, float>::value)
      return d2f(static_cast<double>(x) * 180.0 / 3.14159265358979323846);
    else if constexpr (std::is_floating_point<T>::value)
      return x * 180.0 / 3.14159265358979323846;
    else
      return x;
  }
};
#undef CNB_F_UNOP

// ---- every dtype -------------------------------------------------------------------------------
// ABSOLUTE (:199-236): complex -> real magnitude
template <typename T>
struct Absolute {
  static constexpr bool valid = true;
  template <typename U>
  struct OutOf {
    using type = U;
  };
  template <typename V>
  struct OutOf<cuda::std::complex<V>> {
    using type = V;
  };
  using Out = typename OutOf<T>::type;
  __host__ __device__ Absolute(const void* = nullptr) {}
  __device__ __forceinline__ Out operator()(const T& x) const
  {
    if constexpr (is_complex_v<T>)
      return cuda::std::abs(x);
    else if constexpr (is_signed_int_v<T>)
      return x >= 0 ? x : static_cast<T>(-x);
    else if constexpr (std::is_integral<T>::value)
      return x;
    else if constexpr (is_half_v<T>)
      return __habs(x);
    else
      return fabs(x);
  }
};

// CLIP (:406-422): min/max are two scalar stores of the array dtype
template <typename T>
struct Clip {
  static constexpr bool valid = true;
  using Out = T;
  T lo, hi;
  __host__ Clip(const void* extra)
  {
    if (extra != nullptr) {
      memcpy(&lo, extra, sizeof(T));
      memcpy(&hi, static_cast<const char*>(extra) + sizeof(T), sizeof(T));
    } else {
      memset(&lo, 0, sizeof(T));
      memset(&hi, 0, sizeof(T));
    }
  }
  __device__ __forceinline__ T operator()(const T& x) const
  {
    return lt(x, lo) ? lo : (lt(hi, x) ? hi : x);
  }
};

// CONJ (:424-442), COPY (:444-452; POSITIVE aliases it :146-148)
template <typename T>
struct Conj : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const
  {
    if constexpr (is_complex_v<T>)
      return T(x.real(), -x.imag());
    else
      return x;
  }
};
template <typename T>
struct Copy : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const { return x; }
};

// NEGATIVE (:878-886)
template <typename T>
struct Negative : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const
  {
    if constexpr (is_bool_v<T>)
      return x;  // bool(-int(x))
    else if constexpr (is_half_v<T>)
      return __hneg(x);
    else if constexpr (std::is_integral<T>::value)
      return static_cast<T>(0 - static_cast<std::make_unsigned_t<T>>(x));
    else
      return -x;
  }
};

// SQUARE (:1105-1113)
template <typename T>
struct Square : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const
  {
    if constexpr (is_bool_v<T>)
      return x;
    else if constexpr (is_half_v<T>)
      return f2h(h2f(x) * h2f(x));
    else if constexpr (std::is_integral<T>::value) {
      using W = std::conditional_t<(sizeof(T) < 8), unsigned int, unsigned long long>;
      return static_cast<T>(static_cast<W>(x) * static_cast<W>(x));
    } else
      return x * x;
  }
};

// RECIPROCAL (:921-946): 1/0 -> 0; integers use integer division
template <typename T>
struct Reciprocal : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const
  {
    if constexpr (is_bool_v<T>)
      return x;
    else if constexpr (is_half_v<T>)
      return h2f(x) != 0.0f ? f2h(1.0f / h2f(x)) : f2h(0.0f);
    else if constexpr (is_complex_v<T>)
      return (x.real() != 0 || x.imag() != 0) ? T(1) / x : T(0);
    else if constexpr (std::is_integral<T>::value)
      return x != T(0) ? static_cast<T>(T(1) / x) : T(0);
    else
      return x != T(0) ? T(1) / x : T(0);
  }
};

// SIGN (:982-1033): NaN -> 0; complex: sign of real part, else of imaginary part, result (s, 0)
template <typename V>
__device__ __forceinline__ V sign_of(V x)
{
  if constexpr (std::is_signed<V>::value || std::is_floating_point<V>::value)
    return x > V(0) ? V(1) : (x < V(0) ? V(-1) : V(0));
  else
    return x > V(0) ? V(1) : V(0);
}
template <typename T>
struct Sign : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const
  {
    if constexpr (is_complex_v<T>) {
      if (x.real() != 0) return T(sign_of(x.real()), 0);
      return T(sign_of(x.imag()), 0);
    } else if constexpr (is_half_v<T>)
      return f2h(sign_of<float>(h2f(x)));
    else if constexpr (is_bool_v<T>)
      return x;
    else
      return sign_of<T>(x);
  }
};

// ISFINITE / ISINF / ISNAN (:655-738), LOGICAL_NOT (:858-876), SIGNBIT (:1035-1061)
template <typename T>
struct Isfinite : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& x) const
  {
    if constexpr (is_complex_v<T>)
      return isfinite(x.real()) && isfinite(x.imag());
    else if constexpr (std::is_integral<T>::value)
      return true;
    else
      return isfinite(up(x));
  }
};
template <typename T>
struct Isinf : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& x) const
  {
    if constexpr (is_complex_v<T>)
      return isinf(x.real()) || isinf(x.imag());
    else if constexpr (std::is_integral<T>::value)
      return false;
    else
      return isinf(up(x));
  }
};
template <typename T>
struct Isnan : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& x) const { return isnan_any(x); }
};
template <typename T>
struct LogicalNot : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& x) const { return !truth(x); }
};
template <typename T>
struct Signbit : BoolBase<T> {
  static constexpr bool valid = is_float_v<T>;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& x) const
  {
    if constexpr (is_float_v<T>)
      return signbit(up(x));
    else
      return false;
  }
};

// INVERT (:644-653)
template <typename T>
struct Invert : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value && !is_bool_v<T>;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& x) const
  {
    if constexpr (std::is_integral<T>::value && !is_bool_v<T>)
      return static_cast<T>(~x);
    else
      return x;
  }
};

// REAL (:911-919), IMAG (:634-642)
template <typename T>
struct Real {
  static constexpr bool valid = is_complex_v<T>;
  using Out = typename Absolute<T>::Out;
  __host__ __device__ Real(const void* = nullptr) {}
  __device__ __forceinline__ Out operator()(const T& x) const
  {
    if constexpr (is_complex_v<T>)
      return x.real();
    else
      return x;
  }
};
template <typename T>
struct Imag {
  static constexpr bool valid = is_complex_v<T>;
  using Out = typename Absolute<T>::Out;
  __host__ __device__ Imag(const void* = nullptr) {}
  __device__ __forceinline__ Out operator()(const T& x) const
  {
    if constexpr (is_complex_v<T>)
      return x.imag();
    else
      return x;
  }
};

// ---- two outputs: FREXP (:1190-1216), MODF (:1218-1247) ----------------------------------------
template <typename T>
struct Frexp {
  static constexpr bool valid = is_float_v<T>;
  using Out  = T;
  using Out2 = int32_t;
  __device__ __forceinline__ void operator()(T& o, int32_t& e, const T& x) const
  {
    int ex = 0;
    if constexpr (is_half_v<T>)
      o = f2h(frexpf(h2f(x), &ex));
    else if constexpr (std::is_same<T, float>::value)
      o = frexpf(x, &ex);
    else if constexpr (std::is_same<T, double>::value)
      o = frexp(x, &ex);
    else
      o = x;
    e = ex;
  }
};
template <typename T>
struct Modf {
  static constexpr bool valid = is_float_v<T>;
  using Out  = T;
  using Out2 = T;
  __device__ __forceinline__ void operator()(T& o, T& ip, const T& x) const
  {
    if constexpr (is_half_v<T>) {
      float t;
      o  = f2h(modff(h2f(x), &t));
      ip = f2h(t);
    } else if constexpr (std::is_same<T, float>::value) {
      float t;
      o  = modff(x, &t);
      ip = t;
    } else if constexpr (std::is_same<T, double>::value) {
      double t;
      o  = modf(x, &t);
      ip = t;
    } else {
      o  = x;
      ip = x;
    }
  }
};

}  // namespace uop

template <int OP>
struct UnaryFn;
#define CNB_UN(OPCODE, NAME)      \
  template <>                     \
  struct UnaryFn<OPCODE> {        \
    template <typename T>         \
    using fn = uop::NAME<T>;      \
  };
CNB_UN(CNB_UOP_ABSOLUTE, Absolute)
CNB_UN(CNB_UOP_ARCCOS, Arccos)
CNB_UN(CNB_UOP_ARCCOSH, Arccosh)
CNB_UN(CNB_UOP_ARCSIN, Arcsin)
CNB_UN(CNB_UOP_ARCSINH, Arcsinh)
CNB_UN(CNB_UOP_ARCTAN, Arctan)
CNB_UN(CNB_UOP_ARCTANH, Arctanh)
CNB_UN(CNB_UOP_CBRT, Cbrt)
CNB_UN(CNB_UOP_CEIL, Ceil)
CNB_UN(CNB_UOP_CLIP, Clip)
CNB_UN(CNB_UOP_CONJ, Conj)
CNB_UN(CNB_UOP_COPY, Copy)
CNB_UN(CNB_UOP_COS, Cos)
CNB_UN(CNB_UOP_COSH, Cosh)
CNB_UN(CNB_UOP_DEG2RAD, Deg2rad)
CNB_UN(CNB_UOP_EXP, Exp)
CNB_UN(CNB_UOP_EXP2, Exp2)
CNB_UN(CNB_UOP_EXPM1, Expm1)
CNB_UN(CNB_UOP_FLOOR, Floor)
CNB_UN(CNB_UOP_IMAG, Imag)
CNB_UN(CNB_UOP_INVERT, Invert)
CNB_UN(CNB_UOP_ISFINITE, Isfinite)
CNB_UN(CNB_UOP_ISINF, Isinf)
CNB_UN(CNB_UOP_ISNAN, Isnan)
CNB_UN(CNB_UOP_LOG, Log)
CNB_UN(CNB_UOP_LOG10, Log10)
CNB_UN(CNB_UOP_LOG1P, Log1p)
CNB_UN(CNB_UOP_LOG2, Log2)
CNB_UN(CNB_UOP_LOGICAL_NOT, LogicalNot)
CNB_UN(CNB_UOP_NEGATIVE, Negative)
CNB_UN(CNB_UOP_POSITIVE, Copy)
CNB_UN(CNB_UOP_RAD2DEG, Rad2deg)
CNB_UN(CNB_UOP_REAL, Real)
CNB_UN(CNB_UOP_RECIPROCAL, Reciprocal)
CNB_UN(CNB_UOP_RINT, Rint)
CNB_UN(CNB_UOP_SIGN, Sign)
CNB_UN(CNB_UOP_SIGNBIT, Signbit)
CNB_UN(CNB_UOP_SIN, Sin)
CNB_UN(CNB_UOP_SINH, Sinh)
CNB_UN(CNB_UOP_SQRT, Sqrt)
CNB_UN(CNB_UOP_SQUARE, Square)
CNB_UN(CNB_UOP_TAN, Tan)
CNB_UN(CNB_UOP_TANH, Tanh)
CNB_UN(CNB_UOP_TRUNC, Trunc)
#undef CNB_UN

}  // namespace cnb
