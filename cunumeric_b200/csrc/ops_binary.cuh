// BINARY_OP functors for sm_100a — one struct per CuNumericBinaryOpCode.
//
// Semantics follow the reference's CPU variant (the parity oracle) — cited per op as
// binary_op_util.h:<line> — including the places where it deliberately differs from NumPy
// (SURVEY App. A.4).  Each functor exposes: valid, Out (result type), Rhs2 (second operand type),
// a constructor from the task's extra scalars, and a __device__ operator().
#pragma once

#include "ops_math.cuh"

namespace cnb {
namespace bop {

template <typename T>
struct Base {
  using Out  = T;
  using Rhs2 = T;
  __host__ __device__ Base() {}
  __host__ Base(const void*) {}
};
template <typename T>
struct BoolBase {
  using Out  = bool;
  using Rhs2 = T;
  __host__ __device__ BoolBase() {}
  __host__ BoolBase(const void*) {}
};

// integer arithmetic happens in the promoted type and wraps back to T like the C++ functors
template <typename T>
__device__ __forceinline__ T wrap(long long v)
{
  return static_cast<T>(v);
}

// ---- ADD / SUBTRACT / MULTIPLY: std::plus/minus/multiplies<T> (:169, :874, :788) -------------
template <typename T>
struct Add : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (is_bool_v<T>)
      return a || b;  // bool(int(a)+int(b))
    else if constexpr (is_half_v<T>)
      return f2h(h2f(a) + h2f(b));
    else if constexpr (std::is_integral<T>::value)
      return static_cast<T>(static_cast<std::make_unsigned_t<T>>(a) +
                            static_cast<std::make_unsigned_t<T>>(b));
    else
      return a + b;
  }
};
template <typename T>
struct Subtract : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (is_bool_v<T>)
      return a != b;  // bool(int(a)-int(b))
    else if constexpr (is_half_v<T>)
      return f2h(h2f(a) - h2f(b));
    else if constexpr (std::is_integral<T>::value)
      return static_cast<T>(static_cast<std::make_unsigned_t<T>>(a) -
                            static_cast<std::make_unsigned_t<T>>(b));
    else
      return a - b;
  }
};
template <typename T>
struct Multiply : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (is_bool_v<T>)
      return a && b;
    else if constexpr (is_half_v<T>)
      return f2h(h2f(a) * h2f(b));
    else if constexpr (std::is_integral<T>::value) {
      using W = std::conditional_t<(sizeof(T) < 8), unsigned int, unsigned long long>;
      return static_cast<T>(static_cast<W>(a) * static_cast<W>(b));
    } else
      return a * b;
  }
};

// ---- DIVIDE (:266-282): integers (incl. bool) divide in double ------------------------------
template <typename T>
struct Divide {
  static constexpr bool valid = true;
  using Out  = std::conditional_t<std::is_integral<T>::value, double, T>;
  using Rhs2 = T;
  __host__ __device__ Divide(const void* = nullptr) {}
  __device__ __forceinline__ Out operator()(const T& a, const T& b) const
  {
    if constexpr (std::is_integral<T>::value)
      return static_cast<double>(a) / static_cast<double>(b);
    else if constexpr (is_half_v<T>)
      return f2h(h2f(a) / h2f(b));
    else
      return a / b;
  }
};

// ---- comparisons (:285, :820, :432, :438, :572, :578) ----------------------------------------
template <typename T>
struct Equal : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const { return eq(a, b); }
};
template <typename T>
struct NotEqual : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const { return !eq(a, b); }
};
template <typename T>
struct Less : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const { return lt(a, b); }
};
template <typename T>
struct LessEqual : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const { return le(a, b); }
};
template <typename T>
struct Greater : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const { return lt(b, a); }
};
template <typename T>
struct GreaterEqual : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const { return le(b, a); }
};

// ---- logical (:651-706): complex uses only .real() -------------------------------------------
template <typename T>
struct LogicalAnd : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const
  {
    return truth(a) && truth(b);
  }
};
template <typename T>
struct LogicalOr : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const
  {
    return truth(a) || truth(b);
  }
};
template <typename T>
struct LogicalXor : BoolBase<T> {
  static constexpr bool valid = true;
  using BoolBase<T>::BoolBase;
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const
  {
    return truth(a) != truth(b);
  }
};

// ---- MAXIMUM / MINIMUM (:709-722): std::max/std::min, NOT NaN-propagating ---------------------
template <typename T>
struct Maximum : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const { return lt(a, b) ? b : a; }
};
template <typename T>
struct Minimum : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const { return lt(b, a) ? b : a; }
};

// ---- bitwise (:201-237), shifts (:555-569, :864-871) -----------------------------------------
template <typename T>
struct BitwiseAnd : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    return static_cast<T>(a & b);
  }
};
template <typename T>
struct BitwiseOr : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    return static_cast<T>(a | b);
  }
};
template <typename T>
struct BitwiseXor : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    return static_cast<T>(a ^ b);
  }
};
template <typename T>
struct LeftShift : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value && !is_bool_v<T>;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    // host variant: (a << b) * (b >= 0); shifts happen in the promoted type
    using P = decltype(a << b);
    if constexpr (std::is_signed<T>::value) {
      if (b < 0) return T(0);
    }
    using UP = std::make_unsigned_t<P>;
    return static_cast<T>(static_cast<P>(static_cast<UP>(static_cast<P>(a)) << b));
  }
};
template <typename T>
struct RightShift : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value && !is_bool_v<T>;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    return static_cast<T>(a >> b);
  }
};

// ---- FLOOR_DIVIDE (:383-429), MOD (:724-785), FMOD (:318-348) ---------------------------------
template <typename T>
__device__ __forceinline__ auto floor_div_signed(T a, T b)
{
  auto q = a / b;  // promoted type for small ints
  return q - static_cast<decltype(q)>((((a < 0) != (b < 0)) && q * b != a));
}

template <typename F>
__device__ __forceinline__ F real_mod(F a, F b)
{
  F res = fmod(a, b);
  if (res != F(0)) {
    if ((b < F(0)) != (res < F(0))) res += b;
  } else {
    res = copysign(F(0), b);
  }
  return res;
}

template <typename T>
struct FloorDivide : Base<T> {
  static constexpr bool valid = !is_complex_v<T>;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (is_signed_int_v<T>)
      return static_cast<T>(floor_div_signed(a, b));
    else if constexpr (std::is_integral<T>::value)
      return static_cast<T>(a / b);
    else if constexpr (is_half_v<T>)
      // no fp16 specialisation in the reference: half division, then half floor (two roundings)
      return f2h(floorf(h2f(f2h(h2f(a) / h2f(b)))));
    else if constexpr (is_complex_v<T>)
      return a;
    else
      return floor(a / b);
  }
};
template <typename T>
struct Mod : Base<T> {
  static constexpr bool valid = !is_complex_v<T>;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (is_signed_int_v<T>) {
      auto q = floor_div_signed(a, b);
      return static_cast<T>((a - b * q) * (b != 0));
    } else if constexpr (std::is_integral<T>::value)
      return static_cast<T>(a % b);
    else if constexpr (is_half_v<T>)
      return f2h(real_mod<float>(h2f(a), h2f(b)));
    else if constexpr (is_complex_v<T>)
      return a;
    else
      return real_mod<T>(a, b);
  }
};
template <typename T>
struct Fmod : Base<T> {
  static constexpr bool valid = !is_bool_v<T> && !is_complex_v<T>;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (std::is_integral<T>::value)
      return static_cast<T>(a % b);
    else if constexpr (is_half_v<T>)
      return f2h(fmodf(h2f(a), h2f(b)));
    else if constexpr (is_complex_v<T>)
      return a;
    else
      return fmod(a, b);
  }
};

// ---- GCD / LCM (:350-381, :504-526) ----------------------------------------------------------
template <typename T>
__device__ __forceinline__ T gcd_euclid(T a, T b)
{
  if constexpr (is_bool_v<T>) {
    return a || b;
  } else {
    while (b != 0) {
      T r = static_cast<T>(a % b);
      a   = b;
      b   = r;
    }
    if constexpr (std::is_signed<T>::value) return a >= 0 ? a : static_cast<T>(-a);
    return a;
  }
}
template <typename T>
struct Gcd : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (std::is_integral<T>::value)
      return gcd_euclid<T>(a, b);
    else
      return a;
  }
};
template <typename T>
struct Lcm : Base<T> {
  static constexpr bool valid = std::is_integral<T>::value;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (is_bool_v<T>) {
      return a && b;
    } else if constexpr (std::is_integral<T>::value) {
      T r = gcd_euclid<T>(a, b);
      if (r == 0) return 0;
      r = static_cast<T>(a / r * b);
      if constexpr (std::is_signed<T>::value) return r >= 0 ? r : static_cast<T>(-r);
      return r;
    } else
      return a;
  }
};

// ---- float-only binary math (:175-198, :445-469, :240-263, :794-817): fp16 lifts to fp32 ------
#define CNB_FLOAT_BINOP(NAME, EXPR_F, EXPR_D)                                   \
  template <typename T>                                                         \
  struct NAME : Base<T> {                                                       \
    static constexpr bool valid = is_float_v<T>;                                \
    using Base<T>::Base;                                                        \
    __device__ __forceinline__ T operator()(const T& a_, const T& b_) const     \
    {                                                                           \
      if constexpr (is_half_v<T>) {                                             \
        float a = h2f(a_), b = h2f(b_);                                         \
        return f2h(EXPR_F);                                                     \
      } else if constexpr (std::is_same<T, float>::value) {                     \
        float a = a_, b = b_;                                                   \
        return EXPR_F;                                                          \
      } else if constexpr (std::is_same<T, double>::value) {                    \
        double a = a_, b = b_;                                                  \
        return EXPR_D;                                                          \
      } else                                                                    \
        return a_;                                                              \
    }                                                                           \
  };
CNB_FLOAT_BINOP(Arctan2, atan2f(a, b), atan2(a, b))
CNB_FLOAT_BINOP(Hypot, hypotf(a, b), hypot(a, b))
CNB_FLOAT_BINOP(Copysign, copysignf(a, b), copysign(a, b))
CNB_FLOAT_BINOP(Nextafter, nextafterf(a, b), nextafter(a, b))
// LOGADDEXP (:584-603), LOGADDEXP2 (:618-636)
CNB_FLOAT_BINOP(Logaddexp,
                (a == b ? a + logf(2.0f) : fmaxf(a, b) + log1pf(expf(-fabsf(a - b)))),
                (a == b ? a + log(2.0) : fmax(a, b) + log1p(exp(-fabs(a - b)))))
CNB_FLOAT_BINOP(Logaddexp2,
                (a == b ? a + 1.0f : fmaxf(a, b) + log2f(1.0f + exp2f(-fabsf(a - b)))),
                (a == b ? a + 1.0 : fmax(a, b) + log2(1.0 + exp2(-fabs(a - b)))))
#undef CNB_FLOAT_BINOP

// ---- LDEXP (:529-552): rhs is int32 ------------------------------------------------------------
template <typename T>
struct Ldexp {
  static constexpr bool valid = is_float_v<T>;
  using Out  = T;
  using Rhs2 = int32_t;
  __host__ __device__ Ldexp(const void* = nullptr) {}
  __device__ __forceinline__ T operator()(const T& a, const int32_t& b) const
  {
    if constexpr (is_half_v<T>)
      return f2h(ldexpf(h2f(a), b));
    else if constexpr (std::is_same<T, float>::value)
      return ldexpf(a, b);
    else if constexpr (std::is_same<T, double>::value)
      return ldexp(a, b);
    else
      return a;
  }
};

// a**b for integer-valued a and integer b, in double
__device__ __forceinline__ double pow_integral(double a, long long b)
{
  unsigned long long e = b < 0 ? 0ull - static_cast<unsigned long long>(b)
                               : static_cast<unsigned long long>(b);
  double r = 1.0, base = a;
  while (e != 0) {
    if (e & 1ull) r *= base;
    e >>= 1;
    if (e != 0) base *= base;
  }
  return b < 0 ? 1.0 / r : r;
}

// ---- POWER (:826-861): everything but fp16/complex goes through double ------------------------
template <typename T>
struct Power : Base<T> {
  static constexpr bool valid = true;
  using Base<T>::Base;
  __device__ __forceinline__ T operator()(const T& a, const T& b) const
  {
    if constexpr (is_half_v<T>)
      return f2h(powf(h2f(a), h2f(b)));
    else if constexpr (is_complex_v<T>)
      return cuda::std::pow(a, b);
    else if constexpr (std::is_integral<T>::value)
      // integer operands: the reference's pow(double, double) is exact whenever the result is
      // representable (glibc pow is < 1 ulp), and the cast truncates — a 2-ulp device pow() would
      // turn 3**3 into 26.  Exponentiation by squaring in double is exact below 2**53.
      return static_cast<T>(pow_integral(static_cast<double>(a), static_cast<long long>(b)));
    else if constexpr (std::is_same<T, float>::value)
      return d2f(pow(static_cast<double>(a), static_cast<double>(b)));
    else
      return static_cast<T>(pow(static_cast<double>(a), static_cast<double>(b)));
  }
};
// ---- FLOAT_POWER (:291-315): d, D, and F -> D -------------------------------------------------
template <typename T>
struct FloatPower {
  static constexpr bool valid =
    std::is_same<T, double>::value || std::is_same<T, c128>::value || std::is_same<T, c64>::value;
  using Out  = std::conditional_t<std::is_same<T, c64>::value, c128, T>;
  using Rhs2 = T;
  __host__ __device__ FloatPower(const void* = nullptr) {}
  __device__ __forceinline__ Out operator()(const T& a, const T& b) const
  {
    if constexpr (std::is_same<T, double>::value)
      return pow(a, b);
    else if constexpr (std::is_same<T, c128>::value)
      return cuda::std::pow(a, b);
    else if constexpr (std::is_same<T, c64>::value)
      return cuda::std::pow(c128(a.real(), a.imag()), c128(b.real(), b.imag()));
    else
      return a;
  }
};

// ---- ISCLOSE (:471-501): rtol, atol arrive as two extra float64 scalar stores -----------------
template <typename T>
struct Isclose {
  static constexpr bool valid = true;
  using Out  = bool;
  using Rhs2 = T;
  double rtol, atol;
  __host__ Isclose(const void* extra)
  {
    rtol = extra ? static_cast<const double*>(extra)[0] : 0.0;
    atol = extra ? static_cast<const double*>(extra)[1] : 0.0;
  }
  __device__ __forceinline__ bool operator()(const T& a, const T& b) const
  {
    if constexpr (is_complex_v<T>) {
      return static_cast<double>(cuda::std::abs(a - b)) <=
             atol + rtol * static_cast<double>(cuda::std::abs(b));
    } else if constexpr (std::is_integral<T>::value) {
      return fabs(static_cast<double>(a) - static_cast<double>(b)) <=
             atol + rtol * fabs(static_cast<double>(b));
    } else {
      auto fa = up(a);
      auto fb = up(b);
      if (isinf(fa) || isinf(fb)) return fa == fb;
      return fabs(static_cast<double>(fa) - static_cast<double>(fb)) <=
             atol + rtol * static_cast<double>(fabs(fb));
    }
  }
};

}  // namespace bop

// opcode -> functor template
template <int OP>
struct BinaryFn;
#define CNB_BIN(OPCODE, NAME)            \
  template <>                            \
  struct BinaryFn<OPCODE> {              \
    template <typename T>                \
    using fn = bop::NAME<T>;             \
  };
CNB_BIN(CNB_BINOP_ADD, Add)
CNB_BIN(CNB_BINOP_ARCTAN2, Arctan2)
CNB_BIN(CNB_BINOP_BITWISE_AND, BitwiseAnd)
CNB_BIN(CNB_BINOP_BITWISE_OR, BitwiseOr)
CNB_BIN(CNB_BINOP_BITWISE_XOR, BitwiseXor)
CNB_BIN(CNB_BINOP_COPYSIGN, Copysign)
CNB_BIN(CNB_BINOP_DIVIDE, Divide)
CNB_BIN(CNB_BINOP_EQUAL, Equal)
CNB_BIN(CNB_BINOP_FLOAT_POWER, FloatPower)
CNB_BIN(CNB_BINOP_FLOOR_DIVIDE, FloorDivide)
CNB_BIN(CNB_BINOP_FMOD, Fmod)
CNB_BIN(CNB_BINOP_GCD, Gcd)
CNB_BIN(CNB_BINOP_GREATER, Greater)
CNB_BIN(CNB_BINOP_GREATER_EQUAL, GreaterEqual)
CNB_BIN(CNB_BINOP_HYPOT, Hypot)
CNB_BIN(CNB_BINOP_ISCLOSE, Isclose)
CNB_BIN(CNB_BINOP_LCM, Lcm)
CNB_BIN(CNB_BINOP_LDEXP, Ldexp)
CNB_BIN(CNB_BINOP_LEFT_SHIFT, LeftShift)
CNB_BIN(CNB_BINOP_LESS, Less)
CNB_BIN(CNB_BINOP_LESS_EQUAL, LessEqual)
CNB_BIN(CNB_BINOP_LOGADDEXP, Logaddexp)
CNB_BIN(CNB_BINOP_LOGADDEXP2, Logaddexp2)
CNB_BIN(CNB_BINOP_LOGICAL_AND, LogicalAnd)
CNB_BIN(CNB_BINOP_LOGICAL_OR, LogicalOr)
CNB_BIN(CNB_BINOP_LOGICAL_XOR, LogicalXor)
CNB_BIN(CNB_BINOP_MAXIMUM, Maximum)
CNB_BIN(CNB_BINOP_MINIMUM, Minimum)
CNB_BIN(CNB_BINOP_MOD, Mod)
CNB_BIN(CNB_BINOP_MULTIPLY, Multiply)
CNB_BIN(CNB_BINOP_NEXTAFTER, Nextafter)
CNB_BIN(CNB_BINOP_NOT_EQUAL, NotEqual)
CNB_BIN(CNB_BINOP_POWER, Power)
CNB_BIN(CNB_BINOP_RIGHT_SHIFT, RightShift)
CNB_BIN(CNB_BINOP_SUBTRACT, Subtract)
#undef CNB_BIN

}  // namespace cnb
