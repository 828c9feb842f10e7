// Reduction functors for SCALAR_UNARY_RED / UNARY_RED on sm_100a — one struct per
// CuNumericUnaryRedCode.  Semantics follow unary/unary_red_util.h:104-601 of the reference:
//   VAL    the value the task folds into the output store (bool / uint64 / T / Argval<T>)
//   Acc    what a thread carries while reducing (== VAL, except fp16 SUM-like ops carry fp32 and
//          round once per partial; the reference accumulates in fp16, contract is n*eps)
//   convert(x, index)   unary_red_util.h `convert`
//   fold(a, b)          the Legion reduction op's fold, made order-independent:
//                         MAX/MIN keep `a` unless b is strictly better (NaN never wins, as in the
//                         CPU fold `if (b > a) a = b`), ARG* break ties towards the LOWER index
//                         (what the sequential CPU fold of arg.inl:43-50 produces).
#pragma once

#include "ops_math.cuh"

#include <cfloat>
#include <climits>

namespace cnb {
namespace red {

// ---- identities of the Legion reduction ops (external to the reference) ---------
template <typename T>
__host__ __device__ __forceinline__ T lowest_of()
{
  if constexpr (is_half_v<T>)
    return __ushort_as_half(0xfc00);  // -inf
  else if constexpr (std::is_same<T, float>::value)
    return -INFINITY;
  else if constexpr (std::is_same<T, double>::value)
    return -static_cast<double>(INFINITY);
  else if constexpr (is_bool_v<T>)
    return false;
  else
    return cuda::std::numeric_limits<T>::lowest();
}
template <typename T>
__host__ __device__ __forceinline__ T highest_of()
{
  if constexpr (is_half_v<T>)
    return __ushort_as_half(0x7c00);  // +inf
  else if constexpr (std::is_same<T, float>::value)
    return INFINITY;
  else if constexpr (std::is_same<T, double>::value)
    return static_cast<double>(INFINITY);
  else if constexpr (is_bool_v<T>)
    return true;
  else
    return cuda::std::numeric_limits<T>::max();
}

template <typename T>
struct AccOf {
  using type = T;
};
template <>
struct AccOf<__half> {
  using type = float;
};

template <typename T>
__device__ __forceinline__ T mul_wrap(T a, T b)
{
  if constexpr (is_bool_v<T>)
    return a && b;
  else if constexpr (std::is_integral<T>::value) {
    using W = std::conditional_t<(sizeof(T) < 8), unsigned int, unsigned long long>;
    return static_cast<T>(static_cast<W>(a) * static_cast<W>(b));
  } else
    return a * b;
}
template <typename T>
__device__ __forceinline__ T add_wrap(T a, T b)
{
  if constexpr (is_bool_v<T>)
    return a || b;
  else if constexpr (std::is_integral<T>::value)
    return static_cast<T>(static_cast<std::make_unsigned_t<T>>(a) +
                          static_cast<std::make_unsigned_t<T>>(b));
  else
    return a + b;
}
template <typename T>
__device__ __forceinline__ bool nonzero(const T& x)
{
  if constexpr (is_complex_v<T>)
    return x.real() != 0 || x.imag() != 0;  // rhs != RHS(0)
  else if constexpr (is_half_v<T>)
    return h2f(x) != 0.0f;
  else
    return x != T(0);
}

// Common shape: T input, Val folded into the store, Acc carried by threads.
template <typename T, typename V, typename A = V>
struct Shape {
  using In  = T;
  using Val = V;
  using Acc = A;
  static constexpr bool needs_index = false;
  __device__ __forceinline__ static Val finish(const Acc& a) { return static_cast<Val>(a); }
  __device__ __forceinline__ static Acc lift(const Val& v) { return static_cast<Acc>(v); }
};
template <typename T>
struct HalfShape : Shape<T, T, typename AccOf<T>::type> {
  using Acc = typename AccOf<T>::type;
  __device__ __forceinline__ static T finish(const Acc& a)
  {
    if constexpr (is_half_v<T>)
      return f2h(a);
    else
      return a;
  }
  __device__ __forceinline__ static Acc lift(const T& v)
  {
    if constexpr (is_half_v<T>)
      return h2f(v);
    else
      return v;
  }
};

template <typename T>
struct All : Shape<T, bool> {
  static constexpr bool valid = !std::is_same<T, c128>::value;
  __host__ __device__ All(const void*) {}
  __device__ __forceinline__ static bool identity() { return true; }
  __device__ __forceinline__ bool convert(const T& x, long long) const { return nonzero(x); }
  __device__ __forceinline__ static bool fold(bool a, bool b) { return a && b; }
};
template <typename T>
struct Any : Shape<T, bool> {
  static constexpr bool valid = !std::is_same<T, c128>::value;
  __host__ __device__ Any(const void*) {}
  __device__ __forceinline__ static bool identity() { return false; }
  __device__ __forceinline__ bool convert(const T& x, long long) const { return nonzero(x); }
  __device__ __forceinline__ static bool fold(bool a, bool b) { return a || b; }
};
template <typename T>
struct CountNonzero : Shape<T, unsigned long long> {
  static constexpr bool valid = true;
  __host__ __device__ CountNonzero(const void*) {}
  __device__ __forceinline__ static unsigned long long identity() { return 0ull; }
  __device__ __forceinline__ unsigned long long convert(const T& x, long long) const
  {
    return nonzero(x) ? 1ull : 0ull;
  }
  __device__ __forceinline__ static unsigned long long fold(unsigned long long a,
                                                            unsigned long long b)
  {
    return a + b;
  }
};
// CONTAINS (scalar path only, scalar_unary_red_template.inl:89-91)
template <typename T>
struct Contains : Shape<T, bool> {
  static constexpr bool valid = true;
  T to_find;
  __host__ Contains(const void* extra)
  {
    if (extra)
      memcpy(&to_find, extra, sizeof(T));
    else
      memset(&to_find, 0, sizeof(T));
  }
  __device__ __forceinline__ static bool identity() { return false; }
  __device__ __forceinline__ bool convert(const T& x, long long) const { return eq(x, to_find); }
  __device__ __forceinline__ static bool fold(bool a, bool b) { return a || b; }
};

template <typename T, bool IS_MAX, bool SKIP_NAN>
struct MinMax : Shape<T, T> {
  static constexpr bool valid = !is_complex_v<T> && (!SKIP_NAN || is_float_v<T>);
  __host__ __device__ MinMax(const void*) {}
  __device__ __forceinline__ static T identity()
  {
    return IS_MAX ? lowest_of<T>() : highest_of<T>();
  }
  __device__ __forceinline__ T convert(const T& x, long long) const
  {
    if constexpr (SKIP_NAN) {
      if (isnan_any(x)) return identity();
    }
    return x;
  }
  __device__ __forceinline__ static T fold(const T& a, const T& b)
  {
    if constexpr (IS_MAX)
      return lt(a, b) ? b : a;  // if (b > a) a = b
    else
      return lt(b, a) ? b : a;  // if (b < a) a = b
  }
};

template <typename T, bool SKIP_NAN>
struct Sum : HalfShape<T> {
  using Acc = typename HalfShape<T>::Acc;
  static constexpr bool valid = !SKIP_NAN || is_float_v<T> || is_complex_v<T>;
  __host__ __device__ Sum(const void*) {}
  __device__ __forceinline__ static Acc identity() { return Acc(0); }
  __device__ __forceinline__ Acc convert(const T& x, long long) const
  {
    if constexpr (SKIP_NAN) {
      if (isnan_any(x)) return Acc(0);
    }
    return HalfShape<T>::lift(x);
  }
  __device__ __forceinline__ static Acc fold(const Acc& a, const Acc& b) { return add_wrap(a, b); }
};
template <typename T, bool SKIP_NAN>
struct Prod : HalfShape<T> {
  using Acc = typename HalfShape<T>::Acc;
  static constexpr bool valid = !std::is_same<T, c128>::value &&
                                (!SKIP_NAN || is_float_v<T> || std::is_same<T, c64>::value);
  __host__ __device__ Prod(const void*) {}
  __device__ __forceinline__ static Acc identity() { return Acc(1); }
  __device__ __forceinline__ Acc convert(const T& x, long long) const
  {
    if constexpr (SKIP_NAN) {
      if (isnan_any(x)) return Acc(1);
    }
    return HalfShape<T>::lift(x);
  }
  __device__ __forceinline__ static Acc fold(const Acc& a, const Acc& b) { return mul_wrap(a, b); }
};
// SUM_SQUARES (:273-294) and VARIANCE (:296-317; mu is subtracted on the scalar path,
// scalar_unary_red_template.inl:95-96)
template <typename T, bool CENTERED>
struct SumSquares : HalfShape<T> {
  using Acc = typename HalfShape<T>::Acc;
  static constexpr bool valid = true;
  T mu;
  __host__ SumSquares(const void* extra)
  {
    if (CENTERED && extra)
      memcpy(&mu, extra, sizeof(T));
    else
      memset(&mu, 0, sizeof(T));
  }
  __device__ __forceinline__ static Acc identity() { return Acc(0); }
  __device__ __forceinline__ Acc convert(const T& x, long long) const
  {
    if constexpr (is_half_v<T>) {
      float d = h2f(x);
      if constexpr (CENTERED) d = h2f(f2h(d - h2f(mu)));
      return h2f(f2h(d * d));
    } else if constexpr (is_bool_v<T>) {
      bool d = CENTERED ? (x != mu) : x;
      return d;
    } else {
      T d = x;
      if constexpr (CENTERED) {
        if constexpr (std::is_integral<T>::value)
          d = static_cast<T>(static_cast<std::make_unsigned_t<T>>(x) -
                             static_cast<std::make_unsigned_t<T>>(mu));
        else
          d = x - mu;
      }
      return mul_wrap(d, d);
    }
  }
  __device__ __forceinline__ static Acc fold(const Acc& a, const Acc& b) { return add_wrap(a, b); }
};

// ARGMAX / ARGMIN / NANARGMAX / NANARGMIN (:319-463)
template <typename T, bool IS_MAX, bool SKIP_NAN>
struct ArgMinMax : Shape<T, Argval<T>> {
  using V = Argval<T>;
  static constexpr bool valid       = !is_complex_v<T> && (!SKIP_NAN || is_float_v<T>);
  static constexpr bool needs_index = true;
  __host__ ArgMinMax(const void*) {}
  __device__ __forceinline__ static V identity()
  {
    V v;
    v.arg   = LLONG_MIN;
    v.value = IS_MAX ? lowest_of<T>() : highest_of<T>();
    return v;
  }
  __device__ __forceinline__ V convert(const T& x, long long index) const
  {
    if constexpr (SKIP_NAN) {
      if (isnan_any(x)) return identity();
    }
    V v;
    v.arg   = index;
    v.value = x;
    return v;
  }
  // In-thread fast path for elements visited in INCREASING index order: a strictly better value
  // replaces the incumbent, so the first occurrence survives ties, NaN never wins and an element
  // equal to the identity value never displaces the identity — exactly the sequential CPU fold.
  __device__ __forceinline__ static bool better(const T& incumbent, const T& x)
  {
    return IS_MAX ? lt(incumbent, x) : lt(x, incumbent);
  }
  __device__ __forceinline__ static V fold(const V& a, const V& b)
  {
    // A strictly better value wins.  NaN never wins (the CPU fold is `if (b > a) a = b`).
    // Equal values: the identity (arg == LLONG_MIN) survives, because the sequential fold only
    // replaces on a strictly better value; between two real elements the LOWER index wins, which
    // is the first occurrence the sequential fold keeps.
    if constexpr (is_float_v<T>) {
      if (isnan_any(b.value)) return a;
      if (isnan_any(a.value)) return b;
    }
    const bool better = IS_MAX ? lt(a.value, b.value) : lt(b.value, a.value);
    if (better) return b;
    const bool worse = IS_MAX ? lt(b.value, a.value) : lt(a.value, b.value);
    if (worse) return a;
    if (a.arg == LLONG_MIN) return a;
    if (b.arg == LLONG_MIN) return b;
    return (b.arg < a.arg) ? b : a;
  }
};

}  // namespace red

template <int OP>
struct RedFn;
#define CNB_RED(OPCODE, ...)        \
  template <>                       \
  struct RedFn<OPCODE> {            \
    template <typename T>           \
    using fn = __VA_ARGS__;         \
  };
CNB_RED(CNB_RED_ALL, red::All<T>)
CNB_RED(CNB_RED_ANY, red::Any<T>)
CNB_RED(CNB_RED_ARGMAX, red::ArgMinMax<T, true, false>)
CNB_RED(CNB_RED_ARGMIN, red::ArgMinMax<T, false, false>)
CNB_RED(CNB_RED_CONTAINS, red::Contains<T>)
CNB_RED(CNB_RED_COUNT_NONZERO, red::CountNonzero<T>)
CNB_RED(CNB_RED_MAX, red::MinMax<T, true, false>)
CNB_RED(CNB_RED_MIN, red::MinMax<T, false, false>)
CNB_RED(CNB_RED_NANARGMAX, red::ArgMinMax<T, true, true>)
CNB_RED(CNB_RED_NANARGMIN, red::ArgMinMax<T, false, true>)
CNB_RED(CNB_RED_NANMAX, red::MinMax<T, true, true>)
CNB_RED(CNB_RED_NANMIN, red::MinMax<T, false, true>)
CNB_RED(CNB_RED_NANPROD, red::Prod<T, true>)
CNB_RED(CNB_RED_NANSUM, red::Sum<T, true>)
CNB_RED(CNB_RED_PROD, red::Prod<T, false>)
CNB_RED(CNB_RED_SUM, red::Sum<T, false>)
CNB_RED(CNB_RED_SUM_SQUARES, red::SumSquares<T, false>)
CNB_RED(CNB_RED_VARIANCE, red::SumSquares<T, true>)
#undef CNB_RED

}  // namespace cnb
