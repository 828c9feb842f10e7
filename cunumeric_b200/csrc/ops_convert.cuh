// CONVERT functor (astype) for sm_100a.  Semantics follow the reference's CPU variant,
// unary/convert_util.h:46-206: a C++ static_cast, except that anything <-> fp16 goes through double
// (:77-102), complex -> real keeps the real part and complex -> bool is `re || im` (:62-74), and
// the NaN-aware flavours replace NaN by 1 (NAN_PROD) or 0 (NAN_SUM) (:104-206).
#pragma once

#include "ops_math.cuh"

namespace cnb {

// static_cast between arithmetic types, with the slow F2F.F32.F64 replaced by d2f()
template <typename D, typename S>
__device__ __forceinline__ D cast_to(const S& s)
{
  if constexpr (std::is_same<D, float>::value && std::is_same<S, double>::value)
    return d2f(s);
  else
    return static_cast<D>(s);
}

template <typename D, typename S>
__device__ __forceinline__ D convert_plain(const S& s)
{
  if constexpr (is_complex_v<S>) {
    if constexpr (is_complex_v<D>)
      return D(cast_to<typename D::value_type>(s.real()),
               cast_to<typename D::value_type>(s.imag()));
    else if constexpr (is_half_v<D>)
      return d2h(static_cast<double>(s.real()));
    else
      return cast_to<D>(s.real());
  } else if constexpr (is_half_v<S>) {
    // (every fp16 value is exact in fp32, so fp32 stands in for the reference's double here)
    const float v = h2f(s);
    if constexpr (is_complex_v<D>)
      return D(static_cast<typename D::value_type>(v), 0);
    else
      return static_cast<D>(v);
  } else if constexpr (is_half_v<D>) {
    return d2h(static_cast<double>(s));
  } else if constexpr (is_complex_v<D>) {
    return D(cast_to<typename D::value_type>(s), 0);
  } else {
    return cast_to<D>(s);
  }
}

template <int NAN_OP, typename D, typename S>
struct ConvertFn {
  // convert_template.inl:62-89: same-type pairs are never dispatched; NaN-aware flavours only for
  // floating / complex sources
  static constexpr bool valid =
    !std::is_same<D, S>::value &&
    (NAN_OP == CNB_CONVERT_NAN_NOOP || is_float_v<S> || is_complex_v<S>);
  using O0 = D;
  using O1 = Unused;
  using I0 = S;
  using I1 = Unused;
  using I2 = Unused;
  __device__ __forceinline__ void operator()(D& o, Unused&, const S& s, const Unused&,
                                             const Unused&) const
  {
    if constexpr (NAN_OP == CNB_CONVERT_NAN_NOOP) {
      if constexpr (is_complex_v<S> && is_bool_v<D>)
        o = (s.real() != 0) || (s.imag() != 0);
      else
        o = convert_plain<D, S>(s);
    } else {
      if (isnan_any(s)) {
        if constexpr (is_half_v<D>)
          o = f2h(NAN_OP == CNB_CONVERT_NAN_PROD ? 1.0f : 0.0f);
        else
          o = static_cast<D>(NAN_OP == CNB_CONVERT_NAN_PROD ? 1 : 0);
      } else {
        o = convert_plain<D, S>(s);
      }
    }
  }
};

}  // namespace cnb
