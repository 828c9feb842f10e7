// TEST INFRASTRUCTURE ONLY — UNARY_OP / WHERE / CONVERT legs of oracle/_ref/libcunumeric_ref.so.
// Arithmetic: the reference's UnaryOp / MultiOutUnaryOp / ConvertOp functors, included by path
//   (/root/reference/src/cunumeric/unary/unary_op_util.h:194-1247, unary/convert_util.h:46-206).
// Loop shapes: unary/unary_op.cc:30-47 (single output), :80-100 (FREXP/MODF),
//   ternary/where.cc:36-41, unary/convert.cc (same shape as unary_op.cc); OpenMP twins *_omp.cc.
#include "ref_common.h"
#include "cunumeric/unary/unary_op_util.h"
#include "cunumeric/unary/convert_util.h"

using namespace cunumeric;
using ref::Code;

namespace {

template <UnaryOpCode OP, Code CODE>
int run(const void* inv, void* outv, size_t n, const void* extra, int nthreads, bool query_only)
{
  if constexpr (!UnaryOp<OP, CODE>::valid) {
    return ref::ERR_INVALID;
  } else {
    using FN  = UnaryOp<OP, CODE>;
    using ARG = typename FN::T;
    using RES = std::result_of_t<FN(ARG)>;  // unary_op_template.inl:43
    if (query_only) return ref::code_of<std::decay_t<RES>>::value;
    std::vector<legate::Store> args;
    if (OP == UnaryOpCode::CLIP) {
      // CLIP reads min,max as two scalar stores of the array dtype (unary_op_util.h:411-416)
      args.emplace_back(extra);
      args.emplace_back(static_cast<const char*>(extra) + sizeof(legate::legate_type_of<CODE>));
    }
    FN func{args};
    auto in  = static_cast<const ARG*>(inv);
    auto out = static_cast<std::decay_t<RES>*>(outv);
    if (nthreads > 1) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
      for (size_t idx = 0; idx < n; ++idx) out[idx] = func(in[idx]);
    } else {
      for (size_t idx = 0; idx < n; ++idx) out[idx] = func(in[idx]);
    }
    return ref::code_of<std::decay_t<RES>>::value;
  }
}

template <UnaryOpCode OP>
int by_type(int code, const void* a, void* o, size_t n, const void* extra, int nthreads, bool q)
{
  return ref::type_dispatch(
    code, [&](auto tag) { return run<OP, decltype(tag)::value>(a, o, n, extra, nthreads, q); });
}

int dispatch(int op, int code, const void* a, void* o, size_t n, const void* extra, int nthreads,
             bool q)
{
#define CASE(NAME) \
  case UnaryOpCode::NAME: return by_type<UnaryOpCode::NAME>(code, a, o, n, extra, nthreads, q);
  switch (static_cast<UnaryOpCode>(op)) {
    CASE(ABSOLUTE) CASE(ARCCOS) CASE(ARCCOSH) CASE(ARCSIN) CASE(ARCSINH) CASE(ARCTAN)
    CASE(ARCTANH) CASE(CBRT) CASE(CEIL) CASE(CLIP) CASE(CONJ) CASE(COPY) CASE(COS) CASE(COSH)
    CASE(DEG2RAD) CASE(EXP) CASE(EXP2) CASE(EXPM1) CASE(FLOOR) CASE(IMAG) CASE(INVERT)
    CASE(ISFINITE) CASE(ISINF) CASE(ISNAN) CASE(LOG) CASE(LOG10) CASE(LOG1P) CASE(LOG2)
    CASE(LOGICAL_NOT) CASE(NEGATIVE) CASE(RAD2DEG) CASE(REAL) CASE(RECIPROCAL) CASE(RINT)
    CASE(SIGN) CASE(SIGNBIT) CASE(SIN) CASE(SINH) CASE(SQRT) CASE(SQUARE) CASE(TAN) CASE(TANH)
    CASE(TRUNC)
    // POSITIVE is dispatched to the COPY functor (unary_op_util.h:146-148)
    case UnaryOpCode::POSITIVE: return by_type<UnaryOpCode::COPY>(code, a, o, n, extra, nthreads, q);
    default: break;
  }
#undef CASE
  return ref::ERR_BADCODE;
}

template <UnaryOpCode OP, Code CODE>
int run_multi(const void* inv, void* o1, void* o2, size_t n, bool query_only)
{
  if constexpr (!MultiOutUnaryOp<OP, CODE>::valid) {
    return ref::ERR_INVALID;
  } else {
    using FN = MultiOutUnaryOp<OP, CODE>;
    if (query_only) return ref::code_of<typename FN::RHS2>::value;
    FN func{};
    auto in   = static_cast<const typename FN::RHS1*>(inv);
    auto lhs  = static_cast<typename FN::LHS*>(o1);
    auto rhs2 = static_cast<typename FN::RHS2*>(o2);
    // unary_op.cc:92-98
    for (size_t idx = 0; idx < n; ++idx) lhs[idx] = func(in[idx], &rhs2[idx]);
    return ref::code_of<typename FN::RHS2>::value;
  }
}

template <ConvertCode NAN_OP, Code DST, Code SRC>
int run_convert(const void* inv, void* outv, size_t n, int nthreads)
{
  // convert_template.inl:62-89: SRC==DST is not dispatched; NaN-aware conversion only for
  // floating/complex sources.
  constexpr bool src_fp = legate::is_floating_point<SRC>::value || legate::is_complex<SRC>::value;
  if constexpr (SRC == DST || (NAN_OP != ConvertCode::NOOP && !src_fp)) {
    return ref::ERR_INVALID;
  } else {
    using FN  = ConvertOp<NAN_OP, DST, SRC>;
    using S   = legate::legate_type_of<SRC>;
    using D   = legate::legate_type_of<DST>;
    FN func{};
    auto in  = static_cast<const S*>(inv);
    auto out = static_cast<D*>(outv);
    if (nthreads > 1) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
      for (size_t idx = 0; idx < n; ++idx) out[idx] = func(in[idx]);
    } else {
      for (size_t idx = 0; idx < n; ++idx) out[idx] = func(in[idx]);
    }
    return static_cast<int>(DST);
  }
}

template <ConvertCode NAN_OP>
int convert_by_types(int dst, int src, const void* in, void* out, size_t n, int nthreads)
{
  return ref::type_dispatch(dst, [&](auto dtag) {
    return ref::type_dispatch(src, [&](auto stag) {
      return run_convert<NAN_OP, decltype(dtag)::value, decltype(stag)::value>(in, out, n, nthreads);
    });
  });
}

}  // namespace

extern "C" {

int ref_unary_out_code(int op, int code)
{
  return dispatch(op, code, nullptr, nullptr, 0, nullptr, 1, true);
}

// out[i] = UnaryOp<op,code>(in[i]). extra: CLIP -> {min, max} packed as two values of the dtype.
int ref_unary_op(int op, int code, const void* in, void* out, size_t n, const void* extra,
                 int nthreads)
{
  return dispatch(op, code, in, out, n, extra, nthreads, false);
}

// FREXP (out2 int32) / MODF (out2 same dtype). Returns dtype code of out2.
int ref_unary_multiout(int op, int code, const void* in, void* out1, void* out2, size_t n)
{
  const bool q = (in == nullptr);
  return ref::type_dispatch(code, [&](auto tag) {
    constexpr Code C = decltype(tag)::value;
    switch (static_cast<UnaryOpCode>(op)) {
      case UnaryOpCode::FREXP: return run_multi<UnaryOpCode::FREXP, C>(in, out1, out2, n, q);
      case UnaryOpCode::MODF: return run_multi<UnaryOpCode::MODF, C>(in, out1, out2, n, q);
      default: return ref::ERR_BADCODE;
    }
  });
}

// GETARG: Argval<T> (16-byte struct {int64 arg; T value}) -> int64. The reference dispatches on
// the OUT dtype (int64) and reads `.arg` (unary_op_template.inl:171, unary_op_util.h:624-632).
int ref_getarg(const void* in, void* out, size_t n)
{
  using FN = UnaryOp<UnaryOpCode::GETARG, Code::INT64>;
  std::vector<legate::Store> args;
  FN func{args};
  auto src = static_cast<const typename FN::T*>(in);
  auto dst = static_cast<int64_t*>(out);
  for (size_t idx = 0; idx < n; ++idx) dst[idx] = func(src[idx]);
  return static_cast<int>(Code::INT64);
}

int ref_convert(int nan_op, int dst, int src, const void* in, void* out, size_t n, int nthreads)
{
  switch (static_cast<ConvertCode>(nan_op)) {
    case ConvertCode::NOOP: return convert_by_types<ConvertCode::NOOP>(dst, src, in, out, n, nthreads);
    case ConvertCode::PROD: return convert_by_types<ConvertCode::PROD>(dst, src, in, out, n, nthreads);
    case ConvertCode::SUM: return convert_by_types<ConvertCode::SUM>(dst, src, in, out, n, nthreads);
  }
  return ref::ERR_BADCODE;
}

// out[i] = mask[i] ? in1[i] : in2[i]  (ternary/where.cc:36-41); pure select, any 1/2/4/8/16-byte type
int ref_where(int code, const void* mask, const void* in1, const void* in2, void* out, size_t n,
              int nthreads)
{
  return ref::type_dispatch(code, [&](auto tag) {
    using VAL    = legate::legate_type_of<decltype(tag)::value>;
    auto m       = static_cast<const bool*>(mask);
    auto a       = static_cast<const VAL*>(in1);
    auto b       = static_cast<const VAL*>(in2);
    auto o       = static_cast<VAL*>(out);
    if (nthreads > 1) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
      for (size_t idx = 0; idx < n; ++idx) o[idx] = m[idx] ? a[idx] : b[idx];
    } else {
      for (size_t idx = 0; idx < n; ++idx) o[idx] = m[idx] ? a[idx] : b[idx];
    }
    return static_cast<int>(decltype(tag)::value);
  });
}
}
