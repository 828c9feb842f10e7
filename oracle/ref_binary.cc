// TEST INFRASTRUCTURE ONLY — BINARY_OP leg of oracle/_ref/libcunumeric_ref.so.
// Arithmetic: the reference's BinaryOp<OP,CODE> functors, included by path
//   (/root/reference/src/cunumeric/binary/binary_op_util.h:163-877).
// Loop shape: the dense CPU body, binary/binary_op.cc:43-49; the OpenMP twin,
//   binary/binary_op_omp.cc:44 (`#pragma omp parallel for schedule(static)`).
#include "ref_common.h"
#include "cunumeric/binary/binary_op_util.h"

using namespace cunumeric;
using ref::Code;

namespace {

// s1/s2: element strides of the operands (1 = dense, 0 = a promoted scalar read through a
// stride-0 accessor, the way Future-backed scalar operands reach the reference's generic loop,
// binary_op.cc:46-49)
template <BinaryOpCode OP, Code CODE>
int run(const void* in1v, const void* in2v, void* outv, size_t n, const double* extra, int nthreads,
        bool query_only, size_t s1 = 1, size_t s2 = 1)
{
  if constexpr (!BinaryOp<OP, CODE>::valid) {
    return ref::ERR_INVALID;
  } else {
    using FN   = BinaryOp<OP, CODE>;
    using RHS1 = legate::legate_type_of<CODE>;
    using RHS2 = rhs2_of_binary_op<OP, CODE>;
    using LHS  = std::result_of_t<FN(RHS1, RHS2)>;  // binary_op_template.inl:41
    if (query_only) return ref::code_of<LHS>::value;
    std::vector<legate::Store> args;
    if (OP == BinaryOpCode::ISCLOSE) {
      args.emplace_back(&extra[0]);
      args.emplace_back(&extra[1]);
    }
    FN func{args};
    auto in1 = static_cast<const RHS1*>(in1v);
    auto in2 = static_cast<const RHS2*>(in2v);
    auto out = static_cast<LHS*>(outv);
    if (nthreads > 1) {
#pragma omp parallel for schedule(static) num_threads(nthreads)
      for (size_t idx = 0; idx < n; ++idx) out[idx] = func(in1[idx * s1], in2[idx * s2]);
    } else {
      for (size_t idx = 0; idx < n; ++idx) out[idx] = func(in1[idx * s1], in2[idx * s2]);
    }
    return ref::code_of<LHS>::value;
  }
}

template <BinaryOpCode OP>
int by_type(int code, const void* a, const void* b, void* o, size_t n, const double* extra,
            int nthreads, bool q, size_t s1, size_t s2)
{
  return ref::type_dispatch(code, [&](auto tag) {
    return run<OP, decltype(tag)::value>(a, b, o, n, extra, nthreads, q, s1, s2);
  });
}

int dispatch(int op, int code, const void* a, const void* b, void* o, size_t n,
             const double* extra, int nthreads, bool q, size_t s1 = 1, size_t s2 = 1)
{
#define CASE(NAME)        \
  case BinaryOpCode::NAME: \
    return by_type<BinaryOpCode::NAME>(code, a, b, o, n, extra, nthreads, q, s1, s2);
  switch (static_cast<BinaryOpCode>(op)) {
    CASE(ADD) CASE(ARCTAN2) CASE(BITWISE_AND) CASE(BITWISE_OR) CASE(BITWISE_XOR) CASE(COPYSIGN)
    CASE(DIVIDE) CASE(EQUAL) CASE(FLOAT_POWER) CASE(FLOOR_DIVIDE) CASE(FMOD) CASE(GCD)
    CASE(GREATER) CASE(GREATER_EQUAL) CASE(HYPOT) CASE(ISCLOSE) CASE(LCM) CASE(LDEXP)
    CASE(LEFT_SHIFT) CASE(LESS) CASE(LESS_EQUAL) CASE(LOGADDEXP) CASE(LOGADDEXP2)
    CASE(LOGICAL_AND) CASE(LOGICAL_OR) CASE(LOGICAL_XOR) CASE(MAXIMUM) CASE(MINIMUM) CASE(MOD)
    CASE(MULTIPLY) CASE(NEXTAFTER) CASE(NOT_EQUAL) CASE(POWER) CASE(RIGHT_SHIFT) CASE(SUBTRACT)
  }
#undef CASE
  return ref::ERR_BADCODE;
}

// BINARY_RED: the dense CPU body of binary/binary_red.cc:38-47 — early exit on the first failing
// pair, result folded into `out` with ProdReduction<bool> (out &= result).
template <BinaryOpCode OP, Code CODE>
int run_red(const void* in1v, const void* in2v, bool* out, size_t n, const double* extra)
{
  if constexpr (!BinaryOp<OP, CODE>::valid) {
    return ref::ERR_INVALID;
  } else {
    using FN  = BinaryOp<OP, CODE>;
    using ARG = legate::legate_type_of<CODE>;
    std::vector<legate::Store> args;
    if (OP == BinaryOpCode::ISCLOSE) {
      args.emplace_back(&extra[0]);
      args.emplace_back(&extra[1]);
    }
    FN func{args};
    auto in1 = static_cast<const ARG*>(in1v);
    auto in2 = static_cast<const ARG*>(in2v);
    for (size_t idx = 0; idx < n; ++idx)
      if (!func(in1[idx], in2[idx])) {
        *out = *out && false;
        return 0;
      }
    *out = *out && true;
    return 0;
  }
}

}  // namespace

extern "C" {

// *out &= all(op(in1[i], in2[i])); op must be EQUAL or ISCLOSE (binary_op_util.h reduce_op_dispatch)
int ref_binary_red(int op, int code, const void* in1, const void* in2, bool* out, size_t n,
                   const double* extra)
{
  return ref::type_dispatch(code, [&](auto tag) {
    constexpr Code CODE = decltype(tag)::value;
    switch (static_cast<BinaryOpCode>(op)) {
      case BinaryOpCode::EQUAL: return run_red<BinaryOpCode::EQUAL, CODE>(in1, in2, out, n, extra);
      case BinaryOpCode::ISCLOSE:
        return run_red<BinaryOpCode::ISCLOSE, CODE>(in1, in2, out, n, extra);
      default: return static_cast<int>(ref::ERR_BADCODE);
    }
  });
}

// Returns the dtype code of the result the reference functor produces, or <0.
int ref_binary_out_code(int op, int code)
{
  return dispatch(op, code, nullptr, nullptr, nullptr, 0, nullptr, 1, true);
}

// out[i] = BinaryOp<op,code>(in1[i], in2[i]) over n dense elements. extra = {rtol, atol}.
int ref_binary_op(int op, int code, const void* in1, const void* in2, void* out, size_t n,
                  const double* extra, int nthreads)
{
  return dispatch(op, code, in1, in2, out, n, extra, nthreads, false);
}

// Same, with per-operand element strides (0 = broadcast scalar operand).
int ref_binary_op_strided(int op, int code, const void* in1, size_t s1, const void* in2, size_t s2,
                          void* out, size_t n, const double* extra, int nthreads)
{
  return dispatch(op, code, in1, in2, out, n, extra, nthreads, false, s1, s2);
}
}
