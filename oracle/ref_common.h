// TEST INFRASTRUCTURE ONLY — driver glue for oracle/_ref/libcunumeric_ref.so.
//
// The arithmetic comes from the reference's own functor headers, included BY PATH from
// /root/reference/src (never copied).  This header only adds the (opcode, dtype) run-time
// dispatch that legate.core's double_dispatch/type_dispatch provide in the real build.
#pragma once

#include "legate.h"

#include <omp.h>

namespace ref {

using Code = legate::Type::Code;

template <Code C>
using code_c = std::integral_constant<Code, C>;

// Run-time dtype code -> compile-time tag (stand-in for legate::type_dispatch).
template <typename F>
int type_dispatch(int code, F&& f)
{
  switch (static_cast<Code>(code)) {
    case Code::BOOL: return f(code_c<Code::BOOL>{});
    case Code::INT8: return f(code_c<Code::INT8>{});
    case Code::INT16: return f(code_c<Code::INT16>{});
    case Code::INT32: return f(code_c<Code::INT32>{});
    case Code::INT64: return f(code_c<Code::INT64>{});
    case Code::UINT8: return f(code_c<Code::UINT8>{});
    case Code::UINT16: return f(code_c<Code::UINT16>{});
    case Code::UINT32: return f(code_c<Code::UINT32>{});
    case Code::UINT64: return f(code_c<Code::UINT64>{});
    case Code::FLOAT16: return f(code_c<Code::FLOAT16>{});
    case Code::FLOAT32: return f(code_c<Code::FLOAT32>{});
    case Code::FLOAT64: return f(code_c<Code::FLOAT64>{});
    case Code::COMPLEX64: return f(code_c<Code::COMPLEX64>{});
    case Code::COMPLEX128: return f(code_c<Code::COMPLEX128>{});
  }
  return -2;
}

// C++ result type -> dtype code (so Python can allocate the output the functor really produces).
template <typename T>
struct code_of;
#define REF_CODE_OF(T, C)                               \
  template <>                                           \
  struct code_of<T> {                                   \
    static constexpr int value = static_cast<int>(C);   \
  };
REF_CODE_OF(bool, Code::BOOL)
REF_CODE_OF(int8_t, Code::INT8)
REF_CODE_OF(int16_t, Code::INT16)
REF_CODE_OF(int32_t, Code::INT32)
REF_CODE_OF(int64_t, Code::INT64)
REF_CODE_OF(long long, Code::INT64)
REF_CODE_OF(uint8_t, Code::UINT8)
REF_CODE_OF(uint16_t, Code::UINT16)
REF_CODE_OF(uint32_t, Code::UINT32)
REF_CODE_OF(uint64_t, Code::UINT64)
REF_CODE_OF(unsigned long long, Code::UINT64)
REF_CODE_OF(__half, Code::FLOAT16)
REF_CODE_OF(float, Code::FLOAT32)
REF_CODE_OF(double, Code::FLOAT64)
REF_CODE_OF(::complex<float>, Code::COMPLEX64)
REF_CODE_OF(::complex<double>, Code::COMPLEX128)
REF_CODE_OF(std::complex<float>, Code::COMPLEX64)
REF_CODE_OF(std::complex<double>, Code::COMPLEX128)
#undef REF_CODE_OF

constexpr int ERR_INVALID = -1;  // the reference marks this (op, dtype) `valid = false`
constexpr int ERR_BADCODE = -2;

}  // namespace ref
