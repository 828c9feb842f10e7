// Shared device/host declarations for the sm_100a hot path.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cuda/std/complex>

#include <cstdint>
#include <cstdio>
#include <type_traits>

#include "../../include/cunumeric_b200.h"

namespace cnb {

using c64  = cuda::std::complex<float>;
using c128 = cuda::std::complex<double>;

// Argval<T> of the reference (arg.h:23-59): {int64 arg; T arg_value}; 16 bytes for every T <= 8 B.
template <typename T>
struct alignas(8) Argval {
  long long arg;
  T value;
};
static_assert(sizeof(Argval<bool>) == 16 && sizeof(Argval<double>) == 16 &&
              sizeof(Argval<__half>) == 16, "Argval must be 16 bytes");

// dtype code <-> C++ type
template <int CODE>
struct TypeOf;
#define CNB_TYPE(CODE, T)  \
  template <>              \
  struct TypeOf<CODE> {    \
    using type = T;        \
  };
CNB_TYPE(CNB_BOOL, bool)
CNB_TYPE(CNB_INT8, int8_t)
CNB_TYPE(CNB_INT16, int16_t)
CNB_TYPE(CNB_INT32, int32_t)
CNB_TYPE(CNB_INT64, int64_t)
CNB_TYPE(CNB_UINT8, uint8_t)
CNB_TYPE(CNB_UINT16, uint16_t)
CNB_TYPE(CNB_UINT32, uint32_t)
CNB_TYPE(CNB_UINT64, uint64_t)
CNB_TYPE(CNB_FLOAT16, __half)
CNB_TYPE(CNB_FLOAT32, float)
CNB_TYPE(CNB_FLOAT64, double)
CNB_TYPE(CNB_COMPLEX64, c64)
CNB_TYPE(CNB_COMPLEX128, c128)
#undef CNB_TYPE
template <int CODE>
using type_of = typename TypeOf<CODE>::type;

template <typename T>
struct CodeOf;
#define CNB_CODE(T, CODE)                 \
  template <>                             \
  struct CodeOf<T> {                      \
    static constexpr int value = CODE;    \
  };
CNB_CODE(bool, CNB_BOOL)
CNB_CODE(int8_t, CNB_INT8)
CNB_CODE(int16_t, CNB_INT16)
CNB_CODE(int32_t, CNB_INT32)
CNB_CODE(int64_t, CNB_INT64)
CNB_CODE(long long, CNB_INT64)
CNB_CODE(uint8_t, CNB_UINT8)
CNB_CODE(uint16_t, CNB_UINT16)
CNB_CODE(uint32_t, CNB_UINT32)
CNB_CODE(uint64_t, CNB_UINT64)
CNB_CODE(unsigned long long, CNB_UINT64)
CNB_CODE(__half, CNB_FLOAT16)
CNB_CODE(float, CNB_FLOAT32)
CNB_CODE(double, CNB_FLOAT64)
CNB_CODE(c64, CNB_COMPLEX64)
CNB_CODE(c128, CNB_COMPLEX128)
#undef CNB_CODE
template <typename T>
struct CodeOf<Argval<T>> {
  static constexpr int value = CNB_ARGVAL_BASE + CodeOf<T>::value;
};

template <typename T>
struct is_complex_t : std::false_type {};
template <>
struct is_complex_t<c64> : std::true_type {};
template <>
struct is_complex_t<c128> : std::true_type {};
template <typename T>
inline constexpr bool is_complex_v = is_complex_t<T>::value;
template <typename T>
inline constexpr bool is_half_v = std::is_same<T, __half>::value;
template <typename T>
inline constexpr bool is_bool_v = std::is_same<T, bool>::value;
// "floating" in the reference's sense incl. fp16 (unary_op_util.h:186-188)
template <typename T>
inline constexpr bool is_float_v = std::is_floating_point<T>::value || is_half_v<T>;
template <typename T>
inline constexpr bool is_int_v = std::is_integral<T>::value;  // includes bool
template <typename T>
inline constexpr bool is_signed_int_v = std::is_integral<T>::value && std::is_signed<T>::value;

inline size_t dtype_size(int code)
{
  if (code >= CNB_ARGVAL_BASE) return 16;
  switch (code) {
    case CNB_BOOL:
    case CNB_INT8:
    case CNB_UINT8: return 1;
    case CNB_INT16:
    case CNB_UINT16:
    case CNB_FLOAT16: return 2;
    case CNB_INT32:
    case CNB_UINT32:
    case CNB_FLOAT32: return 4;
    case CNB_INT64:
    case CNB_UINT64:
    case CNB_FLOAT64:
    case CNB_COMPLEX64: return 8;
    case CNB_COMPLEX128: return 16;
  }
  return 0;
}

// run-time dtype code -> compile-time tag
template <int C>
using code_c = std::integral_constant<int, C>;

template <typename F>
int type_dispatch(int code, F&& f)
{
  switch (code) {
    case CNB_BOOL: return f(code_c<CNB_BOOL>{});
    case CNB_INT8: return f(code_c<CNB_INT8>{});
    case CNB_INT16: return f(code_c<CNB_INT16>{});
    case CNB_INT32: return f(code_c<CNB_INT32>{});
    case CNB_INT64: return f(code_c<CNB_INT64>{});
    case CNB_UINT8: return f(code_c<CNB_UINT8>{});
    case CNB_UINT16: return f(code_c<CNB_UINT16>{});
    case CNB_UINT32: return f(code_c<CNB_UINT32>{});
    case CNB_UINT64: return f(code_c<CNB_UINT64>{});
    case CNB_FLOAT16: return f(code_c<CNB_FLOAT16>{});
    case CNB_FLOAT32: return f(code_c<CNB_FLOAT32>{});
    case CNB_FLOAT64: return f(code_c<CNB_FLOAT64>{});
    case CNB_COMPLEX64: return f(code_c<CNB_COMPLEX64>{});
    case CNB_COMPLEX128: return f(code_c<CNB_COMPLEX128>{});
  }
  return CNB_ERR_BAD_ARG;
}

// ---- error plumbing (runtime.cu) ------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();

// ---- launch accounting / tracing (runtime.cu) -------------------------------------------------
// Every kernel launch goes through a LaunchScope: it counts the launch and, while a trace is
// active (cnb_trace_start), brackets it with CUDA events on the launching stream so bench.py can
// report per-kernel durations measured live.
void set_task_tag(int task, int op, int dtype);
struct LaunchScope {
  LaunchScope(cudaStream_t stream, int kernel_kind, long long elems, long long bytes);
  ~LaunchScope();
  cudaStream_t stream_;
  int slot_;
};
enum { KERNEL_ELEMENTWISE = 1, KERNEL_SCALAR_RED = 2, KERNEL_AXIS_COL = 3, KERNEL_AXIS_ROW = 4 };

#define CNB_CUDA(expr)                                  \
  do {                                                  \
    int _rc = ::cnb::check_cuda((expr), #expr);         \
    if (_rc != CNB_OK) return _rc;                      \
  } while (0)

}  // namespace cnb
