// UNARY_OP with two outputs (FREXP / MODF), GETARG, WHERE and FILL.
#include "cnb_elementwise.cuh"
#include "ops_unary.cuh"

namespace cnb {
namespace {

template <class F, class T>
struct MultiOutAdapter {
  using O0 = typename F::Out;
  using O1 = typename F::Out2;
  using I0 = T;
  using I1 = Unused;
  using I2 = Unused;
  F f;
  __device__ __forceinline__ void operator()(O0& o, O1& o2, const I0& a, const Unused&,
                                             const Unused&) const
  {
    f(o, o2, a);
  }
};

template <template <typename> class FN>
int multiout_by_type(const cnb_store_t* out, const cnb_store_t* out2, const cnb_store_t* in,
                     cudaStream_t stream)
{
  return type_dispatch(in->dtype, [&](auto tag) -> int {
    using T = type_of<decltype(tag)::value>;
    using F = FN<T>;
    if constexpr (!F::valid) {
      return set_error(CNB_ERR_INVALID_OP, "FREXP/MODF not valid for dtype %d", in->dtype);
    } else {
      if (out->dtype != CodeOf<typename F::Out>::value || out2->dtype != CodeOf<typename F::Out2>::value)
        return set_error(CNB_ERR_BAD_ARG, "FREXP/MODF output dtypes (%d, %d) do not match dtype %d",
                         out->dtype, out2->dtype, in->dtype);
      MultiOutAdapter<F, T> ad{F{}};
      return ew_launch(ad, out, out2, in, nullptr, nullptr, stream);
    }
  });
}

// GETARG (unary_op_util.h:624-632): Argval<T> -> int64; every Argval<T> is 16 bytes with `arg` first
struct GetargFn {
  using O0 = long long;
  using O1 = Unused;
  using I0 = Argval<long long>;
  using I1 = Unused;
  using I2 = Unused;
  __device__ __forceinline__ void operator()(O0& o, Unused&, const I0& a, const Unused&,
                                             const Unused&) const
  {
    o = a.arg;
  }
};

// WHERE (where.cu:24-31): pure select on N-byte payloads
template <int N>
struct Blob {
  unsigned char b[N];
};
template <int N>
struct WhereFn {
  using V = Pack<unsigned char, N>;
  using O0 = V;
  using O1 = Unused;
  using I0 = bool;
  using I1 = V;
  using I2 = V;
  __device__ __forceinline__ void operator()(O0& o, Unused&, const bool& m, const V& a,
                                             const V& b) const
  {
    o = m ? a : b;
  }
};

template <int N>
struct FillFn {
  using V = Pack<unsigned char, N>;
  using O0 = V;
  using O1 = Unused;
  using I0 = Unused;
  using I1 = Unused;
  using I2 = Unused;
  V value;
  __device__ __forceinline__ void operator()(O0& o, Unused&, const Unused&, const Unused&,
                                             const Unused&) const
  {
    o = value;
  }
};

template <typename F>
int size_dispatch(size_t n, F&& f)
{
  switch (n) {
    case 1: return f(std::integral_constant<int, 1>{});
    case 2: return f(std::integral_constant<int, 2>{});
    case 4: return f(std::integral_constant<int, 4>{});
    case 8: return f(std::integral_constant<int, 8>{});
    case 16: return f(std::integral_constant<int, 16>{});
  }
  return set_error(CNB_ERR_BAD_ARG, "unsupported item size %zu", n);
}

}  // namespace

int unary_multiout(int op, const cnb_store_t* out, const cnb_store_t* out2, const cnb_store_t* in,
                   cudaStream_t stream)
{
  if (op == CNB_UOP_FREXP) return multiout_by_type<uop::Frexp>(out, out2, in, stream);
  if (op == CNB_UOP_MODF) return multiout_by_type<uop::Modf>(out, out2, in, stream);
  return set_error(CNB_ERR_BAD_ARG, "not a two-output unary op: %d", op);
}

int unary_getarg(const cnb_store_t* out, const cnb_store_t* in, cudaStream_t stream)
{
  if (in->dtype < CNB_ARGVAL_BASE || out->dtype != CNB_INT64)
    return set_error(CNB_ERR_BAD_ARG, "GETARG expects Argval input and int64 output (got %d -> %d)",
                     in->dtype, out->dtype);
  return ew_launch(GetargFn{}, out, nullptr, in, nullptr, nullptr, stream);
}

int where_select(const cnb_store_t* out, const cnb_store_t* mask, const cnb_store_t* in1,
                 const cnb_store_t* in2, cudaStream_t stream)
{
  if (mask->dtype != CNB_BOOL) return set_error(CNB_ERR_BAD_ARG, "WHERE mask must be bool");
  if (in1->dtype != out->dtype || in2->dtype != out->dtype)
    return set_error(CNB_ERR_BAD_ARG, "WHERE operands must share out's dtype (%d, %d -> %d)",
                     in1->dtype, in2->dtype, out->dtype);
  return size_dispatch(dtype_size(out->dtype), [&](auto n) -> int {
    return ew_launch(WhereFn<decltype(n)::value>{}, out, nullptr, mask, in1, in2, stream);
  });
}

int fill_value(const cnb_store_t* out, const void* value, cudaStream_t stream)
{
  if (value == nullptr) return set_error(CNB_ERR_BAD_ARG, "FILL without a value");
  return size_dispatch(dtype_size(out->dtype), [&](auto n) -> int {
    constexpr int N = decltype(n)::value;
    FillFn<N> fn;
    memcpy(fn.value.raw, value, N);
    return ew_launch(fn, out, nullptr, nullptr, nullptr, nullptr, stream);
  });
}

}  // namespace cnb
