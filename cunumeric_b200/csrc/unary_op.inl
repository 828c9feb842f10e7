// UNARY_OP dispatch body, included by unary_g*.cu (one TU per opcode group).
#include "cnb_elementwise.cuh"
#include "ops_unary.cuh"

#include <cstring>

namespace cnb {
namespace {

template <class F, class T>
struct UnAdapter {
  using O0 = typename F::Out;
  using O1 = Unused;
  using I0 = T;
  using I1 = Unused;
  using I2 = Unused;
  F f;
  __device__ __forceinline__ void operator()(O0& o, Unused&, const I0& a, const Unused&,
                                             const Unused&) const
  {
    o = f(a);
  }
};

template <int OP>
int unary_by_type(const cnb_store_t* out, const cnb_store_t* in, const void* extra,
                  cudaStream_t stream)
{
  return type_dispatch(in->dtype, [&](auto tag) -> int {
    using T = type_of<decltype(tag)::value>;
    using F = typename UnaryFn<OP>::template fn<T>;
    if constexpr (!F::valid) {
      return set_error(CNB_ERR_INVALID_OP, "UNARY_OP %d is not valid for dtype %d", OP, in->dtype);
    } else {
      if (out->dtype != CodeOf<typename F::Out>::value)
        return set_error(CNB_ERR_BAD_ARG, "UNARY_OP %d on dtype %d: out dtype %d, expected %d", OP,
                         in->dtype, out->dtype, CodeOf<typename F::Out>::value);
      UnAdapter<F, T> ad{F(extra)};
      return ew_launch(ad, out, nullptr, in, nullptr, nullptr, stream);
    }
  });
}

}  // namespace

int CNB_UN_GROUP_NAME(int op, const cnb_store_t* out, const cnb_store_t* in, const void* extra,
                      cudaStream_t stream)
{
  switch (op) {
#define X(OPCODE) \
  case OPCODE: return unary_by_type<OPCODE>(out, in, extra, stream);
    CNB_UN_GROUP_OPS(X)
#undef X
  }
  return 1;  // not in this group
}

}  // namespace cnb
