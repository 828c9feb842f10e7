// BINARY_OP dispatch body, included by binary_g*.cu (one translation unit per opcode group so the
// template instantiations compile in parallel).  Each TU defines CNB_BIN_GROUP_NAME and the list
// CNB_BIN_GROUP_OPS(X).
#include "cnb_elementwise.cuh"
#include "ops_binary.cuh"

namespace cnb {
namespace {

template <class F, class T>
struct BinAdapter {
  using O0 = typename F::Out;
  using O1 = Unused;
  using I0 = T;
  using I1 = typename F::Rhs2;
  using I2 = Unused;
  F f;
  __device__ __forceinline__ void operator()(O0& o, Unused&, const I0& a, const I1& b,
                                             const Unused&) const
  {
    o = f(a, b);
  }
};

template <int OP>
int binary_by_type(const cnb_store_t* out, const cnb_store_t* in1, const cnb_store_t* in2,
                   const void* extra, cudaStream_t stream)
{
  return type_dispatch(in1->dtype, [&](auto tag) -> int {
    using T = type_of<decltype(tag)::value>;
    using F = typename BinaryFn<OP>::template fn<T>;
    if constexpr (!F::valid) {
      return set_error(CNB_ERR_INVALID_OP, "BINARY_OP %d is not valid for dtype %d", OP, in1->dtype);
    } else {
      if (out->dtype != CodeOf<typename F::Out>::value)
        return set_error(CNB_ERR_BAD_ARG, "BINARY_OP %d on dtype %d: out dtype %d, expected %d", OP,
                         in1->dtype, out->dtype, CodeOf<typename F::Out>::value);
      if (in2->dtype != CodeOf<typename F::Rhs2>::value)
        return set_error(CNB_ERR_BAD_ARG, "BINARY_OP %d on dtype %d: in2 dtype %d, expected %d", OP,
                         in1->dtype, in2->dtype, CodeOf<typename F::Rhs2>::value);
      BinAdapter<F, T> ad{F(extra)};
      return ew_launch(ad, out, nullptr, in1, in2, nullptr, stream);
    }
  });
}

}  // namespace

int CNB_BIN_GROUP_NAME(int op, const cnb_store_t* out, const cnb_store_t* in1,
                       const cnb_store_t* in2, const void* extra, cudaStream_t stream)
{
  switch (op) {
#define X(OPCODE) \
  case OPCODE: return binary_by_type<OPCODE>(out, in1, in2, extra, stream);
    CNB_BIN_GROUP_OPS(X)
#undef X
  }
  return 1;  // not in this group
}

}  // namespace cnb
