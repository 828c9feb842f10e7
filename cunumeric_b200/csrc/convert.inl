// CONVERT dispatch body, included by convert_g*.cu: one TU per group of SOURCE dtypes.
#include "cnb_elementwise.cuh"
#include "ops_convert.cuh"

namespace cnb {
namespace {

template <int NAN_OP, int SRC>
int convert_to(const cnb_store_t* out, const cnb_store_t* in, cudaStream_t stream)
{
  using S = type_of<SRC>;
  return type_dispatch(out->dtype, [&](auto dtag) -> int {
    using D = type_of<decltype(dtag)::value>;
    using F = ConvertFn<NAN_OP, D, S>;
    if constexpr (!F::valid) {
      return set_error(CNB_ERR_INVALID_OP, "CONVERT nan_op %d: %d -> %d is not dispatched", NAN_OP,
                       in->dtype, out->dtype);
    } else {
      return ew_launch(F{}, out, nullptr, in, nullptr, nullptr, stream);
    }
  });
}

template <int SRC>
int convert_from(int nan_op, const cnb_store_t* out, const cnb_store_t* in, cudaStream_t stream)
{
  switch (nan_op) {
    case CNB_CONVERT_NAN_NOOP: return convert_to<CNB_CONVERT_NAN_NOOP, SRC>(out, in, stream);
    case CNB_CONVERT_NAN_PROD: return convert_to<CNB_CONVERT_NAN_PROD, SRC>(out, in, stream);
    case CNB_CONVERT_NAN_SUM: return convert_to<CNB_CONVERT_NAN_SUM, SRC>(out, in, stream);
  }
  return set_error(CNB_ERR_BAD_ARG, "unknown nan_op %d", nan_op);
}

}  // namespace

int CNB_CVT_GROUP_NAME(int nan_op, const cnb_store_t* out, const cnb_store_t* in,
                       cudaStream_t stream)
{
  switch (in->dtype) {
#define X(CODE) \
  case CODE: return convert_from<CODE>(nan_op, out, in, stream);
    CNB_CVT_GROUP_SRCS(X)
#undef X
  }
  return 1;  // source dtype not in this group
}

}  // namespace cnb
