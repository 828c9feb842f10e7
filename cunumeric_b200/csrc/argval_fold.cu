// Combine of per-rank arg-reduction partials ("arg-reductions: ncclAllGather of the 16-byte
// (idx, val) partials, then a local fold with lowest-global-index tie-break").  `gathered` holds
// world x n Argval<T> partials in rank order; out[i] = fold over the ranks of gathered[r][i] with the
// very fold the reduction kernels use (ops_reduce.cuh ArgMinMax::fold: a strictly better value wins,
// NaN never wins, the identity survives ties, between real elements the LOWER global index wins) —
// the result is what one GPU would have produced.  One launch instead of two allreduces and a dozen
// small elementwise kernels.
#include "cnb_reduce.cuh"
#include "ops_reduce.cuh"

namespace cnb {
int ensure_init();

namespace {
template <class R>
__global__ void __launch_bounds__(256)
argval_fold_kernel(typename R::Val* out, const typename R::Val* gathered, int world, long long n)
{
  using V = typename R::Val;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    V acc;
    ld_bytes<sizeof(V)>(&acc, reinterpret_cast<const char*>(gathered + i));
    for (int r = 1; r < world; ++r) {
      V p;
      ld_bytes<sizeof(V)>(&p, reinterpret_cast<const char*>(gathered + (long long)r * n + i));
      acc = R::fold(acc, p);
    }
    st_bytes<sizeof(V)>(reinterpret_cast<char*>(out + i), &acc);
  }
}

template <int OP>
int fold_by_type(int elem_dtype, void* out, const void* gathered, int world, long long n, cudaStream_t s)
{
  return type_dispatch(elem_dtype, [&](auto tag) -> int {
    using T = type_of<decltype(tag)::value>;
    using R = typename RedFn<OP>::template fn<T>;
    if constexpr (!R::valid) {
      return set_error(CNB_ERR_INVALID_OP, "arg-reduction %d is not valid for dtype %d", OP, elem_dtype);
    } else {
      static_assert(sizeof(typename R::Val) == 16, "Argval is 16 bytes");
      const long long want = (n + 255) / 256;
      const unsigned grid  = (unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sm_count() * 8));
      {
        LaunchScope scope(s, KERNEL_ELEMENTWISE, n, (long long)(world + 1) * n * 16);
        argval_fold_kernel<R><<<grid, 256, 0, s>>>(static_cast<typename R::Val*>(out),
                                                   static_cast<const typename R::Val*>(gathered), world, n);
      }
      return check_cuda(cudaGetLastError(), "argval_fold_kernel launch");
    }
  });
}
}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" int cnb_argval_fold(int32_t op, int32_t elem_dtype, void* out, const void* gathered,
                               int32_t world, int64_t n, void* stream)
{
  int rc = ensure_init();
  if (rc != CNB_OK) return rc;
  if (out == nullptr || gathered == nullptr || world < 1 || n < 0)
    return set_error(CNB_ERR_BAD_ARG, "cnb_argval_fold: bad argument");
  if (n == 0) return CNB_OK;
  set_task_tag(CNB_OP_UNARY_RED, op, elem_dtype);
  auto s = static_cast<cudaStream_t>(stream);
  switch (op) {
    case CNB_RED_ARGMAX: return fold_by_type<CNB_RED_ARGMAX>(elem_dtype, out, gathered, world, n, s);
    case CNB_RED_ARGMIN: return fold_by_type<CNB_RED_ARGMIN>(elem_dtype, out, gathered, world, n, s);
    case CNB_RED_NANARGMAX: return fold_by_type<CNB_RED_NANARGMAX>(elem_dtype, out, gathered, world, n, s);
    case CNB_RED_NANARGMIN: return fold_by_type<CNB_RED_NANARGMIN>(elem_dtype, out, gathered, world, n, s);
  }
  return set_error(CNB_ERR_BAD_ARG, "cnb_argval_fold: %d is not an arg-reduction", op);
}
