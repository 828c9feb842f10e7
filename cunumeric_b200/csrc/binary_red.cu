// BINARY_RED (array_equal / allclose): out &= all(op(in1, in2)) over the rect, op in {EQUAL, ISCLOSE}
// (reference: src/cunumeric/binary/binary_red.cu:25-103, binary_red_template.inl:31-75; the Python
// side pre-fills `out` with True and reduces with ProdReduction<bool>, deferred.py:3330-3364).
// One pass: every thread ANDs its elements, the CTA votes with __syncthreads_and, and only a CTA that
// saw a mismatch stores `false` (an idempotent store — no atomics, no ticket needed).
#include "cnb_elementwise.cuh"
#include "ops_binary.cuh"

namespace cnb {
namespace {

template <class F, class T>
struct BinRedIo {
  using O0 = Unused;
  using O1 = Unused;
  using I0 = T;
  using I1 = T;
  using I2 = Unused;
};

template <class F, class T>
__global__ void __launch_bounds__(EW_THREADS)
binary_red_kernel(const __grid_constant__ EwPlan plan, const F f, bool* out)
{
  using S = EwShape<BinRedIo<F, T>>;
  constexpr int E = S::E, U = S::U, TILE = S::TILE;
  const int tid = threadIdx.x;
  bool ok       = true;
  for (long long tile = blockIdx.x; tile < plan.num_tiles; tile += gridDim.x) {
    long long row = 0, ct = tile;
    if (plan.rows > 1) {
      row = tile / plan.tiles_per_row;
      ct  = tile - row * plan.tiles_per_row;
    }
    long long off1 = 0, off2 = 0;
    if (plan.n_outer > 0) {
      long long q = row;
#pragma unroll
      for (int d = EW_MAX_OUTER - 1; d >= 0; --d) {
        const long long qq = q / plan.outer[d];
        const long long i  = q - qq * plan.outer[d];
        q                  = qq;
        off1 += i * plan.op[2].outer_stride[d];
        off2 += i * plan.op[3].outer_stride[d];
      }
    }
    const long long col0 = ct * TILE;
    if (plan.vec && col0 + TILE <= plan.inner) {
      Pack<T, E> a[U], b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = col0 + (long long)(u * EW_THREADS + tid) * E;
        ew_load_vec<T, E>(a[u], plan.op[2], off1, e);
        ew_load_vec<T, E>(b[u], plan.op[3], off2, e);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int i = 0; i < E; ++i) ok = ok && f(a[u][i], b[u][i]);
    } else {
      for (long long e = col0 + tid; e < min(col0 + (long long)TILE, plan.inner); e += EW_THREADS) {
        Pack<T, 1> a, b;
        ew_load_one<T>(a, plan.op[2], off1, e);
        ew_load_one<T>(b, plan.op[3], off2, e);
        ok = ok && f(a[0], b[0]);
      }
    }
  }
  if (__syncthreads_and(ok ? 1 : 0) == 0 && tid == 0) *out = false;
}

template <template <typename> class FN>
int binary_red_by_type(const cnb_store_t* out, const cnb_store_t* in1, const cnb_store_t* in2,
                       const void* extra, cudaStream_t stream)
{
  return type_dispatch(in1->dtype, [&](auto tag) -> int {
    using T = type_of<decltype(tag)::value>;
    using F = FN<T>;
    using S = EwShape<BinRedIo<F, T>>;
    if (in2->dtype != in1->dtype)
      return set_error(CNB_ERR_BAD_ARG, "BINARY_RED operands must share a dtype (%d vs %d)",
                       in1->dtype, in2->dtype);
    EwArg args[EW_MAX_OPS] = {{nullptr, 0, true},
                              {nullptr, 0, true},
                              {in1, (int)sizeof(T), false},
                              {in2, (int)sizeof(T), false},
                              {nullptr, 0, false}};
    const int cb          = ew_cmin(16, (int)sizeof(T) * S::E);
    int chunk[EW_MAX_OPS] = {0, 0, cb, cb, 0};
    EwPlan plan;
    int rc = ew_make_plan(plan, args, EW_MAX_OPS, chunk, S::TILE);
    if (rc <= 0) return rc;  // empty intersection: nothing to fold (binary_red_template.inl:48-51)
    auto kernel = binary_red_kernel<F, T>;
    int grid    = ew_grid_size(reinterpret_cast<const void*>(kernel), plan.num_tiles, 32);
    {
      LaunchScope scope(stream, KERNEL_SCALAR_RED, plan.inner * plan.rows,
                        ew_algorithmic_bytes(plan, args, EW_MAX_OPS));
      kernel<<<grid, EW_THREADS, 0, stream>>>(plan, F(extra), static_cast<bool*>(out->ptr));
    }
    return check_cuda(cudaGetLastError(), "binary_red_kernel launch");
  });
}

}  // namespace

int binary_red(int op, const cnb_store_t* out, const cnb_store_t* in1, const cnb_store_t* in2,
               const void* extra, cudaStream_t stream)
{
  if (out->dtype != CNB_BOOL || out->ptr == nullptr)
    return set_error(CNB_ERR_BAD_ARG, "BINARY_RED output must be a bool store");
  switch (op) {
    case CNB_BINOP_EQUAL: return binary_red_by_type<bop::Equal>(out, in1, in2, extra, stream);
    case CNB_BINOP_ISCLOSE: return binary_red_by_type<bop::Isclose>(out, in1, in2, extra, stream);
  }
  // binary_op_util.h:149-161 `reduce_op_dispatch` handles EQUAL and ISCLOSE only
  return set_error(CNB_ERR_INVALID_OP, "BINARY_RED supports EQUAL and ISCLOSE only (got %d)", op);
}

}  // namespace cnb
