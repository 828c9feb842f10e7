// SCALAR_UNARY_RED for sm_100a, included by scalar_red_g*.cu (one TU per opcode group).
//
// Single pass: persistent CTAs accumulate with 128-bit loads (U independent loads in flight per
// thread) -> warp shuffle -> shared memory -> one partial per CTA in global scratch; the CTA that
// takes the last ticket folds the partials in CTA order and folds the total into the output store
// (reduction-accessor semantics).  No second launch, no global atomics on data, no host round trip:
// the reference needs an H2D identity copy, <=1024 atomics and a <<<1,1>>> copy kernel per call
// (scalar_reduction.cuh:47-76, cuda_help.h:255-282).
#include "cnb_reduce.cuh"
#include "ops_reduce.cuh"

namespace cnb {

struct RedScratch {
  char* partials;        // >= grid * 16 bytes
  unsigned int* ticket;  // zero on entry, reset to zero by the last CTA
  int max_grid;
};
int red_acquire_scratch(RedScratch& s, cudaStream_t stream);

namespace {

template <class R>
struct RedIo {  // operand typing for EwShape: only the input participates
  using O0 = Unused;
  using O1 = Unused;
  using I0 = typename R::In;
  using I1 = Unused;
  using I2 = Unused;
};

template <class R>
__global__ void __launch_bounds__(RED_THREADS)
scalar_red_kernel(const __grid_constant__ EwPlan plan, const R r, typename R::Val* out,
                  char* partials_raw, unsigned int* ticket, const int ordered)
{
  using T   = typename R::In;
  using Acc = typename R::Acc;
  using S   = EwShape<RedIo<R>>;
  constexpr int E = S::E, U = S::U, TILE = S::TILE;

  const int tid        = threadIdx.x;
  const bool has_where = plan.op[3].ptr != nullptr;
  Acc acc              = R::identity();

  for (long long tile = blockIdx.x; tile < plan.num_tiles; tile += gridDim.x) {
    long long row = 0, ct = tile;
    if (plan.rows > 1) {
      row = tile / plan.tiles_per_row;
      ct  = tile - row * plan.tiles_per_row;
    }
    long long off_in = 0, off_w = 0, off_ix = 0;
    if (plan.n_outer > 0) {
      long long q = row;
#pragma unroll
      for (int d = EW_MAX_OUTER - 1; d >= 0; --d) {
        const long long qq = q / plan.outer[d];
        const long long i  = q - qq * plan.outer[d];
        q                  = qq;
        off_in += i * plan.op[2].outer_stride[d];
        off_w += i * plan.op[3].outer_stride[d];
        off_ix += i * plan.op[4].outer_stride[d];
      }
    }
    const long long col0 = ct * TILE;
    const long long ix0  = reinterpret_cast<long long>(plan.op[4].ptr) + off_ix;

    if (plan.vec && col0 + TILE <= plan.inner) {
      Pack<T, E> a[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = col0 + (long long)(u * RED_THREADS + tid) * E;
        ld_bytes<sizeof(T) * E>(a[u].raw, plan.op[2].ptr + off_in + e * (long long)sizeof(T));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = col0 + (long long)(u * RED_THREADS + tid) * E;
#pragma unroll
        for (int i = 0; i < E; ++i)
          red_visit(r, acc, a[u][i], ordered != 0,
                    [&] { return ix0 + (e + i) * plan.op[4].inner_stride; });
      }
    } else {
      constexpr int B = ew_cmax(4, ew_cmin(16, 128 / (int)sizeof(T)));
      constexpr int N = E * U;
#pragma unroll 1
      for (int j0 = 0; j0 < N; j0 += B) {
        Pack<T, 1> a[B];
        bool m[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
          const long long e = col0 + (long long)(j0 + j) * RED_THREADS + tid;
          m[j]              = (j0 + j < N) && (e < plan.inner);
          if (m[j]) {
            ld_bytes<sizeof(T)>(a[j].raw, plan.op[2].ptr + off_in + e * plan.op[2].inner_stride);
            if (has_where)
              m[j] = *reinterpret_cast<const unsigned char*>(plan.op[3].ptr + off_w +
                                                             e * plan.op[3].inner_stride) != 0;
          }
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
          if (m[j]) {
            const long long e = col0 + (long long)(j0 + j) * RED_THREADS + tid;
            red_visit(r, acc, a[j][0], ordered != 0,
                      [&] { return ix0 + e * plan.op[4].inner_stride; });
          }
        }
      }
    }
  }

  __shared__ RawSmem<Acc, RED_WARPS> smem;
  __shared__ bool is_last;
  Acc* partials = reinterpret_cast<Acc*>(partials_raw);
  acc           = block_reduce<R>(acc, smem.ptr());
  if (tid == 0) {
    partials[blockIdx.x] = acc;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last              = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    Acc total = R::identity();
    // fold the per-CTA partials in CTA order (each thread takes a contiguous slice)
    const int per = (gridDim.x + RED_THREADS - 1) / RED_THREADS;
    const int lo  = tid * per;
    const int hi  = min(lo + per, (int)gridDim.x);
    for (int i = lo; i < hi; ++i) {
      Acc p;
      ld_bytes<sizeof(Acc)>(&p, reinterpret_cast<const char*>(partials + i));
      total = R::fold(total, p);
    }
    total = block_reduce<R>(total, smem.ptr());
    if (tid == 0) {
      // out.reduce(0, result): fold into the caller's pre-filled store
      const Acc prev = R::lift(*out);
      *out           = R::finish(R::fold(prev, total));
      *ticket        = 0;
    }
  }
}

template <int OP>
int scalar_red_by_type(const cnb_store_t* out, const cnb_store_t* in, const cnb_store_t* where,
                       const int64_t* origin, const int64_t* gshape, const void* extra,
                       cudaStream_t stream)
{
  return type_dispatch(in->dtype, [&](auto tag) -> int {
    using T = type_of<decltype(tag)::value>;
    using R = typename RedFn<OP>::template fn<T>;
    if constexpr (!R::valid) {
      return set_error(CNB_ERR_INVALID_OP, "SCALAR_UNARY_RED %d is not valid for dtype %d", OP,
                       in->dtype);
    } else {
      using S = EwShape<RedIo<R>>;
      if (out->dtype != CodeOf<typename R::Val>::value)
        return set_error(CNB_ERR_BAD_ARG, "SCALAR_UNARY_RED %d on dtype %d: out dtype %d, expected %d",
                         OP, in->dtype, out->dtype, CodeOf<typename R::Val>::value);
      if (out->ptr == nullptr) return set_error(CNB_ERR_BAD_ARG, "null output");
      if (where != nullptr && where->dtype != CNB_BOOL)
        return set_error(CNB_ERR_BAD_ARG, "where mask must be bool");
      // pseudo-operand carrying the GLOBAL row-major flat index (unary_red_util.h:342-351)
      cnb_store_t index_store = *in;
      long long flat_origin   = 0;
      {
        long long stride = 1;
        for (int d = in->ndim - 1; d >= 0; --d) {
          index_store.strides[d] = stride;
          flat_origin += (origin ? origin[d] : 0) * stride;
          stride *= (gshape ? gshape[d] : in->shape[d]);
        }
        index_store.ptr = reinterpret_cast<void*>(flat_origin);
      }
      EwArg args[EW_MAX_OPS] = {{nullptr, 0, true},
                                {nullptr, 0, true},
                                {in, (int)sizeof(T), false},
                                {where, 1, false},
                                {R::needs_index ? &index_store : nullptr, 0, false, true}};
      int chunk[EW_MAX_OPS] = {0, 0, ew_cmin(16, (int)sizeof(T) * S::E), 1, 0};
      EwPlan plan;
      int rc = ew_make_plan(plan, args, EW_MAX_OPS, chunk, S::TILE);
      if (rc < 0) return rc;
      if (rc == 0) return CNB_OK;  // empty rect: the store keeps its pre-filled value
      if (where != nullptr) plan.vec = 0;
      if (plan.op[2].inner_stride != (long long)sizeof(T)) plan.vec = 0;
      // arg-reductions: does every thread meet elements in increasing GLOBAL index order?  True
      // when the canonical (memory-order) iteration is also row-major in index space.
      int ordered = 0;
      if (R::needs_index) {
        ordered               = plan.op[4].inner_stride > 0 ? 1 : 0;
        long long inner_span  = plan.op[4].inner_stride * plan.inner;
        for (int d = EW_MAX_OUTER - 1; d >= 0 && ordered; --d) {
          if (plan.outer[d] == 1) continue;
          if (plan.op[4].outer_stride[d] < inner_span) ordered = 0;
          inner_span = plan.op[4].outer_stride[d] * plan.outer[d];
        }
      }
      RedScratch scratch;
      rc = red_acquire_scratch(scratch, stream);
      if (rc != CNB_OK) return rc;
      auto kernel = scalar_red_kernel<R>;
      int grid    = ew_grid_size(reinterpret_cast<const void*>(kernel), plan.num_tiles, 32);
      if (grid > scratch.max_grid) grid = scratch.max_grid;
      {
        LaunchScope scope(stream, KERNEL_SCALAR_RED, plan.inner * plan.rows,
                          ew_algorithmic_bytes(plan, args, EW_MAX_OPS));
        kernel<<<grid, RED_THREADS, 0, stream>>>(plan, R(extra),
                                                 static_cast<typename R::Val*>(out->ptr),
                                                 scratch.partials, scratch.ticket, ordered);
      }
      return check_cuda(cudaGetLastError(), "scalar_red_kernel launch");
    }
  });
}

}  // namespace

int CNB_SRED_GROUP_NAME(int op, const cnb_store_t* out, const cnb_store_t* in,
                        const cnb_store_t* where, const int64_t* origin, const int64_t* gshape,
                        const void* extra, cudaStream_t stream)
{
  switch (op) {
#define X(OPCODE) \
  case OPCODE: return scalar_red_by_type<OPCODE>(out, in, where, origin, gshape, extra, stream);
    CNB_SRED_GROUP_OPS(X)
#undef X
  }
  return 1;
}

}  // namespace cnb
