// Host side of the elementwise engine: iteration-space canonicalisation and grid sizing.
#include "cnb_elementwise.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

namespace cnb {

int ew_make_plan(EwPlan& plan, const EwArg* args, int nargs, const int* chunk_bytes, int tile_elems)
{
  // reference shape: first present store (the output) defines the iteration space; every other
  // operand must have the same extents (alignment constraints, deferred.py:3325-3326).
  const cnb_store_t* ref = nullptr;
  for (int k = 0; k < nargs; ++k)
    if (args[k].store != nullptr) {
      ref = args[k].store;
      break;
    }
  if (ref == nullptr) return set_error(CNB_ERR_BAD_ARG, "elementwise task without stores");
  int ndim = ref->ndim;
  if (ndim < 0 || ndim > CNB_MAX_DIM) return set_error(CNB_ERR_BAD_ARG, "ndim %d out of range", ndim);

  long long shape[CNB_MAX_DIM + 1];
  long long strides[EW_MAX_OPS][CNB_MAX_DIM + 1];
  for (int d = 0; d < ndim; ++d) shape[d] = ref->shape[d];
  for (int k = 0; k < nargs; ++k) {
    const cnb_store_t* s = args[k].store;
    if (s == nullptr) {
      for (int d = 0; d < ndim; ++d) strides[k][d] = 0;
      continue;
    }
    if (s->ndim != ndim) return set_error(CNB_ERR_BAD_ARG, "operand %d: ndim %d != %d", k, s->ndim, ndim);
    if (s->ptr == nullptr && !args[k].index_only)
      return set_error(CNB_ERR_BAD_ARG, "operand %d: null pointer", k);
    for (int d = 0; d < ndim; ++d) {
      if (s->shape[d] != shape[d])
        return set_error(CNB_ERR_BAD_ARG, "operand %d: extent[%d]=%lld != %lld", k, d,
                         (long long)s->shape[d], shape[d]);
      strides[k][d] = s->strides[d];
      if (args[k].is_output && s->strides[d] == 0 && shape[d] > 1)
        return set_error(CNB_ERR_BAD_ARG, "output operand %d is broadcast along dim %d", k, d);
    }
  }
  if (ndim == 0) {  // 0-d stores are (1,)
    ndim     = 1;
    shape[0] = 1;
    for (int k = 0; k < nargs; ++k) strides[k][0] = 0;
  }
  for (int d = 0; d < ndim; ++d) {
    if (shape[d] < 0) return set_error(CNB_ERR_BAD_ARG, "negative extent");
    if (shape[d] == 0) return 0;  // empty rect: nothing to do (binary_op_template.inl:48)
  }

  // 1. drop unit dims
  int n = 0;
  for (int d = 0; d < ndim; ++d) {
    if (shape[d] == 1) continue;
    shape[n] = shape[d];
    for (int k = 0; k < nargs; ++k) strides[k][n] = strides[k][d];
    ++n;
  }
  if (n == 0) {
    n        = 1;
    shape[0] = 1;
    for (int k = 0; k < nargs; ++k) strides[k][0] = 0;
  }
  // 2. order dims by the output's |stride|, largest first (elementwise work is order-free), so a
  //    uniformly transposed task still streams along its contiguous dimension.
  int order[CNB_MAX_DIM + 1];
  for (int d = 0; d < n; ++d) order[d] = d;
  int okey = 0;
  while (okey < nargs && args[okey].store == nullptr) ++okey;
  std::stable_sort(order, order + n, [&](int a, int b) {
    return std::llabs(strides[okey][a]) > std::llabs(strides[okey][b]);
  });
  {
    long long s2[CNB_MAX_DIM + 1], st2[EW_MAX_OPS][CNB_MAX_DIM + 1];
    for (int d = 0; d < n; ++d) {
      s2[d] = shape[order[d]];
      for (int k = 0; k < nargs; ++k) st2[k][d] = strides[k][order[d]];
    }
    for (int d = 0; d < n; ++d) {
      shape[d] = s2[d];
      for (int k = 0; k < nargs; ++k) strides[k][d] = st2[k][d];
    }
  }
  // 3. merge neighbouring dims that are jointly contiguous for every operand
  int m = 0;
  for (int d = 1; d < n; ++d) {
    bool mergeable = true;
    for (int k = 0; k < nargs; ++k)
      if (strides[k][m] != shape[d] * strides[k][d]) {
        mergeable = false;
        break;
      }
    if (mergeable) {
      shape[m] *= shape[d];
      for (int k = 0; k < nargs; ++k) strides[k][m] = strides[k][d];
    } else {
      ++m;
      shape[m] = shape[d];
      for (int k = 0; k < nargs; ++k) strides[k][m] = strides[k][d];
    }
  }
  n = m + 1;

  // 4. fill the plan: last dim is the inner one
  plan.inner   = shape[n - 1];
  plan.n_outer = n - 1;
  plan.rows    = 1;
  for (int d = 0; d < EW_MAX_OUTER; ++d) plan.outer[d] = 1;
  for (int d = 0; d < n - 1; ++d) {
    const int slot   = EW_MAX_OUTER - (n - 1) + d;  // right-aligned so slot 2 is the fastest outer
    plan.outer[slot] = shape[d];
    plan.rows *= shape[d];
  }
  bool vec = true;
  for (int k = 0; k < nargs; ++k) {
    EwOperand& o = plan.op[k];
    o.ptr        = args[k].store ? static_cast<char*>(args[k].store->ptr) : nullptr;
    o.inner_stride = strides[k][n - 1];
    for (int d = 0; d < EW_MAX_OUTER; ++d) o.outer_stride[d] = 0;
    for (int d = 0; d < n - 1; ++d) o.outer_stride[EW_MAX_OUTER - (n - 1) + d] = strides[k][d];
    if (args[k].store == nullptr || args[k].index_only) continue;
    const bool bcast = (o.inner_stride == 0) && !args[k].is_output;
    const long long align = bcast ? args[k].itemsize : chunk_bytes[k];
    if (!bcast && o.inner_stride != args[k].itemsize) vec = false;
    if (reinterpret_cast<uintptr_t>(o.ptr) % align != 0) vec = false;
    for (int d = 0; d < EW_MAX_OUTER; ++d)
      if (o.outer_stride[d] % align != 0) vec = false;
  }
  plan.vec = vec ? 1 : 0;
  // strided path: rows may be shifted by up to one 128-byte line of the output so that warp
  // stores are line-aligned (see ew_kernel); reserve the slack
  plan.out_pad = 0;
  if (!vec && args[0].store != nullptr && args[0].is_output &&
      plan.op[0].inner_stride == args[0].itemsize && args[0].itemsize < 128)
    plan.out_pad = 128 / args[0].itemsize - 1;
  plan.tiles_per_row = (plan.inner + plan.out_pad + tile_elems - 1) / tile_elems;
  plan.num_tiles     = plan.tiles_per_row * plan.rows;
  return 1;
}

long long ew_algorithmic_bytes(const EwPlan& plan, const EwArg* args, int nargs)
{
  long long total = 0;
  for (int k = 0; k < nargs; ++k) {
    if (args[k].store == nullptr || args[k].index_only) continue;
    long long distinct = (plan.op[k].inner_stride != 0) ? plan.inner : 1;
    for (int d = 0; d < EW_MAX_OUTER; ++d)
      if (plan.op[k].outer_stride[d] != 0) distinct *= plan.outer[d];
    total += distinct * args[k].itemsize;
  }
  return total;
}

int ew_grid_size(const void* kernel, long long num_tiles, int max_ctas_per_sm);

int ew_grid_size(const void* kernel, long long num_tiles, bool vec)
{
  static const int forced = [] {
    const char* e = getenv("CNB_EW_CTAS_PER_SM");
    return e ? atoi(e) : 0;
  }();
  // Fewer, fatter CTAs stream better: with ~128 bytes of loads in flight per thread a few resident
  // CTAs per SM already cover the HBM latency, and more concurrent tile streams only add DRAM page
  // conflicts.  Measured on B200 (profiles/r01_ctas_per_sm.md): 3 CTAs/SM is within 2-3 % of the
  // best setting for every kernel variant tried (2 is slightly better for some fp32/fp64 kernels
  // but costs 15-20 % on WHERE fp16/fp64 and scalar-operand fp16; 4+ loses 2-5 % across the board).
  // CNB_EW_CTAS_PER_SM overrides.
  (void)vec;
  return ew_grid_size(kernel, num_tiles, forced > 0 ? forced : 3);
}

int ew_grid_size(const void* kernel, long long num_tiles, int max_ctas_per_sm)
{
  // persistent CTAs: SMs x resident CTAs per SM (queried once per kernel)
  static std::mutex mu;
  static std::unordered_map<const void*, int> cache;
  int per_sm = 0;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(kernel);
    if (it != cache.end()) per_sm = it->second;
  }
  if (per_sm == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, EW_THREADS, 0) != cudaSuccess ||
        per_sm <= 0)
      per_sm = 4;
    std::lock_guard<std::mutex> g(mu);
    cache[kernel] = per_sm;
  }
  per_sm = std::min(per_sm, max_ctas_per_sm);
  long long resident = static_cast<long long>(sm_count()) * per_sm;
  return static_cast<int>(std::max<long long>(1, std::min<long long>(num_tiles, resident)));
}

}  // namespace cnb
