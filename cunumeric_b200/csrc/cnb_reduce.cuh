// Block-level reduction primitives shared by the scalar and axis reduction kernels:
// warp butterfly over 32-bit words (any trivially copyable Acc up to 16 bytes), then one shared-
// memory exchange across the CTA's warps.  The fold order is fixed, so results are deterministic
// for a given launch geometry.
#pragma once

#include "cnb_elementwise.cuh"

namespace cnb {

constexpr int RED_THREADS = 256;
constexpr int RED_WARPS   = RED_THREADS / 32;

template <typename A>
__device__ __forceinline__ A shfl_xor_any(const A& v, int lane_mask)
{
  constexpr int W = (sizeof(A) + 3) / 4;
  unsigned w[W];
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = 0;
  memcpy(w, &v, sizeof(A));
#pragma unroll
  for (int i = 0; i < W; ++i) w[i] = __shfl_xor_sync(0xffffffffu, w[i], lane_mask);
  A r;
  memcpy(&r, w, sizeof(A));
  return r;
}

template <class R, typename A>
__device__ __forceinline__ A warp_reduce(A v)
{
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    A o = shfl_xor_any(v, m);
    // lower lane id is always the left operand so every lane computes the same ordered fold
    v = ((threadIdx.x & m) == 0) ? R::fold(v, o) : R::fold(o, v);
  }
  return v;
}

// All threads of the CTA call this; the result is valid in thread 0.
template <class R, typename A>
__device__ __forceinline__ A block_reduce(A v, A* smem /* RED_WARPS entries */)
{
  v = warp_reduce<R>(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();  // protect smem reuse across calls
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    A t = (lane < RED_WARPS) ? smem[lane] : R::identity();
    // only lanes < RED_WARPS hold data; reduce those in lane order
#pragma unroll
    for (int m = RED_WARPS / 2; m > 0; m >>= 1) {
      A o = shfl_xor_any(t, m);
      t   = ((lane & m) == 0) ? R::fold(t, o) : R::fold(o, t);
    }
    v = t;
  }
  return v;
}

// Fold one input element into a thread's accumulator.  `index` is only evaluated by the
// arg-reductions; `ordered` promises that this thread visits elements in increasing index order,
// which lets arg-reductions use the cheap strictly-better test instead of the general tie-breaking
// fold.
template <class R, class IndexFn>
__device__ __forceinline__ void red_visit(const R& r, typename R::Acc& acc,
                                          const typename R::In& x, bool ordered, IndexFn&& index)
{
  if constexpr (R::needs_index) {
    if (ordered) {
      if (R::better(acc.value, x)) {
        acc.value = x;
        acc.arg   = index();
      }
    } else {
      acc = R::fold(acc, r.convert(x, index()));
    }
  } else {
    acc = R::fold(acc, r.convert(x, 0));
  }
}

// raw storage for an Acc in shared memory (Acc may have a non-trivial default constructor)
template <typename A, int N>
struct alignas(16) RawSmem {
  unsigned char raw[sizeof(A) * N];
  __device__ __forceinline__ A* ptr() { return reinterpret_cast<A*>(raw); }
};

}  // namespace cnb
