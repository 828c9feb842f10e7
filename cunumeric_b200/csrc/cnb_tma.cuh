// Bulk asynchronous copies (the TMA engine's 1-D mode, SASS UBLKCP) and the shared-memory
// mbarriers that track them — thin wrappers over the sm_90+/sm_100a PTX.  Used by the kernels that
// stream long contiguous runs through a shared-memory ring (axis_red.inl ROW mode): one elected
// producer lane keeps STAGES x CHUNK bytes in flight per SM with a handful of instructions, and the
// loads of the next output element overlap the cross-warp fold of the current one.
#pragma once

#include <cstdint>

namespace cnb {

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// make freshly initialised barriers visible to the async proxy before the first bulk copy
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
    "selp.u32 %0, 1, 0, p;\n"
    "}\n"
    : "=r"(ok)
    : "r"(bar), "r"(parity)
    : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  while (!mbar_try_wait(bar, parity)) {}
}

// global -> shared bulk copy; src, dst and bytes must be multiples of 16.  Completion is reported
// to `bar` as `bytes` transaction bytes.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
  asm volatile(
    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
    "l"(src), "r"(bytes), "r"(bar)
    : "memory");
}

// ---- 2-D tiled mode (cp.async.bulk.tensor, SASS UTMALDG.2D): a box of a pitched 2-D tensor described
// by a CUtensorMap lands densely in shared memory.  x / y are element coordinates of the box origin;
// x * element size must be a multiple of 16 bytes (measured on B200: an odd fp64 column raises
// "illegal instruction"), rows are free.  Out-of-bounds parts of the box are zero-filled and still
// count towards the transaction bytes.  `policy` is an L2 cache policy (see l2_evict_last).
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tensor_map, int x, int y, uint32_t bar,
                                            unsigned long long policy)
{
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
    "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
    "l"(reinterpret_cast<unsigned long long>(tensor_map)), "r"(x), "r"(y), "r"(bar), "l"(policy)
    : "memory");
}

// Halo rows / columns of a tile are fetched again by the neighbouring tiles within a few
// microseconds: evict_last keeps them in L2 across that gap (measured on the 5-point stencil:
// DRAM reads 15.2 GB -> 13.7 GB per sweep of a 12.8 GB grid, 4.9 -> 4.0 ms).
__device__ __forceinline__ unsigned long long l2_evict_last()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// named barrier over a subset of the CTA's threads (count must be a multiple of 32)
__device__ __forceinline__ void named_bar_sync(int id, int count)
{
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

}  // namespace cnb
