// Microbenchmark: which load/store cache hints give the best streaming throughput on B200 for the
// elementwise kernel shape (persistent CTAs, 128-bit accesses, U vectors in flight per thread)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o copy_variants copy_variants.cu
// Prints GB/s (read+write bytes) for a 1:1 copy, a 2:1 add and a 1:2 widening convert.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

enum { LD_DEFAULT, LD_CS, LD_NC_NOALLOC, LD_CG, LD_LU };
enum { ST_DEFAULT, ST_CS, ST_CG, ST_WT };

template <int M> __device__ __forceinline__ uint4 ld(const uint4* p) {
  uint4 v;
  if constexpr (M == LD_DEFAULT) v = *p;
  else if constexpr (M == LD_CS) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (M == LD_NC_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else if constexpr (M == LD_CG) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  else asm volatile("ld.global.lu.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
template <int M> __device__ __forceinline__ void st(uint4* p, uint4 v) {
  if constexpr (M == ST_DEFAULT) *p = v;
  else if constexpr (M == ST_CS) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else if constexpr (M == ST_CG) asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  else asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// NIN inputs of n vectors each, NOUT outputs of n vectors each per "element group"
template <int LM, int SM, int U, int NIN, int NOUT>
__global__ void __launch_bounds__(256) stream_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ o, uint4* __restrict__ o2, long long nvec)
{
  const long long tile = 256LL * U;
  const long long ntiles = nvec / tile;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    uint4 x[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = t * tile + u * 256 + threadIdx.x;
      x[u] = ld<LM>(a + i);
      if (NIN > 1) y[u] = ld<LM>(b + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = t * tile + u * 256 + threadIdx.x;
      uint4 r = x[u];
      if (NIN > 1) { r.x += y[u].x; r.y += y[u].y; r.z += y[u].z; r.w += y[u].w; }
      st<SM>(o + i, r);
      if (NOUT > 1) { r.x ^= 1; st<SM>(o2 + i, r); }
    }
  }
}

template <int LM, int SM, int U, int NIN, int NOUT>
float run(const char* name, const uint4* a, const uint4* b, uint4* o, uint4* o2, long long nvec, int ctas_per_sm, int sms)
{
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto k = stream_kernel<LM, SM, U, NIN, NOUT>;
  int grid = sms * ctas_per_sm;
  for (int i = 0; i < 3; ++i) k<<<grid, 256>>>(a, b, o, o2, nvec);
  float best = 1e9, tot = 0;
  for (int i = 0; i < 10; ++i) {
    CK(cudaEventRecord(e0));
    k<<<grid, 256>>>(a, b, o, o2, nvec);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; tot += ms;
  }
  double bytes = double(nvec) * 16 * (NIN + NOUT);
  printf("%-34s U=%d ctas=%d  in=%d out=%d  best %7.1f GB/s  mean %7.1f GB/s\n", name, U, ctas_per_sm, NIN, NOUT, bytes / best / 1e6, bytes / (tot / 10) / 1e6);
  return best;
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const long long nvec = (1LL << 30) / 4;  // 4 GiB per array
  uint4 *a, *b, *o, *o2;
  CK(cudaMalloc(&a, nvec * 16)); CK(cudaMalloc(&b, nvec * 16)); CK(cudaMalloc(&o, nvec * 16)); CK(cudaMalloc(&o2, nvec * 16));
  CK(cudaMemset(a, 1, nvec * 16)); CK(cudaMemset(b, 2, nvec * 16));
  // reference: cudaMemcpyAsync D2D
  {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9, tot = 0;
    for (int i = 0; i < 13; ++i) {
      CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(o, a, nvec * 16, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (i >= 3) { best = ms < best ? ms : best; tot += ms; }
    }
    printf("%-34s best %7.1f GB/s  mean %7.1f GB/s\n", "cudaMemcpyAsync D2D 4 GiB", nvec * 32.0 / best / 1e6, nvec * 32.0 / (tot / 10) / 1e6);
    best = 1e9; tot = 0;
    for (int i = 0; i < 13; ++i) {
      CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(o, 0, nvec * 16)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (i >= 3) { best = ms < best ? ms : best; tot += ms; }
    }
    printf("%-34s best %7.1f GB/s  mean %7.1f GB/s\n", "cudaMemsetAsync 4 GiB", nvec * 16.0 / best / 1e6, nvec * 16.0 / (tot / 10) / 1e6);
  }
#define ROW(LM, SM, U, NIN, NOUT, C) run<LM, SM, U, NIN, NOUT>(#LM "+" #SM, a, b, o, o2, nvec, C, sms)
  for (int c : {2, 3, 4, 8}) {
    ROW(LD_DEFAULT, ST_DEFAULT, 8, 1, 1, c);
    ROW(LD_CS, ST_CS, 8, 1, 1, c);
    ROW(LD_DEFAULT, ST_CS, 8, 1, 1, c);
    ROW(LD_NC_NOALLOC, ST_DEFAULT, 8, 1, 1, c);
    ROW(LD_NC_NOALLOC, ST_CS, 8, 1, 1, c);
    ROW(LD_CG, ST_CG, 8, 1, 1, c);
    ROW(LD_LU, ST_CS, 8, 1, 1, c);
    ROW(LD_DEFAULT, ST_WT, 8, 1, 1, c);
    ROW(LD_DEFAULT, ST_DEFAULT, 4, 1, 1, c);
    ROW(LD_CS, ST_CS, 4, 1, 1, c);
    ROW(LD_DEFAULT, ST_DEFAULT, 16, 1, 1, c);
  }
  printf("--- 2 inputs : 1 output (add)\n");
  for (int c : {2, 3, 4}) {
    ROW(LD_DEFAULT, ST_DEFAULT, 4, 2, 1, c);
    ROW(LD_CS, ST_CS, 4, 2, 1, c);
    ROW(LD_DEFAULT, ST_CS, 4, 2, 1, c);
    ROW(LD_NC_NOALLOC, ST_CS, 4, 2, 1, c);
    ROW(LD_LU, ST_CS, 4, 2, 1, c);
  }
  printf("--- 1 input : 2 outputs (widening)\n");
  for (int c : {2, 3, 4}) {
    ROW(LD_DEFAULT, ST_DEFAULT, 8, 1, 2, c);
    ROW(LD_CS, ST_CS, 8, 1, 2, c);
    ROW(LD_DEFAULT, ST_CS, 8, 1, 2, c);
    ROW(LD_DEFAULT, ST_WT, 8, 1, 2, c);
  }
  return 0;
}
