// Harness: runs the GENERATED TMA stencil kernel (copied from cunumeric_b200/_fused_cache) outside
// the Python runtime, to separate kernel-code effects from runtime effects.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "gen_stencil_kernel_v1.inc"
#define KERNEL fused_d0709ca0964f3dba1925_tma
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1);} } while (0)
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 40000, iters = 20;
  const int cshift = getenv("CSHIFT") ? atoi(getenv("CSHIFT")) : 3;
  EncodeTiled enc; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const long long p = n + 2;
  char *a, *b; double* sc;
  CK(cudaMalloc(&a, p * p * 8)); CK(cudaMalloc(&b, p * p * 8)); CK(cudaMalloc(&sc, 8));
  CK(cudaMemset(a, 0, p * p * 8)); CK(cudaMemset(b, 0, p * p * 8));
  double f = 0.2; CK(cudaMemcpy(sc, &f, 8, cudaMemcpyHostToDevice));
  auto kern = KERNEL;
  const int smem = S * STAGE;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int trial = 0; trial < 3; ++trial) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) {
      char* in = i & 1 ? b : a; char* out = i & 1 ? a : b;
      TParams P; memset(&P, 0, sizeof(P));
      cuuint64_t dims[2] = {(cuuint64_t)p, (cuuint64_t)p}; cuuint64_t st[1] = {(cuuint64_t)p * 8};
      cuuint32_t box[2] = {132, 10}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc((CUtensorMap*)&P.maps[0], CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, in, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { fprintf(stderr, "encode %d\n", (int)r); return 1; }
      P.inner = n; P.rows = n; P.cshift = cshift;
      P.tiles_x = (n + cshift + TC - 1) / TC; P.num_tiles = P.tiles_x * (((n + TR * V - 1) / (TR * V)) * V);
      const int gx = ((0 - cshift) >= 0 ? (0 - cshift) / 2 * 2 : -(((cshift) + 1) / 2 * 2));
      P.gx[0] = gx; P.gshift[0] = 0 - cshift - gx; P.gy[0] = 0;
      P.out[0].ptr = out + p * 8 + 8; P.out[0].row_stride = p * 8;
      P.scalar[0] = (const char*)sc;
      kern<<<296, 256, smem>>>(P);
    }
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("generated kernel in harness, cshift=%d: %.3f ms/iter\n", cshift, ms / iters);
  }
  return 0;
}
