// Micro-benchmark / correctness probe for the TMA-staged 2-D tile pipeline the fused-chain generator
// emits for shifted views of one pitched buffer (the 5-point Jacobi chain of examples/stencil.py:
// out.center = 0.2 * ((((c + n) + e) + w) + s), one read + one write of the grid per point).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o stencil_tma stencil_tma.cu
//   ./stencil_tma [N=40000] [iters=20]
//
// Every (TR, TC, STAGES) configuration is first checked bit-for-bit against a CPU evaluation on a
// ragged grid (N = 1002; the row pitch must be a multiple of 16 bytes for a tensor map, i.e. N even), then timed on the full grid with the two buffers alternating.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));   \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
  uint32_t ok;
  do {
    asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar)
{
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
    "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
    : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar,
                                                 unsigned long long policy)
{
  asm volatile(
    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
    "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar), "l"(policy)
    : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_last()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first()
{
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void st_hint(double* p, double v, unsigned long long policy)
{
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y)
{
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                 reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int THREADS = 256;
__host__ __device__ constexpr int align128(int x) { return (x + 127) / 128 * 128; }

// One CTA owns tiles blockIdx.x, +gridDim.x, ...  Thread 0 keeps S tile loads in flight (TMA, 2-D
// boxes with a halo); all 256 threads pull their operands out of shared memory into registers,
// hand the stage back, compute and store straight to global memory.  Tiles are aligned to 128-byte
// lines of the OUTPUT buffer (so every warp store covers whole lines), and the box origin of a load
// is rounded down to a 16-byte boundary of the input row — the tensor engine rejects boxes that
// start at an odd fp64 column (measured: "illegal instruction" for x = 1, fine for x = 0 / -2).
template <int TR, int TC, int S>
__global__ void __launch_bounds__(THREADS)
stencil_tma(const __grid_constant__ CUtensorMap in_map, double* __restrict__ out, int n, long long pitch,
            int tiles_x, int ntiles, double factor, int colmajor, int hint, unsigned int* counter)
{
  constexpr int IW = TC + 4, IH = TR + 2;     // 2 halo columns each side (1 needed + 1 for alignment)
  constexpr int IN_BYTES  = IW * IH * 8;
  constexpr int IN_STRIDE = align128(IN_BYTES);
  constexpr int GROUPS    = THREADS / TC;   // thread groups stacked over the rows of a tile
  constexpr int RPT       = TR / GROUPS;    // consecutive rows per thread
  static_assert(THREADS % TC == 0 && TR % GROUPS == 0, "tile shape");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[S];
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_async_smem();
  }
  __syncthreads();
  // each CTA walks a contiguous range of tiles in COLUMN-major order: consecutive tiles are vertical
  // neighbours, so the two halo rows of a tile were fetched by the tile before it (L2 hits)
  const unsigned long long pol_ld = policy_evict_last(), pol_st = policy_evict_first();
  // Tile order: row-major over (band, tx) "runs" of V vertically consecutive tiles; a CTA takes whole
  // runs round-robin.  Inside a run the halo rows of a tile were fetched by the same CTA a moment
  // ago (certain L2 hits); the set of tiles in flight on the chip is a band of V*TR rows, i.e. a
  // few DRAM pages and TLB entries wide.
  const int V       = colmajor < 1 ? 1 : colmajor;
  const int tiles_y = ntiles / tiles_x;
  const int bands   = (tiles_y + V - 1) / V;
  const int nruns   = bands * tiles_x;
  const int first = 0, stride = 1;
  // dynamic scheduling (hint & 8): thread 0 draws run numbers from a global counter when it ISSUES the
  // loads of a run's first tile and leaves them in a shared ring for the consumers, so the CTAs
  // always work on a compact window of runs however unevenly they progress
  __shared__ int run_ring[32];
  const bool dynamic = (hint & 8) != 0;
  auto tile_xy = [&](int k, int& tx, int& ty) -> bool {
    const int run = dynamic ? run_ring[(k / V) & 31] : blockIdx.x + (k / V) * gridDim.x;
    if (run >= nruns) return false;
    const int band = run / tiles_x;
    tx = run - band * tiles_x;
    ty = band * V + (k % V);
    return true;   // (ty may be >= tiles_y in the last band: an empty tile, nothing loaded or stored)
  };
  auto issue = [&](int k) {
    int tx, ty;
    if (dynamic && k % V == 0) run_ring[(k / V) & 31] = (int)atomicAdd(counter, 1u);
    if (tile_xy(k, tx, ty)) {
      const int s  = k % S;
      mbar_expect_tx(smem_u32(&full[s]), IN_BYTES);
      if (hint & 1)
        tma_load_2d_hint(smem_u32(smem + s * IN_STRIDE), &in_map, tx * TC - 2, ty * TR, smem_u32(&full[s]), pol_ld);
      else
        tma_load_2d(smem_u32(smem + s * IN_STRIDE), &in_map, tx * TC - 2, ty * TR, smem_u32(&full[s]));
    }
  };
  if (tid == 0)
    for (int k = 0; k < S; ++k) issue(k);
  __syncthreads();
  const int cx = tid % TC, r0 = (tid / TC) * RPT;
  for (int k = 0;; ++k) {
    int tx, ty;
    if (!tile_xy(k, tx, ty)) break;
    const int s = k % S;
    mbar_wait(smem_u32(&full[s]), (k / S) & 1);
    const double* t = reinterpret_cast<const double*>(smem + s * IN_STRIDE);
    double mid[RPT + 2], west[RPT], east[RPT];
#pragma unroll
    for (int j = 0; j < RPT + 2; ++j) mid[j] = t[(r0 + j) * IW + cx + 2];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      west[j] = t[(r0 + j + 1) * IW + cx + 1];
      east[j] = t[(r0 + j + 1) * IW + cx + 3];
    }
    __syncthreads();                 // everybody has its operands: the stage can be refilled
    if (tid == 0) issue(k + S);
    const int col = tx * TC + cx;            // buffer column of this thread's outputs
    const int row = 1 + ty * TR + r0;        // buffer row of its first output
    if (col >= 1 && col <= n) {
      double* o = out + (long long)row * pitch + col;
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        // center + north + east + west + south, then 0.2 * average (examples/stencil.py:44-46)
        const double avg = (((mid[j + 1] + mid[j]) + east[j]) + west[j]) + mid[j + 2];
        if (row + j <= n) {
          if (hint & 2)
            st_hint(o + (long long)j * pitch, factor * avg, pol_st);
          else if (hint & 4)
            __stcs(o + (long long)j * pitch, factor * avg);
          else
            o[(long long)j * pitch] = factor * avg;
        }
      }
    }
  }
}

// the scalar-load path the generator emits today for misaligned views (one point per thread-step)
__global__ void __launch_bounds__(256) stencil_ldg(const double* __restrict__ in, double* __restrict__ out, int n,
                                                    long long pitch, double factor)
{
  const long long total = (long long)n * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n, c = i - r * n;
    const double* p   = in + (r + 1) * pitch + c + 1;
    const double avg  = (((p[0] + p[-pitch]) + p[1]) + p[-1]) + p[pitch];
    out[(r + 1) * pitch + c + 1] = factor * avg;
  }
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiled g_encode = nullptr;
static int g_hint = 0, g_v = 1;
static CUtensorMapL2promotion g_promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;

static CUtensorMap make_map(void* base, uint64_t w, uint64_t h, uint64_t pitch_bytes, uint32_t bw, uint32_t bh)
{
  CUtensorMap m;
  cuuint64_t dims[2]    = {w, h};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2]     = {bw, bh};
  cuuint32_t es[2]      = {1, 1};
  CUresult r = g_encode(&m, getenv("U64") ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        g_promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d (w=%llu h=%llu pitch=%llu box=%ux%u)\n", (int)r,
            (unsigned long long)w, (unsigned long long)h, (unsigned long long)pitch_bytes, bw, bh);
    exit(1);
  }
  return m;
}

template <int TR, int TC, int S>
static void launch(double* in, double* out, int n, int ctas_per_sm, cudaStream_t st)
{
  const long long pitch = n + 2;
  CUtensorMap im = make_map(in, n + 2, n + 2, pitch * 8, TC + 4, TR + 2);
  const int tiles_x = (n + 1 + TC - 1) / TC, tiles_y = (n + TR - 1) / TR;
  const int smem    = S * align128((TC + 4) * (TR + 2) * 8);
  auto kern         = stencil_tma<TR, TC, S>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int sms = 0, occ = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
  if (ctas_per_sm > 0 && occ > ctas_per_sm) occ = ctas_per_sm;
  const int ntiles = tiles_x * tiles_y;
  const int grid   = ntiles < sms * occ ? ntiles : sms * occ;
  const int colmajor = g_v;
  const int hint = g_hint;
  static unsigned int* counter = nullptr;
  if (!counter) CK(cudaMalloc(&counter, 4));
  CK(cudaMemsetAsync(counter, 0, 4, st));
  kern<<<grid, THREADS, smem, st>>>(im, out, n, pitch, tiles_x, ntiles, 0.2, colmajor, hint, counter);
}

static void fill_grid(std::vector<double>& g, int n)
{
  const long long p = n + 2;
  for (long long r = 0; r < p; ++r)
    for (long long c = 0; c < p; ++c) g[r * p + c] = 0.001 * ((r * 131 + c * 17) % 1009) - 0.3;
}

template <int TR, int TC, int S>
static bool check(int n)
{
  const long long p = n + 2;
  std::vector<double> h(p * p), ref(p * p, -7.0), got(p * p);
  fill_grid(h, n);
  for (long long r = 1; r <= n; ++r)
    for (long long c = 1; c <= n; ++c) {
      const double* q = &h[r * p + c];
      const double avg = (((q[0] + q[-p]) + q[1]) + q[-1]) + q[p];
      ref[r * p + c]   = 0.2 * avg;
    }
  double *din, *dout;
  CK(cudaMalloc(&din, p * p * 8));
  CK(cudaMalloc(&dout, p * p * 8));
  CK(cudaMemcpy(din, h.data(), p * p * 8, cudaMemcpyHostToDevice));
  std::vector<double> init(p * p, -7.0);
  CK(cudaMemcpy(dout, init.data(), p * p * 8, cudaMemcpyHostToDevice));
  launch<TR, TC, S>(din, dout, n, 0, 0);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(got.data(), dout, p * p * 8, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (long long i = 0; i < p * p; ++i)
    if (got[i] != ref[i] && bad++ < 5)
      fprintf(stderr, "  mismatch at (%lld,%lld): got %.17g want %.17g\n", i / p, i % p, got[i], ref[i]);
  CK(cudaFree(din));
  CK(cudaFree(dout));
  return bad == 0;
}

template <int TR, int TC, int S>
static void bench(double* a, double* b, int n, int iters, int ctas)
{
  bool ok = check<TR, TC, S>(1002) && check<TR, TC, S>(254);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("TMA TR=%-3d TC=%-3d S=%-2d hint=%d V=%d : %s ", TR, TC, S, g_hint, g_v, ok ? "exact" : "WRONG");
  double best = 1e9;
  for (int trial = 0; trial < 4; ++trial) {
    for (int i = 0; i < 3; ++i) launch<TR, TC, S>(i & 1 ? b : a, i & 1 ? a : b, n, ctas, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) launch<TR, TC, S>(i & 1 ? b : a, i & 1 ? a : b, n, ctas, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double per = ms / iters;
    printf(" %.3f", per);
    if (per < best) best = per;
  }
  printf(" ms/iter | best %.1f GB/s (16 B/pt)\n", 16.0 * n * n / best / 1e6);
  fflush(stdout);
}

int main(int argc, char** argv)
{
  const int n     = argc > 1 ? atoi(argv[1]) : 40000;
  const int iters = argc > 2 ? atoi(argv[2]) : 20;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&g_encode), cudaEnableDefault, &q));
  if (g_encode == nullptr) {
    fprintf(stderr, "cuTensorMapEncodeTiled not available\n");
    return 1;
  }
  const long long p = n + 2;
  double *a, *b;
  const size_t offa = getenv("OFFA") ? atol(getenv("OFFA")) : 0, offb = getenv("OFFB") ? atol(getenv("OFFB")) : 0;
  if (getenv("ALLOC_ASYNC")) {
    cudaMemPool_t pool;
    CK(cudaDeviceGetDefaultMemPool(&pool, 0));
    uint64_t thr = UINT64_MAX;
    CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    CK(cudaMallocAsync((void**)&a, p * p * 8 + (2 << 20), 0));
    CK(cudaMallocAsync((void**)&b, p * p * 8 + (2 << 20), 0));
    printf("cudaMallocAsync blocks\n");
  } else {
    CK(cudaMalloc(&a, p * p * 8 + (2 << 20)));
    CK(cudaMalloc(&b, p * p * 8 + (2 << 20)));
  }
  a = (double*)((char*)a + offa);
  b = (double*)((char*)b + offb);
  printf("a=%p b=%p\n", (void*)a, (void*)b);
  CK(cudaMemset(a, 0, p * p * 8));
  CK(cudaMemset(b, 0, p * p * 8));
  // baseline: the scalar-load kernel shape
  {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) stencil_ldg<<<148 * 8, 256>>>(a, b, n, p, 0.2);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) stencil_ldg<<<148 * 8, 256>>>(i & 1 ? b : a, i & 1 ? a : b, n, p, 0.2);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("LDG one point per thread-step       :        %.3f ms/iter  %.1f GB/s (16 B/pt)\n", ms / iters,
           16.0 * n * n / (ms / iters) / 1e6);
  }
  const int cfg = getenv("CONFIG") ? atoi(getenv("CONFIG")) : -1;
  if (getenv("PROMO")) {
    const int p = atoi(getenv("PROMO"));
    g_promo = p == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : p == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
              : p == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    printf("L2 promotion %d\n", p);
  }
  const int hints[] = {1, 9};
  const int vs[] = {1, 8};
  for (int h : hints) for (int v : vs) {
    g_hint = h; g_v = v;
    if (cfg < 0 || cfg == 1) bench<8, 128, 6>(a, b, n, iters, 0);
    if (cfg < 0 || cfg == 2) bench<8, 128, 8>(a, b, n, iters, 0);
    if (cfg < 0 || cfg == 3) bench<16, 128, 4>(a, b, n, iters, 0);
    if (cfg < 0 || cfg == 5) bench<32, 128, 2>(a, b, n, iters, 0);
    if (cfg < 0 || cfg == 6) bench<4, 128, 16>(a, b, n, iters, 0);
  }
  return 0;
}
