// Elementwise engine for BINARY_OP / UNARY_OP / WHERE / CONVERT / FILL on sm_100a.
//
// One kernel template serves every layout the task contract allows (dense, row-pitched views,
// stride-0 broadcast operands, permuted strides):
//   * the launcher canonicalises the iteration space on the host (drop unit dims, order dims by the
//     output's stride, merge dims that are jointly contiguous for every operand) into
//     rows x inner, so a dense task of any rank is ONE row and a stencil view is `N` pitched rows;
//   * CTAs are persistent (grid = SMs x resident CTAs) and walk tiles with a grid-stride loop; the
//     only integer division is one per TILE (row = tile / tiles_per_row), never per element —
//     the reference's generic kernel pays a 64-bit div+mod per element (pitches.h:46-55);
//   * a tile takes the 128-bit vector path when every operand is inner-contiguous (or broadcast)
//     and 16-byte aligned: each thread issues U independent 16-byte loads per operand before any
//     use (MLP), lanes are consecutive 16-byte chunks (4 sectors/request); otherwise it takes the
//     strided path (coalesced scalar accesses, same unroll).
#pragma once

#include "cnb_common.cuh"

namespace cnb {

constexpr int EW_THREADS   = 256;
constexpr int EW_MAX_OPS   = 5;  // <= 2 outputs + 3 inputs
constexpr int EW_MAX_OUTER = 3;

struct Unused {};  // placeholder operand type

template <typename T>
inline constexpr int ew_size = std::is_same<T, Unused>::value ? 0 : int(sizeof(T));

struct EwOperand {
  char* ptr;
  long long inner_stride;                // bytes; 0 = broadcast along the inner dim
  long long outer_stride[EW_MAX_OUTER];  // bytes, slowest dim first
};

struct EwPlan {
  long long inner;                // elements in the innermost (fastest) dim
  long long outer[EW_MAX_OUTER];  // outer extents, slowest first, padded with 1
  long long rows;                 // product of outer
  long long tiles_per_row;
  long long num_tiles;
  int n_outer;
  int vec;      // 1 = vector path legal for full tiles
  int out_pad;  // elements of slack per row for the store-alignment shift of the strided path
  EwOperand op[EW_MAX_OPS];  // outputs first, then inputs
};

struct EwArg {  // host-side description of one operand before canonicalisation
  const cnb_store_t* store;
  int itemsize;
  bool is_output;
  bool index_only = false;  // pseudo-operand (strides in index units): exempt from vector checks
};

// Builds the plan. Returns 1 if there is work, 0 if the iteration space is empty, <0 on error.
// `chunk_bytes[k]` = bytes one thread moves per vector access for operand k (alignment unit).
int ew_make_plan(EwPlan& plan, const EwArg* args, int nargs, const int* chunk_bytes, int tile_elems);

// ---------------------------------------------------------------------------------------------
template <typename T, int E>
struct alignas((sizeof(T) * E >= 16) ? 16 : sizeof(T) * E) Pack {
  unsigned char raw[sizeof(T) * E];
  __device__ __forceinline__ T& operator[](int i) { return reinterpret_cast<T*>(raw)[i]; }
  __device__ __forceinline__ const T& operator[](int i) const
  {
    return reinterpret_cast<const T*>(raw)[i];
  }
};

template <int BYTES>
__device__ __forceinline__ void ld_bytes(void* dst, const char* src)
{
  if constexpr (BYTES >= 16) {
    static_assert(BYTES % 16 == 0, "");
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i)
      reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else if constexpr (BYTES == 8) {
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
  } else if constexpr (BYTES == 4) {
    *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
  } else if constexpr (BYTES == 2) {
    *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<const uint16_t*>(src);
  } else {
    static_assert(BYTES == 1, "");
    *reinterpret_cast<uint8_t*>(dst) = *reinterpret_cast<const uint8_t*>(src);
  }
}

template <int BYTES>
__device__ __forceinline__ void st_bytes(char* dst, const void* src)
{
  if constexpr (BYTES >= 16) {
#pragma unroll
    for (int i = 0; i < BYTES / 16; ++i)
      reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  } else if constexpr (BYTES == 8) {
    *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(src);
  } else if constexpr (BYTES == 4) {
    *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(src);
  } else if constexpr (BYTES == 2) {
    *reinterpret_cast<uint16_t*>(dst) = *reinterpret_cast<const uint16_t*>(src);
  } else {
    *reinterpret_cast<uint8_t*>(dst) = *reinterpret_cast<const uint8_t*>(src);
  }
}

constexpr int ew_cmax(int a, int b) { return a > b ? a : b; }
constexpr int ew_cmin(int a, int b) { return a < b ? a : b; }
constexpr int ew_min_nz(int a, int b) { return a == 0 ? b : (b == 0 ? a : (a < b ? a : b)); }

// Per-functor tiling constants: E elements per vector chunk, U chunks per thread per tile.
template <class Fn>
struct EwShape {
  using O0 = typename Fn::O0;
  using O1 = typename Fn::O1;
  using I0 = typename Fn::I0;
  using I1 = typename Fn::I1;
  using I2 = typename Fn::I2;
  static constexpr int max_size =
    ew_cmax(ew_cmax(ew_cmax(ew_size<O0>, ew_size<O1>), ew_cmax(ew_size<I0>, ew_size<I1>)),
            ew_size<I2>);
  static constexpr int min_size = ew_min_nz(
    ew_min_nz(ew_min_nz(ew_size<O0>, ew_size<O1>), ew_min_nz(ew_size<I0>, ew_size<I1>)),
    ew_size<I2>);
  // Elements per chunk.  STORES must be exactly one <=16-byte access per lane, so that a warp's
  // store instruction covers one contiguous 512-byte span: several 16-byte stores per lane at a
  // 64-byte lane stride write half sectors that L2 has to merge (measured: int8->int64 convert at
  // 26 % of the roofline).  Loads may be wider than 16 bytes per chunk (the halves of a sector are
  // served by L1) but no input moves more than 64 bytes per chunk.  Kernels without outputs
  // (reductions) read 16 bytes of the narrowest input.
  static constexpr int max_out = ew_cmax(ew_size<O0>, ew_size<O1>);
  static constexpr int max_in  = ew_cmax(ew_cmax(ew_size<I0>, ew_size<I1>), ew_size<I2>);
  static constexpr int E =
    max_out == 0 ? ew_cmax(1, 16 / min_size)
                 : ew_cmax(1, max_in == 0 ? 16 / max_out : ew_cmin(16 / max_out, 64 / max_in));
  // chunks per thread per tile: keep ~128 bytes of LOADS in flight per thread (all inputs
  // together), so unary / widening-convert kernels get the same memory-level parallelism as the
  // two-input kernels
  static constexpr int in_bytes = ew_size<I0> + ew_size<I1> + ew_size<I2>;
  static constexpr int U =
    in_bytes == 0 ? 4 : ew_cmax(1, ew_cmin(8, 128 / (E * in_bytes)));
  static constexpr int TILE = EW_THREADS * E * U;
};

template <typename T, int E>
__device__ __forceinline__ void ew_load_vec(Pack<T, E>& r, const EwOperand& o, long long row_off,
                                            long long elem)
{
  if constexpr (!std::is_same<T, Unused>::value) {
    if (o.inner_stride == 0) {
      Pack<T, 1> s;
      ld_bytes<sizeof(T)>(s.raw, o.ptr + row_off);
#pragma unroll
      for (int i = 0; i < E; ++i) r[i] = s[0];
    } else {
      ld_bytes<sizeof(T) * E>(r.raw, o.ptr + row_off + elem * (long long)sizeof(T));
    }
  }
}

template <typename T>
__device__ __forceinline__ void ew_load_one(Pack<T, 1>& r, const EwOperand& o, long long row_off,
                                            long long elem)
{
  if constexpr (!std::is_same<T, Unused>::value)
    ld_bytes<sizeof(T)>(r.raw, o.ptr + row_off + elem * o.inner_stride);
}

// Two instantiations per functor, chosen by the launcher, so that neither path drags the other's
// registers along (one combined kernel needed 76-141 registers once the strided path batched its
// loads; capping it at 64 spilled in the vector path):
//   VEC = true : plan.vec tasks — 128-bit path for full tiles, a plain element loop for the one
//                ragged tile at the end of a row;
//   VEC = false: everything else — the batched strided path.
template <class Fn, bool VEC>
__global__ void __launch_bounds__(EW_THREADS)
ew_kernel(const __grid_constant__ EwPlan plan, const Fn fn)
{
  using S  = EwShape<Fn>;
  using O0 = typename Fn::O0;
  using O1 = typename Fn::O1;
  using I0 = typename Fn::I0;
  using I1 = typename Fn::I1;
  using I2 = typename Fn::I2;
  constexpr int E = S::E, U = S::U, TILE = S::TILE;
  constexpr bool HAS_O1 = !std::is_same<O1, Unused>::value;

  const int tid = threadIdx.x;
  for (long long tile = blockIdx.x; tile < plan.num_tiles; tile += gridDim.x) {
    long long row = 0, ct = tile;
    if (plan.rows > 1) {
      row = tile / plan.tiles_per_row;
      ct  = tile - row * plan.tiles_per_row;
    }
    // row -> byte offset of the row start for every operand
    long long off[EW_MAX_OPS];
#pragma unroll
    for (int k = 0; k < EW_MAX_OPS; ++k) off[k] = 0;
    if (plan.n_outer == 1) {
#pragma unroll
      for (int k = 0; k < EW_MAX_OPS; ++k) off[k] = row * plan.op[k].outer_stride[EW_MAX_OUTER - 1];
    } else if (plan.n_outer > 1) {
      long long r = row;
#pragma unroll
      for (int d = EW_MAX_OUTER - 1; d >= 0; --d) {
        const long long q = r / plan.outer[d];
        const long long i = r - q * plan.outer[d];
        r                 = q;
#pragma unroll
        for (int k = 0; k < EW_MAX_OPS; ++k) off[k] += i * plan.op[k].outer_stride[d];
      }
    }
    const long long col0 = ct * TILE;

    if constexpr (VEC) {
      if (col0 + TILE > plan.inner) {
        // ragged last tile of a row: one element per thread per step (all operands are
        // inner-contiguous or broadcast here)
        for (long long e = col0 + tid; e < plan.inner; e += EW_THREADS) {
          Pack<I0, 1> a;
          Pack<I1, 1> b;
          Pack<I2, 1> c;
          ew_load_one<I0>(a, plan.op[2], off[2], e);
          ew_load_one<I1>(b, plan.op[3], off[3], e);
          ew_load_one<I2>(c, plan.op[4], off[4], e);
          Pack<O0, 1> r0;
          Pack<O1, 1> r1;
          fn(r0[0], r1[0], a[0], b[0], c[0]);
          st_bytes<sizeof(O0)>(plan.op[0].ptr + off[0] + e * (long long)sizeof(O0), r0.raw);
          if constexpr (HAS_O1)
            st_bytes<sizeof(O1)>(plan.op[1].ptr + off[1] + e * (long long)sizeof(O1), r1.raw);
        }
        continue;
      }
      // ---- 128-bit vector path: U independent chunk loads per operand, then compute + store
      Pack<I0, E> a[U];
      Pack<I1, E> b[U];
      Pack<I2, E> c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = col0 + (long long)(u * EW_THREADS + tid) * E;
        ew_load_vec<I0, E>(a[u], plan.op[2], off[2], e);
        ew_load_vec<I1, E>(b[u], plan.op[3], off[3], e);
        ew_load_vec<I2, E>(c[u], plan.op[4], off[4], e);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long e = col0 + (long long)(u * EW_THREADS + tid) * E;
        Pack<O0, E> r0;
        Pack<O1, E> r1;
#pragma unroll
        for (int i = 0; i < E; ++i) fn(r0[i], r1[i], a[u][i], b[u][i], c[u][i]);
        st_bytes<sizeof(O0) * E>(plan.op[0].ptr + off[0] + e * (long long)sizeof(O0), r0.raw);
        if constexpr (HAS_O1)
          st_bytes<sizeof(O1) * E>(plan.op[1].ptr + off[1] + e * (long long)sizeof(O1), r1.raw);
      }
    } else {
      // ---- strided path: coalesced element accesses, batches of B independent loads (~64 bytes
      // in flight per thread)
      constexpr int B = S::in_bytes == 0 ? 8 : ew_cmax(2, ew_cmin(16, 64 / S::in_bytes));
      constexpr int N = E * U;  // elements per thread per tile
      // Shift the tile grid of this row so that warp stores start on a 128-byte line of the
      // OUTPUT: a view whose rows start mid-sector (the stencil's `center[:] = work` lands 8 bytes
      // past a sector) otherwise writes two partial sectors per warp store, which L2 can only
      // complete with read-modify-write traffic (measured 56 % of the roofline for that COPY).
      // Loads are indifferent to alignment.  ew_make_plan pads tiles_per_row for the shift.
      long long shift = 0;
      if (plan.out_pad != 0)
        shift = static_cast<long long>(
          (reinterpret_cast<unsigned long long>(plan.op[0].ptr + off[0]) & 127ull) / sizeof(O0));
      const long long tbase = col0 - shift;
#pragma unroll 1
      for (int j0 = 0; j0 < N; j0 += B) {
        Pack<I0, 1> a[B];
        Pack<I1, 1> b[B];
        Pack<I2, 1> c[B];
#pragma unroll
        for (int j = 0; j < B; ++j) {
          const long long e = tbase + (long long)(j0 + j) * EW_THREADS + tid;
          if (j0 + j < N && e >= 0 && e < plan.inner) {
            ew_load_one<I0>(a[j], plan.op[2], off[2], e);
            ew_load_one<I1>(b[j], plan.op[3], off[3], e);
            ew_load_one<I2>(c[j], plan.op[4], off[4], e);
          }
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
          const long long e = tbase + (long long)(j0 + j) * EW_THREADS + tid;
          if (j0 + j < N && e >= 0 && e < plan.inner) {
            Pack<O0, 1> r0;
            Pack<O1, 1> r1;
            fn(r0[0], r1[0], a[j][0], b[j][0], c[j][0]);
            st_bytes<sizeof(O0)>(plan.op[0].ptr + off[0] + e * plan.op[0].inner_stride, r0.raw);
            if constexpr (HAS_O1)
              st_bytes<sizeof(O1)>(plan.op[1].ptr + off[1] + e * plan.op[1].inner_stride, r1.raw);
          }
        }
      }
    }
  }
}

int ew_grid_size(const void* kernel, long long num_tiles, bool vec);
int ew_grid_size(const void* kernel, long long num_tiles, int max_ctas_per_sm);
// bytes the task must move: distinct elements touched per operand x itemsize (a stride-0 scalar
// operand counts once) — the roofline numerator of SURVEY §8(d)
long long ew_algorithmic_bytes(const EwPlan& plan, const EwArg* args, int nargs);

// Launch `Fn` over the stores. Operand order: o0, o1 (may be null), i0, i1, i2 (may be null).
template <class Fn>
int ew_launch(const Fn& fn, const cnb_store_t* o0, const cnb_store_t* o1, const cnb_store_t* i0,
              const cnb_store_t* i1, const cnb_store_t* i2, cudaStream_t stream)
{
  using S = EwShape<Fn>;
  EwArg args[EW_MAX_OPS] = {{o0, ew_size<typename Fn::O0>, true},
                            {o1, ew_size<typename Fn::O1>, true},
                            {i0, ew_size<typename Fn::I0>, false},
                            {i1, ew_size<typename Fn::I1>, false},
                            {i2, ew_size<typename Fn::I2>, false}};
  int chunk[EW_MAX_OPS];
  for (int k = 0; k < EW_MAX_OPS; ++k) chunk[k] = ew_cmin(16, args[k].itemsize * S::E);
  EwPlan plan;
  int rc = ew_make_plan(plan, args, EW_MAX_OPS, chunk, S::TILE);
  if (rc <= 0) return rc;
  auto kernel = plan.vec ? ew_kernel<Fn, true> : ew_kernel<Fn, false>;
  int grid    = ew_grid_size(reinterpret_cast<const void*>(kernel), plan.num_tiles, plan.vec != 0);
  {
    LaunchScope scope(stream, KERNEL_ELEMENTWISE, plan.inner * plan.rows,
                      ew_algorithmic_bytes(plan, args, EW_MAX_OPS));
    kernel<<<grid, EW_THREADS, 0, stream>>>(plan, fn);
  }
  return check_cuda(cudaGetLastError(), "ew_kernel launch");
}

}  // namespace cnb
