// UNARY_RED (single-axis reduction) for sm_100a, included by axis_red_g*.cu.
//
// The launcher canonicalises the task into  kept dims (<=3, merged, fastest last) x axis  using the
// real byte strides, so transposed / sliced inputs pick their mode from the memory layout, not from
// the logical axis number:
//   COLUMN mode (a kept dim is the contiguous one, e.g. axis 0 of a C-order matrix): lanes run
//     along the contiguous kept dim with 128-bit loads, the 8 warps of a CTA stride over the axis,
//     the axis is split across gridDim.y CTAs for occupancy; partials meet in shared memory, split
//     partials in a scratch buffer that the last-arriving CTA of each column tile folds IN SPLIT
//     ORDER.  No atomics on data — the reference finishes every thread with a global atomic /
//     CAS loop (unary_red.cu:298-312, arg.inl:52-82).
//   ROW mode (the axis itself is contiguous, e.g. axis 1): one CTA (long rows) or one warp (short
//     rows) per output element, 128-bit loads along the axis, shuffle + shared-memory fold.
// Results are folded into the caller's pre-filled output store (reduce-accessor semantics).
#include "cnb_reduce.cuh"
#include "cnb_tma.cuh"
#include "ops_reduce.cuh"

#include <algorithm>
#include <cstdlib>

namespace cnb {

void* pool_alloc(size_t nbytes, cudaStream_t stream);
int pool_free(void* p, cudaStream_t stream);

namespace {

constexpr int AX_KEPT = 3;

struct AxisPlan {
  long long kept[AX_KEPT];       // kept extents, slowest first, right-aligned, padded with 1
  long long in_k[AX_KEPT];       // byte strides of `in` over kept dims
  long long out_k[AX_KEPT];      // byte strides of `out`
  long long w_k[AX_KEPT];        // byte strides of `where`
  long long alen, in_a, w_a;     // axis extent and strides
  long long axis_origin;
  long long ncols;               // product of kept
  const char* in;
  const char* where;             // nullptr = no mask
  char* out;
  int vec;                       // vector loads legal along the lane dimension
  // column mode
  long long tiles_fast;          // column tiles along kept[2]
  long long split_len;           // axis rows per split
  int nsplit;
  int warps_x;                   // warps of a CTA laid across the contiguous kept dim (1,2,4,8)
};

// ---------------------------------------------------------------------------------------------
// COLUMN mode.  The CTA's 8 warps are laid out as WX warps across the contiguous kept dim x
// WY = 8/WX row-lanes down the axis.  For wide outputs WX = 8: the whole CTA reads ONE row segment of
// 8*32*V contiguous elements (4 KB for fp32) per step, rows in axis order, UNR rows in flight per
// thread — long contiguous DRAM bursts instead of 512-byte pieces 128 KB apart — and every thread
// owns its V columns outright, so no shared-memory exchange is needed at all.
template <class R, int V>
__device__ __forceinline__ void axis_col_body(const AxisPlan& p, const R& r, char* scratch_raw,
                                              unsigned int* tickets)
{
  using T   = typename R::In;
  using Acc = typename R::Acc;
  using Val = typename R::Val;
  const int tx = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int WX = p.warps_x, WY = RED_WARPS / WX;
  const int wx = warp % WX, wy = warp / WX;
  const int lanes_w = WX * 32 * V;  // columns per CTA tile
  const int lane_c  = (wx * 32 + tx) * V;

  // column tile -> (position along the fast kept dim, index over the slower kept dims)
  const long long tile  = blockIdx.x;
  const long long slow  = tile / p.tiles_fast;
  const long long tfast = tile - slow * p.tiles_fast;
  const long long k1    = slow % p.kept[1];
  const long long k0    = slow / p.kept[1];
  const long long c0    = tfast * lanes_w + lane_c;  // first column of this lane
  const bool active     = c0 < p.kept[2];
  const long long in_base  = k0 * p.in_k[0] + k1 * p.in_k[1] + c0 * p.in_k[2];
  const long long w_base   = k0 * p.w_k[0] + k1 * p.w_k[1] + c0 * p.w_k[2];
  const long long out_base = k0 * p.out_k[0] + k1 * p.out_k[1] + c0 * p.out_k[2];
  const int nvalid = active ? (int)min((long long)V, p.kept[2] - c0) : 0;

  Acc acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = R::identity();

  const long long a_begin = (long long)blockIdx.y * p.split_len;
  const long long a_end   = min(p.alen, a_begin + p.split_len);
  constexpr int UNR       = 4;
  if (active) {
    for (long long a0 = a_begin + wy; a0 < a_end; a0 += (long long)WY * UNR) {
      Pack<T, V> x[UNR];
      bool ok[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const long long a = a0 + (long long)u * WY;
        ok[u]             = a < a_end;
        if (ok[u]) ld_bytes<sizeof(T) * V>(x[u].raw, p.in + in_base + a * p.in_a);
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (ok[u]) {
          const long long a = a0 + (long long)u * WY;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            bool m = v < nvalid;
            if (m && p.where != nullptr)
              m = *reinterpret_cast<const unsigned char*>(p.where + w_base + v * p.w_k[2] +
                                                          a * p.w_a) != 0;
            // each thread walks the axis in increasing order: ordered fast path for arg-reductions
            if (m) red_visit(r, acc[v], x[u][v], true, [&] { return p.axis_origin + a; });
          }
        }
      }
    }
  }

  // fold the WY row-lanes of the CTA in wy order (nothing to do when the CTA is one row wide)
  __shared__ RawSmem<Acc, RED_WARPS * 32 * V> smem;
  if (WY > 1) {
    Acc* sm = smem.ptr();
#pragma unroll
    for (int v = 0; v < V; ++v) sm[wy * lanes_w + lane_c + v] = acc[v];
    __syncthreads();
    if (wy == 0) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        Acc t = sm[lane_c + v];
        for (int w = 1; w < WY; ++w) t = R::fold(t, sm[w * lanes_w + lane_c + v]);
        acc[v] = t;
      }
    }
  }

  if (p.nsplit == 1) {
    if (wy == 0) {
      for (int v = 0; v < nvalid; ++v) {
        Val* o = reinterpret_cast<Val*>(p.out + out_base + v * p.out_k[2]);
        *o     = R::finish(R::fold(R::lift(*o), acc[v]));
      }
    }
    return;
  }

  // split partials -> scratch[split][tile][lanes_w]; the last CTA of the tile folds them in split
  // order
  Acc* scratch            = reinterpret_cast<Acc*>(scratch_raw);
  const long long per_spl = (long long)gridDim.x * lanes_w;
  if (wy == 0) {
#pragma unroll
    for (int v = 0; v < V; ++v)
      scratch[(long long)blockIdx.y * per_spl + tile * lanes_w + lane_c + v] = acc[v];
  }
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[tile], 1u);
    is_last              = (t == (unsigned int)p.nsplit - 1);
  }
  __syncthreads();
  if (is_last && wy == 0) {
    __threadfence();
    for (int v = 0; v < nvalid; ++v) {
      Acc t = R::identity();
      for (int s = 0; s < p.nsplit; ++s) {
        Acc q;
        ld_bytes<sizeof(Acc)>(&q, reinterpret_cast<const char*>(
                                    &scratch[(long long)s * per_spl + tile * lanes_w + lane_c + v]));
        t = R::fold(t, q);
      }
      Val* o = reinterpret_cast<Val*>(p.out + out_base + v * p.out_k[2]);
      *o     = R::finish(R::fold(R::lift(*o), t));
    }
  }
}

// One kernel per vector width so neither drags the other's registers along; 3 CTAs/SM resident
// (the first version compiled both bodies into one kernel: 102 registers, 2 CTAs/SM, 25 %
// occupancy and 56 % of the roofline — profiles/r01_axis_reduction_ncu.md).
template <class R, int V>
__global__ void __launch_bounds__(RED_THREADS, 3)
axis_col_kernel(const __grid_constant__ AxisPlan p, const R r, char* scratch, unsigned int* tickets)
{
  axis_col_body<R, V>(p, r, scratch, tickets);
}

// COLUMN mode, common case (dense C-order matrix reduced along axis 0, no mask): the same
// algorithm as axis_col_body with WX = 8, stripped to what the hot loop needs so that it fits in
// 64 registers (4 CTAs/SM).  Thread t of column tile b owns columns (b*256 + t)*V .. +V for all rows
// of its axis split; the CTA reads one contiguous 256*V-element row segment per step.
template <class R, int V, int UNR>
__global__ void __launch_bounds__(RED_THREADS, 4)
axis_col_fast_kernel(const char* __restrict__ in, char* __restrict__ out, char* scratch_raw,
                     unsigned int* tickets, const R r, long long ncols, long long alen,
                     long long in_a, long long out_stride, long long split_len, int nsplit,
                     long long axis_origin)
{
  using T   = typename R::In;
  using Acc = typename R::Acc;
  using Val = typename R::Val;
  const long long c0      = ((long long)blockIdx.x * RED_THREADS + threadIdx.x) * V;
  const bool active       = c0 < ncols;
  const long long a_begin = (long long)blockIdx.y * split_len;
  const long long a_end   = min(alen, a_begin + split_len);
  Acc acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = R::identity();
  if (active) {
    const char* p = in + c0 * (long long)sizeof(T) + a_begin * in_a;
    long long a   = a_begin;
    for (; a + UNR <= a_end; a += UNR) {
      Pack<T, V> x[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) ld_bytes<sizeof(T) * V>(x[u].raw, p + u * in_a);
      p += UNR * in_a;
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v)
          red_visit(r, acc[v], x[u][v], true, [&] { return axis_origin + a + u; });
    }
    for (; a < a_end; ++a) {
      Pack<T, V> x;
      ld_bytes<sizeof(T) * V>(x.raw, p);
      p += in_a;
#pragma unroll
      for (int v = 0; v < V; ++v) red_visit(r, acc[v], x[v], true, [&] { return axis_origin + a; });
    }
  }
  if (nsplit == 1) {
    if (active) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        Val* o = reinterpret_cast<Val*>(out + (c0 + v) * out_stride);
        *o     = R::finish(R::fold(R::lift(*o), acc[v]));
      }
    }
    return;
  }
  Acc* scratch            = reinterpret_cast<Acc*>(scratch_raw);
  const long long width   = (long long)gridDim.x * RED_THREADS * V;  // padded column count
  const long long my_slot = ((long long)blockIdx.x * RED_THREADS + threadIdx.x) * V;
#pragma unroll
  for (int v = 0; v < V; ++v) scratch[(long long)blockIdx.y * width + my_slot + v] = acc[v];
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[blockIdx.x], 1u);
    is_last              = (t == (unsigned int)nsplit - 1);
  }
  __syncthreads();
  if (is_last && active) {
    __threadfence();
#pragma unroll
    for (int v = 0; v < V; ++v) {
      Acc t = R::identity();
      for (int s = 0; s < nsplit; ++s) {
        Acc q;
        ld_bytes<sizeof(Acc)>(&q, reinterpret_cast<const char*>(&scratch[(long long)s * width + my_slot + v]));
        t = R::fold(t, q);
      }
      Val* o = reinterpret_cast<Val*>(out + (c0 + v) * out_stride);
      *o     = R::finish(R::fold(R::lift(*o), t));
    }
  }
}

template <typename T>
inline constexpr int axis_vec_width =
  (16 / sizeof(T)) > 4 ? 4 : ((16 / sizeof(T)) < 1 ? 1 : int(16 / sizeof(T)));

// ---------------------------------------------------------------------------------------------
// ROW mode: LPR lanes cooperate on one output element (LPR = 32: a warp, LPR = 256: the CTA)
template <class R, int LPR>
__global__ void __launch_bounds__(RED_THREADS)
axis_row_kernel(const __grid_constant__ AxisPlan p, const R r)
{
  using T   = typename R::In;
  using Acc = typename R::Acc;
  using Val = typename R::Val;
  constexpr int V    = (16 / sizeof(T)) < 1 ? 1 : 16 / sizeof(T);
  constexpr int ROWS = RED_THREADS / LPR;  // outputs per CTA iteration
  const int lane     = threadIdx.x % LPR;
  const int sub      = threadIdx.x / LPR;
  __shared__ RawSmem<Acc, RED_WARPS> smem;

  const long long ngroups = (p.ncols + ROWS - 1) / ROWS;
  for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const long long c = g * ROWS + sub;
    const bool active = c < p.ncols;
    long long in_base = 0, w_base = 0, out_base = 0;
    if (active) {
      long long q        = c;
      const long long k2 = q % p.kept[2];
      q /= p.kept[2];
      const long long k1 = q % p.kept[1];
      const long long k0 = q / p.kept[1];
      in_base            = k0 * p.in_k[0] + k1 * p.in_k[1] + k2 * p.in_k[2];
      w_base             = k0 * p.w_k[0] + k1 * p.w_k[1] + k2 * p.w_k[2];
      out_base           = k0 * p.out_k[0] + k1 * p.out_k[1] + k2 * p.out_k[2];
    }
    Acc acc = R::identity();
    if (active) {
      // the row start may be misaligned for 16-byte loads even when the matrix is: peel
      long long a = 0;
      bool vec    = p.vec && V > 1 && p.where == nullptr;
      if (vec) {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(p.in + in_base);
        const int mis        = (int)(addr % 16);
        long long peel       = mis ? (16 - mis) / (long long)sizeof(T) : 0;
        peel                 = min(peel, p.alen);
        if (lane < peel) {
          Pack<T, 1> x;
          ld_bytes<sizeof(T)>(x.raw, p.in + in_base + lane * (long long)sizeof(T));
          red_visit(r, acc, x[0], true, [&] { return p.axis_origin + lane; });
        }
        a = peel;
        constexpr int UNR   = 4;
        const long long nv  = (p.alen - a) / V;  // full vectors
        const char* base    = p.in + in_base + a * (long long)sizeof(T);
        long long i         = lane;
        for (; i + (long long)(UNR - 1) * LPR < nv; i += (long long)UNR * LPR) {
          Pack<T, V> x[UNR];
#pragma unroll
          for (int u = 0; u < UNR; ++u)
            ld_bytes<sizeof(T) * V>(x[u].raw, base + (i + (long long)u * LPR) * (16));
#pragma unroll
          for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int v = 0; v < V; ++v)
              red_visit(r, acc, x[u][v], true,
                        [&] { return p.axis_origin + a + (i + (long long)u * LPR) * V + v; });
        }
        for (; i < nv; i += LPR) {
          Pack<T, V> x;
          ld_bytes<sizeof(T) * V>(x.raw, base + i * 16);
#pragma unroll
          for (int v = 0; v < V; ++v)
            red_visit(r, acc, x[v], true, [&] { return p.axis_origin + a + i * V + v; });
        }
        a += nv * V;
      }
      // scalar remainder / strided / masked path
      for (long long i = a + lane; i < p.alen; i += LPR) {
        bool m = true;
        if (p.where != nullptr)
          m = *reinterpret_cast<const unsigned char*>(p.where + w_base + i * p.w_a) != 0;
        if (m) {
          Pack<T, 1> x;
          ld_bytes<sizeof(T)>(x.raw, p.in + in_base + i * p.in_a);
          red_visit(r, acc, x[0], true, [&] { return p.axis_origin + i; });
        }
      }
    }
    if constexpr (LPR == 32) {
      acc = warp_reduce<R>(acc);
    } else {
      acc = block_reduce<R>(acc, smem.ptr());
    }
    if (active && lane == 0) {
      Val* o = reinterpret_cast<Val*>(p.out + out_base);
      *o     = R::finish(R::fold(R::lift(*o), acc));
    }
  }
}

// ROW mode, long contiguous rows: bulk-copy (TMA) pipeline.  ROWT_CTAS_PER_SM persistent CTAs per
// SM; warp 8's elected lane streams every row of the CTA through a ring of ROWT_STAGES x ROWT_CHUNK
// bytes of shared memory with cp.async.bulk, the 8 consumer warps fold the chunks out of shared
// memory.  The producer runs up to a full ring ahead, so HBM stays busy while the consumers do the
// cross-warp fold and the store of a finished row — the bubble that holds the plain LDG kernel
// at ~90 % of the copy roofline.  Three CTAs per SM (192 KB of ring in total) rather than one fat
// one, so that the row epilogue of one CTA — ~250 dependent instructions for an arg-reduction —
// overlaps the folding done by the other two (one CTA per SM: SUM 104 %, ARGMAX 62 %).  Requirements (checked by the launcher): contiguous axis, no mask,
// 16-byte aligned base / row pitch / row length.
constexpr int ROWT_CONSUMERS = RED_THREADS;
constexpr int ROWT_THREADS   = RED_THREADS + 32;
constexpr int ROWT_STAGES    = 4;
constexpr int ROWT_CTAS_PER_SM = 3;
constexpr int ROWT_CHUNK     = 16384;
constexpr int ROWT_SMEM      = ROWT_STAGES * ROWT_CHUNK + 1024;

template <class R>
__global__ void __launch_bounds__(ROWT_THREADS, ROWT_CTAS_PER_SM)
axis_row_tma_kernel(const __grid_constant__ AxisPlan p, const R r)
{
  using T   = typename R::In;
  using Acc = typename R::Acc;
  using Val = typename R::Val;
  constexpr int V = (16 / sizeof(T)) < 1 ? 1 : 16 / sizeof(T);
  static_assert(sizeof(T) * V == 16, "one 128-bit shared-memory load per step");
  extern __shared__ __align__(1024) unsigned char rowt_smem[];
  // [ring | full barriers | empty barriers | cross-warp partials (2 x 8 Acc)]
  unsigned char* ring = rowt_smem;
  uint64_t* bars      = reinterpret_cast<uint64_t*>(rowt_smem + ROWT_STAGES * ROWT_CHUNK);
  Acc* partials       = reinterpret_cast<Acc*>(rowt_smem + ROWT_STAGES * ROWT_CHUNK + 256);
  static_assert(2 * RED_WARPS * sizeof(Acc) <= 768, "partials fit behind the barriers");
  const uint32_t full0  = smem_u32(bars);
  const uint32_t empty0 = smem_u32(bars + ROWT_STAGES);
  const uint32_t ring0  = smem_u32(ring);

  if (threadIdx.x == 0) {
    for (int s = 0; s < ROWT_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);           // one arrive.expect_tx by the producer
      mbar_init(empty0 + 8 * s, RED_WARPS);  // one arrive per consumer warp
    }
    mbar_fence_init();
  }
  __syncthreads();

  const long long row_bytes = p.alen * (long long)sizeof(T);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int stage      = 0;
  uint32_t phase = 0;

  if (warp == RED_WARPS) {
    // ---- producer
    if (lane != 0) return;
    for (long long c = blockIdx.x; c < p.ncols; c += gridDim.x) {
      long long q        = c;
      const long long k2 = q % p.kept[2];
      q /= p.kept[2];
      const long long k1 = q % p.kept[1];
      const long long k0 = q / p.kept[1];
      const char* src    = p.in + k0 * p.in_k[0] + k1 * p.in_k[1] + k2 * p.in_k[2];
      for (long long off = 0; off < row_bytes; off += ROWT_CHUNK) {
        const uint32_t nb = (uint32_t)min((long long)ROWT_CHUNK, row_bytes - off);
        mbar_wait(empty0 + 8 * stage, phase ^ 1u);  // passes at once on a fresh barrier
        mbar_arrive_expect_tx(full0 + 8 * stage, nb);
        bulk_g2s(ring0 + stage * ROWT_CHUNK, src + off, nb, full0 + 8 * stage);
        if (++stage == ROWT_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    return;
  }

  // ---- consumers
  const int tid = threadIdx.x;
  int parity    = 0;
  for (long long c = blockIdx.x; c < p.ncols; c += gridDim.x) {
    // UNR independent accumulator chains per thread: the fold of one element must not wait for
    // the previous one (with 8 consumer warps per SM a single dependent chain made ARGMAX
    // latency-bound at 64 % of the roofline).  Every chain sees increasing indices.
    constexpr int UNR = ROWT_CHUNK / 16 / ROWT_CONSUMERS;  // 4
    Acc a[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) a[u] = R::identity();
    for (long long off = 0; off < row_bytes; off += ROWT_CHUNK) {
      const int nvec = (int)(min((long long)ROWT_CHUNK, row_bytes - off) >> 4);
      mbar_wait(full0 + 8 * stage, phase);
      const char* buf    = reinterpret_cast<const char*>(ring) + stage * ROWT_CHUNK;
      const long long e0 = p.axis_origin + off / (long long)sizeof(T);
      if (nvec == ROWT_CHUNK / 16) {
        Pack<T, V> x[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) ld_bytes<16>(x[u].raw, buf + (u * ROWT_CONSUMERS + tid) * 16);
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
          for (int u = 0; u < UNR; ++u)
            red_visit(r, a[u], x[u][v], true,
                      [&] { return e0 + ((u * ROWT_CONSUMERS + tid) * V + v); });
      } else {
        for (int i = tid; i < nvec; i += ROWT_CONSUMERS) {
          Pack<T, V> x;
          ld_bytes<16>(x.raw, buf + i * 16);
#pragma unroll
          for (int v = 0; v < V; ++v)
            red_visit(r, a[0], x[v], true, [&] { return e0 + (i * V + v); });
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * stage);
      if (++stage == ROWT_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
    // chains hold interleaved index sets: the general (tie-breaking) fold merges them
    Acc acc = R::fold(R::fold(a[0], a[1]), R::fold(a[2], a[3]));
    static_assert(UNR == 4, "chain merge above is written for 4 chains");
    // fold the row across the 8 consumer warps; partials are double-buffered by row parity so one
    // named barrier per row is enough (a warp can only overwrite buffer b after passing the NEXT
    // row's barrier, which warp 0 reaches after it has read b)
    acc      = warp_reduce<R>(acc);
    Acc* buf = partials + parity * RED_WARPS;
    if (lane == 0) buf[warp] = acc;
    named_bar_sync(1, ROWT_CONSUMERS);
    if (warp == 0) {
      Acc t = (lane < RED_WARPS) ? buf[lane] : R::identity();
#pragma unroll
      for (int m = RED_WARPS / 2; m > 0; m >>= 1) {
        Acc o = shfl_xor_any(t, m);
        t     = ((lane & m) == 0) ? R::fold(t, o) : R::fold(o, t);
      }
      if (lane == 0) {
        long long q        = c;
        const long long k2 = q % p.kept[2];
        q /= p.kept[2];
        const long long k1 = q % p.kept[1];
        const long long k0 = q / p.kept[1];
        Val* o = reinterpret_cast<Val*>(p.out + k0 * p.out_k[0] + k1 * p.out_k[1] + k2 * p.out_k[2]);
        *o     = R::finish(R::fold(R::lift(*o), t));
      }
    }
    parity ^= 1;
  }
}

template <class R>
int launch_axis_row_tma(const AxisPlan& p, int sms, cudaStream_t stream)
{
  auto kernel            = axis_row_tma_kernel<R>;
  static const int ready = [&] {
    return check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           ROWT_SMEM),
                      "axis_row_tma_kernel smem opt-in");
  }();
  if (ready != CNB_OK) return ready;
  const long long g = std::min<long long>(p.ncols, (long long)sms * ROWT_CTAS_PER_SM);
  kernel<<<(unsigned)g, ROWT_THREADS, ROWT_SMEM, stream>>>(p, R(nullptr));
  return check_cuda(cudaGetLastError(), "axis_row_tma_kernel launch");
}

template <int OP>
int axis_red_by_type(int axis, const cnb_store_t* out, const cnb_store_t* in,
                     const cnb_store_t* where, long long axis_origin, cudaStream_t stream)
{
  return type_dispatch(in->dtype, [&](auto tag) -> int {
    using T = type_of<decltype(tag)::value>;
    using R = typename RedFn<OP>::template fn<T>;
    if constexpr (!R::valid || OP == CNB_RED_CONTAINS) {
      return set_error(CNB_ERR_INVALID_OP, "UNARY_RED %d is not valid for dtype %d", OP, in->dtype);
    } else {
      using Acc = typename R::Acc;
      using Val = typename R::Val;
      const int nd = in->ndim;
      if (nd < 1 || nd > CNB_MAX_DIM || axis < 0 || axis >= nd)
        return set_error(CNB_ERR_BAD_ARG, "UNARY_RED: axis %d out of range for ndim %d", axis, nd);
      if (out->ndim != nd)
        return set_error(CNB_ERR_BAD_ARG, "UNARY_RED: out must be promoted to in's rank");
      if (out->dtype != CodeOf<Val>::value)
        return set_error(CNB_ERR_BAD_ARG, "UNARY_RED %d on dtype %d: out dtype %d, expected %d", OP,
                         in->dtype, out->dtype, CodeOf<Val>::value);
      if (where != nullptr && (where->dtype != CNB_BOOL || where->ndim != nd))
        return set_error(CNB_ERR_BAD_ARG, "UNARY_RED: bad where mask");
      for (int d = 0; d < nd; ++d) {
        if (d != axis && out->shape[d] != in->shape[d])
          return set_error(CNB_ERR_BAD_ARG, "UNARY_RED: out extent mismatch on dim %d", d);
        if (where != nullptr && where->shape[d] != in->shape[d])
          return set_error(CNB_ERR_BAD_ARG, "UNARY_RED: where extent mismatch on dim %d", d);
        if (in->shape[d] == 0) return CNB_OK;  // unary_red_template.inl:47
      }

      // ---- canonicalise kept dims
      struct Dim {
        long long n, si, so, sw;
      } dims[CNB_MAX_DIM];
      int nk = 0;
      for (int d = 0; d < nd; ++d) {
        if (d == axis || in->shape[d] == 1) continue;
        dims[nk++] = {in->shape[d], in->strides[d], out->strides[d], where ? where->strides[d] : 0};
      }
      std::stable_sort(dims, dims + nk, [](const Dim& a, const Dim& b) {
        return std::llabs(a.si) > std::llabs(b.si);
      });
      int m = 0;
      for (int d = 1; d < nk; ++d) {
        const Dim& nx = dims[d];
        if (dims[m].si == nx.n * nx.si && dims[m].so == nx.n * nx.so && dims[m].sw == nx.n * nx.sw) {
          dims[m].n *= nx.n;
          dims[m].si = nx.si;
          dims[m].so = nx.so;
          dims[m].sw = nx.sw;
        } else {
          dims[++m] = nx;
        }
      }
      if (nk > 0) nk = m + 1;

      AxisPlan p{};
      for (int d = 0; d < AX_KEPT; ++d) {
        p.kept[d] = 1;
        p.in_k[d] = p.out_k[d] = p.w_k[d] = 0;
      }
      p.ncols = 1;
      for (int d = 0; d < nk; ++d) {
        const int slot = AX_KEPT - nk + d;
        p.kept[slot]   = dims[d].n;
        p.in_k[slot]   = dims[d].si;
        p.out_k[slot]  = dims[d].so;
        p.w_k[slot]    = dims[d].sw;
        p.ncols *= dims[d].n;
      }
      p.alen        = in->shape[axis];
      p.in_a        = in->strides[axis];
      p.w_a         = where ? where->strides[axis] : 0;
      p.axis_origin = axis_origin;
      p.in          = static_cast<const char*>(in->ptr);
      p.where       = where ? static_cast<const char*>(where->ptr) : nullptr;
      p.out         = static_cast<char*>(out->ptr);
      p.nsplit      = 1;
      p.split_len   = p.alen;
      p.tiles_fast  = 1;

      const long long isz = sizeof(T);
      // input once (+ mask) + one read-modify-write of every output element
      const long long algo_bytes = p.ncols * p.alen * (isz + (where ? 1 : 0)) +
                                   2 * p.ncols * (long long)sizeof(Val);
      const bool col_mode = nk > 0 && std::llabs(p.in_k[AX_KEPT - 1]) < std::llabs(p.in_a) &&
                            !(p.alen == 1);
      const int sms = sm_count();
      if (col_mode || (nk > 0 && p.alen == 1)) {
        constexpr int V = axis_vec_width<T>;
        const long long vb = V * isz;
        bool vec = V > 1 && p.in_k[2] == isz && reinterpret_cast<uintptr_t>(p.in) % vb == 0 &&
                   p.in_a % vb == 0 && p.in_k[0] % vb == 0 && p.in_k[1] % vb == 0 &&
                   p.kept[2] % V == 0;
        p.vec               = vec ? 1 : 0;
        if (vec && where == nullptr && nk == 1 && p.ncols >= 4LL * RED_THREADS * V) {
          // fast kernel: dense matrix, contiguous kept dim
          const long long tiles = (p.ncols + (long long)RED_THREADS * V - 1) / ((long long)RED_THREADS * V);
          // one wave: tiles x splits must not exceed the resident CTAs (4 per SM), or a few
          // straggler CTAs double the kernel time
          long long want = std::max<long long>(1, (4LL * sms) / tiles);
          long long maxs = std::max<long long>(1, p.alen / 64);
          int nsplit     = (int)std::max<long long>(1, std::min<long long>(std::min(want, maxs), 65535));
          long long split_len = (p.alen + nsplit - 1) / nsplit;
          nsplit              = (int)((p.alen + split_len - 1) / split_len);
          char* scratch         = nullptr;
          unsigned int* tickets = nullptr;
          if (nsplit > 1) {
            const size_t sbytes = (size_t)nsplit * tiles * RED_THREADS * V * sizeof(Acc);
            const size_t tbytes = (size_t)tiles * sizeof(unsigned int);
            scratch = static_cast<char*>(pool_alloc(sbytes + tbytes, stream));
            if (scratch == nullptr) return CNB_ERR_CUDA;
            tickets = reinterpret_cast<unsigned int*>(scratch + sbytes);
            int rc  = check_cuda(cudaMemsetAsync(tickets, 0, tbytes, stream), "ticket memset");
            if (rc != CNB_OK) return rc;
          }
          dim3 grid((unsigned)tiles, (unsigned)nsplit);
          {
            LaunchScope scope(stream, KERNEL_AXIS_COL, p.ncols * p.alen, algo_bytes);
            axis_col_fast_kernel<R, V, 4><<<grid, RED_THREADS, 0, stream>>>(
              p.in, p.out, scratch, tickets, R(nullptr), p.ncols, p.alen, p.in_a, p.out_k[2],
              split_len, nsplit, p.axis_origin);
          }
          int rc = check_cuda(cudaGetLastError(), "axis_col_fast_kernel launch");
          if (scratch != nullptr) pool_free(scratch, stream);
          return rc;
        }
        const int warp_w    = 32 * (vec ? V : 1);
        int wxs             = RED_WARPS;
        while (wxs > 1 && (long long)(wxs / 2) * warp_w >= p.kept[2]) wxs /= 2;
        p.warps_x           = wxs;
        const int lanes_w   = wxs * warp_w;
        p.tiles_fast        = (p.kept[2] + lanes_w - 1) / lanes_w;
        const long long tiles = p.tiles_fast * p.kept[0] * p.kept[1];
        if (tiles > 0x7fffffffLL) return set_error(CNB_ERR_UNSUPPORTED, "UNARY_RED: too many tiles");
        // split the axis until one full wave (3 CTAs per SM) is in flight, keeping >= 64 rows per
        // split
        long long want = std::max<long long>(1, (3LL * sms) / tiles);
        long long maxs = std::max<long long>(1, p.alen / 64);
        int nsplit     = (int)std::max<long long>(1, std::min<long long>(std::min(want, maxs), 65535));
        p.split_len    = (p.alen + nsplit - 1) / nsplit;
        nsplit         = (int)((p.alen + p.split_len - 1) / p.split_len);
        p.nsplit       = nsplit;
        char* scratch         = nullptr;
        unsigned int* tickets = nullptr;
        if (nsplit > 1) {
          const size_t sbytes = (size_t)nsplit * tiles * lanes_w * sizeof(Acc);
          const size_t tbytes = (size_t)tiles * sizeof(unsigned int);
          scratch = static_cast<char*>(pool_alloc(sbytes + tbytes, stream));
          if (scratch == nullptr) return CNB_ERR_CUDA;
          tickets = reinterpret_cast<unsigned int*>(scratch + sbytes);
          int rc  = check_cuda(cudaMemsetAsync(tickets, 0, tbytes, stream), "ticket memset");
          if (rc != CNB_OK) return rc;
        }
        dim3 grid((unsigned)tiles, (unsigned)nsplit);
        {
          LaunchScope scope(stream, KERNEL_AXIS_COL, p.ncols * p.alen, algo_bytes);
          if (vec)
            axis_col_kernel<R, V><<<grid, RED_THREADS, 0, stream>>>(p, R(nullptr), scratch, tickets);
          else
            axis_col_kernel<R, 1><<<grid, RED_THREADS, 0, stream>>>(p, R(nullptr), scratch, tickets);
        }
        int rc = check_cuda(cudaGetLastError(), "axis_col_kernel launch");
        if (scratch != nullptr) pool_free(scratch, stream);
        return rc;
      } else {
        p.vec = (p.in_a == isz) ? 1 : 0;
        LaunchScope scope(stream, KERNEL_AXIS_ROW, p.ncols * p.alen, algo_bytes);
        static const bool use_tma = [] {
          const char* e = getenv("CNB_AXIS_ROW_TMA");
          return e == nullptr || atoi(e) != 0;
        }();
        const bool tma_ok = use_tma && sizeof(T) <= 16 && p.vec && where == nullptr &&
                            reinterpret_cast<uintptr_t>(p.in) % 16 == 0 && p.in_k[0] % 16 == 0 &&
                            p.in_k[1] % 16 == 0 && p.in_k[2] % 16 == 0 &&
                            (p.alen * isz) % 16 == 0 && p.alen * isz >= ROWT_CHUNK &&
                            p.ncols >= sms / 2;
        if (tma_ok) return launch_axis_row_tma<R>(p, sms, stream);
        if (p.alen >= 2048) {
          auto kernel = axis_row_kernel<R, RED_THREADS>;
          long long g = std::min<long long>(p.ncols, (long long)sms * 8);
          kernel<<<(unsigned)std::max<long long>(1, g), RED_THREADS, 0, stream>>>(p, R(nullptr));
        } else {
          auto kernel    = axis_row_kernel<R, 32>;
          long long ngrp = (p.ncols + RED_WARPS - 1) / RED_WARPS;
          long long g    = std::min<long long>(ngrp, (long long)sms * 8);
          kernel<<<(unsigned)std::max<long long>(1, g), RED_THREADS, 0, stream>>>(p, R(nullptr));
        }
        return check_cuda(cudaGetLastError(), "axis_row_kernel launch");
      }
    }
  });
}

}  // namespace

int CNB_ARED_GROUP_NAME(int op, int axis, const cnb_store_t* out, const cnb_store_t* in,
                        const cnb_store_t* where, long long axis_origin, cudaStream_t stream)
{
  switch (op) {
#define X(OPCODE) \
  case OPCODE: return axis_red_by_type<OPCODE>(axis, out, in, where, axis_origin, stream);
    CNB_ARED_GROUP_OPS(X)
#undef X
  }
  return 1;
}

}  // namespace cnb
