// C ABI task entry points (include/cunumeric_b200.h): argument validation + dispatch to the
// per-opcode-group translation units.
#include "cnb_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace cnb {
int ensure_init();
void* pool_alloc(size_t nbytes, cudaStream_t stream);
int pool_free(void* p, cudaStream_t stream);

#define CNB_DECL_BIN(N) \
  int binary_group##N(int, const cnb_store_t*, const cnb_store_t*, const cnb_store_t*, const void*, cudaStream_t);
CNB_DECL_BIN(1) CNB_DECL_BIN(2) CNB_DECL_BIN(3) CNB_DECL_BIN(4) CNB_DECL_BIN(5)
#define CNB_DECL_UN(N) \
  int unary_group##N(int, const cnb_store_t*, const cnb_store_t*, const void*, cudaStream_t);
CNB_DECL_UN(1) CNB_DECL_UN(2) CNB_DECL_UN(3) CNB_DECL_UN(4) CNB_DECL_UN(5)
#define CNB_DECL_CVT(N) int convert_group##N(int, const cnb_store_t*, const cnb_store_t*, cudaStream_t);
CNB_DECL_CVT(1) CNB_DECL_CVT(2) CNB_DECL_CVT(3) CNB_DECL_CVT(4)
#define CNB_DECL_SRED(N)                                                                      \
  int scalar_red_group##N(int, const cnb_store_t*, const cnb_store_t*, const cnb_store_t*,    \
                          const int64_t*, const int64_t*, const void*, cudaStream_t);
CNB_DECL_SRED(1) CNB_DECL_SRED(2) CNB_DECL_SRED(3)
#define CNB_DECL_ARED(N)                                                                      \
  int axis_red_group##N(int, int, const cnb_store_t*, const cnb_store_t*, const cnb_store_t*, \
                        long long, cudaStream_t);
CNB_DECL_ARED(1) CNB_DECL_ARED(2) CNB_DECL_ARED(3) CNB_DECL_ARED(4) CNB_DECL_ARED(5)

int unary_multiout(int, const cnb_store_t*, const cnb_store_t*, const cnb_store_t*, cudaStream_t);
int unary_getarg(const cnb_store_t*, const cnb_store_t*, cudaStream_t);
int where_select(const cnb_store_t*, const cnb_store_t*, const cnb_store_t*, const cnb_store_t*,
                 cudaStream_t);
int fill_value(const cnb_store_t*, const void*, cudaStream_t);
int binary_red(int, const cnb_store_t*, const cnb_store_t*, const cnb_store_t*, const void*,
               cudaStream_t);

namespace {
int check_store(const cnb_store_t* s, const char* name)
{
  if (s == nullptr) return set_error(CNB_ERR_BAD_ARG, "%s store is NULL", name);
  if (s->ndim < 0 || s->ndim > CNB_MAX_DIM)
    return set_error(CNB_ERR_BAD_ARG, "%s store: ndim %d out of range", name, s->ndim);
  if (dtype_size(s->dtype) == 0)
    return set_error(CNB_ERR_BAD_ARG, "%s store: unknown dtype %d", name, s->dtype);
  return CNB_OK;
}
#define CNB_CHECK(expr)          \
  do {                           \
    int _rc = (expr);            \
    if (_rc != CNB_OK) return _rc; \
  } while (0)

std::mutex g_redop_mu;
std::map<int32_t, int32_t> g_argval_types;  // type_uid -> element dtype code

int axis_red_dispatch(int op, int axis, const cnb_store_t* out, const cnb_store_t* in,
                      const cnb_store_t* where, long long axis_origin, cudaStream_t s)
{
  int rc;
  if ((rc = axis_red_group1(op, axis, out, in, where, axis_origin, s)) != 1) return rc;
  if ((rc = axis_red_group2(op, axis, out, in, where, axis_origin, s)) != 1) return rc;
  if ((rc = axis_red_group3(op, axis, out, in, where, axis_origin, s)) != 1) return rc;
  if ((rc = axis_red_group4(op, axis, out, in, where, axis_origin, s)) != 1) return rc;
  if ((rc = axis_red_group5(op, axis, out, in, where, axis_origin, s)) != 1) return rc;
  if (op == CNB_RED_CONTAINS)
    return set_error(CNB_ERR_INVALID_OP, "CONTAINS exists on the scalar path only");
  return set_error(CNB_ERR_BAD_ARG, "unknown reduction opcode %d", op);
}

// ---- a few very long rows: split each row over several CTAs -------------------------------------
// ROW mode (the axis is the contiguous dim) gives one CTA (or one pipeline slot) per OUTPUT element.
// With fewer outputs than half the SMs — sum(axis=1) of an (8, 1e8) array — most of the chip idles
// (the reference sizes its grid to an occupancy wave regardless of the shape, unary_red.cu:148-161).
// Such a task is run in two stages with the same kernels: stage 1 reduces S segments of every row
// into a [outputs x S] scratch of partials (pre-filled with the identity), stage 2 reduces the
// scratch along S into the caller's (pre-filled) output.  Value reductions only: an arg-reduction's
// partial is an Argval, which no kernel takes as input.
int second_stage_op(int op)
{
  switch (op) {
    case CNB_RED_SUM:
    case CNB_RED_NANSUM:
    case CNB_RED_COUNT_NONZERO: return CNB_RED_SUM;
    case CNB_RED_PROD:
    case CNB_RED_NANPROD: return CNB_RED_PROD;
    case CNB_RED_MAX:
    case CNB_RED_NANMAX: return CNB_RED_MAX;
    case CNB_RED_MIN:
    case CNB_RED_NANMIN: return CNB_RED_MIN;
    case CNB_RED_ALL: return CNB_RED_ALL;
    case CNB_RED_ANY: return CNB_RED_ANY;
    default: return -1;
  }
}

// identity of the SECOND-stage reduction on partials of dtype `code` (what stage 1 folds into)
bool partial_identity(int op2, int code, unsigned char* buf)
{
  std::memset(buf, 0, 16);
  auto put = [&](auto v) { std::memcpy(buf, &v, sizeof(v)); };
  const bool is_max = op2 == CNB_RED_MAX, is_min = op2 == CNB_RED_MIN;
  switch (op2) {
    case CNB_RED_SUM:
    case CNB_RED_ANY: return true;  // all-zero bytes
    case CNB_RED_ALL: put((unsigned char)1); return code == CNB_BOOL;
    case CNB_RED_PROD:
      switch (code) {
        case CNB_BOOL:
        case CNB_INT8:
        case CNB_UINT8: put((unsigned char)1); return true;
        case CNB_INT16:
        case CNB_UINT16: put((uint16_t)1); return true;
        case CNB_INT32:
        case CNB_UINT32: put((uint32_t)1); return true;
        case CNB_INT64:
        case CNB_UINT64: put((uint64_t)1); return true;
        case CNB_FLOAT16: put((uint16_t)0x3c00); return true;
        case CNB_FLOAT32: put(1.0f); return true;
        case CNB_FLOAT64: put(1.0); return true;
        case CNB_COMPLEX64: put(1.0f); return true;  // (1, 0)
        default: return false;
      }
    default: break;
  }
  if (!is_max && !is_min) return false;
  switch (code) {  // MAX: the lowest value, MIN: the highest
    case CNB_BOOL: put((unsigned char)(is_max ? 0 : 1)); return true;
    case CNB_INT8: put((int8_t)(is_max ? INT8_MIN : INT8_MAX)); return true;
    case CNB_INT16: put((int16_t)(is_max ? INT16_MIN : INT16_MAX)); return true;
    case CNB_INT32: put((int32_t)(is_max ? INT32_MIN : INT32_MAX)); return true;
    case CNB_INT64: put((int64_t)(is_max ? INT64_MIN : INT64_MAX)); return true;
    case CNB_UINT8: put((uint8_t)(is_max ? 0 : UINT8_MAX)); return true;
    case CNB_UINT16: put((uint16_t)(is_max ? 0 : UINT16_MAX)); return true;
    case CNB_UINT32: put((uint32_t)(is_max ? 0 : UINT32_MAX)); return true;
    case CNB_UINT64: put((uint64_t)(is_max ? 0 : UINT64_MAX)); return true;
    case CNB_FLOAT16: put((uint16_t)(is_max ? 0xfc00 : 0x7c00)); return true;   // -inf / +inf
    case CNB_FLOAT32: put(is_max ? -HUGE_VALF : HUGE_VALF); return true;
    case CNB_FLOAT64: put(is_max ? -HUGE_VAL : HUGE_VAL); return true;
    default: return false;  // complex MAX / MIN: lexicographic identities, not worth the special case
  }
}

// returns 1 if the task was not split (the caller runs it as usual), else the status of the split run
int try_split_long_rows(int op, int axis, const cnb_store_t* out, const cnb_store_t* in,
                        cudaStream_t s)
{
  static const bool enabled = [] {
    const char* e = getenv("CNB_AXIS_ROW_SPLIT");
    return e == nullptr || atoi(e) != 0;
  }();
  const int op2 = second_stage_op(op);
  if (!enabled || op2 < 0 || in->ndim < 2 || in->ndim + 1 > CNB_MAX_DIM || axis < 0 || axis >= in->ndim)
    return 1;
  const long long isz = (long long)dtype_size(in->dtype);
  const long long vsz = (long long)dtype_size(out->dtype);
  const long long alen = in->shape[axis];
  if (in->strides[axis] != isz || out->dtype >= CNB_NUM_DTYPES) return 1;
  long long nout = 1;
  for (int d = 0; d < in->ndim; ++d)
    if (d != axis) nout *= in->shape[d];
  const int sms = sm_count();
  if (nout <= 0 || nout * 2 >= sms || alen * isz < (32LL << 20)) return 1;   // >= 32 MiB per row
  // every kept dim must be slower than the axis (true ROW mode), else COLUMN mode splits already
  for (int d = 0; d < in->ndim; ++d)
    if (d != axis && in->shape[d] > 1 && std::llabs(in->strides[d]) < alen * isz) return 1;
  unsigned char ident[16];
  if (!partial_identity(op2, out->dtype, ident)) return 1;
  // Stage 1 views every row as a (Q x W) matrix and reduces it along Q — COLUMN mode: all CTAs sweep
  // the row front to back together, so the pages in flight stay a compact window (splitting a row
  // into S far-apart contiguous segments, one CTA each, thrashes the TLB: measured 3.5-4.0 TB/s) —
  // into W partials per row (+1 for the ragged tail); stage 2 reduces the partials along the row.
  const long long W = std::max<long long>(1024, 65536 / isz);        // 64 KiB wide
  const long long Q = alen / W, tail = alen - Q * W;
  const long long P = W + (tail > 0 ? 1 : 0);                          // partials per output
  if (Q < 8) return 1;

  char* scratch = static_cast<char*>(pool_alloc((size_t)(nout * P * vsz), s));
  if (scratch == nullptr) return CNB_ERR_CUDA;
  int rc = CNB_OK;
  {
    cnb_store_t flat{};
    flat.ptr = scratch; flat.dtype = out->dtype; flat.ndim = 1;
    flat.shape[0] = nout * P; flat.strides[0] = vsz;
    rc = fill_value(&flat, ident, s);
  }
  // scratch strides of the kept dims: row-major over them, P partials per output
  long long kstride[CNB_MAX_DIM];
  {
    long long acc = P * vsz;
    for (int d = in->ndim - 1; d >= 0; --d) {
      if (d == axis) continue;
      kstride[d] = acc;
      acc *= in->shape[d];
    }
  }
  // stage 1, one launch per output row: a dense (Q x W) matrix reduced along Q takes the fast COLUMN
  // kernel (every thread owns its columns outright, 7 TB/s); handing all rows to one launch would
  // add a second kept dim and fall to the general column kernel (measured 3.4 TB/s)
  {
    long long idx[CNB_MAX_DIM] = {0, 0, 0, 0};
    for (long long r = 0; r < nout && rc == CNB_OK; ++r) {
      long long in_off = 0, sc_off = 0;
      for (int d = 0; d < in->ndim; ++d)
        if (d != axis) {
          in_off += idx[d] * in->strides[d];
          sc_off += idx[d] * kstride[d];
        }
      cnb_store_t i2{}, o2{};
      i2.ptr = static_cast<char*>(in->ptr) + in_off; i2.dtype = in->dtype; i2.ndim = 2;
      i2.shape[0] = Q; i2.strides[0] = W * isz;
      i2.shape[1] = W; i2.strides[1] = isz;
      o2.ptr = scratch + sc_off; o2.dtype = out->dtype; o2.ndim = 2;
      o2.shape[0] = Q; o2.strides[0] = 0;
      o2.shape[1] = W; o2.strides[1] = vsz;
      rc = axis_red_dispatch(op, 0, &o2, &i2, nullptr, 0, s);
      for (int d = in->ndim - 1; d >= 0; --d) {   // next output row
        if (d == axis) continue;
        if (++idx[d] < in->shape[d]) break;
        idx[d] = 0;
      }
    }
  }
  if (rc == CNB_OK && tail > 0) {
    // the ragged end of every row: one more partial (ROW mode over `tail` elements)
    cnb_store_t i2 = *in, o2 = *out;
    i2.ptr = static_cast<char*>(in->ptr) + Q * W * isz;
    i2.shape[axis] = tail;
    o2.ptr = scratch + W * vsz; o2.shape[axis] = tail;
    for (int d = 0; d < in->ndim; ++d) o2.strides[d] = d == axis ? 0 : kstride[d];
    rc = axis_red_dispatch(op, axis, &o2, &i2, nullptr, 0, s);
  }
  if (rc == CNB_OK) {
    cnb_store_t i3 = *in, o3 = *out;
    i3.ptr = scratch; i3.dtype = out->dtype;
    for (int d = 0; d < in->ndim; ++d) i3.strides[d] = d == axis ? vsz : kstride[d];
    i3.shape[axis] = P;
    o3.shape[axis] = P;
    rc = axis_red_dispatch(op2, axis, &o3, &i3, nullptr, 0, s);
  }
  pool_free(scratch, s);
  return rc;
}

// ---- pitched views that start off a 16-byte boundary ---------------------------------------------
// `x[1:-1, 1:-1].sum(...)`: the rows keep a 16-byte-multiple pitch, but every row starts (and ends)
// inside a 16-byte word, so the whole task falls off the vector / bulk-copy kernels onto the
// element-wise ones (measured on 16382 x 16382 fp32: axis 0 at 0.34, axis 1 at 0.75, full at 0.49 of
// the HBM peak).  The contiguous dim is cut into head | body | tail with the body starting on a
// 16-byte boundary and a multiple of 16 bytes long; the three pieces are reduced by the ordinary
// kernels one after the other.  All of them fold into the caller's pre-filled output
// (reduce-accessor semantics), so pieces along the reduction axis simply accumulate, and pieces of a
// kept dim write disjoint outputs.  Arg-reductions keep their indices through the axis origin.
// Returns 1 if the layout does not call for it.
struct Peel {
  int dim;                 // the contiguous dim
  long long off[3], len[3];
};

bool plan_peel(const cnb_store_t* in, Peel& pl)
{
  static const bool enabled = [] {
    const char* e = getenv("CNB_RED_PEEL");
    return e == nullptr || atoi(e) != 0;
  }();
  if (!enabled || in->ndim < 1 || in->dtype >= CNB_NUM_DTYPES) return false;
  const long long isz = (long long)dtype_size(in->dtype);
  if (isz >= 16) return false;
  long long total = 1;
  int c = -1;
  for (int d = 0; d < in->ndim; ++d) {
    total *= in->shape[d];
    if (in->shape[d] > 1 && in->strides[d] == isz) c = d;
  }
  if (c < 0 || total < (1LL << 20)) return false;
  const long long n = in->shape[c], V = 16 / isz;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(in->ptr);
  if (addr % isz != 0 || n < 64 * V) return false;
  for (int d = 0; d < in->ndim; ++d)
    if (d != c && in->shape[d] > 1 && in->strides[d] % 16 != 0) return false;
  const long long head = (long long)((16 - addr % 16) % 16) / isz;
  const long long body = ((n - head) / V) * V;
  const long long tail = n - head - body;
  if (head == 0 && tail == 0) return false;     // already aligned
  pl.dim = c;
  pl.off[0] = 0;           pl.len[0] = head;
  pl.off[1] = head;        pl.len[1] = body;
  pl.off[2] = head + body; pl.len[2] = tail;
  return true;
}

int try_peel_axis(int op, int axis, const cnb_store_t* out, const cnb_store_t* in,
                  long long axis_origin, cudaStream_t s)
{
  Peel pl;
  if (axis < 0 || axis >= in->ndim || out->ndim != in->ndim || !plan_peel(in, pl)) return 1;
  const long long isz = (long long)dtype_size(in->dtype);
  for (int k = 0; k < 3; ++k) {
    if (pl.len[k] == 0) continue;
    cnb_store_t i2 = *in, o2 = *out;
    i2.ptr = static_cast<char*>(in->ptr) + pl.off[k] * isz;
    i2.shape[pl.dim] = pl.len[k];
    o2.shape[pl.dim] = pl.len[k];
    long long origin = axis_origin;
    if (pl.dim == axis)
      origin += pl.off[k];                                   // same outputs, later part of the axis
    else
      o2.ptr = static_cast<char*>(out->ptr) + pl.off[k] * out->strides[pl.dim];
    // (a few very long rows: the aligned body is what the row split wants to see)
    int rc = try_split_long_rows(op, axis, &o2, &i2, s);
    if (rc == 1) rc = axis_red_dispatch(op, axis, &o2, &i2, nullptr, origin, s);
    if (rc != CNB_OK) return rc;
  }
  return CNB_OK;
}

int scalar_red_dispatch(int op, const cnb_store_t* out, const cnb_store_t* in,
                        const cnb_store_t* where, const int64_t* origin, const int64_t* gshape,
                        const void* extra, cudaStream_t s)
{
  int rc;
  if ((rc = scalar_red_group1(op, out, in, where, origin, gshape, extra, s)) != 1) return rc;
  if ((rc = scalar_red_group2(op, out, in, where, origin, gshape, extra, s)) != 1) return rc;
  if ((rc = scalar_red_group3(op, out, in, where, origin, gshape, extra, s)) != 1) return rc;
  return set_error(CNB_ERR_BAD_ARG, "unknown reduction opcode %d", op);
}

// ---- full reduction of a pitched view -------------------------------------------------------------
// SCALAR_UNARY_RED tiles the innermost dim; its 128-bit path serves FULL tiles only, so a pitched 2-D
// view (every row ends in a partial tile, or starts off a 16-byte boundary) runs mostly on the
// element-wise path (measured: x[1:-1, 1:-1].sum() of 16384^2 fp32 at 0.49 of the HBM peak; cutting
// the view into head | body | tail for the same kernel made it worse, 0.26).  The axis kernels do not
// have that problem: the view is reduced along its contiguous dim into one partial per row (ROW mode,
// bulk-copy pipeline, with the misaligned ends peeled off), then the partials are reduced into the
// caller's pre-filled output.  Value reductions whose partials can be reduced again, no mask.
int try_rows_then_scalar(int op, const cnb_store_t* out, const cnb_store_t* in, cudaStream_t s)
{
  static const bool enabled = [] {
    const char* e = getenv("CNB_RED_PEEL_SCALAR");
    return e == nullptr || atoi(e) != 0;
  }();
  const int op2 = second_stage_op(op);
  if (!enabled || op2 < 0 || in->ndim < 2 || in->dtype >= CNB_NUM_DTYPES || out->dtype >= CNB_NUM_DTYPES)
    return 1;
  const long long isz = (long long)dtype_size(in->dtype);
  const long long vsz = (long long)dtype_size(out->dtype);
  long long total = 1;
  int c = -1;
  for (int d = 0; d < in->ndim; ++d) {
    total *= in->shape[d];
    if (in->shape[d] > 1 && in->strides[d] < 0) return 1;
    if (in->shape[d] > 1 && in->strides[d] == isz) c = d;
  }
  if (c < 0 || total < (1LL << 20)) return 1;
  const long long n = in->shape[c], nrows = total / n;
  if (n * isz < 4096 || nrows < 16) return 1;
  // a dense view is one long row for SCALAR_UNARY_RED already
  {
    int order[CNB_MAX_DIM], k = 0;
    for (int d = 0; d < in->ndim; ++d)
      if (in->shape[d] > 1) order[k++] = d;
    std::sort(order, order + k, [&](int a, int b) {
      return std::llabs(in->strides[a]) < std::llabs(in->strides[b]);
    });
    long long expect = isz;
    bool dense = true;
    for (int i = 0; i < k && dense; ++i) {
      dense = in->strides[order[i]] == expect;
      expect *= in->shape[order[i]];
    }
    if (dense) return 1;
  }
  unsigned char ident[16];
  if (!partial_identity(op2, out->dtype, ident)) return 1;
  char* scratch = static_cast<char*>(pool_alloc((size_t)(nrows * vsz), s));
  if (scratch == nullptr) return CNB_ERR_CUDA;
  cnb_store_t flat{};
  flat.ptr = scratch; flat.dtype = out->dtype; flat.ndim = 1;
  flat.shape[0] = nrows; flat.strides[0] = vsz;
  int rc = fill_value(&flat, ident, s);
  if (rc == CNB_OK) {
    // one partial per row: the output is promoted (stride 0) along the contiguous dim
    cnb_store_t o2 = *in;
    o2.ptr = scratch; o2.dtype = out->dtype;
    long long acc = vsz;
    for (int d = in->ndim - 1; d >= 0; --d) {
      if (d == c) { o2.strides[d] = 0; continue; }
      o2.strides[d] = acc;
      acc *= in->shape[d];
    }
    rc = try_peel_axis(op, c, &o2, in, 0, s);
    if (rc == 1) rc = axis_red_dispatch(op, c, &o2, in, nullptr, 0, s);
  }
  if (rc == CNB_OK) rc = scalar_red_dispatch(op2, out, &flat, nullptr, nullptr, nullptr, nullptr, s);
  pool_free(scratch, s);
  return rc;
}
}  // namespace
}  // namespace cnb

using namespace cnb;

extern "C" {

int cnb_binary_op(int32_t op, const cnb_store_t* out, const cnb_store_t* in1,
                  const cnb_store_t* in2, const void* extra, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  CNB_CHECK(check_store(in1, "in1"));
  CNB_CHECK(check_store(in2, "in2"));
  if (in1->dtype >= CNB_NUM_DTYPES) return set_error(CNB_ERR_BAD_ARG, "BINARY_OP on a struct dtype");
  set_task_tag(CNB_OP_BINARY_OP, op, in1->dtype);
  auto s = (cudaStream_t)stream;
  int rc;
  if ((rc = binary_group1(op, out, in1, in2, extra, s)) != 1) return rc;
  if ((rc = binary_group2(op, out, in1, in2, extra, s)) != 1) return rc;
  if ((rc = binary_group3(op, out, in1, in2, extra, s)) != 1) return rc;
  if ((rc = binary_group4(op, out, in1, in2, extra, s)) != 1) return rc;
  if ((rc = binary_group5(op, out, in1, in2, extra, s)) != 1) return rc;
  return set_error(CNB_ERR_BAD_ARG, "unknown binary opcode %d", op);
}

int cnb_unary_op(int32_t op, const cnb_store_t* out, const cnb_store_t* out2,
                 const cnb_store_t* in, const void* extra, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  CNB_CHECK(check_store(in, "in"));
  set_task_tag(CNB_OP_UNARY_OP, op, in->dtype);
  auto s = (cudaStream_t)stream;
  if (op == CNB_UOP_FREXP || op == CNB_UOP_MODF) {
    CNB_CHECK(check_store(out2, "out2"));
    if (in->dtype >= CNB_NUM_DTYPES) return set_error(CNB_ERR_BAD_ARG, "FREXP/MODF on a struct dtype");
    return unary_multiout(op, out, out2, in, s);
  }
  if (op == CNB_UOP_GETARG) return unary_getarg(out, in, s);
  if (in->dtype >= CNB_NUM_DTYPES) return set_error(CNB_ERR_BAD_ARG, "UNARY_OP on a struct dtype");
  int rc;
  if ((rc = unary_group1(op, out, in, extra, s)) != 1) return rc;
  if ((rc = unary_group2(op, out, in, extra, s)) != 1) return rc;
  if ((rc = unary_group3(op, out, in, extra, s)) != 1) return rc;
  if ((rc = unary_group4(op, out, in, extra, s)) != 1) return rc;
  if ((rc = unary_group5(op, out, in, extra, s)) != 1) return rc;
  return set_error(CNB_ERR_BAD_ARG, "unknown unary opcode %d", op);
}

int cnb_where(const cnb_store_t* out, const cnb_store_t* mask, const cnb_store_t* in1,
              const cnb_store_t* in2, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  CNB_CHECK(check_store(mask, "mask"));
  CNB_CHECK(check_store(in1, "in1"));
  CNB_CHECK(check_store(in2, "in2"));
  set_task_tag(CNB_OP_WHERE, 0, out->dtype);
  return where_select(out, mask, in1, in2, (cudaStream_t)stream);
}

int cnb_convert(int32_t nan_op, const cnb_store_t* out, const cnb_store_t* in, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  CNB_CHECK(check_store(in, "in"));
  if (in->dtype >= CNB_NUM_DTYPES || out->dtype >= CNB_NUM_DTYPES)
    return set_error(CNB_ERR_BAD_ARG, "CONVERT on a struct dtype");
  set_task_tag(CNB_OP_CONVERT, out->dtype, in->dtype);
  auto s = (cudaStream_t)stream;
  int rc;
  if ((rc = convert_group1(nan_op, out, in, s)) != 1) return rc;
  if ((rc = convert_group2(nan_op, out, in, s)) != 1) return rc;
  if ((rc = convert_group3(nan_op, out, in, s)) != 1) return rc;
  if ((rc = convert_group4(nan_op, out, in, s)) != 1) return rc;
  return set_error(CNB_ERR_BAD_ARG, "unknown source dtype %d", in->dtype);
}

int cnb_scalar_unary_red(int32_t op, const cnb_store_t* out, const cnb_store_t* in,
                         const cnb_store_t* where, const int64_t* origin,
                         const int64_t* global_shape, const void* extra, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  CNB_CHECK(check_store(in, "in"));
  if (where != nullptr) CNB_CHECK(check_store(where, "where"));
  if (in->dtype >= CNB_NUM_DTYPES) return set_error(CNB_ERR_BAD_ARG, "reduction on a struct dtype");
  set_task_tag(CNB_OP_SCALAR_UNARY_RED, op, in->dtype);
  auto s = (cudaStream_t)stream;
  if (where == nullptr) {
    const int rc = try_rows_then_scalar(op, out, in, s);
    if (rc != 1) return rc;
  }
  return scalar_red_dispatch(op, out, in, where, origin, global_shape, extra, s);
}

int cnb_unary_red(int32_t op, int32_t axis, const cnb_store_t* out, const cnb_store_t* in,
                  const cnb_store_t* where, int64_t axis_origin, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  CNB_CHECK(check_store(in, "in"));
  if (where != nullptr) CNB_CHECK(check_store(where, "where"));
  if (in->dtype >= CNB_NUM_DTYPES) return set_error(CNB_ERR_BAD_ARG, "reduction on a struct dtype");
  set_task_tag(CNB_OP_UNARY_RED, op, in->dtype);
  auto s = (cudaStream_t)stream;
  if (where == nullptr) {
    int rc = try_peel_axis(op, axis, out, in, axis_origin, s);
    if (rc != 1) return rc;
    rc = try_split_long_rows(op, axis, out, in, s);
    if (rc != 1) return rc;
  }
  return axis_red_dispatch(op, axis, out, in, where, axis_origin, s);
}

int cnb_binary_red(int32_t op, const cnb_store_t* out, const cnb_store_t* in1,
                   const cnb_store_t* in2, const void* extra, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  CNB_CHECK(check_store(in1, "in1"));
  CNB_CHECK(check_store(in2, "in2"));
  if (in1->dtype >= CNB_NUM_DTYPES) return set_error(CNB_ERR_BAD_ARG, "BINARY_RED on a struct dtype");
  set_task_tag(CNB_OP_BINARY_RED, op, in1->dtype);
  return binary_red(op, out, in1, in2, extra, (cudaStream_t)stream);
}

int cnb_fill(const cnb_store_t* out, const void* value, void* stream)
{
  CNB_CHECK(ensure_init());
  CNB_CHECK(check_store(out, "out"));
  set_task_tag(CNB_OP_FILL, 0, out->dtype);
  return fill_value(out, value, (cudaStream_t)stream);
}

// ---- symbols the reference's cffi layer binds (cunumeric_c.h:337-339) ---------------------------
void cunumeric_perform_registration(void) { ensure_init(); }
int cunumeric_has_curand(void) { return 0; }
void cunumeric_register_reduction_op(int32_t type_uid, int32_t elem_type_code)
{
  std::lock_guard<std::mutex> g(g_redop_mu);
  g_argval_types[type_uid] = elem_type_code;
}
int32_t cnb_registered_argval_elem(int32_t type_uid)
{
  std::lock_guard<std::mutex> g(g_redop_mu);
  auto it = g_argval_types.find(type_uid);
  return it == g_argval_types.end() ? -1 : it->second;
}
}
