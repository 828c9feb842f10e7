// C ABI task entry points (include/cunumeric_b200.h): argument validation + dispatch to the
// per-opcode-group translation units.
#include "cnb_common.cuh"

#include <map>
#include <mutex>

namespace cnb {
int ensure_init();

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
  int rc;
  if ((rc = scalar_red_group1(op, out, in, where, origin, global_shape, extra, s)) != 1) return rc;
  if ((rc = scalar_red_group2(op, out, in, where, origin, global_shape, extra, s)) != 1) return rc;
  if ((rc = scalar_red_group3(op, out, in, where, origin, global_shape, extra, s)) != 1) return rc;
  return set_error(CNB_ERR_BAD_ARG, "unknown reduction opcode %d", op);
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
