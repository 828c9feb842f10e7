/* cunumeric_b200 — C ABI of the B200-native hot path (sm_100a) for the cuNumeric NumPy API.
 *
 * This is the drop-in boundary.  In the reference the boundary is a Legate task variant
 *     static void XTask::gpu_variant(legate::TaskContext&)
 * whose arguments arrive positionally in context.inputs()/outputs()/reductions()/scalars()
 * (reference: src/cunumeric/binary/binary_op.h:32-44, cunumeric/deferred.py:3139-3384).  legate.core
 * is not part of the reference tree, so the boundary is moved one level up: every entry point below
 * takes exactly the per-opcode task contract — the same stores, in the same order, with the same
 * scalar arguments — as plain C structs, pointers and sizes.  INTEGRATION.md shows the binding a
 * maintainer adds to cunumeric/deferred.py (ctypes) or to a Legate task body (C++).
 *
 * Conventions
 *   - every compute entry point returns 0 on success, <0 on error (cnb_last_error() gives the
 *     thread-local message).  Invalid (op, dtype) pairs — the ones the reference `assert(false)`s
 *     on (binary_op_template.inl:65-69) — return CNB_ERR_INVALID_OP.
 *   - all launches are asynchronous on the caller's stream (`void* stream` is a cudaStream_t; NULL is
 *     the legacy default stream), like the reference's get_cached_stream() contract
 *     (src/cunumeric/cudalibs.cu:335-338).  Nothing synchronises inside a compute call.
 *   - stores are borrowed for the duration of the call; the library never frees or retains them.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with CNB_ERR_CUDA.
 *   - opcode integers are numerically identical to src/cunumeric/cunumeric_c.h:27-240.
 */
#ifndef CUNUMERIC_B200_H
#define CUNUMERIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CNB_MAX_DIM 4 /* LEGATE_MAX_DIM default (reference README.md:116-117) */

/* legate::Type::Code order (Legion type ids): ? b h i l B H I L e f d F D */
typedef enum cnb_dtype {
  CNB_BOOL = 0,
  CNB_INT8,
  CNB_INT16,
  CNB_INT32,
  CNB_INT64,
  CNB_UINT8,
  CNB_UINT16,
  CNB_UINT32,
  CNB_UINT64,
  CNB_FLOAT16,
  CNB_FLOAT32,
  CNB_FLOAT64,
  CNB_COMPLEX64,
  CNB_COMPLEX128,
  CNB_NUM_DTYPES,
  /* Argval<T> = {int64 arg; T arg_value}, 16 bytes (reference src/cunumeric/arg.h:23-59);
     code = CNB_ARGVAL_BASE + code(T).  The struct dtype runtime.get_argred_type builds
     (cunumeric/runtime.py:125-134). */
  CNB_ARGVAL_BASE = 32
} cnb_dtype_t;

/* src/cunumeric/cunumeric_c.h:27-82 (only the opcodes on the hot path + FILL/BINARY_RED) */
enum {
  CNB_OP_BINARY_OP        = 5,
  CNB_OP_BINARY_RED       = 6,
  CNB_OP_CONVERT          = 11,
  CNB_OP_FILL             = 19,
  CNB_OP_SCALAR_UNARY_RED = 33,
  CNB_OP_UNARY_OP         = 43,
  CNB_OP_UNARY_RED        = 44,
  CNB_OP_WHERE            = 49
};

/* src/cunumeric/cunumeric_c.h:86-134 */
typedef enum cnb_unary_op {
  CNB_UOP_ABSOLUTE = 1, CNB_UOP_ARCCOS, CNB_UOP_ARCCOSH, CNB_UOP_ARCSIN, CNB_UOP_ARCSINH,
  CNB_UOP_ARCTAN, CNB_UOP_ARCTANH, CNB_UOP_CBRT, CNB_UOP_CEIL, CNB_UOP_CLIP, CNB_UOP_CONJ,
  CNB_UOP_COPY, CNB_UOP_COS, CNB_UOP_COSH, CNB_UOP_DEG2RAD, CNB_UOP_EXP, CNB_UOP_EXP2,
  CNB_UOP_EXPM1, CNB_UOP_FLOOR, CNB_UOP_FREXP, CNB_UOP_GETARG, CNB_UOP_IMAG, CNB_UOP_INVERT,
  CNB_UOP_ISFINITE, CNB_UOP_ISINF, CNB_UOP_ISNAN, CNB_UOP_LOG, CNB_UOP_LOG10, CNB_UOP_LOG1P,
  CNB_UOP_LOG2, CNB_UOP_LOGICAL_NOT, CNB_UOP_MODF, CNB_UOP_NEGATIVE, CNB_UOP_POSITIVE,
  CNB_UOP_RAD2DEG, CNB_UOP_REAL, CNB_UOP_RECIPROCAL, CNB_UOP_RINT, CNB_UOP_SIGN, CNB_UOP_SIGNBIT,
  CNB_UOP_SIN, CNB_UOP_SINH, CNB_UOP_SQRT, CNB_UOP_SQUARE, CNB_UOP_TAN, CNB_UOP_TANH,
  CNB_UOP_TRUNC
} cnb_unary_op_t;

/* src/cunumeric/cunumeric_c.h:138-157 */
typedef enum cnb_red_op {
  CNB_RED_ALL = 1, CNB_RED_ANY, CNB_RED_ARGMAX, CNB_RED_ARGMIN, CNB_RED_CONTAINS,
  CNB_RED_COUNT_NONZERO, CNB_RED_MAX, CNB_RED_MIN, CNB_RED_NANARGMAX, CNB_RED_NANARGMIN,
  CNB_RED_NANMAX, CNB_RED_NANMIN, CNB_RED_NANPROD, CNB_RED_NANSUM, CNB_RED_PROD, CNB_RED_SUM,
  CNB_RED_SUM_SQUARES, CNB_RED_VARIANCE
} cnb_red_op_t;

/* src/cunumeric/cunumeric_c.h:161-197 */
typedef enum cnb_binary_op {
  CNB_BINOP_ADD = 1, CNB_BINOP_ARCTAN2, CNB_BINOP_BITWISE_AND, CNB_BINOP_BITWISE_OR,
  CNB_BINOP_BITWISE_XOR, CNB_BINOP_COPYSIGN, CNB_BINOP_DIVIDE, CNB_BINOP_EQUAL,
  CNB_BINOP_FLOAT_POWER, CNB_BINOP_FLOOR_DIVIDE, CNB_BINOP_FMOD, CNB_BINOP_GCD, CNB_BINOP_GREATER,
  CNB_BINOP_GREATER_EQUAL, CNB_BINOP_HYPOT, CNB_BINOP_ISCLOSE, CNB_BINOP_LCM, CNB_BINOP_LDEXP,
  CNB_BINOP_LEFT_SHIFT, CNB_BINOP_LESS, CNB_BINOP_LESS_EQUAL, CNB_BINOP_LOGADDEXP,
  CNB_BINOP_LOGADDEXP2, CNB_BINOP_LOGICAL_AND, CNB_BINOP_LOGICAL_OR, CNB_BINOP_LOGICAL_XOR,
  CNB_BINOP_MAXIMUM, CNB_BINOP_MINIMUM, CNB_BINOP_MOD, CNB_BINOP_MULTIPLY, CNB_BINOP_NEXTAFTER,
  CNB_BINOP_NOT_EQUAL, CNB_BINOP_POWER, CNB_BINOP_RIGHT_SHIFT, CNB_BINOP_SUBTRACT
} cnb_binary_op_t;

/* src/cunumeric/cunumeric_c.h:236-240 */
typedef enum cnb_convert_op {
  CNB_CONVERT_NAN_NOOP = 1, CNB_CONVERT_NAN_PROD, CNB_CONVERT_NAN_SUM
} cnb_convert_op_t;

/* src/cunumeric/cunumeric_c.h:210-213 */
enum { CNB_ARGMAX_REDOP = 1, CNB_ARGMIN_REDOP = 2 };

enum {
  CNB_OK              = 0,
  CNB_ERR_INVALID_OP  = -1, /* (op, dtype) the reference marks `valid = false` */
  CNB_ERR_BAD_ARG     = -2, /* shape / dtype / ndim mismatch between the stores of one task */
  CNB_ERR_CUDA        = -3, /* CUDA runtime error, or no device */
  CNB_ERR_UNSUPPORTED = -4,
  CNB_ERR_COMM        = -5
};

/* A borrowed view of device memory: what a legate::Store accessor over the task's rect exposes
 * (ptr = acc.ptr(rect.lo); per-dimension BYTE strides, 0 for a promoted/broadcast dimension;
 * reference contract: src/cunumeric/mapper.cc:243-245 — no layout constraint, so any affine layout
 * is legal and density is tested by the kernel launcher, binary_op_template.inl:52-59). */
typedef struct cnb_store {
  void* ptr;
  int32_t dtype; /* cnb_dtype_t, or CNB_ARGVAL_BASE + elem code */
  int32_t ndim;  /* 0..CNB_MAX_DIM; 0-d stores are treated as shape (1,) */
  int64_t shape[CNB_MAX_DIM];
  int64_t strides[CNB_MAX_DIM]; /* bytes */
} cnb_store_t;

/* ---------------------------------------------------------------------------------------------
 * Task entry points (one per opcode on the hot path)
 * ------------------------------------------------------------------------------------------- */

/* BINARY_OP — replaces BinaryOpTask::gpu_variant (src/cunumeric/binary/binary_op.cu:85-88;
 * contract deferred.py:3318-3328 <-> binary_op_template.inl:85-94).
 * out dtype must be the functor's result type (bool for compares/logicals/ISCLOSE, float64 for
 * integer DIVIDE, complex128 for complex64 FLOAT_POWER, else in1's dtype); in2 is int32 for LDEXP.
 * extra: ISCLOSE -> host double[2] = {rtol, atol} (the two extra scalar stores); else NULL. */
int cnb_binary_op(int32_t op, const cnb_store_t* out, const cnb_store_t* in1,
                  const cnb_store_t* in2, const void* extra, void* stream);

/* UNARY_OP — replaces UnaryOpTask::gpu_variant (src/cunumeric/unary/unary_op.cu:178-181;
 * contract deferred.py:3152-3165 <-> unary_op_template.inl:177-209).
 * out2 is the second output of FREXP (int32) / MODF (same dtype), NULL otherwise.
 * extra: CLIP -> host {min, max} in the array dtype; else NULL.
 * GETARG: `in` is an Argval store (CNB_ARGVAL_BASE + T), out is int64. */
int cnb_unary_op(int32_t op, const cnb_store_t* out, const cnb_store_t* out2,
                 const cnb_store_t* in, const void* extra, void* stream);

/* WHERE — replaces WhereTask::gpu_variant (src/cunumeric/ternary/where.cu:74-77;
 * contract deferred.py:3374-3384 <-> where_template.inl:65-68). mask is CNB_BOOL. */
int cnb_where(const cnb_store_t* out, const cnb_store_t* mask, const cnb_store_t* in1,
              const cnb_store_t* in2, void* stream);

/* CONVERT — replaces ConvertTask::gpu_variant (src/cunumeric/unary/convert.cu:71-74;
 * contract deferred.py:1371-1378 <-> convert_template.inl:103-105). in.dtype != out.dtype. */
int cnb_convert(int32_t nan_op, const cnb_store_t* out, const cnb_store_t* in, void* stream);

/* SCALAR_UNARY_RED — replaces ScalarUnaryRedTask::gpu_variant
 * (src/cunumeric/unary/scalar_unary_red.cu:26-29; contract deferred.py:3207-3235 <->
 * scalar_unary_red_template.inl:166-190).
 * out: 1-element device store of the reduction VAL type (bool for ALL/ANY/CONTAINS, uint64 for
 *      COUNT_NONZERO, Argval<T> for ARG*, else T), pre-filled by the caller with the identity or
 *      `initial`; the result is FOLDED into it (reduction-accessor semantics, out.reduce(0, v)).
 * where: optional CNB_BOOL store aligned with `in` (NULL = has_where false).
 * origin / global_shape: rect.lo of `in` in the global array and the global shape (scalars[1]);
 *      arg-reductions return the GLOBAL row-major flat index (unary_red_util.h:342-351).
 *      NULL = origin 0 / shape of `in`.
 * extra: CONTAINS -> host value to find; VARIANCE -> host mean; both in the array dtype. */
int cnb_scalar_unary_red(int32_t op, const cnb_store_t* out, const cnb_store_t* in,
                         const cnb_store_t* where, const int64_t* origin,
                         const int64_t* global_shape, const void* extra, void* stream);

/* UNARY_RED — replaces UnaryRedTask::gpu_variant (src/cunumeric/unary/unary_red.cu:342-345;
 * contract deferred.py:3265-3280 <-> unary_red_template.inl:81-95).
 * out: the reduction target PROMOTED to in's shape (ndim == in.ndim, strides[axis] ignored),
 *      pre-filled with the identity; results are folded into it.
 * axis_origin: global coordinate of in's first element along `axis` (arg-reductions store
 *      point[collapsed_dim], unary_red_util.h:334-340). */
int cnb_unary_red(int32_t op, int32_t axis, const cnb_store_t* out, const cnb_store_t* in,
                  const cnb_store_t* where, int64_t axis_origin, void* stream);

/* BINARY_RED — replaces BinaryRedTask::gpu_variant (src/cunumeric/binary/binary_red.cu:105-108;
 * contract deferred.py:3330-3364 <-> binary_red_template.inl:31-75): array_equal / allclose.
 * op is CNB_BINOP_EQUAL or CNB_BINOP_ISCLOSE (extra = host double[2] {rtol, atol}); out is a
 * 1-element CNB_BOOL store pre-filled with true, and is set to false if any pair fails. */
int cnb_binary_red(int32_t op, const cnb_store_t* out, const cnb_store_t* in1,
                   const cnb_store_t* in2, const void* extra, void* stream);

/* FILL — replaces FillTask::gpu_variant (src/cunumeric/nullary/fill.cu; deferred.py:1463-1496).
 * value: host pointer to one element of out's dtype (Argval fill: 16 bytes). */
int cnb_fill(const cnb_store_t* out, const void* value, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Symbols the reference's cffi layer binds (src/cunumeric/cunumeric_c.h:337-339)
 * ------------------------------------------------------------------------------------------- */
void cunumeric_perform_registration(void); /* cunumeric.cc:49-52: here, initialises the device */
int cunumeric_has_curand(void);            /* cunumeric.cc:54-61: 0 — RNG is out of scope */
void cunumeric_register_reduction_op(int32_t type_uid, int32_t elem_type_code);
                                           /* arg_redop_register.cc:60-70 */
/* Redop id registered for (type_uid); 0 if none.  CNB_ARGMAX_REDOP/ARGMIN_REDOP are folded by
 * cnb_comm_allreduce_argval on multi-GPU runs. */
int32_t cnb_registered_argval_elem(int32_t type_uid);

/* ---------------------------------------------------------------------------------------------
 * Device runtime (what legate.core's allocator / StreamPool / Realm copies provide to the
 * reference; src/cunumeric/cudalibs.cu:335-338, device_scalar_reduction_buffer.h:30-36)
 * ------------------------------------------------------------------------------------------- */
int cnb_device_count(void);
int cnb_init(int32_t device);       /* selects the device, creates the memory pool + scratch */
int cnb_sm_count(void);
void* cnb_malloc(size_t nbytes, void* stream); /* stream-ordered pool allocation; NULL on error */
int cnb_free(void* ptr, void* stream);
void* cnb_host_alloc(size_t nbytes);            /* pinned host memory */
int cnb_host_free(void* ptr);
int cnb_memcpy_h2d(void* dst, const void* src, size_t nbytes, void* stream);
int cnb_memcpy_d2h(void* dst, const void* src, size_t nbytes, void* stream);
int cnb_memcpy_d2d(void* dst, const void* src, size_t nbytes, void* stream);
int cnb_memset(void* dst, int value, size_t nbytes, void* stream);
void* cnb_stream_create(void);
int cnb_stream_destroy(void* stream);
int cnb_stream_synchronize(void* stream);
int cnb_device_synchronize(void);
void* cnb_event_create(void);
int cnb_event_destroy(void* event);
int cnb_event_record(void* event, void* stream);
int cnb_event_synchronize(void* event);
int cnb_stream_wait_event(void* stream, void* event);
int cnb_event_elapsed_ms(void* start, void* stop, float* ms);
int cnb_mem_info(size_t* free_bytes, size_t* total_bytes);
uint64_t cnb_launch_count(void);  /* kernels launched by this library since load (diagnostic) */

/* Per-launch timing, measured live with CUDA events on the launching stream (the hook the
 * reference wraps around every task with `Annotation({"OpCode": ...})`, deferred.py:3151). */
typedef struct cnb_trace_record {
  int32_t task;        /* CNB_OP_* of the task that launched the kernel */
  int32_t op;          /* its opcode argument (CONVERT: destination dtype) */
  int32_t dtype;       /* input dtype */
  int32_t kernel_kind; /* 1 elementwise, 2 scalar reduction, 3 axis/column, 4 axis/row */
  int64_t elems;       /* elements in the iteration space */
  int64_t bytes;       /* algorithmic bytes (distinct elements touched x itemsize) */
  float ms;            /* device time of the launch */
} cnb_trace_record_t;
int cnb_trace_start(int32_t capacity);
int cnb_trace_stop(void); /* synchronises; returns the number of records captured, <0 on error */
int cnb_trace_get(int32_t index, cnb_trace_record_t* out);
const char* cnb_last_error(void);
const char* cnb_version(void);

/* ---------------------------------------------------------------------------------------------
 * Fused elementwise chains (hooks in the reference would sit at
 * deferred.py:3139,3302 where UNARY_OP / BINARY_OP tasks are built).  The host layer captures a
 * chain of elementwise tasks over one iteration space, composes the SAME device functors the
 * per-task kernels use into one generated kernel (cunumeric_b200/fusion.py), compiles it for
 * sm_100a and runs it through these three calls.  `plan` is the generated kernel's single
 * by-value parameter (layout defined by the generator); `tag` is the CNB_OP_* recorded in the
 * launch trace (CNB_OP_FUSED). */
#define CNB_OP_FUSED 1000
int cnb_module_load(const void* image, size_t bytes, void** module);
int cnb_module_get_kernel(void* module, const char* name, void** kernel);
/* max_ctas_per_sm: resident CTAs per SM the persistent grid may use (0 = the memory-bound default, 3).
 * n_reductions > 0: the chain ends in that many SCALAR_UNARY_RED tasks (map -> reduce fusion,
 * scalar_unary_red_template.inl:84-118 folded into the producing kernel); the last 16 bytes of
 * `plan` are then overwritten with the reduction scratch {partials, ticket}. */
int cnb_launch_fused(void* kernel, const void* plan, size_t plan_bytes, int64_t num_tiles,
                     int64_t elements, int64_t algorithmic_bytes, int32_t ntasks,
                     int32_t max_ctas_per_sm, int32_t n_reductions, void* stream);
/* TMA-staged flavour of a fused chain (north_star: "TMA-staged tiles for strided and transposed"
 * operands; replaces the per-element div/mod walk of the reference's generic kernels,
 * binary/binary_op.cu:33-52, pitches.h:46-55, for pitched 2-D views).  Every pitched operand buffer
 * is described by one tensor map {base, width x height elements, row pitch, box}; the library
 * encodes them (cuTensorMapEncodeTiled) in front of the generator's parameter tail.  The kernel
 * stages (rows + halo) x (cols + halo) boxes in shared memory with cp.async.bulk.tensor.2d. */
typedef struct cnb_tma_operand {
  void* base;            /* 16-byte aligned start of the buffer */
  int64_t width;         /* elements per buffer row (row pitch / element size) */
  int64_t height;        /* whole rows inside the allocation */
  int64_t pitch_bytes;   /* multiple of 16 */
  int32_t elem_bytes;    /* 1, 2, 4 or 8 */
  int32_t box_width;     /* <= 256 elements, box_width * elem_bytes multiple of 16 */
  int32_t box_height;    /* <= 256 */
  int32_t reserved;
} cnb_tma_operand_t;
int cnb_launch_fused_tma(void* kernel, const cnb_tma_operand_t* ops, int32_t nmaps, const void* tail,
                         size_t tail_bytes, int32_t smem_bytes, int64_t num_tiles, int64_t elements,
                         int64_t algorithmic_bytes, int32_t ntasks, int32_t max_ctas_per_sm,
                         void* stream);
/* dst <- src over [0, nbytes) EXCEPT the pitched window {offset + r * pitch .. + row_bytes, r < rows}.
 * Used when a fused chain redirects a write to a fresh copy of its buffer to remove a
 * write-after-read hazard against its own shifted operands (register renaming at buffer
 * granularity; the stencil's `center[:] = work`, examples/stencil.py:49). */
int cnb_copy_complement(void* dst, const void* src, size_t nbytes, int64_t offset, int64_t rows,
                        int64_t row_bytes, int64_t pitch, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU exchange (one process per GPU).  In the reference these steps are implicit Legion
 * copies / future-map folds; here they are explicit, stream-ordered NCCL calls.
 * ------------------------------------------------------------------------------------------- */
#define CNB_COMM_ID_BYTES 128
/* Fold of per-rank arg-reduction partials gathered in rank order (world x n Argval<T>, 16 bytes each)
 * into n Argvals, with the tie-breaking of ArgmaxReduction / ArgminReduction (arg.inl:43-50: the
 * incumbent survives ties, i.e. the lowest global index).  In the reference Legion folds the partial
 * futures / reduction instances with the registered redop (arg_redop_register.cc:21-46). */
int cnb_argval_fold(int32_t op, int32_t elem_dtype, void* out, const void* gathered, int32_t world,
                    int64_t n, void* stream);
int cnb_comm_unique_id(void* id_out /* CNB_COMM_ID_BYTES */);
void* cnb_comm_init(const void* id, int32_t nranks, int32_t rank);
int cnb_comm_destroy(void* comm);
int cnb_comm_group_start(void);
int cnb_comm_group_end(void);
int cnb_comm_send(void* comm, const void* buf, size_t nbytes, int32_t peer, void* stream);
int cnb_comm_recv(void* comm, void* buf, size_t nbytes, int32_t peer, void* stream);
/* red_op: CNB_RED_SUM / PROD / MAX / MIN (ALL -> MIN, ANY -> MAX on bool bytes) */
int cnb_comm_allreduce(void* comm, const void* send, void* recv, size_t count, int32_t dtype,
                       int32_t red_op, void* stream);
int cnb_comm_allgather(void* comm, const void* send, void* recv, size_t nbytes_per_rank,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CUNUMERIC_B200_H */
