// TEST INFRASTRUCTURE ONLY — SCALAR_UNARY_RED / UNARY_RED legs of oracle/_ref/libcunumeric_ref.so.
// Arithmetic: the reference's UnaryRedOp<OP,CODE> (convert + fold), Argval and Argmax/Argmin
//   reductions, included by path (/root/reference/src/cunumeric/unary/unary_red_util.h:104-601,
//   arg.h:23-96, arg.inl:24-113).
// Loop shapes restated here (the task bodies need legate accessors, which are external):
//   scalar: execution_policy/reduction/scalar_reduction.h:42-47 driving the kernel body
//           unary/scalar_unary_red_template.inl:84-118; OpenMP: scalar_reduction_omp.h:29-41
//   axis:   unary/unary_red.cc:38-46 (one lhs.reduce(point, convert(...)) per input element)
#include "ref_common.h"
#include "cunumeric/unary/unary_red_util.h"

namespace cunumeric {
// Argmax/Argmin identities: Argval(LLONG_MIN, Max/MinReduction<T>::identity)
// (defined by the reference in arg_redop_register.cc:21-30, which needs the legate runtime).
#define REF_ARG_IDENT(T)                                                                       \
  template <>                                                                                  \
  const Argval<T> ArgmaxReduction<T>::identity = Argval<T>(legate::MaxReduction<T>::identity); \
  template <>                                                                                  \
  const Argval<T> ArgminReduction<T>::identity = Argval<T>(legate::MinReduction<T>::identity);
REF_ARG_IDENT(__half)
REF_ARG_IDENT(float)
REF_ARG_IDENT(double)
REF_ARG_IDENT(bool)
REF_ARG_IDENT(int8_t)
REF_ARG_IDENT(int16_t)
REF_ARG_IDENT(int32_t)
REF_ARG_IDENT(int64_t)
REF_ARG_IDENT(uint8_t)
REF_ARG_IDENT(uint16_t)
REF_ARG_IDENT(uint32_t)
REF_ARG_IDENT(uint64_t)
#undef REF_ARG_IDENT
}  // namespace cunumeric

using namespace cunumeric;
using ref::Code;

namespace {

constexpr int MAXD = LEGATE_MAX_DIM;

template <UnaryRedCode OP>
constexpr bool is_arg_op = OP == UnaryRedCode::ARGMAX || OP == UnaryRedCode::ARGMIN ||
                           OP == UnaryRedCode::NANARGMAX || OP == UnaryRedCode::NANARGMIN;

struct Geometry {
  int ndim;
  int64_t extent[MAXD];  // local rect extents
  int64_t origin[MAXD];  // rect.lo in the global index space
  int64_t shape[MAXD];   // global array shape (scalars[1] of SCALAR_UNARY_RED)
};

// pitches.unflatten(idx, origin) (pitches.h:46-55), row-major
inline void unflatten(const Geometry& g, size_t idx, legate::Point<MAXD>& p)
{
  for (int d = g.ndim - 1; d >= 0; --d) {
    p[d] = g.origin[d] + static_cast<int64_t>(idx % g.extent[d]);
    idx /= g.extent[d];
  }
}

template <UnaryRedCode OP, Code CODE>
int run_scalar(const void* inv, const bool* where, size_t volume, const Geometry& g, void* outv,
               const void* extra, int nthreads)
{
  constexpr bool contains = OP == UnaryRedCode::CONTAINS;
  if constexpr (!(UnaryRedOp<OP, CODE>::valid || contains)) {
    return ref::ERR_INVALID;
  } else {
    using RED   = UnaryRedOp<OP, CODE>;
    using LG_OP = typename RED::OP;
    using LHS   = typename RED::VAL;
    using RHS   = legate::legate_type_of<CODE>;
    auto in     = static_cast<const RHS*>(inv);
    RHS scalar_arg{};  // to_find (CONTAINS) or mu (VARIANCE)
    if constexpr (contains || OP == UnaryRedCode::VARIANCE) std::memcpy(&scalar_arg, extra, sizeof(RHS));
    const LHS identity = LG_OP::identity;

    // scalar_unary_red_template.inl:84-118 (dense and sparse bodies compute the same thing)
    auto body = [&](LHS& lhs, size_t idx) {
      bool mask = true;
      if (where != nullptr) mask = where[idx];
      if constexpr (contains) {
        if (mask && (in[idx] == scalar_arg)) lhs = true;
      } else if constexpr (is_arg_op<OP>) {
        legate::Point<MAXD> p{}, shape{};
        unflatten(g, idx, p);
        // RED::convert(point, shape, identity, rhs) flattens point in the GLOBAL shape
        // (unary_red_util.h:342-351); only the first ndim entries participate.
        int64_t flat = 0;
        for (int d = 0; d < g.ndim; ++d) flat = flat * g.shape[d] + p[d];
        legate::Point<1> p1{}, s1{};
        p1[0] = flat;
        s1[0] = 1;
        if (mask) RED::template fold<true>(lhs, RED::convert(p1, s1, identity, in[idx]));
      } else if constexpr (OP == UnaryRedCode::VARIANCE) {
        if (mask) RED::template fold<true>(lhs, RED::convert(in[idx] - scalar_arg, identity));
      } else {
        if (mask) RED::template fold<true>(lhs, RED::convert(in[idx], identity));
      }
    };

    auto out = static_cast<LHS*>(outv);
    if (nthreads > 1) {
      // scalar_reduction_omp.h:29-41
      std::vector<LHS> locals(nthreads, identity);
#pragma omp parallel num_threads(nthreads)
      {
        const int tid = omp_get_thread_num();
        LHS local     = identity;
#pragma omp for schedule(static)
        for (size_t idx = 0; idx < volume; ++idx) body(local, idx);
        locals[tid] = local;
      }
      for (int t = 0; t < nthreads; ++t) LG_OP::template fold<true>(*out, locals[t]);
    } else {
      // scalar_reduction.h:44-46
      LHS result = identity;
      for (size_t idx = 0; idx < volume; ++idx) body(result, idx);
      LG_OP::template fold<true>(*out, result);  // out.reduce(0, result)
    }
    return 0;
  }
}

template <UnaryRedCode OP, Code CODE>
int run_axis(const void* inv, const bool* where, const Geometry& g, int axis, void* outv,
             int nthreads)
{
  if constexpr (!UnaryRedOp<OP, CODE>::valid) {
    return ref::ERR_INVALID;
  } else {
    using RED   = UnaryRedOp<OP, CODE>;
    using LG_OP = typename RED::OP;
    using LHS   = typename RED::VAL;
    using RHS   = legate::legate_type_of<CODE>;
    auto in     = static_cast<const RHS*>(inv);
    auto out    = static_cast<LHS*>(outv);
    // canonical (outer, axis, inner) view of the row-major rect; out is (outer, inner)
    int64_t outer = 1, inner = 1;
    for (int d = 0; d < axis; ++d) outer *= g.extent[d];
    for (int d = axis + 1; d < g.ndim; ++d) inner *= g.extent[d];
    const int64_t alen = g.extent[axis];
    // unary_red.cc:38-46: for every point (row-major order), lhs.reduce(point, convert(point,
    // collapsed_dim, identity, rhs[point])).  Per output element this is a sequential fold over
    // the collapsed coordinate in increasing order, which is what each (o, i) lane does below.
    const LHS identity = LG_OP::identity;
#pragma omp parallel for schedule(static) collapse(2) num_threads(nthreads) if (nthreads > 1)
    for (int64_t o = 0; o < outer; ++o) {
      for (int64_t i = 0; i < inner; ++i) {
        LHS& lhs = out[o * inner + i];
        for (int64_t a = 0; a < alen; ++a) {
          const size_t idx = static_cast<size_t>((o * alen + a) * inner + i);
          bool mask        = true;
          if (where != nullptr) mask = where[idx];
          if (mask) {
            legate::Point<1> p{};
            p[0] = g.origin[axis] + a;  // point[collapsed_dim]
            LG_OP::template fold<true>(lhs, RED::convert(p, 0, identity, in[idx]));
          }
        }
      }
    }
    return 0;
  }
}

template <typename F>
int op_dispatch_red(int op, F&& f)
{
#define CASE(NAME) \
  case UnaryRedCode::NAME: return f(std::integral_constant<UnaryRedCode, UnaryRedCode::NAME>{});
  switch (static_cast<UnaryRedCode>(op)) {
    CASE(ALL) CASE(ANY) CASE(ARGMAX) CASE(ARGMIN) CASE(CONTAINS) CASE(COUNT_NONZERO) CASE(MAX)
    CASE(MIN) CASE(NANARGMAX) CASE(NANARGMIN) CASE(NANMAX) CASE(NANMIN) CASE(NANPROD)
    CASE(NANSUM) CASE(PROD) CASE(SUM) CASE(SUM_SQUARES) CASE(VARIANCE)
  }
#undef CASE
  return ref::ERR_BADCODE;
}

Geometry make_geometry(int ndim, const int64_t* extent, const int64_t* origin, const int64_t* shape)
{
  Geometry g{};
  g.ndim = ndim;
  for (int d = 0; d < ndim; ++d) {
    g.extent[d] = extent[d];
    g.origin[d] = origin ? origin[d] : 0;
    g.shape[d]  = shape ? shape[d] : extent[d];
  }
  return g;
}

}  // namespace

extern "C" {

// Whole-rect reduction folded into *out (which the caller pre-fills with identity / `initial`,
// as deferred.py:3207-3213 does). `in`/`where` are dense row-major over `extent`.
// out points at one VAL: bool / uint64 / T / Argval<T> (16 bytes) depending on op.
int ref_scalar_unary_red(int op, int code, const void* in, const void* where, int ndim,
                         const int64_t* extent, const int64_t* origin, const int64_t* shape,
                         void* out, const void* extra, int nthreads)
{
  Geometry g    = make_geometry(ndim, extent, origin, shape);
  size_t volume = 1;
  for (int d = 0; d < ndim; ++d) volume *= static_cast<size_t>(extent[d]);
  return op_dispatch_red(op, [&](auto optag) {
    return ref::type_dispatch(code, [&](auto tag) {
      return run_scalar<decltype(optag)::value, decltype(tag)::value>(
        in, static_cast<const bool*>(where), volume, g, out, extra, nthreads);
    });
  });
}

// Reduction along `axis`; out has the rect's extents with extent[axis] removed (row-major) and is
// pre-filled by the caller.
int ref_unary_red(int op, int code, const void* in, const void* where, int ndim,
                  const int64_t* extent, const int64_t* origin, int axis, void* out, int nthreads)
{
  Geometry g = make_geometry(ndim, extent, origin, nullptr);
  return op_dispatch_red(op, [&](auto optag) {
    constexpr UnaryRedCode OP = decltype(optag)::value;
    if constexpr (OP == UnaryRedCode::CONTAINS) {
      return ref::ERR_INVALID;  // scalar path only (unary_red_util.h:592-601)
    } else {
      return ref::type_dispatch(code, [&](auto tag) {
        return run_axis<OP, decltype(tag)::value>(in, static_cast<const bool*>(where), g, axis, out,
                                                  nthreads);
      });
    }
  });
}

// Writes LG_OP::identity of (op, code) into *out; returns sizeof(VAL) or <0.
int ref_red_identity(int op, int code, void* out)
{
  return op_dispatch_red(op, [&](auto optag) {
    constexpr UnaryRedCode OP = decltype(optag)::value;
    return ref::type_dispatch(code, [&](auto tag) {
      constexpr Code C = decltype(tag)::value;
      if constexpr (!(UnaryRedOp<OP, C>::valid || OP == UnaryRedCode::CONTAINS)) {
        return ref::ERR_INVALID;
      } else {
        using RED   = UnaryRedOp<OP, C>;
        using LHS   = typename RED::VAL;
        LHS ident   = RED::OP::identity;
        std::memset(out, 0, sizeof(LHS));
        std::memcpy(out, &ident, sizeof(LHS));
        return static_cast<int>(sizeof(LHS));
      }
    });
  });
}
}
