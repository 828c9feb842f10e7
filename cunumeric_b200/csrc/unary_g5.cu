// hyperbolic
#define CNB_UN_GROUP_NAME unary_group5
#define CNB_UN_GROUP_OPS(X) \
  X(CNB_UOP_SINH) X(CNB_UOP_COSH) X(CNB_UOP_TANH) X(CNB_UOP_ARCSINH) X(CNB_UOP_ARCCOSH) \
  X(CNB_UOP_ARCTANH)
#include "unary_op.inl"
