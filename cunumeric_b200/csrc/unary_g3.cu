// exponentials and logarithms
#define CNB_UN_GROUP_NAME unary_group3
#define CNB_UN_GROUP_OPS(X) \
  X(CNB_UOP_EXP) X(CNB_UOP_EXP2) X(CNB_UOP_EXPM1) X(CNB_UOP_LOG) X(CNB_UOP_LOG10) \
  X(CNB_UOP_LOG1P) X(CNB_UOP_LOG2) X(CNB_UOP_SQRT) X(CNB_UOP_CBRT)
#include "unary_op.inl"
