// trigonometric
#define CNB_UN_GROUP_NAME unary_group4
#define CNB_UN_GROUP_OPS(X) \
  X(CNB_UOP_SIN) X(CNB_UOP_COS) X(CNB_UOP_TAN) X(CNB_UOP_ARCSIN) X(CNB_UOP_ARCCOS) \
  X(CNB_UOP_ARCTAN)
#include "unary_op.inl"
