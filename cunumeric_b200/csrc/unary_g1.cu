// structural / all-dtype unary ops
#define CNB_UN_GROUP_NAME unary_group1
#define CNB_UN_GROUP_OPS(X) \
  X(CNB_UOP_ABSOLUTE) X(CNB_UOP_CLIP) X(CNB_UOP_CONJ) X(CNB_UOP_COPY) X(CNB_UOP_POSITIVE) \
  X(CNB_UOP_NEGATIVE) X(CNB_UOP_SQUARE) X(CNB_UOP_RECIPROCAL) X(CNB_UOP_SIGN)
#include "unary_op.inl"
