#define CNB_SRED_GROUP_NAME scalar_red_group1
#define CNB_SRED_GROUP_OPS(X) \
  X(CNB_RED_SUM) X(CNB_RED_PROD) X(CNB_RED_MAX) X(CNB_RED_MIN) X(CNB_RED_SUM_SQUARES) \
  X(CNB_RED_VARIANCE)
#include "scalar_red.inl"
