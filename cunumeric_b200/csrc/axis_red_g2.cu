#define CNB_ARED_GROUP_NAME axis_red_group2
#define CNB_ARED_GROUP_OPS(X) X(CNB_RED_MAX) X(CNB_RED_MIN) X(CNB_RED_COUNT_NONZERO)
#include "axis_red.inl"
