#define CNB_ARED_GROUP_NAME axis_red_group3
#define CNB_ARED_GROUP_OPS(X) X(CNB_RED_SUM_SQUARES) X(CNB_RED_VARIANCE) X(CNB_RED_ALL) X(CNB_RED_ANY)
#include "axis_red.inl"
