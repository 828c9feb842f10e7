#define CNB_ARED_GROUP_NAME axis_red_group1
#define CNB_ARED_GROUP_OPS(X) X(CNB_RED_SUM) X(CNB_RED_PROD)
#include "axis_red.inl"
