#define CNB_ARED_GROUP_NAME axis_red_group4
#define CNB_ARED_GROUP_OPS(X) X(CNB_RED_ARGMAX) X(CNB_RED_ARGMIN)
#include "axis_red.inl"
