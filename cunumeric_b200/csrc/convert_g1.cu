#define CNB_CVT_GROUP_NAME convert_group1
#define CNB_CVT_GROUP_SRCS(X) X(CNB_BOOL) X(CNB_INT8) X(CNB_INT16) X(CNB_INT32) X(CNB_INT64)
#include "convert.inl"
