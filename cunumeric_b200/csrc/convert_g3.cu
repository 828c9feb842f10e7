#define CNB_CVT_GROUP_NAME convert_group3
#define CNB_CVT_GROUP_SRCS(X) X(CNB_FLOAT16) X(CNB_FLOAT32) X(CNB_FLOAT64)
#include "convert.inl"
