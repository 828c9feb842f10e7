#define CNB_CVT_GROUP_NAME convert_group2
#define CNB_CVT_GROUP_SRCS(X) X(CNB_UINT8) X(CNB_UINT16) X(CNB_UINT32) X(CNB_UINT64)
#include "convert.inl"
