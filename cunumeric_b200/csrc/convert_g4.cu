#define CNB_CVT_GROUP_NAME convert_group4
#define CNB_CVT_GROUP_SRCS(X) X(CNB_COMPLEX64) X(CNB_COMPLEX128)
#include "convert.inl"
