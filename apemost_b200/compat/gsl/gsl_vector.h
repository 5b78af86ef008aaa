/* gsl_vector.h -- part of the minimal GSL-compatible header set; see gsl_compat.h */
#ifndef APM_COMPAT_GSL_VECTOR_H_
#define APM_COMPAT_GSL_VECTOR_H_
#include "gsl_compat.h"
#endif
