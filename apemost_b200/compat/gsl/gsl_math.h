/* gsl_math.h -- part of the minimal GSL-compatible header set; see gsl_compat.h */
#ifndef APM_COMPAT_GSL_MATH_H_
#define APM_COMPAT_GSL_MATH_H_
#include "gsl_compat.h"
#endif
