/* gsl_permutation.h -- part of the minimal GSL-compatible header set; see gsl_compat.h */
#ifndef APM_COMPAT_GSL_PERMUTATION_H_
#define APM_COMPAT_GSL_PERMUTATION_H_
#include "gsl_compat.h"
#endif
