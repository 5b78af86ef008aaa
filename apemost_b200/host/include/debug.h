/* debug.h -- the diagnostic macros APEMoST model files use (reference src/debug.h:34-100):
 * active with -DDEBUG / -DVERBOSE, compiled out otherwise. */
#ifndef APM_HOST_DEBUG_H_
#define APM_HOST_DEBUG_H_
#include "apm_host.h"

#ifdef DEBUG
#define IFDEBUG if (1)
#else
#define IFDEBUG if (0)
#endif
#ifdef VERBOSE
#define IFVERBOSE if (1)
#else
#define IFVERBOSE if (0)
#endif
#ifdef SEGV
#define IFSEGV if (1)
#else
#define IFSEGV if (0)
#endif

#define APM_DBG_(fmt, str, var) IFDEBUG { printf("\tDEBUG[%s:%d]: %s: " fmt "\n", __FILE__, __LINE__, str, var); fflush(NULL); }
#define debug(str)          IFDEBUG { printf("\tDEBUG[%s:%d]: %s\n", __FILE__, __LINE__, str); fflush(NULL); }
#define dump_i(str, var)    APM_DBG_("%i", str, var)
#define dump_ui(str, var)   APM_DBG_("%u", str, var)
#define dump_d(str, var)    APM_DBG_("%f", str, var)
#define dump_ul(str, var)   APM_DBG_("%lu", str, var)
#define dump_size(str, var) APM_DBG_("%lu", str, (unsigned long) (var))
#define dump_s(str, var)    APM_DBG_("%s", str, var)
#define dump_p(str, var)    APM_DBG_("%p", str, (void *) (var))
#define dump_v(str, v)      IFDEBUG { printf("\tDEBUG[%s:%d]: %s: ", __FILE__, __LINE__, str); dump_vectorln(v); fflush(NULL); }
#define require(x) (x)
#endif
