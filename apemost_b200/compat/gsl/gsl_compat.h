/*
 * gsl_compat.h -- minimal, from-scratch, GSL-API-compatible subset.
 *
 * APEMoST's plugin ABI is expressed in GSL types: the `mcmc` struct holds
 * gsl_vector* / gsl_matrix* / gsl_rng* members (reference src/mcmc_struct.h:30-106)
 * and every apps/<model>.c reads parameters and data through gsl_vector_get /
 * gsl_matrix_get.  GSL itself is not vendored by the reference (Makefile:10,
 * -lgsl -lgslcblas) and is absent on the build boxes, so the host side of this
 * engine ships the subset of the API that APEMoST touches, written against the
 * public GSL semantics (struct layouts, MT19937 seeding, polar Box-Muller,
 * histogram conventions).  If a real GSL is installed,
 * compile with -DAPM_USE_SYSTEM_GSL and this directory is simply not put on
 * the include path.
 *
 * The same files let the *unmodified* reference sources compile here as the
 * checker binary oracle/_ref (see oracle/Makefile).
 */
#ifndef APM_GSL_COMPAT_H_
#define APM_GSL_COMPAT_H_

#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* the reference builds with -ansi -pedantic, where `inline` is not a keyword */
#if defined(__GNUC__)
#define APM_GSL_INLINE static __inline__
#else
#define APM_GSL_INLINE static inline
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- math --- */
#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
#ifndef M_E
#define M_E 2.71828182845904523536028747135
#endif
#define GSL_MAX(a, b) ((a) > (b) ? (a) : (b))
#define GSL_MIN(a, b) ((a) < (b) ? (a) : (b))
#define GSL_NAN (NAN)
#define GSL_POSINF (INFINITY)
#define GSL_NEGINF (-INFINITY)
#define GSL_DBL_EPSILON 2.2204460492503131e-16

/* --------------------------------------------------------------- errno --- */
enum {
	GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_EDOM = 1,
	GSL_ERANGE = 2, GSL_EFAULT = 3, GSL_EINVAL = 4, GSL_EFAILED = 5,
	GSL_ENOMEM = 8, GSL_EBADLEN = 19, GSL_ENOTSQR = 20, GSL_ESING = 21
};
const char * gsl_strerror(const int gsl_errno);
/* default handler semantics: print and abort() */
void gsl_error(const char * reason, const char * file, int line, int gsl_errno);

/* -------------------------------------------------------------- vector --- */
typedef struct {
	size_t size;
	double * data;
} gsl_block;

typedef struct {
	size_t size;
	size_t stride;
	double * data;
	gsl_block * block;
	int owner;
} gsl_vector;

typedef struct {
	gsl_vector vector;
} _gsl_vector_view;
typedef _gsl_vector_view gsl_vector_view;
typedef struct {
	gsl_vector vector;
} _gsl_vector_const_view;
typedef const _gsl_vector_const_view gsl_vector_const_view;

gsl_vector * gsl_vector_alloc(const size_t n);
gsl_vector * gsl_vector_calloc(const size_t n);
void gsl_vector_free(gsl_vector * v);

APM_GSL_INLINE double gsl_vector_get(const gsl_vector * v, const size_t i) {
#ifndef GSL_RANGE_CHECK_OFF
	if (i >= v->size) {
		gsl_error("index out of range", __FILE__, __LINE__, GSL_EINVAL);
		return 0;
	}
#endif
	return v->data[i * v->stride];
}
APM_GSL_INLINE void gsl_vector_set(gsl_vector * v, const size_t i, double x) {
#ifndef GSL_RANGE_CHECK_OFF
	if (i >= v->size) {
		gsl_error("index out of range", __FILE__, __LINE__, GSL_EINVAL);
		return;
	}
#endif
	v->data[i * v->stride] = x;
}
void gsl_vector_set_all(gsl_vector * v, double x);
void gsl_vector_set_zero(gsl_vector * v);
int gsl_vector_memcpy(gsl_vector * dest, const gsl_vector * src);
int gsl_vector_add(gsl_vector * a, const gsl_vector * b);
int gsl_vector_sub(gsl_vector * a, const gsl_vector * b);
int gsl_vector_mul(gsl_vector * a, const gsl_vector * b);
int gsl_vector_div(gsl_vector * a, const gsl_vector * b);
int gsl_vector_scale(gsl_vector * a, const double x);
int gsl_vector_add_constant(gsl_vector * a, const double x);
double gsl_vector_max(const gsl_vector * v);
double gsl_vector_min(const gsl_vector * v);
void gsl_vector_minmax(const gsl_vector * v, double * min_out, double * max_out);
int gsl_vector_fprintf(FILE * stream, const gsl_vector * v, const char * format);

/* integer vectors (only the alternate calibrators use them) */
typedef struct {
	size_t size;
	size_t stride;
	int * data;
	void * block;
	int owner;
} gsl_vector_int;
gsl_vector_int * gsl_vector_int_alloc(const size_t n);
void gsl_vector_int_free(gsl_vector_int * v);
void gsl_vector_int_set_all(gsl_vector_int * v, int x);
APM_GSL_INLINE int gsl_vector_int_get(const gsl_vector_int * v, const size_t i) {
	return v->data[i * v->stride];
}
APM_GSL_INLINE void gsl_vector_int_set(gsl_vector_int * v, const size_t i, int x) {
	v->data[i * v->stride] = x;
}

/* -------------------------------------------------------------- matrix --- */
typedef struct {
	size_t size1;
	size_t size2;
	size_t tda;
	double * data;
	gsl_block * block;
	int owner;
} gsl_matrix;

gsl_matrix * gsl_matrix_alloc(const size_t n1, const size_t n2);
gsl_matrix * gsl_matrix_calloc(const size_t n1, const size_t n2);
void gsl_matrix_free(gsl_matrix * m);
APM_GSL_INLINE double gsl_matrix_get(const gsl_matrix * m, const size_t i,
		const size_t j) {
#ifndef GSL_RANGE_CHECK_OFF
	if (i >= m->size1 || j >= m->size2) {
		gsl_error("index out of range", __FILE__, __LINE__, GSL_EINVAL);
		return 0;
	}
#endif
	return m->data[i * m->tda + j];
}
APM_GSL_INLINE void gsl_matrix_set(gsl_matrix * m, const size_t i, const size_t j,
		const double x) {
#ifndef GSL_RANGE_CHECK_OFF
	if (i >= m->size1 || j >= m->size2) {
		gsl_error("index out of range", __FILE__, __LINE__, GSL_EINVAL);
		return;
	}
#endif
	m->data[i * m->tda + j] = x;
}
void gsl_matrix_set_all(gsl_matrix * m, double x);
int gsl_matrix_get_col(gsl_vector * v, const gsl_matrix * m, const size_t j);
_gsl_vector_const_view gsl_matrix_const_column(const gsl_matrix * m,
		const size_t j);
int gsl_matrix_fscanf(FILE * stream, gsl_matrix * m);

/* ----------------------------------------------------------------- rng --- */
typedef struct {
	const char * name;
	unsigned long int max;
	unsigned long int min;
	size_t size;
	void (*set)(void * state, unsigned long int seed);
	unsigned long int (*get)(void * state);
	double (*get_double)(void * state);
} gsl_rng_type;

typedef struct {
	const gsl_rng_type * type;
	void * state;
} gsl_rng;

extern const gsl_rng_type * gsl_rng_mt19937;
extern const gsl_rng_type * gsl_rng_default;
extern unsigned long int gsl_rng_default_seed;

const gsl_rng_type * gsl_rng_env_setup(void);
gsl_rng * gsl_rng_alloc(const gsl_rng_type * T);
void gsl_rng_free(gsl_rng * r);
void gsl_rng_set(const gsl_rng * r, unsigned long int seed);
unsigned long int gsl_rng_get(const gsl_rng * r);
double gsl_rng_uniform(const gsl_rng * r);
double gsl_rng_uniform_pos(const gsl_rng * r);

/* ------------------------------------------------------------- randist --- */
double gsl_ran_gaussian(const gsl_rng * r, const double sigma);
double gsl_ran_logistic(const gsl_rng * r, const double a);
double gsl_ran_flat(const gsl_rng * r, const double a, const double b);

/* ------------------------------------------------------------------ sf --- */
double gsl_sf_log(const double x);
double gsl_sf_sin(const double x);
double gsl_sf_cos(const double x);

/* ----------------------------------------------------------- histogram --- */
typedef struct {
	size_t n;
	double * range;
	double * bin;
} gsl_histogram;

gsl_histogram * gsl_histogram_alloc(size_t n);
void gsl_histogram_free(gsl_histogram * h);
int gsl_histogram_set_ranges_uniform(gsl_histogram * h, double xmin, double xmax);
int gsl_histogram_increment(gsl_histogram * h, double x);
int gsl_histogram_accumulate(gsl_histogram * h, double x, double weight);
double gsl_histogram_get(const gsl_histogram * h, size_t i);
int gsl_histogram_get_range(const gsl_histogram * h, size_t i, double * lower,
		double * upper);
double gsl_histogram_max(const gsl_histogram * h);
double gsl_histogram_min(const gsl_histogram * h);
size_t gsl_histogram_bins(const gsl_histogram * h);
double gsl_histogram_sum(const gsl_histogram * h);
double gsl_histogram_mean(const gsl_histogram * h);
double gsl_histogram_sigma(const gsl_histogram * h);
int gsl_histogram_scale(gsl_histogram * h, double scale);
int gsl_histogram_fprintf(FILE * stream, const gsl_histogram * h,
		const char * range_format, const char * bin_format);

/* -------------------------------------------------------------- linalg --- */
typedef struct {
	size_t size;
	size_t * data;
} gsl_permutation;
gsl_permutation * gsl_permutation_alloc(const size_t n);
void gsl_permutation_free(gsl_permutation * p);
int gsl_linalg_LU_decomp(gsl_matrix * A, gsl_permutation * p, int * signum);
int gsl_linalg_LU_solve(const gsl_matrix * LU, const gsl_permutation * p,
		const gsl_vector * b, gsl_vector * x);

#ifdef __cplusplus
}
#endif

#endif /* APM_GSL_COMPAT_H_ */
