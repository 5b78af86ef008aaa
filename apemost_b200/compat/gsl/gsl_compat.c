/*
 * gsl_compat.c -- implementation of the GSL-compatible subset declared in
 * gsl_compat.h.  Written from the documented GSL semantics (SURVEY.md
 * Appendix E); no GSL source was available or consulted.
 */
#include "gsl_compat.h"

/* --------------------------------------------------------------- errno --- */
const char * gsl_strerror(const int e) {
	switch (e) {
	case GSL_SUCCESS: return "success";
	case GSL_FAILURE: return "failure";
	case GSL_EDOM: return "input domain error";
	case GSL_ERANGE: return "output range error";
	case GSL_EINVAL: return "invalid argument supplied by user";
	case GSL_EFAILED: return "generic failure";
	case GSL_ENOMEM: return "malloc failed";
	case GSL_EBADLEN: return "matrix/vector sizes are not conformant";
	case GSL_ENOTSQR: return "matrix not square";
	case GSL_ESING: return "singularity or extremely bad function behavior detected";
	default: return "unknown error code";
	}
}

void gsl_error(const char * reason, const char * file, int line, int gsl_errno) {
	fprintf(stderr, "gsl: %s:%d: ERROR: %s\n", file, line, reason);
	fprintf(stderr, "Default GSL error handler invoked (%s).\n",
			gsl_strerror(gsl_errno));
	fflush(stderr);
	abort();
}
#define APM_GSL_ERROR(reason, code) do { \
	gsl_error(reason, __FILE__, __LINE__, code); return code; } while (0)

/* -------------------------------------------------------------- vector --- */
static gsl_vector * vector_new(const size_t n, int zero) {
	gsl_vector * v = (gsl_vector *) malloc(sizeof(gsl_vector));
	gsl_block * b = (gsl_block *) malloc(sizeof(gsl_block));
	if (v == NULL || b == NULL) {
		gsl_error("failed to allocate space for vector", __FILE__, __LINE__,
				GSL_ENOMEM);
		return NULL;
	}
	b->size = n;
	b->data = (double *) (zero ? calloc(n ? n : 1, sizeof(double)) : malloc(
			(n ? n : 1) * sizeof(double)));
	if (b->data == NULL) {
		gsl_error("failed to allocate space for block", __FILE__, __LINE__,
				GSL_ENOMEM);
		return NULL;
	}
	v->size = n;
	v->stride = 1;
	v->data = b->data;
	v->block = b;
	v->owner = 1;
	return v;
}
gsl_vector * gsl_vector_alloc(const size_t n) {
	return vector_new(n, 0);
}
gsl_vector * gsl_vector_calloc(const size_t n) {
	return vector_new(n, 1);
}
void gsl_vector_free(gsl_vector * v) {
	if (v == NULL)
		return;
	if (v->owner && v->block != NULL) {
		free(v->block->data);
		free(v->block);
	}
	free(v);
}
void gsl_vector_set_all(gsl_vector * v, double x) {
	size_t i;
	for (i = 0; i < v->size; i++)
		v->data[i * v->stride] = x;
}
void gsl_vector_set_zero(gsl_vector * v) {
	gsl_vector_set_all(v, 0.0);
}
int gsl_vector_memcpy(gsl_vector * dest, const gsl_vector * src) {
	size_t i;
	if (dest->size != src->size)
		APM_GSL_ERROR("vector lengths are not equal", GSL_EBADLEN);
	for (i = 0; i < src->size; i++)
		dest->data[i * dest->stride] = src->data[i * src->stride];
	return GSL_SUCCESS;
}
#define VEC_BINOP(name, op) \
int name(gsl_vector * a, const gsl_vector * b) { \
	size_t i; \
	if (a->size != b->size) \
		APM_GSL_ERROR("vectors must have same length", GSL_EBADLEN); \
	for (i = 0; i < a->size; i++) \
		a->data[i * a->stride] op b->data[i * b->stride]; \
	return GSL_SUCCESS; \
}
VEC_BINOP(gsl_vector_add, +=)
VEC_BINOP(gsl_vector_sub, -=)
VEC_BINOP(gsl_vector_mul, *=)
VEC_BINOP(gsl_vector_div, /=)
int gsl_vector_scale(gsl_vector * a, const double x) {
	size_t i;
	for (i = 0; i < a->size; i++)
		a->data[i * a->stride] *= x;
	return GSL_SUCCESS;
}
int gsl_vector_add_constant(gsl_vector * a, const double x) {
	size_t i;
	for (i = 0; i < a->size; i++)
		a->data[i * a->stride] += x;
	return GSL_SUCCESS;
}
void gsl_vector_minmax(const gsl_vector * v, double * min_out, double * max_out) {
	double mn = v->data[0], mx = v->data[0];
	size_t i;
	for (i = 0; i < v->size; i++) {
		double x = v->data[i * v->stride];
		if (x < mn)
			mn = x;
		if (x > mx)
			mx = x;
		if (isnan(x)) {
			mn = x;
			mx = x;
			break;
		}
	}
	*min_out = mn;
	*max_out = mx;
}
double gsl_vector_max(const gsl_vector * v) {
	double mn, mx;
	gsl_vector_minmax(v, &mn, &mx);
	return mx;
}
double gsl_vector_min(const gsl_vector * v) {
	double mn, mx;
	gsl_vector_minmax(v, &mn, &mx);
	return mn;
}
int gsl_vector_fprintf(FILE * stream, const gsl_vector * v, const char * format) {
	size_t i;
	for (i = 0; i < v->size; i++) {
		if (fprintf(stream, format, v->data[i * v->stride]) < 0)
			APM_GSL_ERROR("fprintf failed", GSL_EFAILED);
		if (putc('\n', stream) == EOF)
			APM_GSL_ERROR("putc failed", GSL_EFAILED);
	}
	return GSL_SUCCESS;
}

gsl_vector_int * gsl_vector_int_alloc(const size_t n) {
	gsl_vector_int * v = (gsl_vector_int *) malloc(sizeof(gsl_vector_int));
	v->size = n;
	v->stride = 1;
	v->data = (int *) malloc((n ? n : 1) * sizeof(int));
	v->block = NULL;
	v->owner = 1;
	return v;
}
void gsl_vector_int_free(gsl_vector_int * v) {
	if (v == NULL)
		return;
	free(v->data);
	free(v);
}
void gsl_vector_int_set_all(gsl_vector_int * v, int x) {
	size_t i;
	for (i = 0; i < v->size; i++)
		v->data[i * v->stride] = x;
}

/* -------------------------------------------------------------- matrix --- */
static gsl_matrix * matrix_new(const size_t n1, const size_t n2, int zero) {
	gsl_matrix * m = (gsl_matrix *) malloc(sizeof(gsl_matrix));
	gsl_block * b = (gsl_block *) malloc(sizeof(gsl_block));
	size_t n = n1 * n2;
	if (m == NULL || b == NULL) {
		gsl_error("failed to allocate space for matrix", __FILE__, __LINE__,
				GSL_ENOMEM);
		return NULL;
	}
	b->size = n;
	b->data = (double *) (zero ? calloc(n ? n : 1, sizeof(double)) : malloc(
			(n ? n : 1) * sizeof(double)));
	if (b->data == NULL) {
		gsl_error("failed to allocate space for block", __FILE__, __LINE__,
				GSL_ENOMEM);
		return NULL;
	}
	m->size1 = n1;
	m->size2 = n2;
	m->tda = n2;
	m->data = b->data;
	m->block = b;
	m->owner = 1;
	return m;
}
gsl_matrix * gsl_matrix_alloc(const size_t n1, const size_t n2) {
	return matrix_new(n1, n2, 0);
}
gsl_matrix * gsl_matrix_calloc(const size_t n1, const size_t n2) {
	return matrix_new(n1, n2, 1);
}
void gsl_matrix_free(gsl_matrix * m) {
	if (m == NULL)
		return;
	if (m->owner && m->block != NULL) {
		free(m->block->data);
		free(m->block);
	}
	free(m);
}
void gsl_matrix_set_all(gsl_matrix * m, double x) {
	size_t i, j;
	for (i = 0; i < m->size1; i++)
		for (j = 0; j < m->size2; j++)
			m->data[i * m->tda + j] = x;
}
int gsl_matrix_get_col(gsl_vector * v, const gsl_matrix * m, const size_t j) {
	size_t i;
	if (j >= m->size2)
		APM_GSL_ERROR("column index is out of range", GSL_EINVAL);
	if (v->size != m->size1)
		APM_GSL_ERROR("matrix column size and vector length are not equal",
				GSL_EBADLEN);
	for (i = 0; i < m->size1; i++)
		v->data[i * v->stride] = m->data[i * m->tda + j];
	return GSL_SUCCESS;
}
_gsl_vector_const_view gsl_matrix_const_column(const gsl_matrix * m,
		const size_t j) {
	_gsl_vector_const_view view;
	memset(&view, 0, sizeof(view));
	if (j >= m->size2) {
		gsl_error("column index is out of range", __FILE__, __LINE__, GSL_EINVAL);
		return view;
	}
	view.vector.data = m->data + j;
	view.vector.size = m->size1;
	view.vector.stride = m->tda;
	view.vector.block = m->block;
	view.vector.owner = 0;
	return view;
}
int gsl_matrix_fscanf(FILE * stream, gsl_matrix * m) {
	size_t i, j;
	for (i = 0; i < m->size1; i++) {
		for (j = 0; j < m->size2; j++) {
			double tmp;
			if (fscanf(stream, "%lg", &tmp) != 1)
				APM_GSL_ERROR("fscanf failed", GSL_EFAILED);
			m->data[i * m->tda + j] = tmp;
		}
	}
	return GSL_SUCCESS;
}

/* ----------------------------------------------------------------- rng --- */
#define MT_N 624
#define MT_M 397
typedef struct {
	unsigned long mt[MT_N];
	int mti;
} mt_state_t;

static void mt_set(void * vstate, unsigned long int s) {
	mt_state_t * state = (mt_state_t *) vstate;
	int i;
	if (s == 0)
		s = 4357; /* the generator's own default seed */
	state->mt[0] = s & 0xffffffffUL;
	for (i = 1; i < MT_N; i++) {
		state->mt[i] = (1812433253UL * (state->mt[i - 1] ^ (state->mt[i - 1]
				>> 30)) + (unsigned long) i);
		state->mt[i] &= 0xffffffffUL;
	}
	state->mti = i;
}
static unsigned long int mt_get(void * vstate) {
	mt_state_t * state = (mt_state_t *) vstate;
	unsigned long k;
	unsigned long * const mt = state->mt;
	const unsigned long UPPER = 0x80000000UL, LOWER = 0x7fffffffUL;
#define MT_MAGIC(y) (((y) & 0x1) ? 0x9908b0dfUL : 0)
	if (state->mti >= MT_N) {
		int kk;
		for (kk = 0; kk < MT_N - MT_M; kk++) {
			unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
			mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ MT_MAGIC(y);
		}
		for (; kk < MT_N - 1; kk++) {
			unsigned long y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
			mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ MT_MAGIC(y);
		}
		{
			unsigned long y = (mt[MT_N - 1] & UPPER) | (mt[0] & LOWER);
			mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ MT_MAGIC(y);
		}
		state->mti = 0;
	}
#undef MT_MAGIC
	k = mt[state->mti];
	k ^= (k >> 11);
	k ^= (k << 7) & 0x9d2c5680UL;
	k ^= (k << 15) & 0xefc60000UL;
	k ^= (k >> 18);
	state->mti++;
	return k & 0xffffffffUL;
}
static double mt_get_double(void * vstate) {
	return mt_get(vstate) / 4294967296.0;
}
static const gsl_rng_type mt_type = { "mt19937", 0xffffffffUL, 0,
		sizeof(mt_state_t), &mt_set, &mt_get, &mt_get_double };
const gsl_rng_type * gsl_rng_mt19937 = &mt_type;
const gsl_rng_type * gsl_rng_default = &mt_type;
unsigned long int gsl_rng_default_seed = 0;

const gsl_rng_type * gsl_rng_env_setup(void) {
	unsigned long int seed = 0;
	const char * p = getenv("GSL_RNG_TYPE");
	if (p != NULL) {
		if (strcmp(p, "mt19937") != 0) {
			gsl_error("unknown generator (only mt19937 is provided)", __FILE__,
					__LINE__, GSL_EINVAL);
			return NULL;
		}
		fprintf(stderr, "GSL_RNG_TYPE=%s\n", p);
	}
	gsl_rng_default = gsl_rng_mt19937;
	p = getenv("GSL_RNG_SEED");
	if (p != NULL) {
		seed = strtoul(p, 0, 0);
		fprintf(stderr, "GSL_RNG_SEED=%lu\n", seed);
	}
	gsl_rng_default_seed = seed;
	return gsl_rng_default;
}
gsl_rng * gsl_rng_alloc(const gsl_rng_type * T) {
	gsl_rng * r = (gsl_rng *) malloc(sizeof(gsl_rng));
	if (r == NULL) {
		gsl_error("failed to allocate space for rng struct", __FILE__, __LINE__,
				GSL_ENOMEM);
		return NULL;
	}
	r->state = calloc(1, T->size);
	r->type = T;
	gsl_rng_set(r, gsl_rng_default_seed);
	return r;
}
void gsl_rng_free(gsl_rng * r) {
	if (r == NULL)
		return;
	free(r->state);
	free(r);
}
void gsl_rng_set(const gsl_rng * r, unsigned long int seed) {
	(r->type->set)(r->state, seed);
}
unsigned long int gsl_rng_get(const gsl_rng * r) {
	return (r->type->get)(r->state);
}
double gsl_rng_uniform(const gsl_rng * r) {
	return (r->type->get_double)(r->state);
}
double gsl_rng_uniform_pos(const gsl_rng * r) {
	double x;
	do {
		x = (r->type->get_double)(r->state);
	} while (x == 0);
	return x;
}

/* ------------------------------------------------------------- randist --- */
double gsl_ran_gaussian(const gsl_rng * r, const double sigma) {
	double x, y, r2;
	do {
		/* choose x,y in uniform square (-1,-1) to (+1,+1) */
		x = -1 + 2 * gsl_rng_uniform_pos(r);
		y = -1 + 2 * gsl_rng_uniform_pos(r);
		/* see if it is in the unit circle */
		r2 = x * x + y * y;
	} while (r2 > 1.0 || r2 == 0);
	/* Box-Muller transform */
	return sigma * y * sqrt(-2.0 * log(r2) / r2);
}
double gsl_ran_logistic(const gsl_rng * r, const double a) {
	double x, z;
	do {
		x = gsl_rng_uniform_pos(r);
	} while (x == 1);
	z = log(x / (1 - x));
	return a * z;
}
double gsl_ran_flat(const gsl_rng * r, const double a, const double b) {
	double u = gsl_rng_uniform(r);
	return a * (1 - u) + b * u;
}

/* ------------------------------------------------------------------ sf --- */
double gsl_sf_log(const double x) {
	if (x <= 0.0) {
		gsl_error("domain error", __FILE__, __LINE__, GSL_EDOM);
		return NAN;
	}
	return log(x);
}
double gsl_sf_sin(const double x) {
	return sin(x);
}
double gsl_sf_cos(const double x) {
	return cos(x);
}

/* ----------------------------------------------------------- histogram --- */
gsl_histogram * gsl_histogram_alloc(size_t n) {
	gsl_histogram * h;
	if (n == 0) {
		gsl_error("histogram length n must be positive integer", __FILE__,
				__LINE__, GSL_EDOM);
		return NULL;
	}
	h = (gsl_histogram *) malloc(sizeof(gsl_histogram));
	h->range = (double *) calloc(n + 1, sizeof(double));
	h->bin = (double *) calloc(n, sizeof(double));
	h->n = n;
	return h;
}
void gsl_histogram_free(gsl_histogram * h) {
	if (h == NULL)
		return;
	free(h->range);
	free(h->bin);
	free(h);
}
int gsl_histogram_set_ranges_uniform(gsl_histogram * h, double xmin, double xmax) {
	size_t i;
	const size_t n = h->n;
	if (xmin >= xmax)
		APM_GSL_ERROR("xmin must be less than xmax", GSL_EINVAL);
	for (i = 0; i <= n; i++) {
		double f1 = ((double) (n - i) / (double) n);
		double f2 = ((double) i / (double) n);
		h->range[i] = f1 * xmin + f2 * xmax;
	}
	for (i = 0; i < n; i++)
		h->bin[i] = 0;
	return GSL_SUCCESS;
}
static int hist_find(const size_t n, const double range[], const double x,
		size_t * i) {
	size_t i_linear, lower, upper, mid;
	if (x < range[0])
		return -1;
	if (x >= range[n])
		return +1;
	/* optimize for linear case */
	{
		double u = (x - range[0]) / (range[n] - range[0]);
		i_linear = (size_t) (u * n);
	}
	if (x >= range[i_linear] && x < range[i_linear + 1]) {
		*i = i_linear;
		return 0;
	}
	/* binary search */
	upper = n;
	lower = 0;
	while (upper - lower > 1) {
		mid = (upper + lower) / 2;
		if (x >= range[mid])
			lower = mid;
		else
			upper = mid;
	}
	*i = lower;
	if (x < range[lower] || x >= range[lower + 1])
		APM_GSL_ERROR("x not found in range", GSL_EFAILED);
	return 0;
}
int gsl_histogram_accumulate(gsl_histogram * h, double x, double weight) {
	size_t index = 0;
	int status = hist_find(h->n, h->range, x, &index);
	if (status)
		return GSL_EDOM;
	if (index >= h->n)
		APM_GSL_ERROR("index lies outside valid range of 0 .. n - 1", GSL_EFAILED);
	h->bin[index] += weight;
	return GSL_SUCCESS;
}
int gsl_histogram_increment(gsl_histogram * h, double x) {
	return gsl_histogram_accumulate(h, x, 1.0);
}
double gsl_histogram_get(const gsl_histogram * h, size_t i) {
	if (i >= h->n) {
		gsl_error("index lies outside valid range of 0 .. n - 1", __FILE__,
				__LINE__, GSL_EDOM);
		return 0;
	}
	return h->bin[i];
}
int gsl_histogram_get_range(const gsl_histogram * h, size_t i, double * lower,
		double * upper) {
	if (i >= h->n)
		APM_GSL_ERROR("index lies outside valid range of 0 .. n - 1", GSL_EDOM);
	*lower = h->range[i];
	*upper = h->range[i + 1];
	return GSL_SUCCESS;
}
double gsl_histogram_max(const gsl_histogram * h) {
	return h->range[h->n];
}
double gsl_histogram_min(const gsl_histogram * h) {
	return h->range[0];
}
size_t gsl_histogram_bins(const gsl_histogram * h) {
	return h->n;
}
double gsl_histogram_sum(const gsl_histogram * h) {
	double sum = 0;
	size_t i;
	for (i = 0; i < h->n; i++)
		sum += h->bin[i];
	return sum;
}
double gsl_histogram_mean(const gsl_histogram * h) {
	size_t i;
	long double wmean = 0, W = 0;
	for (i = 0; i < h->n; i++) {
		double xi = (h->range[i + 1] + h->range[i]) / 2;
		double wi = h->bin[i];
		if (wi > 0) {
			W += wi;
			wmean += (xi - wmean) * (wi / W);
		}
	}
	return wmean;
}
double gsl_histogram_sigma(const gsl_histogram * h) {
	size_t i;
	long double wvariance = 0, wmean = 0, W = 0;
	for (i = 0; i < h->n; i++) {
		double xi = (h->range[i + 1] + h->range[i]) / 2;
		double wi = h->bin[i];
		if (wi > 0) {
			W += wi;
			wmean += (xi - wmean) * (wi / W);
		}
	}
	W = 0.0;
	for (i = 0; i < h->n; i++) {
		double xi = ((h->range[i + 1]) + (h->range[i])) / 2;
		double wi = h->bin[i];
		if (wi > 0) {
			const long double delta = (xi - wmean);
			W += wi;
			wvariance += (delta * delta - wvariance) * (wi / W);
		}
	}
	return sqrt((double) wvariance);
}
int gsl_histogram_scale(gsl_histogram * h, double scale) {
	size_t i;
	for (i = 0; i < h->n; i++)
		h->bin[i] *= scale;
	return GSL_SUCCESS;
}
int gsl_histogram_fprintf(FILE * stream, const gsl_histogram * h,
		const char * range_format, const char * bin_format) {
	size_t i;
	for (i = 0; i < h->n; i++) {
		if (fprintf(stream, range_format, h->range[i]) < 0)
			APM_GSL_ERROR("fprintf failed", GSL_EFAILED);
		if (putc(' ', stream) == EOF)
			APM_GSL_ERROR("putc failed", GSL_EFAILED);
		if (fprintf(stream, range_format, h->range[i + 1]) < 0)
			APM_GSL_ERROR("fprintf failed", GSL_EFAILED);
		if (putc(' ', stream) == EOF)
			APM_GSL_ERROR("putc failed", GSL_EFAILED);
		if (fprintf(stream, bin_format, h->bin[i]) < 0)
			APM_GSL_ERROR("fprintf failed", GSL_EFAILED);
		if (putc('\n', stream) == EOF)
			APM_GSL_ERROR("putc failed", GSL_EFAILED);
	}
	return GSL_SUCCESS;
}

/* -------------------------------------------------------------- linalg --- */
gsl_permutation * gsl_permutation_alloc(const size_t n) {
	gsl_permutation * p = (gsl_permutation *) malloc(sizeof(gsl_permutation));
	p->size = n;
	p->data = (size_t *) malloc((n ? n : 1) * sizeof(size_t));
	return p;
}
void gsl_permutation_free(gsl_permutation * p) {
	if (p == NULL)
		return;
	free(p->data);
	free(p);
}
int gsl_linalg_LU_decomp(gsl_matrix * A, gsl_permutation * p, int * signum) {
	const size_t N = A->size1;
	size_t i, j, k;
	if (A->size1 != A->size2)
		APM_GSL_ERROR("LU decomposition requires square matrix", GSL_ENOTSQR);
	if (p->size != N)
		APM_GSL_ERROR("permutation length must match matrix size", GSL_EBADLEN);
	*signum = 1;
	for (i = 0; i < N; i++)
		p->data[i] = i;
	for (j = 0; j + 1 < N; j++) {
		/* find the pivot in column j */
		double max = fabs(gsl_matrix_get(A, j, j));
		size_t i_pivot = j;
		for (i = j + 1; i < N; i++) {
			double aij = fabs(gsl_matrix_get(A, i, j));
			if (aij > max) {
				max = aij;
				i_pivot = i;
			}
		}
		if (i_pivot != j) {
			for (k = 0; k < N; k++) {
				double t = gsl_matrix_get(A, j, k);
				gsl_matrix_set(A, j, k, gsl_matrix_get(A, i_pivot, k));
				gsl_matrix_set(A, i_pivot, k, t);
			}
			{
				size_t t = p->data[j];
				p->data[j] = p->data[i_pivot];
				p->data[i_pivot] = t;
			}
			*signum = -(*signum);
		}
		{
			double ajj = gsl_matrix_get(A, j, j);
			if (ajj != 0.0) {
				for (i = j + 1; i < N; i++) {
					double aij = gsl_matrix_get(A, i, j) / ajj;
					gsl_matrix_set(A, i, j, aij);
					for (k = j + 1; k < N; k++) {
						double aik = gsl_matrix_get(A, i, k);
						double ajk = gsl_matrix_get(A, j, k);
						gsl_matrix_set(A, i, k, aik - aij * ajk);
					}
				}
			}
		}
	}
	return GSL_SUCCESS;
}
int gsl_linalg_LU_solve(const gsl_matrix * LU, const gsl_permutation * p,
		const gsl_vector * b, gsl_vector * x) {
	const size_t N = LU->size1;
	size_t i, j;
	if (LU->size1 != LU->size2)
		APM_GSL_ERROR("LU matrix must be square", GSL_ENOTSQR);
	if (N != p->size || N != b->size || N != x->size)
		APM_GSL_ERROR("sizes are not conformant", GSL_EBADLEN);
	for (i = 0; i < N; i++) {
		if (gsl_matrix_get(LU, i, i) == 0.0)
			APM_GSL_ERROR("matrix is singular", GSL_EDOM);
	}
	/* x = P b */
	for (i = 0; i < N; i++)
		gsl_vector_set(x, i, gsl_vector_get(b, p->data[i]));
	/* forward: L y = P b (unit diagonal) */
	for (i = 0; i < N; i++) {
		double s = gsl_vector_get(x, i);
		for (j = 0; j < i; j++)
			s -= gsl_matrix_get(LU, i, j) * gsl_vector_get(x, j);
		gsl_vector_set(x, i, s);
	}
	/* backward: U x = y */
	for (i = N; i-- > 0;) {
		double s = gsl_vector_get(x, i);
		for (j = i + 1; j < N; j++)
			s -= gsl_matrix_get(LU, i, j) * gsl_vector_get(x, j);
		gsl_vector_set(x, i, s / gsl_matrix_get(LU, i, i));
	}
	return GSL_SUCCESS;
}
