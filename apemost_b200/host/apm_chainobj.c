/*
 * apm_chainobj.c -- the host-side chain object: allocation, the accessors model
 * files call, the `params` / `data` parsers and the beta ladder.
 *
 * Behaviour follows the reference (cited per function); the chain structs are a
 * host mirror of the device state and the carrier of everything that is read
 * from or written to files.
 */
#include <ctype.h>
#include "apm_host.h"

#define APM_NAME_MAX 256

static void * xcalloc(size_t n, size_t size) {
	void * p = calloc(n ? n : 1, size);
	if (p == NULL) {
		fprintf(stderr, "out of memory\n");
		exit(1);
	}
	return p;
}

/* ---- life cycle: reference src/mcmc.c:37-118 ------------------------------- */
mcmc * mcmc_init(const unsigned int n_pars) {
	mcmc * m = (mcmc *) xcalloc(1, sizeof(mcmc));
	m->n_par = n_pars;
	m->prob = -1e+10;      /* no point evaluated yet (src/mcmc.c:47,49) */
	m->prob_best = -1e+10;
	m->prior = 0;
	m->random = NULL;      /* random numbers are drawn on the device (counter RNG per chain) */
	m->params = gsl_vector_alloc(n_pars);
	m->params_best = gsl_vector_alloc(n_pars);
	m->params_step = gsl_vector_calloc(n_pars);
	m->params_min = gsl_vector_calloc(n_pars);
	m->params_max = gsl_vector_calloc(n_pars);
	m->params_descr = (const char **) xcalloc(n_pars, sizeof(char *));
	m->params_accepts = (unsigned long *) xcalloc(n_pars, sizeof(unsigned long));
	m->params_rejects = (unsigned long *) xcalloc(n_pars, sizeof(unsigned long));
	m->files = NULL;
	m->data = NULL;
	m->additional_data = NULL;
	return m;
}

mcmc * mcmc_free(mcmc * m) {
	unsigned int i;
	if (m == NULL)
		return NULL;
	mcmc_dump_close(m);
	gsl_vector_free(m->params);
	gsl_vector_free(m->params_best);
	gsl_vector_free(m->params_step);
	gsl_vector_free(m->params_min);
	gsl_vector_free(m->params_max);
	for (i = 0; i < m->n_par; i++)
		free((void *) m->params_descr[i]);
	free((void *) m->params_descr);
	free(m->params_accepts);
	free(m->params_rejects);
	if (m->data != NULL)
		gsl_matrix_free((gsl_matrix *) m->data);
	m->data = NULL;
	return m;
}

/* invariants of a usable chain: reference src/mcmc.c:120-131 */
void mcmc_check(const mcmc * m) {
	assert(m != NULL);
	assert(m->n_par > 0);
	assert(m->params != NULL && m->params->size == m->n_par);
	assert(m->params_best != NULL && m->params_best->size == m->n_par);
	assert(m->params_step != NULL && m->params_step->size == m->n_par);
	assert(m->params_min != NULL && m->params_max != NULL);
	assert(m->params_descr != NULL);
	assert(m->data != NULL);
	(void) m;
}

gsl_vector * dup_vector(const gsl_vector * v) {
	gsl_vector * c = gsl_vector_alloc(v->size);
	gsl_vector_memcpy(c, v);
	return c;
}

/* reference src/mcmc_internal.h:46-48 */
double mod_double(double x, double div) {
	return x < 0 ? x - div * (int) (x / div - 1) : x - div * (int) (x / div);
}

/* ---- input files ------------------------------------------------------------- */
static FILE * open_or_die(const char * filename) {
	FILE * f = fopen(filename, "r");
	if (f == NULL) {
		fprintf(stderr, "error opening file %s\n", filename);
		perror("file could not be opened");
		exit(1);
	}
	return f;
}

/* number of '\n' in the file = number of parameters / data rows (reference src/utils.c:32-48) */
static unsigned int count_newlines(const char * filename) {
	FILE * f = open_or_die(filename);
	unsigned int n = 0;
	int c;
	while ((c = fgetc(f)) != EOF)
		if (c == '\n')
			n++;
	fclose(f);
	return n;
}

/* whitespace-separated fields on the first line (reference src/utils.c:50-75) */
static unsigned int count_first_line_fields(const char * filename) {
	FILE * f = open_or_die(filename);
	static char line[10000];
	unsigned int fields = 0;
	int in_field = 0;
	const char * p;
	if (fgets(line, sizeof(line), f) == NULL) {
		fprintf(stderr, "error: file %s is empty!", filename);
		exit(1);
	}
	fclose(f);
	for (p = line; *p; p++) {
		if (isspace((unsigned char) *p)) {
			in_field = 0;
		} else if (!in_field) {
			in_field = 1;
			fields++;
		}
	}
	return fields;
}

/* one line of the params file: start min max name step (reference src/mcmc_parser.c:47-95) */
static int parse_param_line(mcmc * m, FILE * f, unsigned int i) {
	double start, lo, hi, step;
	char * name = (char *) xcalloc(APM_NAME_MAX, 1);
	int got = fscanf(f, "%lf\t%lf\t%lf\t%255s\t%lf\n", &start, &lo, &hi, name, &step);
	if (got != 5) {
		fprintf(stderr, "only %d fields matched.\n", got);
		return 1;
	}
	if (name[0] == 0) {
		fprintf(stderr, "description invalid: %s\n", name);
		return 1;
	}
	if (lo > hi) {
		fprintf(stderr, "min(%f) < max(%f)\n", lo, hi);
		return 1;
	}
	if (start > hi) {
		fprintf(stderr, "start(%f) > max(%f)\n", start, hi);
		return 1;
	}
	if (start < lo) {
		fprintf(stderr, "start(%f) < min(%f)\n", start, lo);
		return 1;
	}
	if (step < 0) /* automatic: 10 % of the parameter range */
		step = (hi - lo) * 0.1;
	gsl_vector_set(m->params, i, start);
	gsl_vector_set(m->params_best, i, start);
	gsl_vector_set(m->params_min, i, lo);
	gsl_vector_set(m->params_max, i, hi);
	gsl_vector_set(m->params_step, i, step);
	m->params_descr[i] = name;
	return 0;
}

mcmc * mcmc_load_params(const char * filename) {
	unsigned int n = count_newlines(filename), i;
	mcmc * m = mcmc_init(n);
	FILE * f = open_or_die(filename);
	for (i = 0; i < n; i++) {
		if (parse_param_line(m, f, i) != 0) {
			fprintf(stderr, "Line %d of %s is of incorrect format.\n", i + 1, filename);
			exit(1);
		}
	}
	fclose(f);
	return m;
}

/* rows = '\n' count, columns = fields of the first line, values row-major
 * (reference src/mcmc_parser.c:97-122) */
void mcmc_load_data(mcmc * m, const char * datafilename) {
	unsigned int rows = count_newlines(datafilename);
	unsigned int cols = count_first_line_fields(datafilename);
	gsl_matrix * data = gsl_matrix_alloc(rows, cols);
	FILE * f = open_or_die(datafilename);
	if (gsl_matrix_fscanf(f, data) != 0) {
		fprintf(stderr, "error reading input data. Perhaps inconsistent format?\n");
		fprintf(stderr, "tried to read %d x %d.\n", cols, rows);
		exit(3);
	}
	fclose(f);
	m->data = data;
}

void mcmc_reuse_data(mcmc * m, const mcmc * m_orig) {
	assert(m_orig->data != NULL);
	m->data = m_orig->data;
}

mcmc * mcmc_load(const char * filename, const char * datafilename) {
	mcmc * m = mcmc_load_params(filename);
	mcmc_load_data(m, datafilename);
	return m;
}

/* ---- accessors: reference src/mcmc_gettersetter.c ----------------------------- */
unsigned int get_n_par(const mcmc * m) { return m->n_par; }
gsl_vector * get_params(const mcmc * m) { return m->params; }
double get_params_for(const mcmc * m, const unsigned int i) { return gsl_vector_get(m->params, i); }
void set_params_for(mcmc * m, const double v, const unsigned int i) { gsl_vector_set(m->params, i, v); }
void set_params(mcmc * m, gsl_vector * v) {
	gsl_vector_free(m->params);
	m->params = v;
}
gsl_vector * get_params_best(const mcmc * m) { return m->params_best; }
void set_params_best(mcmc * m, const gsl_vector * v) { gsl_vector_memcpy(m->params_best, v); }
gsl_vector * get_steps(const mcmc * m) { return m->params_step; }
double get_steps_for(const mcmc * m, const unsigned int i) { return gsl_vector_get(m->params_step, i); }
void set_steps_for(mcmc * m, const double v, const unsigned int i) { gsl_vector_set(m->params_step, i, v); }
/* reference src/mcmc_gettersetter.c:263-266 */
double get_steps_for_normalized(const mcmc * m, const unsigned int i) {
	return get_steps_for(m, i) / (get_params_max_for(m, i) - get_params_min_for(m, i));
}
/* reference src/mcmc_gettersetter.c:119-127 */
void reset_accept_rejects(mcmc * m) {
	unsigned int i;
	for (i = 0; i < get_n_par(m); i++) {
		m->params_accepts[i] = 0;
		m->params_rejects[i] = 0;
	}
	m->reject = 0;
	m->accept = 0;
}
gsl_vector * get_params_min(const mcmc * m) { return m->params_min; }
gsl_vector * get_params_max(const mcmc * m) { return m->params_max; }
double get_params_min_for(const mcmc * m, const unsigned int i) { return gsl_vector_get(m->params_min, i); }
double get_params_max_for(const mcmc * m, const unsigned int i) { return gsl_vector_get(m->params_max, i); }
const char ** get_params_descr(const mcmc * m) { return m->params_descr; }
double get_prob(const mcmc * m) { return m->prob; }
void set_prob(mcmc * m, const double v) { m->prob = v; }
double get_prior(const mcmc * m) { return m->prior; }
void set_prior(mcmc * m, const double v) { m->prior = v; }
double get_prob_best(const mcmc * m) { return m->prob_best; }
void set_prob_best(mcmc * m, const double v) { m->prob_best = v; }
const gsl_matrix * get_data(const mcmc * m) { return m->data; }
void set_data(mcmc * m, const gsl_matrix * data) { m->data = data; }
unsigned long get_params_accepts_global(const mcmc * m) { return m->accept; }
unsigned long get_params_rejects_global(const mcmc * m) { return m->reject; }
unsigned long get_params_accepts_for(const mcmc * m, const unsigned int i) { return m->params_accepts[i]; }
unsigned long get_params_rejects_for(const mcmc * m, const unsigned int i) { return m->params_rejects[i]; }

/* reference src/parallel_tempering_beta.c:25-37 */
void set_beta(mcmc * m, const double newbeta) {
	parallel_tempering_mcmc * pt = (parallel_tempering_mcmc *) m->additional_data;
	pt->beta = newbeta;
	pt->swapcount = 0;
}
double get_beta(const mcmc * m) { return ((const parallel_tempering_mcmc *) m->additional_data)->beta; }
unsigned long get_swapcount(const mcmc * m) {
	return ((const parallel_tempering_mcmc *) m->additional_data)->swapcount;
}

void dump_vector(const gsl_vector * v) {
	size_t i;
	for (i = 0; i < v->size; i++)
		printf("%f\t", gsl_vector_get(v, i));
}
void dump_vectorln(const gsl_vector * v) {
	dump_vector(v);
	printf("\n");
}

/* ---- beta ladder: reference src/parallel_tempering_beta.c:53-102.  i counts from the
 * hot end here; get_chain_beta reverses it so that chain 0 has beta = 1. ------------ */
static double unit_pos(unsigned int i, unsigned int n_beta) { return i * 1.0 / (n_beta - 1); }
static double cheb_pos(unsigned int i, unsigned int n_beta) { return (1 - cos(i * M_PI / (n_beta - 1))) / 2; }

double equidistant_beta(const unsigned int i, const unsigned int n_beta, const double beta_0) {
	return beta_0 + i * (1 - beta_0) / (n_beta - 1);
}
double equidistant_temperature(const unsigned int i, const unsigned int n_beta, const double beta_0) {
	return 1 / (1 / beta_0 + i * (1 - 1 / beta_0) / (n_beta - 1));
}
double chebyshev_temperature(const unsigned int i, const unsigned int n_beta, const double beta_0) {
	return 1 / (1 / beta_0 + (1 - 1 / beta_0) / 2 * (1 - cos(i * M_PI / (n_beta - 1))));
}
double chebyshev_beta(const unsigned int i, const unsigned int n_beta, const double beta_0) {
	return beta_0 + (1 - beta_0) / 2 * (1 - cos(i * M_PI / (n_beta - 1)));
}
double equidistant_stepwidth(const unsigned int i, const unsigned int n_beta, const double beta_0) {
	return beta_0 + pow(unit_pos(i, n_beta), 2) * (1 - beta_0);
}
double chebyshev_stepwidth(const unsigned int i, const unsigned int n_beta, const double beta_0) {
	return beta_0 + (1 - beta_0) * pow(cheb_pos(i, n_beta), 2);
}
double hot_chains(const unsigned int i, const unsigned int n_beta, const double beta_0) {
	(void) i;
	(void) n_beta;
	return beta_0;
}

double get_chain_beta(unsigned int i, unsigned int n_beta, double beta_0) {
	if (n_beta == 1)
		return 1.0;
	return BETA_ALIGNMENT(n_beta - i - 1, n_beta, beta_0);
}

/* automatic beta_0: (max_j range_j / (step_j * factor_j)) ^ -1/2, exponent as coded in the
 * reference */
double calc_beta_0(mcmc * m, gsl_vector * stepwidth_factors) {
	gsl_vector * r = dup_vector(get_params_max(m));
	double b;
	gsl_vector_sub(r, get_params_min(m));
	gsl_vector_scale(r, BETA_0_STEPWIDTH);
	gsl_vector_div(r, get_steps(m));
	gsl_vector_div(r, stepwidth_factors);
	b = pow(gsl_vector_max(r), -0.5);
	gsl_vector_free(r);
	return b;
}
