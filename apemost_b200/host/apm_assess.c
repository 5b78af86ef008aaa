/*
 * apm_assess.c -- "what is the acceptance rate at these step widths, and how well do we know it?"
 * over the GPU engine: the measurement every alternate calibrator of the reference is built on
 * (assess_acceptance_rate, reference src/markov_chain.c:100-224).
 *
 * The measurement: take steps of one kind (one parameter, or all of them) in growing batches and
 * keep the log of which were accepted.  The rate is accepts / steps; its accuracy is judged by the
 * largest excursion of the running accept count from the straight line of that slope, over the
 * number of steps.  The closer the rate is to the wanted one the more accuracy is asked for
 * (ACCURACY_DEVIATION_FACTOR of the distance, never better than 0.5 %, clamped to the caller's
 * window), and the batch grows to the length at which that excursion would be small enough.
 * The steps themselves -- the only expensive part -- are the engine's apm_gpu_steps.
 *
 * Two things are the reference's and are kept because calibration_results must come out the same:
 *  - the rate counts the accepts BEFORE the batch's last step (the counter is sampled at the top
 *    of the step loop, :146,161), still divided by the full length;
 *  - the excursion is truncated to an integer before its absolute value is taken (abs(), not
 *    fabs(), :182-183) and starts at 1.
 */
#include "apm_session.h"

#ifndef ACCURACY_DEVIATION_FACTOR
#define ACCURACY_DEVIATION_FACTOR 0.25 /* reference src/markov_chain.c:91-98 */
#endif

/* which of the steps taken so far were accepted */
typedef struct {
	unsigned char * hit;
	unsigned int len;
} accept_log;

/* take the steps that bring the log to `upto` entries: apm_gpu_steps for chain g alone */
static void log_extend(apm_session * s, int g, int kind, const unsigned char * only_g, accept_log * lg,
		unsigned int upto) {
	const unsigned int more = upto - lg->len;
	unsigned char * rows = (unsigned char *) malloc((size_t) more * s->n_chains);
	unsigned int j;
	lg->hit = (unsigned char *) realloc(lg->hit, upto);
	assert(rows != NULL && lg->hit != NULL);
	apm_gpu_check(s, apm_gpu_steps(s->gpu, only_g, kind, (long long) more, rows), "stepping");
	for (j = 0; j < more; j++)
		lg->hit[lg->len + j] = rows[(size_t) j * s->n_chains + g];
	lg->len = upto;
	free(rows);
}

static double rate_before_last_step(const accept_log * lg) {
	unsigned long hits = 0;
	unsigned int j;
	for (j = 0; j + 1 < lg->len; j++)
		hits += lg->hit[j];
	return hits / (double) lg->len;
}

static unsigned int largest_excursion(const accept_log * lg, double rate) {
	unsigned long hits = 0;
	unsigned int worst = 1, j;
	for (j = 0; j < lg->len; j++) {
		int off;
		hits += lg->hit[j] != 0;
		off = (int) (hits - rate * j);
		if (off < 0)
			off = -off;
		if ((unsigned int) off > worst)
			worst = (unsigned int) off;
	}
	return worst;
}

static double wanted_accuracy(double rate, double target, double at_best, double at_worst) {
	double want = (rate < target ? target - rate : rate - target) * ACCURACY_DEVIATION_FACTOR;
	if (want < 0.005)
		want = 0.005;
	if (want < at_best)
		want = at_best;
	if (want > at_worst)
		want = at_worst;
	return want;
}

unsigned int apm_assess_acceptance_rate(apm_session * s, int g, unsigned int param,
		double desired_acceptance_rate, double min_accuracy, double max_accuracy,
		double * acceptance_rate, double * accuracy) {
	mcmc * m = s->chains[g];
	const unsigned int n_par = get_n_par(m);
	const int kind = param < n_par ? (int) param : (int) n_par; /* anything else: full steps */
	unsigned char * only_g = (unsigned char *) calloc(s->n_chains, 1);
	accept_log lg = { NULL, 0 };
	unsigned int length = 40;

	assert(only_g != NULL);
	only_g[g] = 1;
	reset_accept_rejects(m);
	apm_session_push(s, g, 1); /* the widths under test and the zeroed counters */
	for (;;) {
		unsigned int off;
		double want;
		log_extend(s, g, kind, only_g, &lg, length);
		*acceptance_rate = rate_before_last_step(&lg);
		off = largest_excursion(&lg, *acceptance_rate);
		*accuracy = off / 1. / length;
		want = wanted_accuracy(*acceptance_rate, desired_acceptance_rate, min_accuracy, max_accuracy);
		if (*accuracy <= want)
			break;
		assert(off / want >= length);
		length = ((unsigned int) ((off / 1. / want) / 8) + 1) * 8;
	}
	apm_session_pull(s, g, 1);
	free(lg.hit);
	free(only_g);
	return length;
}
