/*
 * apm_calibrate_alt.c -- the reference's alternate step-width calibrator (-DCALIBRATE_ALTERNATE)
 * over the GPU engine (markov_chain_calibrate_alt, reference src/markov_chain_calibrate.c:916-1037).
 *
 * The scheme: sweep over the parameters; measure each one's acceptance rate at its current width
 * (apm_assess.c, i.e. apm_gpu_steps on the device) and move the width by a fraction of the rate's
 * distance from the target; ask for more accurate measurements from sweep to sweep (at most
 * MAX_ACCURACY_IMPROVEMENT times better than the last sweep's average), skipping parameters that
 * are already measured ten times better than that; stop when the largest distance seen in a
 * sweep and the summed accuracies are both small.
 *
 * Kept from the reference, since calibration_results and calibration_progress.data must match:
 *  - the fraction is best * SCALE_LIN_WORST + SCALE_MIN where `best` starts at 1 and is replaced by
 *    a sweep's SUMMED accuracy whenever that sweep's AVERAGE accuracy beats it (:1014-1016);
 *  - a move below -1 becomes -0.9 (:1002-1003);
 *  - the give-up test is iterations > iter_limit * n_par, answered with exit(1) (:1010-1013).
 */
#include "apm_session.h"

#ifndef MAX_ACCURACY_IMPROVEMENT
#define MAX_ACCURACY_IMPROVEMENT 2.8   /* reference src/markov_chain_calibrate.c:916-925 */
#endif
#ifndef SCALE_LIN_WORST
#define SCALE_LIN_WORST 5
#endif
#ifndef SCALE_MIN
#define SCALE_MIN 0.4
#endif

typedef struct {
	apm_session * s;
	int g;
	mcmc * m;
	unsigned int n_par, steps_used;
	double target;
	double * known_to;     /* [n_par] accuracy of each parameter's last measurement */
	double sweep_average;  /* average accuracy of the previous sweep (0 before the first) */
	double best;           /* see the header comment */
	FILE * progress;
} alt_state;

/* one sweep; returns the largest |rate - target| met and leaves the summed accuracy in *summed */
static double alt_sweep(alt_state * a, double * summed) {
	const double ask_for = a->sweep_average / MAX_ACCURACY_IMPROVEMENT;
	double farthest = 0;
	unsigned int i;
	printf("calculating for up to %f accuracy\n", ask_for);
	*summed = 0;
	for (i = 0; i < a->n_par; i++) {
		double rate, accuracy, off, move;
		if (a->known_to[i] < 0.1 * a->sweep_average)
			continue;
		a->steps_used += apm_assess_acceptance_rate(a->s, a->g, i, a->target, ask_for, 1 /* no upper limit */,
				&rate, &accuracy);
		printf("%d: a/r: %f (+-%f); desired: %f; steps: %f\n", i, rate, accuracy, a->target,
				get_steps_for_normalized(a->m, i));
		if (a->progress != NULL)
			fprintf(a->progress, "%d\t%d\t%f\t%f\t%f\n", i + 1, a->steps_used, get_steps_for_normalized(a->m, i),
					rate, accuracy);
		*summed += accuracy;
		a->known_to[i] = accuracy;

		off = rate - a->target;
		move = off * (a->best * SCALE_LIN_WORST + SCALE_MIN);
		if (move < -1)
			move = -0.9;
		if (farthest < (off < 0 ? -off : off))
			farthest = off < 0 ? -off : off;
		set_steps_for(a->m, get_steps_for(a->m, i) * (1 + move), i);
		printf("%d: new steps: %f\n", i, get_steps_for_normalized(a->m, i));
	}
	return farthest;
}

void apm_calibrate_alt(apm_session * s, int g, double desired_acceptance_rate, const double max_ar_deviation,
		const unsigned int iter_limit) {
	alt_state a;
	unsigned int i;
	a.s = s;
	a.g = g;
	a.m = s->chains[g];
	a.n_par = get_n_par(a.m);
	a.steps_used = 0;
	a.target = desired_acceptance_rate;
	a.known_to = (double *) calloc(a.n_par, sizeof(double));
	a.sweep_average = 0;
	a.best = 1;
	a.progress = fopen(apm_out_path("calibration_progress.data"), "w");
	assert(a.known_to != NULL);
	for (i = 0; i < a.n_par; i++)
		a.known_to[i] = 0;

	for (;;) {
		double summed, farthest = alt_sweep(&a, &summed);
		if (a.steps_used > iter_limit * a.n_par) {
			fprintf(stderr, "calibration failed: iteration limit reached\n");
			exit(1);
		}
		a.sweep_average = summed / a.n_par;
		if (a.sweep_average < a.best)
			a.best = summed;
		printf("max deviation: %f; ", farthest);
		if (farthest < max_ar_deviation && summed < max_ar_deviation * 2) {
			printf("small deviation: %f; quitting\n", farthest);
			break;
		}
	}
	if (a.progress != NULL)
		fclose(a.progress);
	free(a.known_to);
	/* the widths chosen last live in the host struct only */
	apm_session_push(s, g, 1);
}
