/*
 * apm_calibrate_alt.c -- the reference's alternate step-width calibrator (-DCALIBRATE_ALTERNATE)
 * over the GPU engine:
 *
 *   assess_acceptance_rate        reference src/markov_chain.c:117-224
 *   markov_chain_calibrate_alt    reference src/markov_chain_calibrate.c:916-1037
 *
 * Both are sequential, data-dependent host algorithms around one inner loop -- "n more steps of
 * this kind, remember which were accepted" -- and that loop is what the engine provides
 * (apm_gpu_steps).  The host keeps the reference's control flow and arithmetic, including its
 * quirks: the acceptance rate is taken from the counter as it stood BEFORE the last step
 * (:146,161 assign `accepts` at the top of the loop body), and the deviation statistics go
 * through the integer abs() (:182-183).
 *
 * CALIBRATE_MULTILIN (apm_calibrate_multilin.c) sits on the same assessment plus a linear
 * regression; CALIBRATE_QUADRATIC is not built.
 */
#include "apm_session.h"

#ifndef ACCURACY_DEVIATION_FACTOR
#define ACCURACY_DEVIATION_FACTOR 0.25 /* reference src/markov_chain.c:91-98 */
#endif
#ifndef MAX_ACCURACY_IMPROVEMENT
#define MAX_ACCURACY_IMPROVEMENT 2.8   /* reference src/markov_chain_calibrate.c:916-925 */
#endif
#ifndef SCALE_LIN_WORST
#define SCALE_LIN_WORST 5
#endif
#ifndef SCALE_MIN
#define SCALE_MIN 0.4
#endif

static double abs_double(double x) {
	return x < 0 ? -x : x;
}

unsigned int apm_assess_acceptance_rate(apm_session * s, int g, unsigned int param,
		double desired_acceptance_rate, double min_accuracy, double max_accuracy,
		double * acceptance_rate, double * accuracy) {
	mcmc * m = s->chains[g];
	const unsigned int n_par = get_n_par(m);
	const int kind = param < n_par ? (int) param : (int) n_par;
	unsigned int i = 0, j, n = 40, maxdev;
	unsigned long accepts;
	double stdev, accept_rate, required_accuracy;
	unsigned char * acceptslog = NULL, * chunk = NULL;
	unsigned char * select = (unsigned char *) calloc(s->n_chains, 1);
	select[g] = 1;

	reset_accept_rejects(m);
	apm_session_push(s, g, 1);
	while (1) {
		acceptslog = (unsigned char *) realloc(acceptslog, n);
		chunk = (unsigned char *) realloc(chunk, (size_t) (n - i) * s->n_chains);
		assert(acceptslog != NULL && chunk != NULL);
		/* for (; i < n; i++) { markov_chain_step[_for]; mcmc_check_best; log the outcome } */
		apm_gpu_check(s, apm_gpu_steps(s->gpu, select, kind, (long long) (n - i), chunk), "stepping");
		for (j = i; j < n; j++)
			acceptslog[j] = chunk[(size_t) (j - i) * s->n_chains + g];
		i = n;
		/* `accepts` holds the counter read before the last step */
		accepts = 0;
		for (j = 0; j + 1 < n; j++)
			accepts += acceptslog[j];
		accept_rate = accepts / (double) n;

		/* get max deviation */
		accepts = 0;
		stdev = 0;
		maxdev = 0 + 1;
		for (j = 0; j < n; j++) {
			int dev;
			if (acceptslog[j] != 0)
				accepts++;
			stdev += pow(accepts - accept_rate * j, 2);
			dev = (int) (accepts - accept_rate * j); /* abs() takes an int */
			if (dev < 0)
				dev = -dev;
			if ((unsigned int) dev > maxdev)
				maxdev = (unsigned int) dev;
		}
		stdev = sqrt(stdev / n) * 2;
		(void) stdev;

		required_accuracy = abs_double(accept_rate - desired_acceptance_rate) * ACCURACY_DEVIATION_FACTOR;
		if (required_accuracy < 0.005)
			required_accuracy = 0.005;
		if (required_accuracy < min_accuracy)
			required_accuracy = min_accuracy;
		if (required_accuracy > max_accuracy)
			required_accuracy = max_accuracy;

		*acceptance_rate = accept_rate;
		*accuracy = maxdev / 1. / n;
		if (*accuracy <= required_accuracy)
			break;
		assert(maxdev / required_accuracy >= n);
		n = ((unsigned int) ((maxdev / 1. / required_accuracy) / 8) + 1) * 8;
	}
	apm_session_pull(s, g, 1);
	free(acceptslog);
	free(chunk);
	free(select);
	return n;
}

void apm_calibrate_alt(apm_session * s, int g, double desired_acceptance_rate, const double max_ar_deviation,
		const unsigned int iter_limit) {
	mcmc * m = s->chains[g];
	unsigned int i, j;
	double current_acceptance_rate, accuracy;
	const unsigned int n_par = get_n_par(m);
	double scale, movedirection, move, max_deviation;
	double worst_accuracy = 0, worst_accuracy_previous = 0, best_worst_accuracy = 1;
	unsigned int iter = 0;
	FILE * progress_plot_file = fopen(apm_out_path("calibration_progress.data"), "w");
	gsl_vector * accuracies = gsl_vector_alloc(n_par);
	gsl_vector_set_all(accuracies, 0);

	/* a point in step-width space is assessed parameter by parameter; each step width moves in
	 * proportion to how far its acceptance rate is from the target, less and less once the
	 * assessments have settled down */
	while (1) {
		max_deviation = 0;
		for (j = 0; j < 1; j++) {
			printf("calculating for up to %f accuracy\n", worst_accuracy_previous / MAX_ACCURACY_IMPROVEMENT);
			worst_accuracy = 0;
			for (i = 0; i < n_par; i++) {
				if (gsl_vector_get(accuracies, i) < 0.1 * worst_accuracy_previous)
					continue;
				iter += apm_assess_acceptance_rate(s, g, i, desired_acceptance_rate,
						worst_accuracy_previous / MAX_ACCURACY_IMPROVEMENT, 1 /* no restriction */,
						&current_acceptance_rate, &accuracy);
				printf("%d: a/r: %f (+-%f); desired: %f; steps: %f\n", i, current_acceptance_rate, accuracy,
						desired_acceptance_rate, get_steps_for_normalized(m, i));
				if (progress_plot_file != NULL)
					fprintf(progress_plot_file, "%d\t%d\t%f\t%f\t%f\n", i + 1, iter,
							get_steps_for_normalized(m, i), current_acceptance_rate, accuracy);
				worst_accuracy += accuracy;
				gsl_vector_set(accuracies, i, accuracy);

				movedirection = current_acceptance_rate - desired_acceptance_rate;
				scale = best_worst_accuracy * SCALE_LIN_WORST + SCALE_MIN;
				assert(scale > 0);
				move = movedirection * scale;
				if (move < -1)
					move = -0.9;
				if (max_deviation < abs_double(movedirection))
					max_deviation = abs_double(movedirection);
				set_steps_for(m, get_steps_for(m, i) * (1 + move), i);
				printf("%d: new steps: %f\n", i, get_steps_for_normalized(m, i));
			}
			if (iter > iter_limit * n_par) {
				fprintf(stderr, "calibration failed: iteration limit reached\n");
				exit(1);
			}
			worst_accuracy_previous = worst_accuracy / n_par;
			if (worst_accuracy_previous < best_worst_accuracy)
				best_worst_accuracy = worst_accuracy;
		}
		printf("max deviation: %f; ", max_deviation);
		if (max_deviation < max_ar_deviation && worst_accuracy < max_ar_deviation * 2) {
			printf("small deviation: %f; quitting\n", max_deviation);
			break;
		}
	}
	if (progress_plot_file != NULL)
		fclose(progress_plot_file);
	gsl_vector_free(accuracies);
	/* the step widths chosen last live in the host struct only */
	apm_session_push(s, g, 1);
}
