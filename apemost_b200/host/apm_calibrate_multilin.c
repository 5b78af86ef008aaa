/*
 * apm_calibrate_multilin.c -- the reference's regression calibrator (-DCALIBRATE_MULTILIN) over
 * the GPU engine:
 *
 *   markov_chain_calibrate_multilinear_regression   reference src/markov_chain_calibrate.c:33-237
 *   linreg_n, calc_deviation                        reference src/gsl_helper.c:189-298
 *
 * The algorithm: assess the acceptance rate at the current step widths, then at n_par more points
 * (all widths shrunk to about 1 % of the best point, one moved towards the target); from then
 * on fit rate = d + k . widths by least squares over all points so far, jump to the widths the fit
 * predicts for the target (an equal share of the gap per parameter, plus a little noise) and
 * assess again, until 100 n_par points or the iteration limit; the best point seen wins.
 * All stepping happens on the device (apm_assess_acceptance_rate -> apm_gpu_steps); the noise
 * comes from the chain's host-side random stream (apm_gpu_host_uniform), which is where the
 * reference's gsl_rng_uniform(get_random(m)) draws come from in its single global stream.
 *
 * Reference behaviour that is kept, because calibration_results must come out the same:
 *  - the first assessment varies parameter 0 only (:70), the later fitted points take full steps
 *    (parameter index -1, :207);
 *  - the "1 %" points are (1 + u) * 0.01 of the best widths, u uniform (:93-94);
 *  - the regression weights are all reset to 1 before every fit (:128), so the weights computed
 *    for each point never matter;
 *  - linreg_n centres column l by the sum of ROW l's first n_par entries divided by the number
 *    of points (gsl_helper.c:181-188 as called from :225) and the fit is solved with that;
 *  - the early exit "we already are where we wanted to jump" needs max |k / sigma| < 0
 *    (:183) and therefore never fires.
 */
#include "apm_session.h"

static double abs_double(double x) {
	return x < 0 ? -x : x;
}

/* reference src/mcmc_gettersetter.c:155-158 */
static void set_steps_for_normalized(mcmc * m, const double new_step, const unsigned int i) {
	set_steps_for(m, new_step * (get_params_max_for(m, i) - get_params_min_for(m, i)), i);
}

/* least squares of y on the columns of x (n points, m columns) the way linreg_n does it */
static gsl_vector * fit_plane(const double * x, const double * y, const double * w, unsigned int n,
		unsigned int m, double * d) {
	gsl_matrix * xx = gsl_matrix_alloc(m, m);
	gsl_vector * xy = gsl_vector_alloc(m);
	gsl_vector * k = gsl_vector_alloc(m);
	gsl_permutation * p = gsl_permutation_alloc(m);
	double * centre = (double *) malloc(m * sizeof(double));
	double ybar = 0, wsum = 0, acc;
	unsigned int i, j, l;
	int signum;

	assert(n >= m);
	for (i = 0; i < n; i++) {
		ybar += y[i] * w[i];
		wsum += w[i];
	}
	ybar = ybar / wsum;
	for (l = 0; l < m; l++) {
		acc = 0;
		for (i = 0; i < m; i++)
			acc += x[l * m + i]; /* row l, not column l */
		centre[l] = acc / n;
	}
	for (l = 0; l < m; l++) {
		for (j = 0; j < m; j++) {
			if (j < l) {
				gsl_matrix_set(xx, l, j, gsl_matrix_get(xx, j, l));
				continue;
			}
			acc = 0;
			for (i = 0; i < n; i++)
				acc += (x[i * m + l] - centre[l]) * (x[i * m + j] - centre[j]);
			gsl_matrix_set(xx, l, j, acc / n);
		}
		acc = 0;
		wsum = 0;
		for (i = 0; i < n; i++) {
			acc += (x[i * m + l] - centre[l]) * (y[i] - ybar) * w[i];
			wsum += w[i];
		}
		gsl_vector_set(xy, l, acc * n / wsum);
	}
	gsl_linalg_LU_decomp(xx, p, &signum);
	gsl_linalg_LU_solve(xx, p, xy, k);

	acc = 0;
	for (i = 0; i < n; i++)
		acc += y[i];
	*d = acc / n;
	for (j = 0; j < m; j++)
		*d -= gsl_vector_get(k, j) * centre[j];

	free(centre);
	gsl_permutation_free(p);
	gsl_vector_free(xy);
	gsl_matrix_free(xx);
	return k;
}

/* weighted rms residual of the fit (calc_deviation) */
static double fit_scatter(const double * x, const double * y, const gsl_vector * k, double d, const double * w,
		unsigned int n, unsigned int m) {
	double sumsq = 0, wsum = 0, z;
	unsigned int i, j;
	for (i = 0; i < n; i++) {
		z = d;
		for (j = 0; j < m; j++)
			z += x[i * m + j] * gsl_vector_get(k, j);
		sumsq += w[i] * pow(y[i] - z, 2);
		wsum += w[i];
	}
	return sqrt(sumsq / wsum);
}

typedef struct {
	unsigned int n, n_par, best;
	double * widths; /* [n_all][n_par], normalised step widths of every assessed point */
	double * rates;  /* [n_all] */
	double target;
} point_log;

/* store the point just assessed; it becomes the best one if its rate is nearer the target */
static void log_point(point_log * pl, const mcmc * m, double rate) {
	unsigned int j;
	for (j = 0; j < pl->n_par; j++)
		pl->widths[pl->n * pl->n_par + j] = get_steps_for_normalized(m, j);
	pl->rates[pl->n] = rate;
	if (pl->n > 0 && abs_double(rate - pl->target) < abs_double(pl->rates[pl->best] - pl->target))
		pl->best = pl->n;
	pl->n++;
}

void apm_calibrate_multilin(apm_session * s, int g, double desired_acceptance_rate, const double max_ar_deviation,
		const unsigned int iter_limit) {
	mcmc * m = s->chains[g];
	const unsigned int n_par = get_n_par(m);
	const unsigned int n_all = 100 * n_par;
	const double calc_accuracy = max_ar_deviation;
	unsigned int i, j, iter = 0;
	double rate, accuracy, next, share, d, sigma, lo, hi, v;
	double * ones = (double *) malloc(n_all * sizeof(double));
	point_log pl;
	gsl_vector * k;

	pl.n = 0;
	pl.n_par = n_par;
	pl.best = 0;
	pl.target = desired_acceptance_rate;
	pl.widths = (double *) malloc((size_t) n_all * n_par * sizeof(double));
	pl.rates = (double *) malloc(n_all * sizeof(double));
	for (i = 0; i < n_all; i++)
		ones[i] = 1.;

	/* the starting point */
	iter += apm_assess_acceptance_rate(s, g, 0, desired_acceptance_rate, 0, calc_accuracy * 5, &rate, &accuracy);
	log_point(&pl, m, rate);

	/* one more point per dimension */
	for (i = 0; i < n_par; i++) {
		const double * from = &pl.widths[pl.best * n_par];
		for (j = 0; j < n_par; j++)
			set_steps_for_normalized(m, from[j] * (1 + apm_session_uniform(s, g)) * 0.01, j);
		next = from[i] * (1 + 5 * (pl.rates[pl.best] - desired_acceptance_rate));
		if (next > 1)
			next = 1;
		if (next < from[i] * 0.01)
			next = from[i] * 0.01;
		set_steps_for_normalized(m, next, i);
		iter += apm_assess_acceptance_rate(s, g, i, desired_acceptance_rate, 0, calc_accuracy * 5, &rate,
				&accuracy);
		log_point(&pl, m, rate);
	}

	while (pl.n < n_all && iter < iter_limit) {
		printf("                                        %d iterations\n", iter);
		k = fit_plane(pl.widths, pl.rates, ones, pl.n, n_par, &d);
		sigma = fit_scatter(pl.widths, pl.rates, k, d, ones, pl.n, n_par);
		printf("d = %g, sigma = %g, k = \n", d, sigma);
		gsl_vector_fprintf(stdout, k, "%g");

		if (sigma < calc_accuracy / 3)
			sigma = calc_accuracy / 3;
		else
			sigma = sigma / 3;

		/* every parameter is to close an equal share of the gap to the target */
		share = (desired_acceptance_rate - d) / n_par;
		for (i = 0; i < n_par; i++) {
			if (gsl_vector_get(k, i) > 0)
				gsl_vector_set(k, i, -5);
			next = share / gsl_vector_get(k, i) + apm_session_uniform(s, g) * sigma / 3;

			lo = hi = pl.widths[i];
			for (j = 1; j < pl.n; j++) {
				v = pl.widths[j * n_par + i];
				if (v < lo)
					lo = v;
				if (v > hi)
					hi = v;
			}
			if (next < lo * 0.01)
				next = lo * 0.01;
			if (next > hi * 100)
				next = hi * 100 + log(next - hi * 100) + 1;
			set_steps_for_normalized(m, next, i);
		}
		gsl_vector_free(k);

		iter += apm_assess_acceptance_rate(s, g, (unsigned int) -1, desired_acceptance_rate, 0, calc_accuracy / 4,
				&rate, &accuracy);
		log_point(&pl, m, rate);
	}

	for (i = 0; i < n_par; i++)
		set_steps_for_normalized(m, pl.widths[pl.best * n_par + i], i);
	free(pl.widths);
	free(pl.rates);
	free(ones);
	/* the step widths chosen last live in the host struct only */
	apm_session_push(s, g, 1);
}
