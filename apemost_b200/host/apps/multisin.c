/*
 * multisin.c -- an example model file in APEMoST's plugin style (reference
 * doc/manual.rst:126-168, contract src/mcmc.h:164,173): K superposed sines plus an offset,
 *
 *     y(x) = offset + sum_k A_k * sin(2 pi (f_k x + phi_k)),      K = (n_par - 1) / 3
 *
 * with the params file listing A_1 f_1 phi_1 ... A_K f_K phi_K offset (phases in cycles).
 * K = 1 is exactly the reference's apps/simplesin.c; K = 5 is the "5 sines" model BASELINE.json
 * names, which the reference tree does not contain -- it is a NEW model.
 *
 * This file is the host half (used by `check` and `eval_multisin.exe --host`); multisin.cuh is
 * the __device__ counterpart the sampler runs.
 */
#include <gsl/gsl_sf.h>

#include "mcmc.h"
#include "parallel_tempering.h"
#include "debug.h"

#ifndef SIGMA
#define SIGMA 0.5
#endif

void calc_model(mcmc * m, const gsl_vector * old_values) {
	const unsigned int n_par = get_n_par(m);
	const unsigned int n_sines = (n_par - 1) / 3;
	const double offset = get_params_for(m, n_par - 1);
	double square_sum = 0;
	unsigned int i, k;

	(void) old_values;
	assert(n_par == 3 * n_sines + 1);
	for (i = 0; i < m->data->size1; i++) {
		const double x = gsl_matrix_get(m->data, i, 0);
		double y = 0, deltay;
		for (k = 0; k < n_sines; k++) {
			const double amplitude = get_params_for(m, 3 * k);
			const double frequency = get_params_for(m, 3 * k + 1);
			const double phase = get_params_for(m, 3 * k + 2);
			y += amplitude * gsl_sf_sin(2.0 * M_PI * (frequency * x + phase));
		}
		deltay = y + offset - gsl_matrix_get(m->data, i, 1);
		square_sum += deltay * deltay;
	}
	set_prob(m, get_beta(m) * square_sum / (-2 * SIGMA * SIGMA));
}

void calc_model_for(mcmc * m, const unsigned int i, const double old_value) {
	(void) i;
	(void) old_value;
	calc_model(m, NULL);
}
