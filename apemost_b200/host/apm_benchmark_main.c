/*
 * apm_benchmark_main.c -- benchmark_<model>.exe <npartialcalc> <ncalc>: evaluates the model at the
 * params file's start values the given number of times so that the model calculation can be
 * timed, and checks that every evaluation reproduces the first one -- the reference's
 * apps/benchmark_main.c:27-79 (there: n calls of calc_model_for, p calls of calc_model, an
 * assert that prob never changes; timing is left to gprof / time(1)).
 *
 * On the GPU an evaluation is one entry of a batched apm_gpu_eval: the n + p evaluations are
 * issued in batches of APM_BENCH_BATCH (default 4096) identical parameter vectors, so the same
 * point is evaluated at every position of the kernel's chain tiles.  Every result must equal
 * the first bit for bit (the likelihood kernel sums in a fixed order), otherwise exit(1) with
 * the reference's "original prob / new prob" lines.  With a linked host calc_model
 * (APM_HAVE_HOST_MODEL) the device value is also compared with the host's to 1e-12 relative.
 * New: the elapsed time and the evaluations (and data rows) per second are printed.
 */
#include <time.h>
#include "apm_session.h"

#ifndef APM_BENCH_BATCH
#define APM_BENCH_BATCH 4096
#endif

static double now_s(void) {
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char ** argv) {
	apm_session * s;
	mcmc * m;
	long long total, done = 0;
	double * p, * b, * prob, * prior, prob0, t0, dt;
	int i, j, batch, n_par, first = 1;
	long long kernel_launches = 0;

	if (argc != 2 + 1) {
		fprintf(stderr, "SYNOPSIS: %s <npartialcalc> <ncalc>\n"
				"\n"
				"This program calculates the model the given number of times\n"
				"to allow benchmarking of the model calculations.\n"
				"\n"
				"\tnpartialcalc\tnumber of calls to calc_model_for\n"
				"\tncalc\tnumber of calls to calc_model\n"
				"\n"
				"On the GPU both kinds are evaluations of the full model, issued in batches of %d.\n",
				argv[0], (int) APM_BENCH_BATCH);
		exit(1);
	}
	assert(atoll(argv[1]) >= 0);
	assert(atoll(argv[2]) >= 0);
	total = atoll(argv[1]) + atoll(argv[2]);

	s = apm_session_open();
	m = s->chains[0];
	n_par = s->n_par;
	set_beta(m, 1.0);
	batch = (int) (total < APM_BENCH_BATCH ? (total > 0 ? total : 1) : APM_BENCH_BATCH);
	p = (double *) calloc((size_t) batch * n_par, sizeof(double));
	b = (double *) calloc(batch, sizeof(double));
	prob = (double *) calloc(batch, sizeof(double));
	prior = (double *) calloc(batch, sizeof(double));
	for (i = 0; i < batch; i++) {
		b[i] = 1.0;
		for (j = 0; j < n_par; j++)
			p[(size_t) i * n_par + j] = gsl_vector_get(get_params(m), j);
	}
	/* calc_model(m, NULL); prob = get_prob(m);  (reference :62-63) */
	apm_gpu_check(s, apm_gpu_eval(s->gpu, 1, p, b, prob, prior), "evaluating the model");
	prob0 = prob[0];
#ifdef APM_HAVE_HOST_MODEL
	calc_model(m, NULL);
	if (fabs(get_prob(m) - prob0) > 1e-12 * fabs(prob0)) {
		fprintf(stderr, "device model %.17g differs from host calc_model %.17g\n", prob0, get_prob(m));
		exit(1);
	}
#endif
	t0 = now_s();
	while (done < total) {
		const int n = (int) (total - done < batch ? total - done : batch);
		apm_gpu_check(s, apm_gpu_eval(s->gpu, n, p, b, prob, prior), "evaluating the model");
		kernel_launches++;
		for (i = 0; i < n; i++) {
			if (prob[i] != prob0) {
				if (first) {
					printf("original prob: %f\n", prob0);
					printf("new prob: %f\n", prob[i]);
					first = 0;
				}
				fprintf(stderr, "evaluation %lld does not reproduce the first one (%.17g != %.17g)\n",
						done + i, prob[i], prob0);
				exit(1);
			}
		}
		done += n;
	}
	dt = now_s() - t0;
	printf("%lld model evaluations in %.6f s (%lld batches): %.6g evaluations/s, %.6g data rows/s; prob = "
			DUMP_FORMAT "\n", total, dt, kernel_launches, dt > 0 ? total / dt : 0.0,
			dt > 0 ? total / dt * (double) m->data->size1 : 0.0, prob0);
	free(p);
	free(b);
	free(prob);
	free(prior);
	apm_session_close(s);
	return 0;
}
