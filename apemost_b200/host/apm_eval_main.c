/*
 * apm_eval_main.c -- eval_<model>.exe: parameter vectors on stdin, "prob<TAB>prior" on
 * stdout ("%.15e"), at beta = 1 -- the reference's parity probe (apps/eval_main.c:27-68),
 * answered by the GPU model; with --host (and a linked calc_model) by the host model.
 */
#include "apm_session.h"

int main(int argc, char ** argv) {
	const int use_host = argc > 1 && strcmp(argv[1], "--host") == 0;
	apm_session * s = apm_session_open();
	mcmc * m = s->chains[0];
	const int zero = 0;
	unsigned int i = 0;
	double v;
	set_beta(m, 1.0);
	while (scanf("%lf", &v) == 1) {
		set_params_for(m, v, i++);
		if (i < get_n_par(m))
			continue;
		i = 0;
		if (use_host) {
#ifdef APM_HAVE_HOST_MODEL
			calc_model(m, NULL);
#else
			fprintf(stderr, "no host calc_model is linked into this binary\n");
			return 1;
#endif
		} else {
			apm_session_calc_model(s, &zero, 1);
		}
		printf(DUMP_FORMAT "\t" DUMP_FORMAT "\n", get_prob(m), get_prior(m));
		fflush(stdout);
	}
	apm_session_close(s);
	return 0;
}
