/*
 * apm_phases.c -- the phases calibrate_first / calibrate_rest / run over the GPU engine.
 *
 * Each phase follows the reference's order of operations (cited per function) with the
 * three CPU hot loops replaced by engine calls:
 *     calc_model(chains[i], NULL)        -> apm_gpu_eval
 *     markov_chain_calibrate(chains[i])  -> apm_gpu_calibrate (all selected chains at once)
 *     run_sampler()'s step/swap loop     -> apm_gpu_run + apm_gpu_read_trace
 * Everything read from or written to disk goes through the chain structs, so the files are
 * produced by the same formatting code paths as in the reference.
 */
#include <pthread.h>
#include <signal.h>
#include <time.h>
#include "apm_session.h"

/* ------------------------------------------------------------------ configuration */
static unsigned circular_mask(void) {
	/* CIRCULAR_PARAMS is a 1-based list (reference src/markov_chain.h:34-46) */
	const unsigned int list[] = { CIRCULAR_PARAMS, 0 };
	unsigned mask = 0;
	unsigned int j;
	for (j = 0; j < sizeof(list) / sizeof(list[0]); j++)
		if (list[j] >= 1 && list[j] <= 32)
			mask |= 1u << (list[j] - 1);
	return mask;
}

/* Several processes (one per GPU) can share one working directory: the N_ENSEMBLES ensembles are
 * dealt out in contiguous blocks by rank, every ensemble keeps its global number (its ens<e>/
 * directory and its random streams), so the files do not depend on how many processes ran.
 * Rank and size come from APM_RANK / APM_WORLD_SIZE or, under torchrun, RANK / WORLD_SIZE;
 * the device from APM_DEVICE or LOCAL_RANK.  Independent ensembles need no communication. */
static int env_int(const char * a, const char * b, int fallback) {
	const char * v = getenv(a);
	if (v == NULL && b != NULL)
		v = getenv(b);
	return v != NULL ? atoi(v) : fallback;
}

static void ensemble_share(int * first, int * count) {
	const int world = env_int("APM_WORLD_SIZE", "WORLD_SIZE", 1), rank = env_int("APM_RANK", "RANK", 0);
	const int n = N_ENSEMBLES;
	if (world < 1 || rank < 0 || rank >= world) {
		fprintf(stderr, "bad rank %d / world size %d\n", rank, world);
		exit(1);
	}
	*first = (int) ((long long) n * rank / world);
	*count = (int) ((long long) n * (rank + 1) / world) - *first;
}

static void engine_config(apm_gpu_config * cfg, int ens_first, int n_ens, int n_beta, int n_par) {
	const char * seed = getenv("GSL_RNG_SEED"); /* the reference's seed variable (src/mcmc.c:27-35) */
	memset(cfg, 0, sizeof(*cfg));
	cfg->device = env_int("APM_DEVICE", "LOCAL_RANK", 0);
	cfg->ensemble_id_offset = ens_first;
	cfg->chain_id_offset = ens_first * n_beta;
	cfg->model_id = APM_MODEL_ID;
	cfg->n_ensembles = n_ens;
	cfg->n_beta = n_beta;
	cfg->n_par = n_par;
	cfg->seed = seed ? strtoull(seed, NULL, 10) : 0;
#if defined(PROPOSAL_UNIFORM)
	cfg->proposal = APM_PROPOSAL_UNIFORM;
#elif defined(PROPOSAL_LOGISTIC)
	cfg->proposal = APM_PROPOSAL_LOGISTIC;
#else
	cfg->proposal = APM_PROPOSAL_GAUSSIAN;
#endif
	cfg->circular_mask = circular_mask();
#ifdef APM_EXACT_SWAP
	cfg->quirks = 0;                    /* statistically exact swap and revert */
#else
	cfg->quirks = APM_QUIRKS_REFERENCE; /* behave exactly like the reference */
#endif
	cfg->path = APM_PATH_AUTO;
#if defined(SIGMA)
	cfg->model_const[0] = SIGMA;        /* apps/simplesin*.c */
#elif defined(HMIN)
	cfg->model_const[0] = HMIN;         /* apps/pulse*.c */
#endif
}

static void calib_config(apm_gpu_calib_cfg * c, int skip) {
	memset(c, 0, sizeof(*c));
	c->burn_in_iterations = BURN_IN_ITERATIONS;
	c->desired_acceptance_rate = TARGET_ACCEPTANCE_RATE;
	c->max_ar_deviation = MAX_AR_DEVIATION;
	c->iter_limit = ITER_LIMIT;
	c->mul = MUL;
	c->adjust_step = DEFAULT_ADJUST_STEP;
	c->skip_calibrate = skip;
	c->iter_readjust = ITER_READJUST;
	c->no_rescaling_limit = NO_RESCALING_LIMIT;
}

/* ------------------------------------------------------------------ session */
void apm_gpu_check(apm_session * s, int rc, const char * what) {
	if (rc == APM_OK)
		return;
	fprintf(stderr, "%s failed: %s\n", what, apm_gpu_last_error(s ? s->gpu : NULL));
	exit(1);
}

mcmc ** apm_ensemble(apm_session * s, int e) {
	return s->chains + (size_t) e * s->n_beta;
}

static void attach_pt_data(mcmc * m) {
	m->additional_data = calloc(1, sizeof(parallel_tempering_mcmc));
	assert(m->additional_data != NULL);
	set_beta(m, 1.0);
}

/* N_BETA chains from the params file sharing chain 0's data table
 * (reference src/parallel_tempering_config.c:95-123) */
mcmc ** setup_chains(void) {
	const unsigned int n_beta = N_BETA;
	mcmc ** chains = (mcmc **) calloc(n_beta, sizeof(mcmc *));
	unsigned int i;
	assert(chains != NULL);
	printf("Initializing %d chains ...\n", n_beta);
	for (i = 0; i < n_beta; i++) {
		chains[i] = mcmc_load_params(PARAMS_FILENAME);
		if (i == 0)
			mcmc_load_data(chains[i], DATA_FILENAME);
		else
			mcmc_reuse_data(chains[i], chains[0]);
		attach_pt_data(chains[i]);
		mcmc_check(chains[i]);
	}
	return chains;
}

apm_session * apm_session_open(void) {
	apm_session * s = (apm_session *) calloc(1, sizeof(*s));
	mcmc ** first;
	apm_gpu_config cfg;
	const gsl_matrix * d;
	size_t n, nv;
	int g;
	assert(s != NULL);
	ensemble_share(&s->ens_first, &s->n_ens);
	if (s->n_ens < 1) {
		printf("no ensemble for this process (N_ENSEMBLES = %d)\n", (int) N_ENSEMBLES);
		exit(0);
	}
	s->n_beta = N_BETA;
	s->n_chains = s->n_ens * s->n_beta;
	s->chains = (mcmc **) calloc(s->n_chains, sizeof(mcmc *));
	first = setup_chains();
	memcpy(s->chains, first, s->n_beta * sizeof(mcmc *));
	free(first);
	for (g = s->n_beta; g < s->n_chains; g++) { /* further ensembles: same params, same table */
		s->chains[g] = mcmc_load_params(PARAMS_FILENAME);
		mcmc_reuse_data(s->chains[g], s->chains[0]);
		attach_pt_data(s->chains[g]);
	}
	s->n_par = get_n_par(s->chains[0]);
	n = s->n_chains;
	nv = n * s->n_par;
	s->beta = (double *) calloc(n, sizeof(double));
	s->prob = (double *) calloc(n, sizeof(double));
	s->prior = (double *) calloc(n, sizeof(double));
	s->prob_best = (double *) calloc(n, sizeof(double));
	s->params = (double *) calloc(nv, sizeof(double));
	s->steps = (double *) calloc(nv, sizeof(double));
	s->params_best = (double *) calloc(nv, sizeof(double));
	s->accept = (unsigned long long *) calloc(n, sizeof(unsigned long long));
	s->reject = (unsigned long long *) calloc(n, sizeof(unsigned long long));
	s->n_iter = (unsigned long long *) calloc(n, sizeof(unsigned long long));
	s->swapcount = (unsigned long long *) calloc(n, sizeof(unsigned long long));
	s->pacc = (unsigned long long *) calloc(nv, sizeof(unsigned long long));
	s->prej = (unsigned long long *) calloc(nv, sizeof(unsigned long long));

	engine_config(&cfg, s->ens_first, s->n_ens, s->n_beta, s->n_par);
	if (apm_gpu_create(&s->gpu, &cfg) != APM_OK) {
		fprintf(stderr, "could not start the GPU engine: %s\n", apm_gpu_last_error(NULL));
		exit(1);
	}
#ifdef RWM
#error "RWM does not compile in the reference either (src/parallel_tempering.c:278 calls markov_chain_step with 2 arguments)"
#endif
#ifdef ADAPT
	/* adapt() once per round before the swap (reference src/parallel_tempering.c:282-302,404) */
	apm_gpu_check(s, apm_gpu_set_adapt(s->gpu, 1, (double) TARGET_ACCEPTANCE_RATE), "enabling ADAPT");
#endif
	/* RANDOMSWAP (reference src/parallel_tempering_interaction.c:47-65,130-131) needs nothing: called
	 * with n_swap = 1 its extra uniform is always below 1.0 / n_swap, so it is the default swap rule
	 * plus one discarded draw -- the same distribution under the engine's counter RNG */
	d = s->chains[0]->data;
	assert(d->tda == d->size2);
	apm_gpu_check(s, apm_gpu_set_data(s->gpu, d->data, (long long) d->size1, (int) d->size2), "uploading the data table");
	apm_gpu_check(s, apm_gpu_set_bounds(s->gpu, get_params_min(s->chains[0])->data,
			get_params_max(s->chains[0])->data), "setting the parameter bounds");
	apm_session_push(s, 0, s->n_chains);
	return s;
}

void apm_session_close(apm_session * s) {
	int g;
	apm_gpu_destroy(s->gpu);
	for (g = s->n_chains - 1; g >= 0; g--) {
		free(s->chains[g]->additional_data);
		if (g != 0)
			set_data(s->chains[g], NULL); /* borrowed from chain 0 */
		free(mcmc_free(s->chains[g]));
	}
	free(s->chains);
	free(s->beta); free(s->prob); free(s->prior); free(s->prob_best); free(s->params); free(s->steps);
	free(s->params_best); free(s->accept); free(s->reject); free(s->n_iter); free(s->swapcount);
	free(s->pacc); free(s->prej);
	free(s);
}

static void chain_io(apm_session * s, int first, apm_gpu_chain_io * io) {
	const size_t o = first, ov = (size_t) first * s->n_par;
	memset(io, 0, sizeof(*io));
	io->beta = s->beta + o;
	io->params = s->params + ov;
	io->steps = s->steps + ov;
	io->prob = s->prob + o;
	io->prior = s->prior + o;
	io->prob_best = s->prob_best + o;
	io->params_best = s->params_best + ov;
	io->accept = s->accept + o;
	io->reject = s->reject + o;
	io->params_accepts = s->pacc + ov;
	io->params_rejects = s->prej + ov;
	io->n_iter = s->n_iter + o;
	io->swapcount = s->swapcount + o;
}

void apm_session_push(apm_session * s, int first, int count) {
	apm_gpu_chain_io io;
	int g, j;
	for (g = first; g < first + count; g++) {
		const mcmc * m = s->chains[g];
		s->beta[g] = get_beta(m);
		s->prob[g] = m->prob;
		s->prior[g] = m->prior;
		s->prob_best[g] = m->prob_best;
		s->accept[g] = m->accept;
		s->reject[g] = m->reject;
		s->n_iter[g] = m->n_iter;
		s->swapcount[g] = get_swapcount(m);
		for (j = 0; j < s->n_par; j++) {
			const size_t k = (size_t) g * s->n_par + j;
			s->params[k] = gsl_vector_get(m->params, j);
			s->steps[k] = gsl_vector_get(m->params_step, j);
			s->params_best[k] = gsl_vector_get(m->params_best, j);
			s->pacc[k] = m->params_accepts[j];
			s->prej[k] = m->params_rejects[j];
		}
	}
	chain_io(s, first, &io);
	apm_gpu_check(s, apm_gpu_set_chains(s->gpu, first, count, &io), "uploading the chain state");
}

void apm_session_pull(apm_session * s, int first, int count) {
	apm_gpu_chain_io io;
	int g, j;
	chain_io(s, first, &io);
	apm_gpu_check(s, apm_gpu_get_chains(s->gpu, first, count, &io), "downloading the chain state");
	for (g = first; g < first + count; g++) {
		mcmc * m = s->chains[g];
		parallel_tempering_mcmc * pt = (parallel_tempering_mcmc *) m->additional_data;
		pt->beta = s->beta[g];
		pt->swapcount = (unsigned long) s->swapcount[g];
		m->prob = s->prob[g];
		m->prior = s->prior[g];
		m->prob_best = s->prob_best[g];
		m->accept = (unsigned long) s->accept[g];
		m->reject = (unsigned long) s->reject[g];
		m->n_iter = (unsigned long) s->n_iter[g];
		for (j = 0; j < s->n_par; j++) {
			const size_t k = (size_t) g * s->n_par + j;
			gsl_vector_set(m->params, j, s->params[k]);
			gsl_vector_set(m->params_step, j, s->steps[k]);
			gsl_vector_set(m->params_best, j, s->params_best[k]);
			m->params_accepts[j] = (unsigned long) s->pacc[k];
			m->params_rejects[j] = (unsigned long) s->prej[k];
		}
	}
}

/* calc_model(chains[g], NULL) for g in which[]: one batched evaluation on the device */
void apm_session_calc_model(apm_session * s, const int * which, int n) {
	double * p = (double *) calloc((size_t) n * s->n_par, sizeof(double));
	double * b = (double *) calloc(n, sizeof(double));
	double * prob = (double *) calloc(n, sizeof(double));
	double * prior = (double *) calloc(n, sizeof(double));
	apm_gpu_chain_io io;
	int i, j;
	for (i = 0; i < n; i++) {
		const mcmc * m = s->chains[which[i]];
		b[i] = get_beta(m);
		for (j = 0; j < s->n_par; j++)
			p[(size_t) i * s->n_par + j] = gsl_vector_get(m->params, j);
	}
	apm_gpu_check(s, apm_gpu_eval(s->gpu, n, p, b, prob, prior), "evaluating the model");
	for (i = 0; i < n; i++) {
		mcmc * m = s->chains[which[i]];
		set_prior(m, prior[i]);
		set_prob(m, prob[i]);
		memset(&io, 0, sizeof(io));
		io.prob = &prob[i];
		io.prior = &prior[i];
		apm_gpu_check(s, apm_gpu_set_chains(s->gpu, which[i], 1, &io), "storing the model value");
	}
	free(p); free(b); free(prob); free(prior);
}

/* gsl_rng_uniform(get_random(chains[g])) for host-side algorithms */
double apm_session_uniform(apm_session * s, int g) {
	double u = 0;
	apm_gpu_check(s, apm_gpu_host_uniform(s->gpu, g, &u), "drawing a random number");
	return u;
}

/* APM_HOST_TIMING=1: where the wall time of a phase goes, on stderr */
static double wall_s(void) {
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}
#define TIMING_MARK(what) do { if (timing) { const double t_now = wall_s(); \
		fprintf(stderr, "[timing] %-28s %8.3f s\n", what, t_now - t_mark); t_mark = t_now; } } while (0)

/* markov_chain_calibrate for the selected chains, all at once on the device.  On failure:
 * the reference's message and exit code (src/markov_chain_calibrate.c:1104-1110,1169-1174) */
static void calibrate_selected(apm_session * s, const unsigned char * select, int skip,
		apm_gpu_calib_progress ** rows_out, long long * n_rows_out) {
	apm_gpu_calib_cfg cfg;
	int * status = (int *) calloc(s->n_chains, sizeof(int));
	long long cap = 0, n_rows = 0;
	apm_gpu_calib_progress * rows = NULL;
	int g, rc, n_sel = 0;
#if defined(CALIBRATE_ALTERNATE) || defined(CALIBRATE_MULTILIN) || defined(CALIBRATE_QUADRATIC)
	if (!skip) {
		/* markov_chain_calibrate = burn_in + the alternate calibrator (the reference's precedence:
		 * MULTILIN, then QUADRATIC, then ALTERNATE, src/markov_chain_calibrate.c:1190-1202), one chain after the
		 * other like the reference's loop (src/parallel_tempering.c:173-197 with one thread) */
		unsigned char * one = (unsigned char *) calloc(s->n_chains, 1);
		for (g = 0; g < s->n_chains; g++) {
			if (!select[g])
				continue;
			one[g] = 1;
			calib_config(&cfg, 1 /* burn_in only */);
			apm_gpu_check(s, apm_gpu_calibrate(s->gpu, one, &cfg, NULL, NULL, 0, NULL), "burn-in");
			one[g] = 0;
			apm_session_pull(s, g, 1);
			apm_set_output_dir(s->ens_first + g / s->n_beta);
#if defined(CALIBRATE_MULTILIN)
			apm_calibrate_multilin(s, g, TARGET_ACCEPTANCE_RATE, MAX_AR_DEVIATION, ITER_LIMIT);
#elif defined(CALIBRATE_QUADRATIC)
			apm_calibrate_quadratic(s, g, TARGET_ACCEPTANCE_RATE, MAX_AR_DEVIATION, ITER_LIMIT);
#else
			apm_calibrate_alt(s, g, TARGET_ACCEPTANCE_RATE, MAX_AR_DEVIATION, ITER_LIMIT);
#endif
		}
		apm_set_output_dir(-1);
		free(one);
		free(status);
		if (rows_out != NULL) {
			*rows_out = NULL; /* calibration_progress.data has been written by the calibrator itself */
			*n_rows_out = 0;
		}
		return;
	}
#endif
	calib_config(&cfg, skip);
	for (g = 0; g < s->n_chains; g++)
		n_sel += select[g] != 0;
	if (rows_out != NULL) {
		cap = ((long long) ITER_LIMIT / ITER_READJUST + 2) * s->n_par * n_sel;
		rows = (apm_gpu_calib_progress *) calloc(cap > 0 ? cap : 1, sizeof(*rows));
	}
	rc = apm_gpu_calibrate(s->gpu, select, &cfg, status, rows, cap, &n_rows);
	if (rc == APM_ECALIB) {
		for (g = 0; g < s->n_chains; g++) {
			if (status[g] == APM_CALIB_STEP_TOO_LARGE)
				fprintf(stderr, "calibration failed: step width became too large (chain %d).\n", g);
			else if (status[g] == APM_CALIB_ITER_LIMIT)
				fprintf(stderr, "calibration failed: limit of %d iterations reached (chain %d).\n",
						(int) ITER_LIMIT, g);
		}
		exit(1);
	}
	apm_gpu_check(s, rc, "calibration");
	free(status);
	if (rows_out != NULL) {
		*rows_out = rows;
		*n_rows_out = n_rows < cap ? n_rows : cap;
	}
}

/* ------------------------------------------------------------------ calibrate_first
 * reference src/parallel_tempering.c:78-95 */
void calibrate_first(void) {
	const int timing = getenv("APM_HOST_TIMING") != NULL;
	double t_mark = wall_s();
	apm_session * s = apm_session_open();
	unsigned char * select = (unsigned char *) calloc(s->n_chains, 1);
	int * which = (int *) calloc(s->n_ens, sizeof(int));
	apm_gpu_calib_progress * rows = NULL;
	long long n_rows = 0;
	int e;

	printf("Starting markov chain calibration\n");
	fflush(stdout);
	for (e = 0; e < s->n_ens; e++) {
		which[e] = e * s->n_beta;
		select[which[e]] = 1;
	}
	TIMING_MARK("start-up (files, CUDA, upload)");
	apm_session_calc_model(s, which, s->n_ens);
	calibrate_selected(s, select, 0, &rows, &n_rows);
	TIMING_MARK("calibrating the first chain");
	apm_session_pull(s, 0, s->n_chains);
	for (e = 0; e < s->n_ens; e++) {
		apm_set_output_dir(s->ens_first + e);
		if (rows != NULL)
			apm_write_calibration_progress(rows, n_rows, which[e]);
		write_calibrations_file(apm_ensemble(s, e), 1);
		write_params_file(apm_ensemble(s, e)[0]);
	}
	apm_set_output_dir(-1);
	free(rows); free(select); free(which);
	TIMING_MARK("writing the files");
	apm_session_close(s);
	TIMING_MARK("engine shutdown");
}

/* ------------------------------------------------------------------ calibrate_rest
 * reference src/parallel_tempering.c:115-207: learn per-parameter step-width factors by
 * calibrating chain 1, choose beta_0, predict every chain's step widths as
 * steps(chain 0) * beta^-1/2 * factors, then calibrate (or only burn in) chains 1..n-1.
 * All ensembles go through each stage together. */
static void place_chain(mcmc * c, const mcmc * c0, double beta, const gsl_vector * factors) {
	set_beta(c, beta);
	gsl_vector_memcpy(get_steps(c), get_steps(c0));
	gsl_vector_scale(get_steps(c), pow(get_beta(c), -0.5));
	if (factors != NULL)
		gsl_vector_mul(get_steps(c), factors);
	set_params(c, dup_vector(get_params_best(c0)));
}

void calibrate_rest(void) {
	const int timing = getenv("APM_HOST_TIMING") != NULL;
	double t_mark = wall_s();
	apm_session * s = apm_session_open();
	const int n_beta = s->n_beta, n_ens = s->n_ens;
	gsl_vector ** factors = (gsl_vector **) calloc(n_ens, sizeof(gsl_vector *));
	double * beta_0 = (double *) calloc(n_ens, sizeof(double));
	unsigned char * select = (unsigned char *) calloc(s->n_chains, 1);
	int * which = (int *) calloc(s->n_chains, sizeof(int));
	apm_gpu_calib_progress * rows = NULL;
	long long n_rows = 0;
	int e, i, n;
#ifdef SKIP_CALIBRATE_ALLCHAINS
	const int skip = 1;
#else
	const int skip = 0;
#endif

	for (e = 0; e < n_ens; e++) {
		apm_set_output_dir(s->ens_first + e);
		read_calibration_file(apm_ensemble(s, e), 1);
		factors[e] = gsl_vector_alloc(s->n_par);
		gsl_vector_set_all(factors[e], 1);
		beta_0[e] = BETA_0;
	}
	printf("Calibrating chains\n");
	fflush(stdout);
	TIMING_MARK("start-up (files, CUDA, upload)");

	if (n_beta > 1) {
		/* the second chain tells how step widths really scale with beta */
		for (e = 0, n = 0; e < n_ens; e++) {
			mcmc ** c = apm_ensemble(s, e);
			const double b0 = beta_0[e] < 0 ? calc_beta_0(c[0], factors[e]) : beta_0[e];
			place_chain(c[1], c[0], get_chain_beta(1, n_beta, b0), NULL);
			which[n++] = e * n_beta + 1;
			select[e * n_beta + 1] = 1;
		}
		apm_session_push(s, 0, s->n_chains);
		apm_session_calc_model(s, which, n);
		printf("Calibrating second chain to infer stepwidth factor\n");
		printf("\tChain %2d - beta = %f\tsteps: ", 1, get_beta(s->chains[1]));
		dump_vectorln(get_steps(s->chains[1]));
		fflush(stdout);
		calibrate_selected(s, select, 0, NULL, NULL);
		TIMING_MARK("calibrating the second chain");
		apm_session_pull(s, 0, s->n_chains);
		for (e = 0; e < n_ens; e++) {
			mcmc ** c = apm_ensemble(s, e);
			gsl_vector_scale(factors[e], pow(get_beta(c[1]), -0.5));
			gsl_vector_mul(factors[e], get_steps(c[0]));
			gsl_vector_div(factors[e], get_steps(c[1]));
		}
	}
	printf("stepwidth factors: ");
	dump_vectorln(factors[0]);
	for (e = 0; e < n_ens; e++) {
		if (beta_0[e] < 0) {
			beta_0[e] = calc_beta_0(apm_ensemble(s, e)[0], factors[e]);
			if (e == 0)
				printf("automatic beta_0: %f\n", beta_0[e]);
		}
	}
	fflush(stdout);

	if (n_beta > 1) {
		memset(select, 0, s->n_chains);
		for (e = 0, n = 0; e < n_ens; e++) {
			mcmc ** c = apm_ensemble(s, e);
			for (i = 1; i < n_beta; i++) {
				place_chain(c[i], c[0], get_chain_beta(i, n_beta, beta_0[e]), factors[e]);
				which[n++] = e * n_beta + i;
				select[e * n_beta + i] = 1;
				if (e == 0) {
					printf("\tChain %2d - beta = %f\tsteps: ", i, get_beta(c[i]));
					dump_vectorln(get_steps(c[i]));
				}
			}
		}
		apm_session_push(s, 0, s->n_chains);
		apm_session_calc_model(s, which, n);
		calibrate_selected(s, select, skip, &rows, &n_rows);
		TIMING_MARK("calibrating chains 1 .. n-1");
		apm_session_pull(s, 0, s->n_chains);
	}
	printf("all chains calibrated.\n");
	for (i = 0; i < n_beta; i++) {
		printf("\tChain %2d - beta = %f \tsteps: ", i, get_beta(s->chains[i]));
		dump_vectorln(get_steps(s->chains[i]));
	}
	for (e = 0; e < n_ens; e++) {
		apm_set_output_dir(s->ens_first + e);
		if (rows != NULL && !skip)
			apm_write_calibration_progress(rows, n_rows, e * n_beta + n_beta - 1);
		write_calibration_summary(apm_ensemble(s, e), n_beta);
		write_calibrations_file(apm_ensemble(s, e), n_beta);
		gsl_vector_free(factors[e]);
	}
	apm_set_output_dir(-1);
	free(rows); free(select); free(which); free(factors); free(beta_0);
	TIMING_MARK("writing the files");
	apm_session_close(s);
	TIMING_MARK("engine shutdown");
}

/* ------------------------------------------------------------------ run
 * reference src/parallel_tempering.c:209-250 (prepare_and_run_sampler), :347-419
 * (run_sampler), :308-345 (dump), :36-52 (report), src/parallel_tempering_run.c:28-58 */
static volatile sig_atomic_t keep_running = 1;
static volatile sig_atomic_t dump_requested = 0;

static void on_sigint(int signalnr) {
	(void) signalnr;
	keep_running = 0;
}
static void on_sigusr(int signalnr) {
	signal(signalnr, on_sigusr);
	dump_requested = 1;
}

/* ------------------------------------------------------------------ dump writer
 * The per-iteration text of run_sampler() (reference src/parallel_tempering.c:396-401,
 * src/mcmc_dump.c:79-88) for one engine call's trace.  At GPU step rates formatting it is the
 * slow part of `run` (two "%e" conversions per chain and iteration), so (a) every file is written
 * by one thread of an OpenMP team -- a file's bytes and their order do not depend on the team --
 * and (b) the job runs on a helper thread while the engine computes the next call's trace into
 * the other buffer. */
typedef struct {
	const apm_session * s;
	FILE ** prob_files;
	const double * t_prob, * t_dl, * t_par;
	long long n_prob_rows, n_par_rows;
	int n_dumped, params_chains;
} write_job;

enum { WR_BLOCK = 256, WR_LINE = 32, WR_SUPER = 16 }; /* rows per formatting item, bytes reserved per line, items per file and pass */

static void * write_trace(void * arg) {
	const write_job * w = (const write_job *) arg;
	const apm_session * s = w->s;
	const int n_chains = s->n_chains, n_par = s->n_par, n_beta = s->n_beta;
	const int n_files = n_chains + w->n_dumped * n_par;
	const long long rows_max = w->n_prob_rows > w->n_par_rows ? w->n_prob_rows : w->n_par_rows;
	static char * text = NULL;   /* [n_files][WR_SUPER][WR_BLOCK * WR_LINE] (only the writer thread uses it) */
	static int * length = NULL;  /* [n_files][WR_SUPER] */
	static size_t text_cap = 0;
	const size_t item_bytes = (size_t) WR_BLOCK * WR_LINE;
	long long base;
	if ((size_t) n_files * WR_SUPER * item_bytes > text_cap) {
		text_cap = (size_t) n_files * WR_SUPER * item_bytes;
		text = (char *) realloc(text, text_cap);
		length = (int *) realloc(length, (size_t) n_files * WR_SUPER * sizeof(int));
		assert(text != NULL && length != NULL);
	}
	/* A file's lines must be written in order, but they can be FORMATTED in any order: the rows are
	 * cut into blocks, (file, block) items are formatted into memory by the whole team -- so the
	 * few "%.15e" parameter files do not become the critical path -- and then every file's blocks
	 * are written out in order. */
	for (base = 0; base < rows_max; base += (long long) WR_SUPER * WR_BLOCK) {
		int item, k;
#pragma omp parallel
		{
#pragma omp for schedule(dynamic, 1)
			for (item = 0; item < n_files * WR_SUPER; item++) {
				const int file = item / WR_SUPER, blk = item % WR_SUPER;
				const long long r0 = base + (long long) blk * WR_BLOCK;
				char * out = text + (size_t) item * item_bytes;
				size_t used = 0;
				long long r, r1;
				if (file < n_chains) {
					/* prob-chain<k>.dump: "%6e\t%6e\n" = prob, prob - prior (apm_fastfmt.c) */
					r1 = r0 + WR_BLOCK < w->n_prob_rows ? r0 + WR_BLOCK : w->n_prob_rows;
					for (r = r0; r < r1; r++)
						used += (size_t) apm_format_prob_line(w->t_prob[r * n_chains + file], w->t_dl[r * n_chains + file],
								out + used);
				} else {
					/* <name>-chain-<i>.prob.dump: one parameter of one dumped chain, "%.15e\n" */
					const int i = (file - n_chains) / n_par, j = (file - n_chains) % n_par;
					r1 = r0 + WR_BLOCK < w->n_par_rows ? r0 + WR_BLOCK : w->n_par_rows;
					for (r = r0; r < r1; r++) {
						const double v = w->t_par[((size_t) r * w->n_dumped + i) * n_par + j];
						int len = apm_format_e15(v, out + used); /* DUMP_FORMAT is "%.15e" */
						if (len == 0)
							len = snprintf(out + used, WR_LINE - 1, DUMP_FORMAT, v);
						used += (size_t) len;
						out[used++] = '\n';
					}
				}
				length[item] = (int) used;
			}
#pragma omp for schedule(dynamic, 1)
			for (k = 0; k < n_files; k++) {
				FILE * f;
				int blk;
				if (k < n_chains) {
					f = w->prob_files[k];
				} else {
					const int i = (k - n_chains) / n_par, j = (k - n_chains) % n_par;
					const mcmc * m = w->params_chains == 2 ? s->chains[i] : s->chains[i * n_beta];
					f = m->files != NULL ? m->files[j] : NULL;
				}
				if (f == NULL)
					continue;
				for (blk = 0; blk < WR_SUPER; blk++)
					if (length[k * WR_SUPER + blk] > 0)
						fwrite(text + (size_t) (k * WR_SUPER + blk) * item_bytes, 1, (size_t) length[k * WR_SUPER + blk], f);
			}
		}
	}
	return NULL;
}

/* the helper thread: lives for the whole run (so does its OpenMP team), takes one job at a time */
typedef struct {
	pthread_t thread;
	pthread_mutex_t lock;
	pthread_cond_t wake, idle;
	write_job job;
	int has_job, busy, quit, started;
} writer_t;

static void * writer_main(void * arg) {
	writer_t * w = (writer_t *) arg;
	pthread_mutex_lock(&w->lock);
	while (1) {
		while (!w->has_job && !w->quit)
			pthread_cond_wait(&w->wake, &w->lock);
		if (!w->has_job && w->quit)
			break;
		w->has_job = 0;
		w->busy = 1;
		pthread_mutex_unlock(&w->lock);
		write_trace(&w->job);
		pthread_mutex_lock(&w->lock);
		w->busy = 0;
		pthread_cond_broadcast(&w->idle);
	}
	pthread_mutex_unlock(&w->lock);
	return NULL;
}

static void writer_start(writer_t * w) {
	memset(w, 0, sizeof(*w));
	pthread_mutex_init(&w->lock, NULL);
	pthread_cond_init(&w->wake, NULL);
	pthread_cond_init(&w->idle, NULL);
	w->started = pthread_create(&w->thread, NULL, writer_main, w) == 0;
}

/* block until the job handed over last (if any) has been written */
static void writer_wait(writer_t * w) {
	if (!w->started)
		return;
	pthread_mutex_lock(&w->lock);
	while (w->has_job || w->busy)
		pthread_cond_wait(&w->idle, &w->lock);
	pthread_mutex_unlock(&w->lock);
}

/* hand over a job (the previous one must have been waited for) */
static void writer_submit(writer_t * w, const write_job * job) {
	if (!w->started) {
		write_trace((void *) job);
		return;
	}
	pthread_mutex_lock(&w->lock);
	w->job = *job;
	w->has_job = 1;
	pthread_cond_signal(&w->wake);
	pthread_mutex_unlock(&w->lock);
}

static void writer_stop(writer_t * w) {
	if (!w->started)
		return;
	writer_wait(w);
	pthread_mutex_lock(&w->lock);
	w->quit = 1;
	pthread_cond_signal(&w->wake);
	pthread_mutex_unlock(&w->lock);
	pthread_join(w->thread, NULL);
	w->started = 0;
}

static void report(apm_session * s) {
	int e, i;
	printf("printing chain parameters: \n");
	for (i = 0; i < s->n_beta; i++) {
		const mcmc * m = s->chains[i];
		printf("\tchain %d: swapped %lu times: ", i, get_swapcount(m));
		printf("\tchain %d: current %f: ", i, get_prob(m));
		dump_vectorln(get_params(m));
		printf("\tchain %d: best %f: ", i, get_prob_best(m));
		dump_vectorln(get_params_best(m));
	}
	printf("\nwriting out visited parameters ");
	for (e = 0; e < s->n_ens; e++)
		for (i = 0; i < s->n_beta; i++)
			mcmc_dump_flush(apm_ensemble(s, e)[i]);
	printf(".done.\n");
	fflush(stdout);
}

static unsigned long gcd_ul(unsigned long a, unsigned long b) {
	while (b != 0) {
		unsigned long t = a % b;
		a = b;
		b = t;
	}
	return a;
}

/* what `analyse` needs, taken from the on-device accumulators at full precision (the text
 * dumps carry 7 digits): per chain beta, n, mean(prob - prior), then mean and variance of
 * every parameter */
static void write_run_statistics(apm_session * s) {
	const size_t n = s->n_chains, nv = n * s->n_par;
	unsigned long long * cnt = (unsigned long long *) calloc(n, sizeof(*cnt));
	double * sdl = (double *) calloc(n, sizeof(double));
	double * sp = (double *) calloc(nv, sizeof(double)), *sp2 = (double *) calloc(nv, sizeof(double));
	int e, k, j;
	apm_gpu_check(s, apm_gpu_get_stats(s->gpu, cnt, sdl, sp, sp2), "reading the accumulators");
	for (e = 0; e < s->n_ens; e++) {
		FILE * f;
		apm_set_output_dir(s->ens_first + e);
		f = fopen(apm_out_path("run_statistics"), "w");
		if (f == NULL)
			continue;
		for (k = 0; k < s->n_beta; k++) {
			const size_t g = (size_t) e * s->n_beta + k;
			const double c = cnt[g] > 0 ? (double) cnt[g] : 1;
			fprintf(f, "%d\t" DUMP_FORMAT "\t%llu\t" DUMP_FORMAT, k, get_beta(s->chains[g]), cnt[g], sdl[g] / c);
			for (j = 0; j < s->n_par; j++) {
				const double mean = sp[g * s->n_par + j] / c;
				fprintf(f, "\t" DUMP_FORMAT "\t" DUMP_FORMAT, mean, sp2[g * s->n_par + j] / c - mean * mean);
			}
			fprintf(f, "\n");
		}
		fclose(f);
	}
	apm_set_output_dir(-1);
	free(cnt); free(sdl); free(sp); free(sp2);
}

/* the marginal statistics `analyse` needs -- per parameter the NBINS-bin histogram of the recorded
 * chain(s) and the batch means behind the Monte Carlo error estimate -- from the on-device
 * accumulators (reference src/analyse.c:115-247 derives them from the
 * parameter dumps).  One file per ensemble:
 *   marginals <n_par> <n_bins> <batch_size> <n_slots>
 *   chain <k> <n_values> <n_batches>                    (once per slot)
 *   counts <n_bins integers>                            (once per parameter)
 *   means <n_batches values, %.17g>                     (once per parameter) */
static void write_run_marginals(apm_session * s, int which_chains, unsigned long batch, int max_batches) {
	const int per_ens = which_chains == 2 ? s->n_beta : 1;
	const size_t slots = (size_t) s->n_ens * per_ens, np = s->n_par;
	unsigned long long * counts = (unsigned long long *) calloc(slots * np * NBINS, sizeof(*counts));
	double * means = (double *) calloc(slots * np * (max_batches > 0 ? max_batches : 1), sizeof(double));
	unsigned long long * nv = (unsigned long long *) calloc(slots, sizeof(*nv)), * nb = (unsigned long long *) calloc(slots, sizeof(*nb));
	int e, k, j, b;
	apm_gpu_check(s, apm_gpu_get_marginals(s->gpu, counts, means, nv, nb), "reading the marginal statistics");
	for (e = 0; e < s->n_ens; e++) {
		FILE * f;
		apm_set_output_dir(s->ens_first + e);
		f = fopen(apm_out_path("run_marginals"), "w");
		if (f == NULL)
			continue;
		fprintf(f, "marginals %d %d %lu %d\n", s->n_par, (int) NBINS, batch, per_ens);
		for (k = 0; k < per_ens; k++) {
			const size_t slot = (size_t) e * per_ens + k;
			const unsigned long long kept = nb[slot] < (unsigned long long) max_batches ? nb[slot] : (unsigned long long) max_batches;
			fprintf(f, "chain %d %llu %llu\n", k, nv[slot], kept);
			for (j = 0; j < s->n_par; j++) {
				fprintf(f, "counts");
				for (b = 0; b < NBINS; b++)
					fprintf(f, " %llu", counts[(slot * np + j) * NBINS + b]);
				fprintf(f, "\nmeans");
				for (b = 0; b < (int) kept; b++)
					fprintf(f, " %.17g", means[(slot * np + j) * max_batches + b]);
				fprintf(f, "\n");
			}
		}
		fclose(f);
	}
	apm_set_output_dir(-1);
	free(counts); free(means); free(nv); free(nb);
}

void prepare_and_run_sampler(const unsigned long max_iterations, int append) {
	const int timing = getenv("APM_HOST_TIMING") != NULL;
	double t_mark = wall_s(), t_a = 0, t_b = 0, t_engine = 0, t_read = 0, t_join = 0;
	long n_calls = 0;
	apm_session * s = apm_session_open();
	const int n_beta = s->n_beta, n_ens = s->n_ens, n_par = s->n_par, n_chains = s->n_chains;
	char * mode = append == 1 ? "a" : "w";
	int n_swap = N_SWAP;
	FILE ** prob_files = (FILE **) calloc(n_chains, sizeof(FILE *));
	FILE ** accept_files = (FILE **) calloc(n_ens, sizeof(FILE *));
	apm_gpu_trace_cfg trace;
	unsigned long iter, interval_rounds;
	long long max_rows, call_cap = 1;
	double * t_prob[2] = { NULL, NULL }, *t_dl[2] = { NULL, NULL }, *t_par[2] = { NULL, NULL };
	size_t t_cap[2] = { 0, 0 };
	write_job job;
	writer_t writer;
	int cur = 0;
	char name[64];
	int e, i, n_dumped;
	/* APM_NO_DUMPS=1: no prob-chain<k>.dump / <name>-chain-<i>.prob.dump text at all -- `analyse` then
	 * works from run_statistics and run_marginals, the on-device accumulators */
	const int no_dumps = getenv("APM_NO_DUMPS") != NULL;
	int marg_which = 0, marg_cap = 0;
	unsigned long marg_batch = 0;

#ifdef DUMP_ALL_CHAINS
	trace.params_chains = 2;
	n_dumped = n_chains;
#else
	trace.params_chains = 1;
	n_dumped = n_ens;
#endif
	trace.prob_every = 1;
	if (no_dumps) {
		trace.params_chains = 0;
		trace.prob_every = 0;
		n_dumped = 0;
	}

	for (e = 0; e < n_ens; e++) {
		mcmc ** c = apm_ensemble(s, e);
		apm_set_output_dir(s->ens_first + e);
		read_calibration_file(c, n_beta);
		for (i = 0; i < (trace.params_chains == 2 ? n_beta : (trace.params_chains == 1 ? 1 : 0)); i++)
			mcmc_open_dump_files(c[i], "-chain", i, mode);
		for (i = 0; i < (no_dumps ? 0 : n_beta); i++) {
			snprintf(name, sizeof(name), "prob-chain%d.dump", i);
			prob_files[e * n_beta + i] = fopen(apm_out_path(name), mode);
			if (prob_files[e * n_beta + i] == NULL) {
				fprintf(stderr, "opening file %s failed\n", name);
				perror("opening file failed");
				exit(1);
			}
			setvbuf(prob_files[e * n_beta + i], NULL, _IOFBF, 1 << 18);
		}
		accept_files[e] = fopen(apm_out_path("acceptance_rate.dump.gnuplot"), "w");
		if (accept_files[e] != NULL) {
			FILE * f = accept_files[e];
			fprintf(f, "# format: iteration | number of accepts for each chain\nplot ");
			for (i = 0; i < n_beta; i++)
				fprintf(f, "\"acceptance_rate.dump\" u 1:%d title \"chain %d, beta = %f\"%s", i + 2, i,
						get_beta(c[i]), i != n_beta - 1 ? ", " : "");
			fprintf(f, "\n");
			fclose(f);
		}
		accept_files[e] = fopen(apm_out_path("acceptance_rate.dump"), mode);
		assert(accept_files[e] != NULL);
	}
	apm_set_output_dir(-1);
	apm_session_push(s, 0, n_chains);
	apm_gpu_check(s, apm_gpu_reset_stats(s->gpu), "resetting the accumulators");

	if (n_swap < 0) {
		n_swap = 2000 / n_beta;
		if (n_swap < 1)
			n_swap = 1;
		printf("automatic n_swap: %d\n", n_swap);
	}
	/* marginal statistics on the device: the histogram bins and the batches analyse would form from
	 * the parameter dumps.  calc_mcmc_error's batch size is sqrt(number of dumped values), known in
	 * advance for a run of a fixed length that starts its files afresh; otherwise (open-ended or
	 * appending) only the dumps can tell and none are kept here. */
	for (e = 0; e < n_ens; e++) {
		apm_set_output_dir(s->ens_first + e);
		remove(apm_out_path("run_marginals"));
	}
	apm_set_output_dir(-1);
	if (max_iterations != 0 && append != 1) {
		const unsigned long total = (unsigned long) ((max_iterations + n_swap - 1) / n_swap) * n_swap;
#ifdef HISTOGRAMS_ALLCHAINS
		marg_which = 2; /* one histogram per parameter over all chains' values: that many values more */
		marg_batch = (unsigned long) sqrt((double) total * n_beta);
#else
		marg_which = 1;
		marg_batch = (unsigned long) sqrt((double) total);
#endif
		marg_cap = marg_batch > 0 ? (int) (total / marg_batch) + 2 : 0;
		apm_gpu_check(s, apm_gpu_set_marginals(s->gpu, marg_which, NBINS, marg_batch, marg_cap),
				"setting up the marginal statistics");
	}
	/* rounds between two acceptance_rate.dump rows: iter advances by n_swap per round and a
	 * row is due whenever iter % PRINT_PROB_INTERVAL == 0 */
	interval_rounds = PRINT_PROB_INTERVAL / gcd_ul(PRINT_PROB_INTERVAL, (unsigned long) n_swap);
	/* bound the trace of one engine call to ~256 MB of host memory */
	max_rows = (long long) (16u << 20) / n_chains;
	if (max_rows < n_swap)
		max_rows = n_swap;

	signal(SIGINT, on_sigint);
	signal(SIGUSR1, on_sigusr);
	signal(SIGUSR2, on_sigusr);
	keep_running = 1;
	dump_requested = 0;
	writer_start(&writer);
	iter = s->chains[0]->n_iter;
	printf("starting the analysis\n");
	fflush(stdout);
	TIMING_MARK("start-up (files, CUDA, upload)");

	while (keep_running && (max_iterations == 0 || iter < max_iterations)) {
		/* one engine call: up to the next status row, the end of the run, the trace bound and
		 * about half a second of device time, so that signals are honoured promptly */
		const unsigned long done_rounds = iter / n_swap;
		long long rounds = (long long) (interval_rounds - done_rounds % interval_rounds);
		long long n_prob_rows = 0, n_par_rows = 0;
		struct timespec t0, t1;
		double secs;
		size_t need;
		if (max_iterations != 0) {
			const long long left = (long long) ((max_iterations - iter + n_swap - 1) / n_swap);
			if (rounds > left)
				rounds = left;
		}
		if (rounds > call_cap)
			rounds = call_cap;
		if (rounds * n_swap > max_rows)
			rounds = max_rows / n_swap;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		t_a = wall_s();
		apm_gpu_check(s, apm_gpu_run(s->gpu, rounds, n_swap, &trace), "sampling");
		t_b = wall_s();
		t_engine += t_b - t_a;
		n_calls++;
		need = (size_t) rounds * n_swap;
		if (need > t_cap[cur]) { /* (buffer `cur` is idle: its job was waited for one call ago) */
			free(t_prob[cur]); free(t_dl[cur]); free(t_par[cur]);
			t_cap[cur] = need;
			t_prob[cur] = (double *) malloc(need * n_chains * sizeof(double));
			t_dl[cur] = (double *) malloc(need * n_chains * sizeof(double));
			t_par[cur] = (double *) malloc(need * n_dumped * n_par * sizeof(double) + 8);
			assert(t_prob[cur] != NULL && t_dl[cur] != NULL && t_par[cur] != NULL);
		}
		if (!no_dumps)
			apm_gpu_check(s, apm_gpu_read_trace(s->gpu, t_prob[cur], t_dl[cur], t_par[cur], &n_prob_rows, &n_par_rows),
					"reading the trace");
		t_read += wall_s() - t_b;
		clock_gettime(CLOCK_MONOTONIC, &t1);
		secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
		if (secs < 0.2 && call_cap < (1ll << 40))
			call_cap *= 2;
		else if (secs > 0.5 && call_cap > 1) /* SIGINT / SIGUSR1 are honoured between calls: <= ~0.5 s */
			call_cap /= 2;

		/* hand the trace to the writer; the previous job (other buffer, same files) must be done */
		t_a = wall_s();
		writer_wait(&writer);
		t_join += wall_s() - t_a;
		job.s = s;
		job.prob_files = prob_files;
		job.t_prob = t_prob[cur];
		job.t_dl = t_dl[cur];
		job.t_par = t_par[cur];
		job.n_prob_rows = n_prob_rows;
		job.n_par_rows = n_par_rows;
		job.n_dumped = n_dumped;
		job.params_chains = trace.params_chains;
		if (!no_dumps)
			writer_submit(&writer, &job);
		cur ^= 1;
		iter += (unsigned long) rounds * n_swap;

		if (iter % PRINT_PROB_INTERVAL == 0) { /* dump() */
			const mcmc * m = s->chains[0];
			apm_session_pull(s, 0, n_chains);
			if (dump_requested) {
				report(s);
				dump_requested = 0;
				writer_wait(&writer);
				for (i = 0; i < n_chains; i++)
					if (prob_files[i] != NULL)
						fflush(prob_files[i]);
			}
			for (e = 0; e < n_ens; e++) {
				fprintf(accept_files[e], "%lu", iter);
				for (i = 0; i < n_beta; i++)
					fprintf(accept_files[e], "\t%lu", get_params_accepts_global(apm_ensemble(s, e)[i]));
				fprintf(accept_files[e], "\n");
				fflush(accept_files[e]);
			}
			printf("iteration: %lu, a/r: %.3f(%lu/%lu), v:", iter,
					(double) m->accept / (double) (m->accept + m->reject), m->accept, m->reject);
			dump_vector(get_params(m));
			printf(" [%.0f chain-steps/s]\r", (double) rounds * n_swap * n_chains / (secs > 0 ? secs : 1));
			fflush(stdout);
		}
	}
	TIMING_MARK("sampling loop");
	if (timing)
		fprintf(stderr, "[timing]   of which: engine %.3f s, reading the trace %.3f s, waiting for the writer %.3f s "
				"(%ld engine calls)\n", t_engine, t_read, t_join, n_calls);
	writer_stop(&writer);
	TIMING_MARK("waiting for the dump writer");
	apm_session_pull(s, 0, n_chains);
	for (e = 0; e < n_ens; e++)
		fclose(accept_files[e]);
	for (i = 0; i < n_chains; i++)
		if (prob_files[i] != NULL)
			fclose(prob_files[i]);
	printf("handled %lu iterations on %d chains\n", iter, n_chains);
	report(s);
	write_run_statistics(s);
	if (marg_which != 0 && iter == (unsigned long) ((max_iterations + n_swap - 1) / n_swap) * n_swap)
		write_run_marginals(s, marg_which, marg_batch, marg_cap); /* (not after an interrupted run: wrong batch size) */
	for (i = 0; i < 2; i++) {
		free(t_prob[i]);
		free(t_dl[i]);
		free(t_par[i]);
	}
	free(prob_files);
	free(accept_files);
	TIMING_MARK("closing files, statistics");
	apm_session_close(s);
	TIMING_MARK("engine shutdown");
}
