/*
 * apm_session.h -- internal to the host layer: one "session" = the chain structs of
 * every ensemble (the file-facing mirror) + one engine handle (the device state).
 */
#ifndef APM_SESSION_H_
#define APM_SESSION_H_

#include "apm_host.h"
#include "apemost_gpu.h"

#define APM_PATH_MAX 1024

/* the device model this binary is built for: the Makefile derives it from the target
 * name (simplesin.exe -> APM_MODEL_SIMPLESIN, ...); a model file with its own
 * apps/<name>.cuh is built as APM_MODEL_USER */
#ifndef APM_MODEL_ID
#define APM_MODEL_ID APM_MODEL_SIMPLESIN
#endif
#ifndef APM_MODEL_NAME
#define APM_MODEL_NAME "simplesin"
#endif

typedef struct {
	int ens_first;           /* global number of this process's first ensemble */
	int n_ens, n_beta, n_par, n_chains; /* n_ens: ensembles handled by this process */
	mcmc ** chains;          /* [n_chains]: chains[e * n_beta + k], k = 0 is beta = 1 */
	apm_gpu * gpu;
	/* flat staging buffers, chain-major like the ABI */
	double * beta, * params, * steps, * prob, * prior, * prob_best, * params_best;
	unsigned long long * accept, * reject, * pacc, * prej, * n_iter, * swapcount;
} apm_session;

apm_session * apm_session_open(void);
void apm_session_close(apm_session * s);
/* chain structs -> device / device -> chain structs, chains [first, first + count) */
void apm_session_push(apm_session * s, int first, int count);
void apm_session_pull(apm_session * s, int first, int count);
/* calc_model(chains[g], NULL) for the listed chains, on the device; results land in the
 * structs and on the device */
void apm_session_calc_model(apm_session * s, const int * which, int n);
void apm_gpu_check(apm_session * s, int rc, const char * what);
mcmc ** apm_ensemble(apm_session * s, int e); /* = &chains[e * n_beta] */

/* apm_assess.c: assess_acceptance_rate (reference src/markov_chain.c:117-224) for chain g: the
 * measurement the three alternate calibrators are built on; returns the steps it took */
unsigned int apm_assess_acceptance_rate(apm_session * s, int g, unsigned int param, double desired_acceptance_rate,
		double min_accuracy, double max_accuracy, double * acceptance_rate, double * accuracy);
/* apm_calibrate_alt.c: -DCALIBRATE_ALTERNATE (reference src/markov_chain_calibrate.c:927-1037) for
 * chain g, after its burn-in */
void apm_calibrate_alt(apm_session * s, int g, double desired_acceptance_rate, const double max_ar_deviation,
		const unsigned int iter_limit);
/* apm_calibrate_quadratic.c: -DCALIBRATE_QUADRATIC (reference src/markov_chain_calibrate.c:239-914) */
void apm_calibrate_quadratic(apm_session * s, int g, double desired_acceptance_rate, const double max_ar_deviation,
		const unsigned int iter_limit);

/* apm_calibrate_multilin.c: -DCALIBRATE_MULTILIN (reference src/markov_chain_calibrate.c:33-237,
 * src/gsl_helper.c:189-298) for chain g, after its burn-in */
void apm_calibrate_multilin(apm_session * s, int g, double desired_acceptance_rate, const double max_ar_deviation,
		const unsigned int iter_limit);
double apm_session_uniform(apm_session * s, int g);

/* apm_fastfmt.c: "%6e" without printf (declines with 0 where it cannot guarantee printf's bytes) */
int apm_format_e6(double v, char * out);
int apm_format_e15(double v, char * out); /* "%.15e" */
int apm_format_prob_line(double prob, double dl, char * buf);

/* apm_files.c */
void apm_set_output_dir(int ensemble);
const char * apm_out_path(const char * name);
void apm_write_calibration_progress(const apm_gpu_calib_progress * rows, long long n, int chain);

#endif
