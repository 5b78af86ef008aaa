/*
 * apm_host.h -- the C host layer of apemost-b200: what an APEMoST model file
 * (apps/<model>.c) and an APEMoST user see, re-implemented over the GPU engine's
 * C ABI (include/apemost_gpu.h).
 *
 * The reference is one C program whose plugin is bound at link time (reference
 * Makefile:54-55).  This layer keeps that shape: `make <model>.exe` links this
 * host layer, the (unmodified) apps/<model>.c if it is available, and
 * libapemost_gpu.so, and the resulting binary offers the reference's phases
 *   check | calibrate_first | calibrate_rest | run [--append] | analyse
 * (reference apps/generic_main.c:112-159) on the same `params` / `data` input
 * files with byte-compatible output files.
 *
 * What lives where
 *   - this header      the `mcmc` chain object, layout-compatible with reference
 *                      src/mcmc_struct.h:30-106, the accessors model files use
 *                      (src/mcmc_gettersetter.h), the plugin contract
 *                      (src/mcmc.h:164,173), the phase entry points
 *                      (src/parallel_tempering.h:55-63) and the compile-time
 *                      configuration macros with the reference's defaults
 *                      (src/define_defaults.h);
 *   - mcmc.h, parallel_tempering.h, debug.h, ... next to it are one-line
 *     forwarders so that a model file's #include lines resolve unchanged.
 *
 * The Metropolis steps, swaps and calibration decisions are NOT here: the host
 * only parses files, keeps the chain structs as a mirror of the device state,
 * enqueues engine calls and writes text.
 */
#ifndef APM_HOST_H_
#define APM_HOST_H_

#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_vector.h>
#include <gsl/gsl_matrix.h>
#include <gsl/gsl_rng.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- compile-time configuration (pass with CCFLAGS="-DN_BETA=12 ...") ------ */
#ifndef N_BETA
#define N_BETA 20
#endif
#ifndef BETA_0
#define BETA_0 -0.001            /* < 0: automatic (calc_beta_0) */
#endif
#ifndef BURN_IN_ITERATIONS
#define BURN_IN_ITERATIONS 10000
#endif
#ifndef ITER_LIMIT
#define ITER_LIMIT 100000
#endif
#ifndef MUL
#define MUL 0.85
#endif
#ifndef N_SWAP
#define N_SWAP -30               /* < 0: automatic, 2000 / N_BETA */
#endif
#ifndef TARGET_ACCEPTANCE_RATE
#define TARGET_ACCEPTANCE_RATE 0.50
#endif
#ifndef MAX_AR_DEVIATION
#define MAX_AR_DEVIATION 0.01
#endif
#ifndef PARAMS_FILENAME
#define PARAMS_FILENAME "params"
#endif
#ifndef DATA_FILENAME
#define DATA_FILENAME "data"
#endif
#ifndef CALIBRATION_FILE
#define CALIBRATION_FILE "calibration_results"
#endif
#ifndef MAX_ITERATIONS
#define MAX_ITERATIONS 0         /* 0: run until SIGINT */
#endif
#ifndef PRINT_PROB_INTERVAL
#define PRINT_PROB_INTERVAL 1000
#endif
#ifndef DEFAULT_ADJUST_STEP
#define DEFAULT_ADJUST_STEP 0.5
#endif
#ifndef NO_RESCALING_LIMIT
#define NO_RESCALING_LIMIT 15
#endif
#ifndef ITER_READJUST
#define ITER_READJUST 200
#endif
#ifndef CIRCULAR_PARAMS
#define CIRCULAR_PARAMS 0        /* 1-based list, e.g. -DCIRCULAR_PARAMS=1,3 */
#endif
#ifndef BETA_ALIGNMENT
#define BETA_ALIGNMENT chebyshev_beta
#endif
#ifndef BETA_0_STEPWIDTH
#define BETA_0_STEPWIDTH 1.0
#endif
#ifndef NBINS
#define NBINS 200
#endif
#ifndef GNUPLOT_STYLE
#define GNUPLOT_STYLE "with histeps"
#endif
#ifndef N_ENSEMBLES
#define N_ENSEMBLES 1            /* new: independent PT ensembles run side by side */
#endif
#define DUMP_FORMAT "%.15e"

/* ---- the chain object (field order = reference src/mcmc_struct.h:30-106) --- */
typedef struct {
	unsigned int n_par;
	unsigned long accept;
	unsigned long reject;
	double prob;
	double prior;
	double prob_best;
	gsl_rng * random;
	gsl_vector * params;
	gsl_vector * params_best;
	FILE ** files;
	const char ** params_descr;
	unsigned long * params_accepts;
	unsigned long * params_rejects;
	gsl_vector * params_step;
	gsl_vector * params_min;
	gsl_vector * params_max;
	const gsl_matrix * data;
	unsigned long n_iter;
	void * additional_data;
} mcmc;

/* hung off mcmc.additional_data (reference src/parallel_tempering_beta.h:65-76) */
typedef struct {
	double beta;
	unsigned long swapcount;
} parallel_tempering_mcmc;

/* ---- the plugin contract (reference src/mcmc.h:164,173).  The host copies of
 * these functions are optional here: the sampler evaluates the model's
 * __device__ counterpart; a linked calc_model() is used by `check` and by
 * eval_<model>.exe --host to cross-check device against host. ------------- */
void calc_model(mcmc * m, const gsl_vector * old_values);
void calc_model_for(mcmc * m, const unsigned int i, const double old_value);

/* ---- chain object life cycle and input files (reference src/mcmc.c:37-118,
 * src/mcmc_parser.c:47-181) ------------------------------------------------ */
mcmc * mcmc_init(const unsigned int n_pars);
mcmc * mcmc_free(mcmc * m);
void mcmc_check(const mcmc * m);
mcmc * mcmc_load_params(const char * filename);
void mcmc_load_data(mcmc * m, const char * datafilename);
void mcmc_reuse_data(mcmc * m, const mcmc * m_orig);
mcmc * mcmc_load(const char * filename, const char * datafilename);
gsl_vector * dup_vector(const gsl_vector * v);
double mod_double(double x, double div);

/* ---- accessors used by model files and tools (reference
 * src/mcmc_gettersetter.h:25-109, src/parallel_tempering_beta.h:27-37) ------ */
unsigned int get_n_par(const mcmc * m);
gsl_vector * get_params(const mcmc * m);
double get_params_for(const mcmc * m, const unsigned int i);
void set_params_for(mcmc * m, const double v, const unsigned int i);
void set_params(mcmc * m, gsl_vector * v);       /* takes ownership, frees the old vector */
gsl_vector * get_params_best(const mcmc * m);
void set_params_best(mcmc * m, const gsl_vector * v);
gsl_vector * get_steps(const mcmc * m);
double get_steps_for(const mcmc * m, const unsigned int i);
void set_steps_for(mcmc * m, const double v, const unsigned int i);
double get_steps_for_normalized(const mcmc * m, const unsigned int i); /* step / (max - min) */
void reset_accept_rejects(mcmc * m);
gsl_vector * get_params_min(const mcmc * m);
gsl_vector * get_params_max(const mcmc * m);
double get_params_min_for(const mcmc * m, const unsigned int i);
double get_params_max_for(const mcmc * m, const unsigned int i);
const char ** get_params_descr(const mcmc * m);
double get_prob(const mcmc * m);
void set_prob(mcmc * m, const double v);
double get_prior(const mcmc * m);
void set_prior(mcmc * m, const double v);
double get_prob_best(const mcmc * m);
void set_prob_best(mcmc * m, const double v);
const gsl_matrix * get_data(const mcmc * m);
void set_data(mcmc * m, const gsl_matrix * data);
unsigned long get_params_accepts_global(const mcmc * m);
unsigned long get_params_rejects_global(const mcmc * m);
unsigned long get_params_accepts_for(const mcmc * m, const unsigned int i);
unsigned long get_params_rejects_for(const mcmc * m, const unsigned int i);
void set_beta(mcmc * m, const double newbeta);   /* also zeroes the swap count, as the reference */
double get_beta(const mcmc * m);
unsigned long get_swapcount(const mcmc * m);
void dump_vector(const gsl_vector * v);
void dump_vectorln(const gsl_vector * v);

/* ---- beta ladder (reference src/parallel_tempering_beta.c:53-102) ---------- */
double equidistant_beta(const unsigned int i, const unsigned int n_beta, const double beta_0);
double equidistant_temperature(const unsigned int i, const unsigned int n_beta, const double beta_0);
double chebyshev_temperature(const unsigned int i, const unsigned int n_beta, const double beta_0);
double chebyshev_beta(const unsigned int i, const unsigned int n_beta, const double beta_0);
double equidistant_stepwidth(const unsigned int i, const unsigned int n_beta, const double beta_0);
double chebyshev_stepwidth(const unsigned int i, const unsigned int n_beta, const double beta_0);
double hot_chains(const unsigned int i, const unsigned int n_beta, const double beta_0);
double get_chain_beta(unsigned int i, unsigned int n_beta, double beta_0);
double calc_beta_0(mcmc * m, gsl_vector * stepwidth_factors);

/* ---- calibration / parameter files (reference
 * src/parallel_tempering_config.c:28-202) ----------------------------------- */
mcmc ** setup_chains(void);
void read_calibration_file(mcmc ** chains, unsigned int n_chains);
void write_calibrations_file(mcmc ** chains, const unsigned int n_chains);
void write_calibration_summary(mcmc ** chains, unsigned int n_chains);
void write_params_file(mcmc * m);

/* ---- dump files (reference src/mcmc_dump.c:60-113) ------------------------- */
void mcmc_open_dump_files(mcmc * m, const char * suffix, int index, char * mode);
void mcmc_dump_current(const mcmc * m);
void mcmc_dump_flush(const mcmc * m);
void mcmc_dump_close(mcmc * m);

/* ---- the phases (reference src/parallel_tempering.h:55-63) ------------------ */
void calibrate_first(void);
void calibrate_rest(void);
void prepare_and_run_sampler(const unsigned long max_iterations, int append);
void analyse_marginal_distributions(void);
void analyse_data_probability(void);
/* new: print the compiled configuration, check the input files, and cross-check the
 * device model against a linked host calc_model() (reference apps/generic_main.c:283-370,
 * apps/benchmark_main.c:63-77) */
void check(void);

#ifdef __cplusplus
}
#endif
#endif /* APM_HOST_H_ */
