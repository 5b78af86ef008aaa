/*
 * apm_oracle.h -- CPU oracle for the APEMoST hot path.  TEST INFRASTRUCTURE.
 *
 * A plain-C (GSL-free) restatement of the reference algorithm: calc_model of
 * the shipped models, markov_chain_step / _step_for, check_accept,
 * mcmc_check_best, burn_in, markov_chain_calibrate_orig, tempering_interaction
 * and the run_sampler loop.  Every function in apm_oracle.c cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (apemost_b200/, include/) never does.
 *
 * Pinning ("parity pinned"): tests/test_oracle_vs_ref.py and the fixtures in
 * tests/golden/ check this restatement against
 *   - the reference's only known-answer vector (doc/manual.rst:189-213),
 *   - the beta ladder printed at doc/manual.rst:419-440,
 *   - the mod_double values of the reference's tests/tests.c:152-160,
 *   - outputs of the UNMODIFIED reference built here as oracle/_ref: eval_<model>
 *     on grids of parameter vectors, and byte-for-byte identical dump files of
 *     complete calibrate/run phases when the oracle is put in ORC_RNG_MT19937 mode
 *     (same global MT19937 stream, same draw order as the single-threaded reference).
 *
 * The oracle has two random-number modes.  ORC_RNG_MT19937 replays the
 * reference's process-global GSL generator so that dumps can be compared with
 * oracle/_ref byte for byte.  ORC_RNG_PHILOX uses the engine's counter-based
 * per-chain streams (Philox4x32-10), so that the CUDA engine and the oracle can
 * be compared trajectory for trajectory.  Everything except where the random
 * numbers come from is the same code in both modes.
 *
 * The API mirrors include/apemost_gpu.h (orc_ instead of apm_gpu_) so that the
 * parity tests read the same on both sides.
 */
#ifndef APM_ORACLE_H_
#define APM_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_RNG_MT19937 0
#define ORC_RNG_PHILOX  1

/* same numbering as APM_MODEL_* / APM_PROPOSAL_* / APM_QUIRK_* */
#define ORC_MODEL_SIMPLESIN   0
#define ORC_MODEL_SIMPLESIN5  1
#define ORC_MODEL_NORMAL      2
#define ORC_MODEL_PULSE_VROT  3
#define ORC_MODEL_SIMPLESIN2  4
#define ORC_MODEL_PULSE       5
#define ORC_MODEL_BERNOULLI   6

#define ORC_QUIRK_STALE_PROB_ON_SWAP    1u
#define ORC_QUIRK_STALE_PRIOR_ON_REJECT 2u

typedef struct orc_engine orc_engine;

typedef struct {
	int model_id;
	int n_ensembles;
	int n_beta;
	int n_par;
	unsigned long long seed;
	int proposal;
	unsigned circular_mask;
	unsigned quirks;
	int rng_kind;            /* ORC_RNG_* */
	int chain_id_offset;
	int ensemble_id_offset;
	double model_const[4];
	int n_threads;           /* OpenMP threads over chains in orc_run/orc_calibrate (PHILOX mode only;
	                            MT19937 mode is single-threaded like a race-free reference run) */
} orc_config;

typedef struct {
	double * beta;
	double * params;
	double * steps;
	double * prob;
	double * prior;
	double * prob_best;
	double * params_best;
	unsigned long long * accept;
	unsigned long long * reject;
	unsigned long long * params_accepts;
	unsigned long long * params_rejects;
	unsigned long long * n_iter;
	unsigned long long * swapcount;
	unsigned long long * rng_counter;
} orc_chain_io;

typedef struct {
	int prob_every;
	int params_chains;
} orc_trace_cfg;

typedef struct {
	unsigned long long burn_in_iterations;
	double desired_acceptance_rate;
	double max_ar_deviation;
	unsigned long long iter_limit;
	double mul;
	double adjust_step;
	int skip_calibrate;
	int iter_readjust;
	int no_rescaling_limit;
} orc_calib_cfg;

typedef struct {
	int chain;
	int param;
	unsigned long long iter;
	double step_normalised;
	double accept_rate;
} orc_calib_progress;

int orc_create(orc_engine ** e, const orc_config * cfg);
int orc_destroy(orc_engine * e);
int orc_set_data(orc_engine * e, const double * rowmajor, long long n_rows, int n_cols);
int orc_set_bounds(orc_engine * e, const double * pmin, const double * pmax);
int orc_set_chains(orc_engine * e, int first, int count, const orc_chain_io * in);
int orc_get_chains(orc_engine * e, int first, int count, orc_chain_io * out);
int orc_eval(orc_engine * e, int n, const double * params, const double * beta,
		double * prob_out, double * prior_out);
int orc_run(orc_engine * e, long long n_rounds, int n_swap, const orc_trace_cfg * trace);
int orc_read_trace(orc_engine * e, double * prob, double * prob_minus_prior,
		double * params, long long * n_prob_rows, long long * n_param_rows);
int orc_calibrate(orc_engine * e, const unsigned char * select,
		const orc_calib_cfg * cfg, int * status, orc_calib_progress * progress,
		long long progress_capacity, long long * n_progress);
/* the inner loop of assess_acceptance_rate (ref src/markov_chain.c:143-172); mirrors apm_gpu_steps */
int orc_steps(orc_engine * e, const unsigned char * select, int kind, long long n_steps,
		unsigned char * accepted);
/* -DADAPT (ref src/parallel_tempering.c:282-302) and -DRANDOMSWAP
 * (ref src/parallel_tempering_interaction.c:47-65,130-131) */
int orc_set_adapt(orc_engine * e, int enabled, double target_acceptance_rate);
int orc_set_random_swap(orc_engine * e, int enabled);
int orc_set_marginals(orc_engine * e, int which_chains, int n_bins, unsigned long long batch_size, int max_batches);
int orc_get_marginals(orc_engine * e, unsigned long long * counts, double * batch_means,
		unsigned long long * n_values, unsigned long long * n_batches); /* mirror apm_gpu_*_marginals */
int orc_reset_stats(orc_engine * e);
int orc_get_stats(orc_engine * e, unsigned long long * n, double * sum_dl,
		double * sum_params, double * sum_params_sq);

/* building blocks exported for unit tests */
void orc_calc_model(int model_id, const double * model_const, const double * params,
		int n_par, const double * data, long long n_rows, int n_cols, double beta,
		double * prob, double * prior);
double orc_mod_double(double x, double div);            /* src/mcmc_internal.h:46-48 */
double orc_get_chain_beta(unsigned i, unsigned n_beta, double beta_0); /* parallel_tempering_beta.c:85-90 */
double orc_calc_beta_0(int n_par, const double * pmin, const double * pmax,
		const double * steps, const double * stepwidth_factors); /* parallel_tempering_beta.c:92-102 */
void orc_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4]);
/* the two 53-bit uniforms in (0,1) of one Philox block */
void orc_philox_uniforms(unsigned long long seed, unsigned c0, unsigned long long step,
		unsigned purpose, unsigned idx, unsigned attempt, double * u0, double * u1);
/* evidence by the reference's rectangle rule, analyse.c:50-93 */
double orc_evidence(int n_beta, const double * beta, const double * mean_dl);
/* GSL-compatible MT19937 access (pins the shim and the oracle to the same stream) */
void orc_mt_seed(orc_engine * e, unsigned long seed);
double orc_mt_uniform(orc_engine * e);
int orc_host_uniform(orc_engine * e, int g, double * u); /* mirrors apm_gpu_host_uniform */

#ifdef __cplusplus
}
#endif
#endif
