/*
 * apemost_gpu.h -- C ABI of the B200 parallel-tempering MCMC engine.
 *
 * This is the drop-in boundary for APEMoST's hot path.  APEMoST has no FFI of
 * its own (it is one C program whose plugin is bound at link time, reference
 * Makefile:54-55), so the entry points below are the calls a maintainer would
 * put in place of the reference's in-process hot loops; each entry cites the
 * reference code it replaces.  Plain pointers and sizes only: no CUDA, torch
 * or GSL types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success, a negative APM_E* code otherwise;
 *     apm_gpu_last_error() gives the text (the reference exit(1)s / asserts
 *     instead, SURVEY.md section 8b "Error convention"; the host wrapper
 *     turns codes back into the same stderr text + exit(1));
 *   - the caller owns every host buffer; the library copies on set_* and owns
 *     all device memory;
 *   - a handle is bound to one CUDA device and is not re-entrant;
 *   - chains are numbered g = ensemble * n_beta + k, k = position in the beta
 *     ladder (k = 0 is beta = 1), matching chains[k] in the reference
 *     (src/parallel_tempering_config.c:95-123) with an ensemble dimension
 *     added in front;
 *   - matrices are row-major: params[g * n_par + j], data[row * n_cols + col]
 *     (= gsl_matrix with tda = n_cols, reference src/mcmc_parser.c:97-122).
 *
 * There is NO CPU fallback: without a usable CUDA device apm_gpu_create fails
 * with APM_ENODEVICE.
 */
#ifndef APEMOST_GPU_H_
#define APEMOST_GPU_H_

#ifdef __cplusplus
extern "C" {
#endif

#define APM_GPU_ABI_VERSION 6

/* ---- error codes -------------------------------------------------------- */
#define APM_OK          0
#define APM_EINVAL     -1   /* bad argument */
#define APM_ENODEVICE  -2   /* no CUDA device / wrong architecture */
#define APM_ECUDA      -3   /* a CUDA runtime call failed */
#define APM_ENOMEM     -4
#define APM_ESTATE     -5   /* call order violated (e.g. run before set_data) */
#define APM_ECALIB     -6   /* calibration failed for >= 1 chain (see status[]) */
#define APM_ENCCL      -7   /* an NCCL call failed */

/* ---- built-in device models: __device__ counterparts of apps/<model>.c --- */
#define APM_MODEL_SIMPLESIN    0  /* reference apps/simplesin.c:12-38 */
#define APM_MODEL_SIMPLESIN5   1  /* reference apps/simplesin5.c:15-41 (formula) */
#define APM_MODEL_NORMAL       2  /* reference apps/normal.c:8-34 (data-free) */
#define APM_MODEL_PULSE_VROT   3  /* reference apps/pulse_vrot.c:12-65 */
#define APM_MODEL_SIMPLESIN2   4  /* reference apps/simplesin2.c */
#define APM_MODEL_PULSE        5  /* reference apps/pulse.c */
#define APM_MODEL_BERNOULLI    6  /* reference apps/bernoulli_example.c:9-51 (n_par = data columns, 2..4) */
#define APM_MODEL_BERNOULLI_EXAMPLE APM_MODEL_BERNOULLI
#define APM_MODEL_USER         100 /* apps/<model>.cuh compiled in with -DAPM_USER_MODEL_HEADER */

/* ---- proposal distribution: reference src/mcmc_gettersetter.c:290-306 ---- */
#define APM_PROPOSAL_GAUSSIAN  0
#define APM_PROPOSAL_LOGISTIC  1  /* -DPROPOSAL_LOGISTIC */
#define APM_PROPOSAL_UNIFORM   2  /* -DPROPOSAL_UNIFORM  */

/* ---- reference quirks.  Set = behave exactly as
 *      the reference does; clear = the statistically exact behaviour. ------- */
#define APM_QUIRK_STALE_PROB_ON_SWAP    1u /* swap exchanges params only, prob stays
                                              (parallel_tempering_interaction.c:99-123) */
#define APM_QUIRK_STALE_PRIOR_ON_REJECT 2u /* revert() restores prob but not prior
                                              (markov_chain.c:313-315,383) */
#define APM_QUIRKS_REFERENCE (APM_QUIRK_STALE_PROB_ON_SWAP | APM_QUIRK_STALE_PRIOR_ON_REJECT)

/* ---- kernel path selection ---------------------------------------------- */
#define APM_PATH_AUTO    0  /* cluster or fused if the data table fits in shared memory, else grid or tiled */
#define APM_PATH_TILED   1  /* (chain tile x row split) likelihood kernel + control kernel per step */
#define APM_PATH_FUSED   2  /* one persistent launch per run: CTA per ensemble, warp per chain */
#define APM_PATH_CLUSTER 3  /* fused, with every ensemble spread over a thread-block cluster (2..8 CTAs,
                               warp groups per chain, table by TMA multicast, swaps through distributed
                               shared memory): for fewer ensembles than SMs.  apm_gpu_run only;
                               calibration takes the fused path */

#define APM_PATH_GRID    4  /* mid-size tables (too large for one SM's shared memory, at most 512 chains):
                               one cooperative launch per run, the table partitioned over the shared
                               memories of all SMs, two grid barriers per step; runs, calibration and
                               apm_gpu_steps */

/* ---- calibration status per chain --------------------------------------- */
#define APM_CALIB_OK             0
#define APM_CALIB_STEP_TOO_LARGE 1  /* "step width became too large", markov_chain_calibrate.c:1104-1110 */
#define APM_CALIB_ITER_LIMIT     2  /* "limit of iterations reached", markov_chain_calibrate.c:1169-1174 */
#define APM_CALIB_NOT_SELECTED   -1

typedef struct apm_gpu apm_gpu;

typedef struct {
	int device;               /* CUDA device ordinal */
	int model_id;             /* APM_MODEL_* */
	int n_ensembles;          /* independent PT ensembles on this device */
	int n_beta;               /* ladder length (N_BETA, define_defaults.h:24-26); may exceed the
	                             reference's 99 (parallel_tempering.c:366) */
	int n_par;                /* parameters per chain; must match the model */
	unsigned long long seed;  /* counter-RNG key (replaces GSL_RNG_SEED, src/mcmc.c:27-35) */
	int proposal;             /* APM_PROPOSAL_* */
	unsigned circular_mask;   /* bit j set: parameter j wraps (CIRCULAR_PARAMS lists j+1,
	                             markov_chain.h:34-46, markov_chain.c:241-262) */
	unsigned quirks;          /* APM_QUIRK_* bit set */
	int path;                 /* APM_PATH_* */
	int chain_id_offset;      /* global id of local chain 0: keeps RNG streams distinct when
	                             ensembles are sharded over several GPUs */
	int ensemble_id_offset;   /* same, for the per-ensemble swap stream */
	double model_const[4];    /* model constants; 0 = default.  [0] = SIGMA (simplesin*) / HMIN (pulse*) */
} apm_gpu_config;

/* Per-chain state, struct of nullable array pointers (NULL = leave alone /
 * do not fetch).  Mirrors the scalar and vector members of `mcmc`
 * (reference src/mcmc_struct.h:30-106) plus parallel_tempering_mcmc
 * {beta, swapcount} (src/parallel_tempering_beta.h:65-76). */
typedef struct {
	double * beta;                       /* [count] */
	double * params;                     /* [count][n_par] */
	double * steps;                      /* [count][n_par]   params_step */
	double * prob;                       /* [count] */
	double * prior;                      /* [count] */
	double * prob_best;                  /* [count] */
	double * params_best;                /* [count][n_par] */
	unsigned long long * accept;         /* [count] */
	unsigned long long * reject;         /* [count] */
	unsigned long long * params_accepts; /* [count][n_par] */
	unsigned long long * params_rejects; /* [count][n_par] */
	unsigned long long * n_iter;         /* [count] */
	unsigned long long * swapcount;      /* [count] */
	unsigned long long * rng_counter;    /* [count] per-chain counter-RNG position */
} apm_gpu_chain_io;

typedef struct {
	int prob_every;    /* record (prob, prob - prior) of every chain each prob_every-th step;
	                      1 = what prob-chain<k>.dump holds (parallel_tempering.c:399-401); 0 = off */
	int params_chains; /* whose parameter vectors to record every step: 0 none, 1 chain 0 of each
	                      ensemble (<name>-chain-0.prob.dump, mcmc_dump.c:79-88), 2 all (DUMP_ALL_CHAINS) */
} apm_gpu_trace_cfg;

/* calibration constants: arguments of markov_chain_calibrate()
 * (reference src/markov_chain_calibrate.c:1182-1204, defaults define_defaults.h) */
typedef struct {
	unsigned long long burn_in_iterations; /* BURN_IN_ITERATIONS 10000 */
	double desired_acceptance_rate;        /* TARGET_ACCEPTANCE_RATE 0.5 (also the global target) */
	double max_ar_deviation;               /* MAX_AR_DEVIATION 0.01 */
	unsigned long long iter_limit;         /* ITER_LIMIT 100000 */
	double mul;                            /* MUL 0.85 */
	double adjust_step;                    /* DEFAULT_ADJUST_STEP 0.5 */
	int skip_calibrate;                    /* 1 = burn_in only (SKIP_CALIBRATE_ALLCHAINS) */
	int iter_readjust;                     /* ITER_READJUST 200 (0 = default) */
	int no_rescaling_limit;                /* NO_RESCALING_LIMIT 15 (0 = default) */
} apm_gpu_calib_cfg;

/* one row of calibration_progress.data (markov_chain_calibrate.c:1143-1146) */
typedef struct {
	int chain;
	int param;
	unsigned long long iter;
	double step_normalised;
	double accept_rate;
} apm_gpu_calib_progress;

/* ---- lifecycle ---------------------------------------------------------- */
int apm_gpu_create(apm_gpu ** handle, const apm_gpu_config * cfg);
int apm_gpu_destroy(apm_gpu * handle);
const char * apm_gpu_last_error(const apm_gpu * handle); /* handle may be NULL: last create error */
int apm_gpu_abi_version(void);
int apm_gpu_model_n_par(int model_id);   /* 0 = any */
int apm_gpu_model_n_cols(int model_id);  /* 0 = data-free, -1 = one column per parameter */

/* ---- inputs -------------------------------------------------------------
 * set_data replaces mcmc_load_data + mcmc_reuse_data (src/mcmc_parser.c:97-146):
 * one table per device, shared read-only by every chain.  The device keeps the columns the
 * model reads (at most 4; two-column models: rows of 2 doubles = the gsl_matrix layout).  row_offset/n_rows_total
 * describe a contiguous shard of a larger table (data-sharded mode). */
int apm_gpu_set_data(apm_gpu * h, const double * rowmajor, long long n_rows,
		int n_cols);
int apm_gpu_set_bounds(apm_gpu * h, const double * params_min,
		const double * params_max); /* [n_par] each; params file columns 2,3 */
int apm_gpu_set_chains(apm_gpu * h, int first, int count,
		const apm_gpu_chain_io * in);
int apm_gpu_get_chains(apm_gpu * h, int first, int count,
		apm_gpu_chain_io * out);

/* ---- calc_model for n parameter vectors (parity hook; replaces the plugin
 * call at src/markov_chain.c:376 and apps/eval_main.c:60-63) --------------- */
int apm_gpu_eval(apm_gpu * h, int n, const double * params /*[n][n_par]*/,
		const double * beta /*[n]*/, double * prob_out /*[n]*/,
		double * prior_out /*[n]*/);

/* ---- the sampler: n_rounds x { n_swap x markov_chain_step + mcmc_check_best
 * + append for every chain; one tempering_interaction per ensemble }
 * (replaces the loop at src/parallel_tempering.c:392-409) ------------------ */
int apm_gpu_run(apm_gpu * h, long long n_rounds, int n_swap,
		const apm_gpu_trace_cfg * trace);
/* copies out what the last apm_gpu_run recorded; any pointer may be NULL.
 * prob/prob_minus_prior: [n_steps/prob_every][n_chains];
 * params: [n_steps][n_dumped][n_par], n_dumped = 0 / n_ensembles / n_chains */
int apm_gpu_read_trace(apm_gpu * h, double * prob, double * prob_minus_prior,
		double * params, long long * n_prob_rows, long long * n_param_rows);

/* ---- markov_chain_calibrate for a selection of chains, all concurrently
 * (replaces src/markov_chain_calibrate.c:1182-1204 -> burn_in markov_chain.c:34-79
 * -> markov_chain_calibrate_orig :1039-1180).  select[g] != 0 picks chain g
 * (NULL = all).  status[g] receives APM_CALIB_*; returns APM_ECALIB if any
 * selected chain failed.  progress (may be NULL) receives up to
 * progress_capacity rows; *n_progress the number produced. */
int apm_gpu_calibrate(apm_gpu * h, const unsigned char * select,
		const apm_gpu_calib_cfg * cfg, int * status,
		apm_gpu_calib_progress * progress, long long progress_capacity,
		long long * n_progress);

/* ---- n_steps x { markov_chain_step_for(kind) -- or markov_chain_step if kind == n_par --;
 * mcmc_check_best } for the selected chains (NULL = all), in lockstep: the inner loop of
 * assess_acceptance_rate (reference src/markov_chain.c:143-172), on which the alternate
 * calibrators are built (src/markov_chain_calibrate.c:33-1037).  accepted[step * n_chains + g]
 * receives 1 where chain g's step was accepted (what assess_acceptance_rate keeps in its bit
 * field), 0 otherwise and for chains not selected. --------------------------------------- */
int apm_gpu_steps(apm_gpu * h, const unsigned char * select, int kind,
		long long n_steps, unsigned char * accepted /*[n_steps][n_chains]*/);

/* ---- one uniform in (0, 1) from chain g's random stream, for host-side algorithms: what the
 * regression calibrator draws with gsl_rng_uniform(get_random(m)) between its assessments
 * (reference src/markov_chain_calibrate.c:93-94,155).  A stream of its own per chain (counter
 * RNG, purpose 4), so the draws do not shift the chain's step streams; host work only. ------- */
int apm_gpu_host_uniform(apm_gpu * h, int g, double * u);

/* ---- adapt() as compiled with -DADAPT (reference src/parallel_tempering.c:282-302, called
 * once per round before the swap, :404): a chain whose per-parameter accept + reject counter
 * sums have reached 20000 scales all its step widths by 0.99 (accepts / rejects below
 * target - 0.05) or 1 / 0.99 (above target + 0.05) and resets its counters beyond 100000.
 * Off by default, like the reference. ------------------------------------------------------ */
int apm_gpu_set_adapt(apm_gpu * h, int enabled, double target_acceptance_rate);

/* ---- on-device accumulators: what analyse needs
 * without the text round trip.  Per chain: n = recorded steps, sum_dl =
 * sum of (prob - prior) (analyse.c:50-93 divides its mean by beta), and
 * first/second moments of every parameter. ------------------------------- */
int apm_gpu_reset_stats(apm_gpu * h);
int apm_gpu_get_stats(apm_gpu * h, unsigned long long * n /*[n_chains]*/,
		double * sum_dl /*[n_chains]*/, double * sum_params /*[n_chains][n_par]*/,
		double * sum_params_sq /*[n_chains][n_par]*/);

/* ---- marginal statistics on the device: what the reference's analyse
 * derives from <name>-chain-<i>.prob.dump, accumulated while the sampler runs instead of parsed from
 * hundreds of MB of text (calc_marginal_distribution + calc_mcmc_error, reference src/analyse.c:115-247,
 * create_hist src/histogram.c:34-43).  For the chosen chains (which_chains: 0 off, 1 = chain 0 of every
 * ensemble, 2 = every chain; one "slot" per chosen chain, in chain order) and every parameter:
 *   counts       gsl_histogram_increment's bins: n_bins uniform bins over [params_min, params_max], the
 *                last edge widened by (max - min) / 10000 like create_hist; edges are computed with
 *                gsl_histogram_set_ranges_uniform's expressions, so a value lands in the same bin;
 *   batch_means  the means calc_mcmc_error forms: a batch closes when (values so far) % batch_size ==
 *                batch_size - 1 and its mean is its sum / batch_size (so the first batch holds
 *                batch_size - 1 values; the reference's arithmetic, kept); the first max_batches are kept.
 * set_marginals (re)allocates and zeroes the accumulators; they then collect every recorded step of
 * every apm_gpu_run until the next call.  get_marginals: any pointer may be NULL.
 *   counts [slots][n_par][n_bins], batch_means [slots][n_par][max_batches], n_values / n_batches [slots] */
int apm_gpu_set_marginals(apm_gpu * h, int which_chains, int n_bins, unsigned long long batch_size,
		int max_batches);
int apm_gpu_get_marginals(apm_gpu * h, unsigned long long * counts, double * batch_means,
		unsigned long long * n_values, unsigned long long * n_batches);

/* ---- multi-GPU, data-sharded likelihood: every rank
 * holds all chains and a contiguous row shard; per step the per-chain partial
 * sums are all-reduced with NCCL (fp64 sum) and every rank takes the identical
 * accept/swap decision.  nccl_unique_id is the 128-byte ncclUniqueId produced
 * by apm_gpu_nccl_unique_id() on rank 0 and distributed by the caller. ------ */
int apm_gpu_nccl_unique_id(unsigned char id_out[128]);
int apm_gpu_nccl_init(apm_gpu * h, const unsigned char id[128], int rank,
		int n_ranks);

/* ---- multi-GPU, ladder split (third row): the n_beta_total rungs of
 * every ensemble are dealt out over n_ranks GPUs in contiguous blocks (rank r holds rungs
 * [n_beta_total * r / n_ranks, n_beta_total * (r + 1) / n_ranks); the handle must have been
 * created with that many as n_beta, the same n_ensembles / seed / offsets on every rank).
 * Chains step independently; once per round the ranks trade the swap-relevant state of
 * their boundary chains (prob, beta, prior, best, params: 4 + 2 n_par doubles per ensemble)
 * with ncclSend/ncclRecv, take the identical decision for a pair that straddles two GPUs
 * (tempering_interaction, src/parallel_tempering_interaction.c:125-141) and each updates
 * the chain it owns.  Results equal the single-GPU run bit for bit. ------------------- */
int apm_gpu_ladder_init(apm_gpu * h, const unsigned char id[128], int rank,
		int n_ranks, int n_beta_total);

/* ---- introspection used by the tests and bench.py ----------------------- */
/* number of kernels this handle has launched so far */
long long apm_gpu_launch_count(const apm_gpu * h);
/* device time in ms of the likelihood kernels of the last run/eval (CUDA events
 * on the engine's stream), and how many launches that covers */
int apm_gpu_last_kernel_ms(const apm_gpu * h, double * loglik_ms,
		long long * loglik_launches, double * total_ms);
/* which path the last run used: APM_PATH_TILED, _FUSED, _CLUSTER or _GRID */
int apm_gpu_last_path(const apm_gpu * h);
/* per_launch != 0: bracket every likelihood launch of the tiled path with CUDA events (what
 * apm_gpu_last_kernel_ms reports as loglik_ms / loglik_launches) and enqueue the launches one by one.
 * Default 0: no per-launch events, and a round of the sampler (n_swap x {likelihood, control kernel})
 * is replayed as one CUDA graph -- the step counter lives on the device, the host has nothing to do
 * per step; loglik_launches is then reported as 0. */
int apm_gpu_set_timing(apm_gpu * h, int per_launch);
/* microbenchmark: sustained FP64 FMA issue rate of this device, in
 * FP64 instructions (lane-operations) per second; used as the roofline peak */
int apm_gpu_measure_fp64_peak(int device, double seconds, double * instr_per_s);
/* the same per clock: FP64 lane-operations per SM per clock (64 by the architecture), from clock64()
 * inside a short DFMA kernel -- independent of the clocks, which sag under the long pure-DFMA load the
 * per-second microbenchmark applies; multiplied by the SM count and the SM clock during a run it is that
 * run's FP64 issue peak */
int apm_gpu_measure_fp64_per_clock(int device, double * lanes_per_sm_per_clock, int * n_sm);

#ifdef __cplusplus
}
#endif

#endif /* APEMOST_GPU_H_ */
