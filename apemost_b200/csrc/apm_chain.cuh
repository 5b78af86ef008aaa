// apm_chain.cuh -- device-side chain state and the per-chain Metropolis / parallel
// tempering / calibration logic.  These __device__ functions restate, one to one, the
// reference's control code; both kernel families (the tiled path's control kernel and
// the fused persistent kernel) call them, so the semantics exist exactly once.
#pragma once

#include "apm_math.cuh"
#include "apm_models.cuh"

namespace apm {

typedef unsigned long long u64;

// what evaluation is pending for a chain
#define PEND_NONE (-1)
// 0 .. n_par-1 : single-parameter step (markov_chain_step_for), n_par : full step

// calibration state machine phases (markov_chain_calibrate = burn_in + _orig)
enum {
	CAL_IDLE = 0, CAL_BURN_A = 1, CAL_BURN_B = 2, CAL_SINGLE = 3, CAL_FULL = 4, CAL_DONE = 5,
	CAL_STEPS = 6 // apm_gpu_steps: a fixed number of steps of one kind, accept log kept
};

struct CalState {
	int phase;
	int status;
	int sub;           // steps done in the current block / parameter index within an iteration
	int nchecks_without_rescaling;
	int rescaled;
	int reached_perfection;
	u64 iter;
	double rat_limit;
	double saved_steps[APM_MAX_PAR];
	// CAL_STEPS
	long long steps_left, steps_done;
	u64 last_count;
};

struct CalibCfgDev {
	u64 burn_in_iterations;
	double desired_acceptance_rate;
	double max_ar_deviation;
	u64 iter_limit;
	double mul;
	double adjust_step;
	int skip_calibrate;
	int iter_readjust;
	int no_rescaling_limit;
	// apm_gpu_steps (steps_n > 0): steps_n x { step of kind steps_kind; mcmc_check_best } instead of a
	// calibration; steps_kind = parameter index (markov_chain_step_for) or n_par (markov_chain_step)
	int steps_kind;
	long long steps_n;
};

struct ProgressRow {
	int chain;
	int param;
	u64 iter;
	double step_normalised;
	double accept_rate;
};

// All per-chain arrays are chain-major: v[g * n_par + j].
struct DevState {
	int n_chains, n_beta, n_ens, n_par;
	u64 seed;
	int proposal;
	unsigned circular_mask, quirks;
	int chain_id_offset, ensemble_id_offset;
	// ladder split over several GPUs (otherwise n_beta_total = id_stride = n_beta, k_offset = 0):
	// this device holds rungs [k_offset, k_offset + n_beta) of every ensemble's n_beta_total
	int n_beta_total, id_stride, k_offset;
	int g_base;        // added to a chain index where it is reported (progress rows); != 0 only in
	                   // the fused path, whose DevState is indexed by position in the ensemble
	double model_const[4];
	// -DADAPT (reference src/parallel_tempering.c:282-302): 0 = off
	int adapt;
	double adapt_target;
	// mcmc struct members (reference src/mcmc_struct.h:30-106)
	double * params, *params_best, *steps;
	double * prob, *prior, *prob_best;
	u64 * accept, *reject, *pacc, *prej, *n_iter;
	// parallel_tempering_mcmc (reference src/parallel_tempering_beta.h:65-76)
	double * beta;
	u64 * swapcount;
	// bounds (identical for all chains: one params file)
	double * pmin, *pmax;
	// counter-RNG positions
	u64 * rng_ctr;     // per chain
	u64 * swap_round;  // per ensemble
	// pending proposal
	double * prop;     // [n_chains][n_par]
	int * pend;        // [n_chains]
	// likelihood partial sums written by the tiled kernel: [n_chains][n_splits]
	double * partial;
	int n_splits;
	// accumulators
	u64 * stat_n;
	double * stat_sum_dl, *stat_sum_p, *stat_sum_p2;
	// calibration
	CalState * cal;
	ProgressRow * progress;
	long long progress_cap;
	unsigned long long * progress_n;
	int * n_active;    // chains still calibrating (device counter)
	unsigned char * alog; // apm_gpu_steps: [steps][n_chains] 1 = the step was accepted
	// compacted list of the chains whose proposal is pending (calibration: the likelihood
	// kernel only visits these).  Two buffers / counters used alternately: the control kernel
	// of step s fills [w] for the next likelihood launch and clears [1 - w].
	int * act_idx;     // [2][n_chains]
	int * act_n;       // [2]
	// trace of the current run
	double * tr_prob, *tr_dl, *tr_params;
	int tr_prob_every, tr_params_chains, tr_dumped;
	// marginal statistics of the recorded chains (apm_gpu_set_marginals): what the
	// reference's analyse derives from <name>-chain-<i>.prob.dump (src/analyse.c:115-247) -- per
	// parameter a histogram of marg_bins uniform bins over [min, max] and the means of consecutive
	// batches of marg_batch values.  marg_mode: 0 off, 1 rung 0 of every ensemble, 2 every chain.
	int marg_mode, marg_bins, marg_cap;
	unsigned long long marg_batch;
	unsigned long long * marg_counts; // [slots][n_par][marg_bins]
	double * marg_bsum;               // [slots][n_par]  running sum of the open batch
	double * marg_means;              // [slots][n_par][marg_cap]
	unsigned long long * marg_n;      // [slots] values seen
	unsigned long long * marg_nb;     // [slots] batches closed
	// tiled path: the run's step counter lives on the device, so that a whole round of launches can be
	// replayed as a CUDA graph without any per-step argument from the host.  run_ctr[0] = steps recorded
	// so far in this apm_gpu_run call, run_ctr[1] = the control kernel's "blocks done" ticket.
	unsigned long long * run_ctr;
};

// global number of chain g = its random stream (independent of how chains are spread over
// devices, processes, ensembles per device or kernel paths)
APM_D uint32_t chain_rng_id(const DevState & S, int g) {
	return (uint32_t) (S.chain_id_offset + (g / S.n_beta) * S.id_stride + S.k_offset + g % S.n_beta);
}

// ---- proposal: reference src/markov_chain.c:226-270 (do_step_for) ------------------------
// Kept out of line: Philox + log + sqrt + cos, called from several places of the fused / cluster
// kernels, would otherwise be inlined into (and bloat) their step loops.
// (first_attempt > 0: the caller has tried the attempts before it -- all outside the bounds of a
// parameter that is not circular -- and the redraws carry on from there)
__device__ __noinline__ double propose_coordinate_core(u64 seed, uint32_t id, u64 ctr, int i, double old_value,
		double step, int proposal, bool wrap, double mn, double mx, unsigned first_attempt = 0) {
	unsigned attempt = first_attempt;
	double u0, u1, new_value;
	philox_uniforms(seed, id, ctr, PURPOSE_JUMP, (uint32_t) i, attempt++, u0, u1);
	new_value = old_value + jump_from_uniforms(proposal, step, u0, u1);
	if (new_value > mx || new_value < mn) {
		if (wrap) {
			// CIRCULAR_PARAMS lists this parameter (:257-260)
			new_value = mn + mod_double(new_value - mn, mx - mn);
		} else {
			// redraw until inside the bounds (:235-240; :241-262 for the unlisted parameters)
			do {
				philox_uniforms(seed, id, ctr, PURPOSE_JUMP, (uint32_t) i, attempt++, u0, u1);
				new_value = old_value + jump_from_uniforms(proposal, step, u0, u1);
			} while (new_value > mx || new_value < mn);
		}
	}
	return new_value;
}

APM_D double propose_coordinate(const DevState & S, int g, u64 ctr, int i, double old_value,
		double step, unsigned first_attempt = 0) {
	return propose_coordinate_core(S.seed, chain_rng_id(S, g), ctr, i, old_value, step, S.proposal,
			(S.circular_mask >> i) & 1u, S.pmin[i], S.pmax[i], first_attempt);
}

// do_step (:272-277) / do_step_for for the pending kind; writes S.prop[g]
APM_D void chain_propose(const DevState & S, int g, int kind) {
	const int n = S.n_par;
	const u64 ctr = S.rng_ctr[g];
	const double * p = S.params + (size_t) g * n;
	const double * st = S.steps + (size_t) g * n;
	double * q = S.prop + (size_t) g * n;
	for (int i = 0; i < n; i++) {
		if (kind == n || kind == i)
			q[i] = propose_coordinate(S, g, ctr, i, p[i], st[i]);
		else
			q[i] = p[i];
	}
	S.pend[g] = kind;
}

// mcmc_check_best: reference src/mcmc_calculate.c:35-41
APM_D void chain_check_best(const DevState & S, int g) {
	if (S.prob[g] > S.prob_best[g]) {
		S.prob_best[g] = S.prob[g];
		for (int i = 0; i < S.n_par; i++)
			S.params_best[(size_t) g * S.n_par + i] = S.params[(size_t) g * S.n_par + i];
	}
}

// restart_from_best: reference src/markov_chain.c:29-32
APM_D void chain_restart_from_best(const DevState & S, int g) {
	for (int i = 0; i < S.n_par; i++)
		S.params[(size_t) g * S.n_par + i] = S.params_best[(size_t) g * S.n_par + i];
	S.prob[g] = S.prob_best[g];
}

// reset_accept_rejects: reference src/mcmc_gettersetter.c:119-127
APM_D void chain_reset_accept_rejects(const DevState & S, int g) {
	for (int i = 0; i < S.n_par; i++) {
		S.pacc[(size_t) g * S.n_par + i] = 0;
		S.prej[(size_t) g * S.n_par + i] = 0;
	}
	S.reject[g] = 0;
	S.accept[g] = 0;
}

// The second half of markov_chain_step / markov_chain_step_for (reference
// src/markov_chain.c:369-386, 317-333) once the model's running sum for the proposal
// is known: set_prior/set_prob, check_accept (:282-311), bookkeeping.
// `pre_logu`, when given, is log(u0) of this step's accept draw, computed ahead of time by another
// warp from the same counter (the draw depends on the chain's id and step counter only).
// (chain_finalize_value: the same from the point where the proposal's prob and prior are known)
APM_D void chain_finalize_value(const DevState & S, int g, double prob_new, double prior_new, const double * pre_logu) {
	const int n = S.n_par;
	const int kind = S.pend[g];
	const double * q = S.prop + (size_t) g * n;
	double * p = S.params + (size_t) g * n;
	const double prob_old = S.prob[g];
	int accepted;
	if (prob_new == prob_old)
		accepted = 1;
	else if (prob_new > prob_old)
		accepted = 1;
	else {
		// get_next_alog_urandom: reference src/mcmc_gettersetter.c:307-309
		double logu;
		if (pre_logu != nullptr) {
			logu = *pre_logu;
		} else {
			double u0, u1;
			philox_uniforms(S.seed, chain_rng_id(S, g), S.rng_ctr[g], PURPOSE_ACCEPT, 0, 0, u0, u1);
			logu = log(u0);
		}
		accepted = logu < (prob_new - prob_old) ? 1 : 0;
	}
	if (accepted) {
		S.prob[g] = prob_new;
		S.prior[g] = prior_new;
		if (kind == n) {
			for (int i = 0; i < n; i++)
				p[i] = q[i];
			S.accept[g]++; // inc_params_accepts: src/mcmc_gettersetter.c:98-103
			for (int i = 0; i < n; i++)
				S.pacc[(size_t) g * n + i]++;
		} else {
			p[kind] = q[kind];
			S.pacc[(size_t) g * n + kind]++;
		}
	} else {
		// revert() restores prob only (:313-315)
		if (S.quirks & 2u)
			S.prior[g] = prior_new;
		if (kind == n) {
			S.reject[g]++;
			for (int i = 0; i < n; i++)
				S.prej[(size_t) g * n + i]++;
		} else {
			S.prej[(size_t) g * n + kind]++;
		}
	}
	S.rng_ctr[g]++;
	S.pend[g] = PEND_NONE;
}

template<class M>
APM_D void chain_finalize(const DevState & S, int g, double sum, const double * pre_logu = nullptr) {
	const double * q = S.prop + (size_t) g * S.n_par;
	double prior_new = S.prior[g];
	if (M::HAS_PRIOR)
		prior_new = M::prior(q, S.n_par, S.model_const);
	chain_finalize_value(S, g, M::finish(S.beta[g], sum, prior_new, q, S.model_const), prior_new, pre_logu);
}

// ---- marginal statistics ---------------------------------------------------
// which slot of the marg_* arrays chain g feeds (-1: none); same rule as the parameter dump
APM_D int marg_slot(const DevState & S, int g) {
	if (S.marg_mode == 2)
		return g;
	if (S.marg_mode == 1 && S.k_offset + g % S.n_beta == 0)
		return g / S.n_beta;
	return -1;
}
// edge k of gsl_histogram_set_ranges_uniform(h, lo, hi) with create_hist's widened last edge
// (reference src/histogram.c:34-43): the same expressions, so the same doubles
APM_D double marg_edge(int k, int nb, double lo, double hi) {
	const double f1 = (double) (nb - k) / (double) nb, f2 = (double) k / (double) nb;
	double e = add_rn(mul_rn(f1, lo), mul_rn(f2, hi));
	if (k == nb)
		e = add_rn(e, (hi - lo) / 10000);
	return e;
}
// one value of parameter i of the chain in `slot`: gsl_histogram_increment (the bin with
// edge[b] <= v < edge[b + 1]; outside all bins: not counted) and calc_mcmc_error's batches
// (reference src/analyse.c:115-142: a batch closes when (values so far) % batch == batch - 1, its
// mean is its sum over the batch SIZE).  seen / closed = the slot's counters before this step.
APM_D void marg_add_value(const DevState & S, int slot, int i, double v, unsigned long long seen,
		unsigned long long closed) {
	const int nb = S.marg_bins, n = S.n_par;
	const double lo = S.pmin[i], hi = S.pmax[i], top = marg_edge(nb, nb, lo, hi);
	if (v >= lo && v < top) {
		int b = (int) ((v - lo) / (top - lo) * nb);
		b = b < 0 ? 0 : (b > nb - 1 ? nb - 1 : b);
		while (b > 0 && v < marg_edge(b, nb, lo, hi))
			b--;
		while (b < nb - 1 && v >= marg_edge(b + 1, nb, lo, hi))
			b++;
		S.marg_counts[((size_t) slot * n + i) * nb + b]++;
	}
	double & bs = S.marg_bsum[(size_t) slot * n + i];
	bs += v;
	if (S.marg_batch > 0 && (seen + 1) % S.marg_batch == S.marg_batch - 1) {
		if (closed < (unsigned long long) S.marg_cap)
			S.marg_means[((size_t) slot * n + i) * S.marg_cap + closed] = bs / (double) S.marg_batch;
		bs = 0;
	}
}
// after all parameters of the step went in
APM_D void marg_step_done(const DevState & S, int slot, unsigned long long seen, unsigned long long closed) {
	S.marg_n[slot] = seen + 1;
	if (S.marg_batch > 0 && (seen + 1) % S.marg_batch == S.marg_batch - 1)
		S.marg_nb[slot] = closed + 1;
}
APM_D void chain_marginals(const DevState & S, int g, const double * params) {
	const int slot = marg_slot(S, g);
	if (slot < 0)
		return;
	const unsigned long long seen = S.marg_n[slot], closed = S.marg_nb[slot];
	for (int i = 0; i < S.n_par; i++)
		marg_add_value(S, slot, i, params[i], seen, closed);
	marg_step_done(S, slot, seen, closed);
}

// the bookkeeping of one sampler iteration after the step (reference
// src/parallel_tempering.c:396-401): check_best, append (n_iter++), the prob-chain
// line and the parameter dump, plus the on-device accumulators
APM_D void chain_record(const DevState & S, int g, long long step_index) {
	const int n = S.n_par;
	chain_check_best(S, g);
	S.n_iter[g]++;
	const double prob = S.prob[g], dl = S.prob[g] - S.prior[g];
	if (S.tr_prob_every > 0 && S.tr_prob != nullptr && step_index % S.tr_prob_every == 0) {
		long long row = step_index / S.tr_prob_every;
		S.tr_prob[row * S.n_chains + g] = prob;
		S.tr_dl[row * S.n_chains + g] = dl;
	}
	if (S.tr_params != nullptr) {
		int slot = -1;
		if (S.tr_params_chains == 2)
			slot = g;
		else if (S.tr_params_chains == 1 && S.k_offset + g % S.n_beta == 0)
			slot = g / S.n_beta; // rung 0 of the whole ladder (k_offset != 0: a block of a split ladder)
		if (slot >= 0)
			for (int i = 0; i < n; i++)
				S.tr_params[((size_t) step_index * S.tr_dumped + slot) * n + i] =
						S.params[(size_t) g * n + i];
	}
	S.stat_n[g]++;
	S.stat_sum_dl[g] += dl;
	for (int i = 0; i < n; i++) {
		double v = S.params[(size_t) g * n + i];
		S.stat_sum_p[(size_t) g * n + i] += v;
		S.stat_sum_p2[(size_t) g * n + i] += v * v;
	}
	if (S.marg_mode)
		chain_marginals(S, g, S.params + (size_t) g * n);
}

// chain_finalize_value's state transition + chain_record for a FULL step whose outcome was decided
// elsewhere (free_run_kernel: the lanes that play a chain hand each step's outcome to a
// book-keeping thread, which writes it down one batch behind).  The chain's point after the step
// (params_after, prob_after, prior_after) comes with the outcome; S.params / S.prob / S.prior
// belong to the deciding lanes and are not touched here.
APM_D void chain_book_step(const DevState & S, int g, int accepted, double prob_after, double prior_after,
		const double * params_after, long long step_index) {
	const int n = S.n_par;
	if (accepted) {
		S.accept[g]++;
		for (int i = 0; i < n; i++)
			S.pacc[(size_t) g * n + i]++;
	} else {
		S.reject[g]++;
		for (int i = 0; i < n; i++)
			S.prej[(size_t) g * n + i]++;
	}
	S.rng_ctr[g]++;
	if (prob_after > S.prob_best[g]) { // mcmc_check_best
		S.prob_best[g] = prob_after;
		for (int i = 0; i < n; i++)
			S.params_best[(size_t) g * n + i] = params_after[i];
	}
	S.n_iter[g]++;
	const double dl = prob_after - prior_after;
	if (S.tr_prob_every > 0 && S.tr_prob != nullptr && step_index % S.tr_prob_every == 0) {
		long long row = step_index / S.tr_prob_every;
		S.tr_prob[row * S.n_chains + g] = prob_after;
		S.tr_dl[row * S.n_chains + g] = dl;
	}
	if (S.tr_params != nullptr) {
		int slot = -1;
		if (S.tr_params_chains == 2)
			slot = g;
		else if (S.tr_params_chains == 1 && S.k_offset + g % S.n_beta == 0)
			slot = g / S.n_beta;
		if (slot >= 0)
			for (int i = 0; i < n; i++)
				S.tr_params[((size_t) step_index * S.tr_dumped + slot) * n + i] = params_after[i];
	}
	S.stat_n[g]++;
	S.stat_sum_dl[g] += dl;
	for (int i = 0; i < n; i++) {
		const double v = params_after[i];
		S.stat_sum_p[(size_t) g * n + i] += v;
		S.stat_sum_p2[(size_t) g * n + i] += v * v;
	}
	if (S.marg_mode)
		chain_marginals(S, g, params_after);
}

// chain_book_step for a whole batch of `ns` consecutive full steps of chain g (free_run_kernel's
// book-keepers): the counters move once per batch, the accumulators are carried in registers over
// the batch -- added to in step order, so the sums are the ones step-by-step book-keeping gives.
// Step j's outcome: ring[j * stride_j + f * stride_f], f = 0 accepted, 1 prob, 2 prior, 3 + i the
// point after the step.  mcmc_check_best has been done by whoever decided the steps.
APM_D void chain_book_batch(const DevState & S, int g, int ns, const double * ring, size_t stride_j, size_t stride_f,
		long long step0) {
	const int n = S.n_par;
	constexpr int NR = 4; // parameters whose sums are carried in registers
	unsigned long long acc = 0;
	double sum_dl = S.stat_sum_dl[g];
	double sp[NR], sp2[NR];
#pragma unroll
	for (int i = 0; i < NR; i++) {
		sp[i] = i < n ? S.stat_sum_p[(size_t) g * n + i] : 0.0;
		sp2[i] = i < n ? S.stat_sum_p2[(size_t) g * n + i] : 0.0;
	}
	int slot = -1;
	if (S.tr_params != nullptr) {
		if (S.tr_params_chains == 2)
			slot = g;
		else if (S.tr_params_chains == 1 && S.k_offset + g % S.n_beta == 0)
			slot = g / S.n_beta;
	}
	const bool trace_prob = S.tr_prob_every > 0 && S.tr_prob != nullptr;
	for (int j = 0; j < ns; j++) {
		const double * e = ring + (size_t) j * stride_j;
		const long long step_index = step0 + j;
		const double prob = e[stride_f], prior = e[2 * stride_f], dl = prob - prior;
		acc += e[0] != 0.0;
		if (trace_prob && step_index % S.tr_prob_every == 0) {
			const long long row = step_index / S.tr_prob_every;
			S.tr_prob[row * S.n_chains + g] = prob;
			S.tr_dl[row * S.n_chains + g] = dl;
		}
		sum_dl += dl;
#pragma unroll
		for (int i = 0; i < NR; i++)
			if (i < n) {
				const double v = e[(size_t) (3 + i) * stride_f];
				sp[i] += v;
				sp2[i] += v * v;
				if (slot >= 0)
					S.tr_params[((size_t) step_index * S.tr_dumped + slot) * n + i] = v;
			}
		for (int i = NR; i < n; i++) {
			const double v = e[(size_t) (3 + i) * stride_f];
			S.stat_sum_p[(size_t) g * n + i] += v;
			S.stat_sum_p2[(size_t) g * n + i] += v * v;
			if (slot >= 0)
				S.tr_params[((size_t) step_index * S.tr_dumped + slot) * n + i] = v;
		}
		if (S.marg_mode) {
			double pa[APM_MAX_PAR];
			for (int i = 0; i < n; i++)
				pa[i] = e[(size_t) (3 + i) * stride_f];
			chain_marginals(S, g, pa);
		}
	}
	const unsigned long long rej = (unsigned long long) ns - acc;
	S.accept[g] += acc;
	S.reject[g] += rej;
	for (int i = 0; i < n; i++) {
		S.pacc[(size_t) g * n + i] += acc;
		S.prej[(size_t) g * n + i] += rej;
	}
	S.rng_ctr[g] += (unsigned long long) ns;
	S.n_iter[g] += (unsigned long long) ns;
	S.stat_n[g] += (unsigned long long) ns;
	S.stat_sum_dl[g] = sum_dl;
#pragma unroll
	for (int i = 0; i < NR; i++)
		if (i < n) {
			S.stat_sum_p[(size_t) g * n + i] = sp[i];
			S.stat_sum_p2[(size_t) g * n + i] = sp2[i];
		}
}

// ---- warp-cooperative forms of chain_finalize / chain_record (fused and cluster paths) ------
// The same state transitions, element for element, with the per-parameter work spread over
// lanes 0..n-1 of the calling warp and the scalar work on lane 0, so that the serial tail of a
// step is a few dozen instructions instead of several hundred.  `sum` and `pre_logu` need to be
// valid on lane 0 only.  All 32 lanes must call.
// (pre_prior, when given: the proposal's prior, computed ahead by another warp -- it depends on the
// proposal only)
template<class M>
APM_D void chain_finalize_warp(const DevState & S, int g, double sum, const double * pre_logu, int lane,
		const double * pre_prior = nullptr) {
	const int n = S.n_par;
	const int kind = S.pend[g];
	const double * q = S.prop + (size_t) g * n;
	double * p = S.params + (size_t) g * n;
	int accepted = 0;
	if (lane == 0) {
		const double prob_old = S.prob[g];
		const double prior_old = S.prior[g];
		double prior_new = prior_old;
		if (M::HAS_PRIOR)
			prior_new = pre_prior != nullptr ? *pre_prior : M::prior(q, n, S.model_const);
		const double prob_new = M::finish(S.beta[g], sum, prior_new, q, S.model_const);
		if (prob_new == prob_old)
			accepted = 1;
		else if (prob_new > prob_old)
			accepted = 1;
		else {
			double logu;
			if (pre_logu != nullptr) {
				logu = *pre_logu;
			} else {
				double u0, u1;
				philox_uniforms(S.seed, chain_rng_id(S, g), S.rng_ctr[g], PURPOSE_ACCEPT, 0, 0, u0, u1);
				logu = log(u0);
			}
			accepted = logu < (prob_new - prob_old) ? 1 : 0;
		}
		if (accepted) {
			S.prob[g] = prob_new;
			S.prior[g] = prior_new;
			if (kind == n)
				S.accept[g]++;
		} else {
			if (S.quirks & 2u)
				S.prior[g] = prior_new;
			if (kind == n)
				S.reject[g]++;
		}
	}
	accepted = __shfl_sync(0xffffffffu, accepted, 0);
	if (lane < n && (kind == n || kind == lane)) {
		if (accepted) {
			p[lane] = q[lane];
			S.pacc[(size_t) g * n + lane]++;
		} else {
			S.prej[(size_t) g * n + lane]++;
		}
	}
	__syncwarp();
	if (lane == 0) {
		S.rng_ctr[g]++;
		S.pend[g] = PEND_NONE;
	}
	__syncwarp();
}

APM_D void chain_record_warp(const DevState & S, int g, long long step_index, int lane) {
	const int n = S.n_par; // <= APM_MAX_PAR = 16, so lanes n .. n + 3 exist
	const double prob = S.prob[g], dl = S.prob[g] - S.prior[g];
	const bool better = prob > S.prob_best[g]; // mcmc_check_best
	const double v = lane < n ? S.params[(size_t) g * n + lane] : 0.0;
	const int mslot = S.marg_mode ? marg_slot(S, g) : -1;
	const unsigned long long m_seen = mslot >= 0 ? S.marg_n[mslot] : 0, m_closed = mslot >= 0 ? S.marg_nb[mslot] : 0;
	__syncwarp();
	if (mslot >= 0) { // the marginal statistics: a parameter per lane, the counters by one lane afterwards
		if (lane < n)
			marg_add_value(S, mslot, lane, v, m_seen, m_closed);
		__syncwarp();
		if (lane == 0)
			marg_step_done(S, mslot, m_seen, m_closed);
	}
	if (lane < n) {
		if (better)
			S.params_best[(size_t) g * n + lane] = v;
		if (S.tr_params != nullptr) {
			int slot = -1;
			if (S.tr_params_chains == 2)
				slot = g;
			else if (S.tr_params_chains == 1 && S.k_offset + g % S.n_beta == 0)
				slot = g / S.n_beta;
			if (slot >= 0)
				S.tr_params[((size_t) step_index * S.tr_dumped + slot) * n + lane] = v;
		}
		S.stat_sum_p[(size_t) g * n + lane] += v;
		S.stat_sum_p2[(size_t) g * n + lane] += v * v;
	} else if (lane == n) {
		if (better)
			S.prob_best[g] = prob;
		S.n_iter[g]++;
	} else if (lane == n + 1) {
		S.stat_n[g]++;
	} else if (lane == n + 2) {
		S.stat_sum_dl[g] += dl;
	} else if (lane == n + 3) {
		if (S.tr_prob_every > 0 && S.tr_prob != nullptr && step_index % S.tr_prob_every == 0) {
			long long row = step_index / S.tr_prob_every;
			S.tr_prob[row * S.n_chains + g] = prob;
			S.tr_dl[row * S.n_chains + g] = dl;
		}
	}
	__syncwarp();
}

// the state transition of a full step whose outcome is already known (cluster path: every warp
// of a chain's group takes the accept decision redundantly in registers, one warp writes it
// down): what chain_finalize does after the decision.  prop = this lane's proposed coordinate.
APM_D void chain_apply_step_warp(const DevState & S, int g, int accepted, double prob_new, double prior_new,
		double prop, int lane) {
	const int n = S.n_par;
	if (lane < n) {
		if (accepted) {
			S.params[(size_t) g * n + lane] = prop;
			S.pacc[(size_t) g * n + lane]++;
		} else {
			S.prej[(size_t) g * n + lane]++;
		}
	} else if (lane == n) {
		if (accepted) {
			S.prob[g] = prob_new;
			S.prior[g] = prior_new;
		} else if (S.quirks & 2u) {
			S.prior[g] = prior_new;
		}
	} else if (lane == n + 1) {
		if (accepted)
			S.accept[g]++;
		else
			S.reject[g]++;
	} else if (lane == n + 2) {
		S.rng_ctr[g]++;
	}
	__syncwarp();
}

// K steps' worth of a chain's random draws in one go, lane-parallel: lane j * (n + 1) + d computes,
// for step counter ctr0 + j, the first-attempt jump of coordinate d (d < n) or log(u) of the accept
// test (d == n).  out[j * (n + 1) + d].  Same values as propose_coordinate / chain_finalize draw.
APM_D void chain_draw_batch(const DevState & S, int g, u64 ctr0, int K, int lane, double * out) {
	const int n = S.n_par;
	const int j = lane / (n + 1), d = lane - j * (n + 1);
	if (j >= K)
		return;
	const uint32_t id = chain_rng_id(S, g);
	const bool is_jump = d < n;
	double u0, u1;
	philox_uniforms(S.seed, id, ctr0 + (u64) j, is_jump ? PURPOSE_JUMP : PURPOSE_ACCEPT, is_jump ? (uint32_t) d : 0u, 0,
			u0, u1);
	if (S.proposal == 0) {
		const double lg = log(u0);
		double r = lg;
		if (is_jump) // = jump_from_uniforms(0, sigma, u0, u1), the same operations in the same order
			r = S.steps[(size_t) g * n + d] * (sqrt(-2.0 * lg) * cos(2.0 * 3.14159265358979323846 * u1));
		out[lane] = r;
	} else {
		out[lane] = is_jump ? jump_from_uniforms(S.proposal, S.steps[(size_t) g * n + d], u0, u1) : log(u0);
	}
}

// ---- a warp that plays a whole chain (fused and grid paths): steps with batched draws ---------
// `draws` is the chain's 32-double buffer holding the draws of step counters draw_base ..
// draw_base + K - 1 (chain_draw_batch), K = 32 / (n_par + 1).
// the first proposal after the chain's state has been (re)loaded: draws a fresh batch
APM_D void chain_first_proposal_warp(const DevState & L, int g, double * draws, u64 & draw_base, int K, int lane) {
	const int n = L.n_par;
	draw_base = L.rng_ctr[g];
	chain_draw_batch(L, g, draw_base, K, lane, draws);
	__syncwarp();
	if (lane < n) {
		const double x = L.params[(size_t) g * n + lane];
		double v = x + draws[lane];
		if (v > L.pmax[lane] || v < L.pmin[lane])
			v = propose_coordinate(L, g, draw_base, lane, x, L.steps[(size_t) g * n + lane]);
		L.prop[(size_t) g * n + lane] = v;
	}
	if (lane == 0)
		L.pend[g] = n;
	__syncwarp();
}

// markov_chain_step's second half for the pending full step (sum valid on lane 0) and, if
// propose_next, the next step's proposal; the accept draw and the first-attempt jumps come
// from the batch, which is refilled when used up
template<class M>
APM_D void chain_step_tail_warp(const DevState & L, int g, double sum, double * draws, u64 & draw_base, int K,
		bool propose_next, int lane) {
	const int n = L.n_par;
	const u64 ctr = L.rng_ctr[g];
	chain_finalize_warp<M>(L, g, sum, draws + (size_t) (ctr - draw_base) * (n + 1) + n, lane);
	if (!propose_next)
		return;
	if (ctr + 1 >= draw_base + (u64) K) {
		draw_base = ctr + 1;
		chain_draw_batch(L, g, draw_base, K, lane, draws);
		__syncwarp();
	}
	if (lane < n) {
		const double x = L.params[(size_t) g * n + lane];
		double v = x + draws[(size_t) (ctr + 1 - draw_base) * (n + 1) + lane];
		if (v > L.pmax[lane] || v < L.pmin[lane])
			v = propose_coordinate(L, g, ctr + 1, lane, x, L.steps[(size_t) g * n + lane]);
		L.prop[(size_t) g * n + lane] = v;
	}
	if (lane == 0)
		L.pend[g] = n;
	__syncwarp();
}

// adapt() as compiled with -DADAPT (reference src/parallel_tempering.c:282-302): once per
// round, before the swap, every chain nudges all its step widths by 1 % from the sums of its
// per-parameter accept / reject counters (accepts / REJECTS, as coded: SURVEY App. D 7)
APM_D void chain_adapt(const DevState & S, int g) {
	const int n = S.n_par;
	u64 sa = 0, sr = 0;
	for (int i = 0; i < n; i++) {
		sa += S.pacc[(size_t) g * n + i];
		sr += S.prej[(size_t) g * n + i];
	}
	if (sa + sr < 20000)
		return;
	const double ratio = (double) sa * 1.0 / (double) sr;
	if (ratio < S.adapt_target - 0.05) {
		for (int i = 0; i < n; i++)
			S.steps[(size_t) g * n + i] *= 0.99;
	} else if (ratio > S.adapt_target + 0.05) {
		for (int i = 0; i < n; i++)
			S.steps[(size_t) g * n + i] *= 1 / 0.99;
	}
	if (sa + sr > 100000)
		chain_reset_accept_rejects(S, g);
}

// tempering_interaction for one ensemble: reference
// src/parallel_tempering_interaction.c:125-141 -> decide_swap_now :87-97 ->
// check_swap_probability :25-42 -> do_swap :99-123.  One thread per ensemble.
//
// With the ladder split over GPUs the pair (a, a + 1) may straddle two devices.  Each device
// then holds one chain of the pair and a copy of the other's swap-relevant state ("pack":
// prob, beta, prior, prob_best, params[n], params_best[n]) received from its neighbour; both
// evaluate the same expression on the same doubles with the same uniform, reach the same
// decision and each updates the chain it owns.  pack_prev = last rung of the previous device,
// pack_next = first rung of the next one, [n_ens][LADDER_PACK(n)] each.
#define LADDER_PACK(n) (4 + 2 * (n))

// The two draws of a round (which pair, and the logarithm the exchange is tested against) depend on
// the ensemble's id and its round counter only: a caller with idle threads may take them ahead of
// time (ensemble_swap_draws) and hand them over as `drawn`.
APM_D void ensemble_swap_draws(const DevState & S, int ens, double * drawn) {
	const uint32_t id = (uint32_t) (S.ensemble_id_offset + ens);
	const u64 round = S.swap_round[ens];
	double u_pick, u_test, dummy;
	philox_uniforms(S.seed, id, round, PURPOSE_SWAP_PICK, 0, 0, u_pick, dummy);
	philox_uniforms(S.seed, id, round, PURPOSE_SWAP_TEST, 0, 0, u_test, dummy);
	drawn[0] = u_pick;
	drawn[1] = log(u_test);
}

APM_D void ensemble_swap(const DevState & S, int ens, const double * pack_prev = nullptr,
		const double * pack_next = nullptr, const double * drawn = nullptr) {
	const int n_beta = S.n_beta, n = S.n_par, total = S.n_beta_total;
	if (total == 1)
		return;
	const int base = ens * n_beta;
	const uint32_t id = (uint32_t) (S.ensemble_id_offset + ens);
	const u64 round = S.swap_round[ens];
	double u_pick, log_u_test;
	if (drawn != nullptr) {
		u_pick = drawn[0];
		log_u_test = drawn[1];
	} else {
		double u_test, dummy;
		philox_uniforms(S.seed, id, round, PURPOSE_SWAP_PICK, 0, 0, u_pick, dummy);
		philox_uniforms(S.seed, id, round, PURPOSE_SWAP_TEST, 0, 0, u_test, dummy);
		log_u_test = log(u_test);
	}
	S.swap_round[ens] = round + 1;
	const int a = (int) (total * 1000 * u_pick) % (total - 1); // rung of the whole ladder
	const int la = a - S.k_offset, lb = la + 1;                  // positions on this device
	const bool own_a = la >= 0 && la < n_beta, own_b = lb >= 0 && lb < n_beta;
	if (!own_a && !own_b)
		return;
	if ((!own_a && pack_prev == nullptr) || (!own_b && pack_next == nullptr))
		return;
	const int ga = base + la, gb = base + lb;
	const double * pa = own_a ? nullptr : pack_prev + (size_t) ens * LADDER_PACK(n);
	const double * pb = own_b ? nullptr : pack_next + (size_t) ens * LADDER_PACK(n);
	const double a_prob = own_a ? S.prob[ga] : pa[0], b_prob = own_b ? S.prob[gb] : pb[0];
	const double a_beta = own_a ? S.beta[ga] : pa[1], b_beta = own_b ? S.beta[gb] : pb[1];
	const double a_prior = own_a ? S.prior[ga] : pa[2], b_prior = own_b ? S.prior[gb] : pb[2];
	const double a_best = own_a ? S.prob_best[ga] : pa[3], b_best = own_b ? S.prob_best[gb] : pb[3];
	const double * a_params = own_a ? S.params + (size_t) ga * n : pa + 4;
	const double * b_params = own_b ? S.params + (size_t) gb * n : pb + 4;
	const double * a_params_best = own_a ? S.params_best + (size_t) ga * n : pa + 4 + n;
	const double * b_params_best = own_b ? S.params_best + (size_t) gb * n : pb + 4 + n;
	double r;
	if (S.quirks & 1u) {
		r = a_beta * b_prob / b_beta + b_beta * a_prob / a_beta - (a_prob + b_prob);
	} else {
		double la_ = (a_prob - a_prior) / a_beta;
		double lb_ = (b_prob - b_prior) / b_beta;
		r = (a_beta - b_beta) * (lb_ - la_);
	}
	if (r > log_u_test) {
		for (int i = 0; i < n; i++) {
			const double ta = a_params[i], tb = b_params[i];
			if (own_a)
				S.params[(size_t) ga * n + i] = tb;
			if (own_b)
				S.params[(size_t) gb * n + i] = ta;
		}
		if (!(S.quirks & 1u)) {
			const double la_ = (a_prob - a_prior) / a_beta;
			const double lb_ = (b_prob - b_prior) / b_beta;
			if (own_a) {
				S.prior[ga] = b_prior;
				S.prob[ga] = b_prior + a_beta * lb_;
			}
			if (own_b) {
				S.prior[gb] = a_prior;
				S.prob[gb] = a_prior + b_beta * la_;
			}
		}
		// the better of the two bests goes to both
		if (a_best > b_best) {
			if (own_b) {
				S.prob_best[gb] = a_best;
				for (int i = 0; i < n; i++)
					S.params_best[(size_t) gb * n + i] = a_params_best[i];
			}
		} else if (own_a) {
			S.prob_best[ga] = b_best;
			for (int i = 0; i < n; i++)
				S.params_best[(size_t) ga * n + i] = b_params_best[i];
		}
		if (own_a)
			S.swapcount[ga]++; // inc_swapcount(chains[candidate]) (:139)
	}
}

// ---- calibration state machine ----------------------------------------------------------
// One call = "the step that was pending has been finalised; decide what this chain
// does next".  Unrolls, per chain, the loops of burn_in (reference
// src/markov_chain.c:34-79) and markov_chain_calibrate_orig
// (src/markov_chain_calibrate.c:1039-1180) so that thousands of chains calibrate
// concurrently, each at its own position in the algorithm.
APM_D void cal_finish(const DevState & S, int g, int status) {
	S.cal[g].phase = CAL_DONE;
	S.cal[g].status = status;
	atomicSub(S.n_active, 1);
}

APM_D void cal_enter_orig(const DevState & S, int g, const CalibCfgDev & cfg) {
	CalState & c = S.cal[g];
	const int n = S.n_par;
	if (cfg.skip_calibrate) { // SKIP_CALIBRATE_ALLCHAINS: burn_in only (parallel_tempering.c:189-195)
		cal_finish(S, g, 0);
		return;
	}
	c.rat_limit = pow(cfg.desired_acceptance_rate, 1.0 / n); // (:1058)
	for (int i = 0; i < n; i++)
		S.steps[(size_t) g * n + i] *= cfg.adjust_step;          // (:1060)
	chain_reset_accept_rejects(S, g);                           // (:1062)
	c.iter = 0;
	c.sub = 0;
	c.nchecks_without_rescaling = 0;
	c.reached_perfection = 0;
	c.phase = CAL_SINGLE;
}

// leave BURN_A (first half done or empty)
APM_D void cal_burn_midpoint(const DevState & S, int g, const CalibCfgDev & cfg) {
	CalState & c = S.cal[g];
	const int n = S.n_par;
	chain_restart_from_best(S, g);                             // markov_chain.c:59
	for (int i = 0; i < n; i++)
		S.steps[(size_t) g * n + i] *= 0.5;                    // :60
	c.sub = 0;
	if (c.iter < cfg.burn_in_iterations) {
		c.phase = CAL_BURN_B;
	} else {
		for (int i = 0; i < n; i++)
			S.steps[(size_t) g * n + i] = c.saved_steps[i];    // :73
		cal_enter_orig(S, g, cfg);
	}
}

// the counter whose change tells whether a step of `kind` was accepted: this is how
// assess_acceptance_rate keeps its log (reference src/markov_chain.c:146-172)
APM_D u64 steps_counter(const DevState & S, int g, int kind) {
	return kind == S.n_par ? S.accept[g] : S.pacc[(size_t) g * S.n_par + kind];
}

APM_D void cal_begin(const DevState & S, int g, const CalibCfgDev & cfg) {
	CalState & c = S.cal[g];
	const int n = S.n_par;
	c.status = 0;
	c.iter = 0;
	c.sub = 0;
	if (cfg.steps_n > 0) {
		c.phase = CAL_STEPS;
		c.sub = cfg.steps_kind;
		c.steps_left = cfg.steps_n;
		c.steps_done = 0;
		c.last_count = steps_counter(S, g, cfg.steps_kind);
		return;
	}
	for (int i = 0; i < n; i++) {
		c.saved_steps[i] = S.steps[(size_t) g * n + i];                 // markov_chain.c:38
		S.steps[(size_t) g * n + i] = (S.pmax[i] - S.pmin[i]) * 0.1;    // :39-41
	}
	if (c.iter < cfg.burn_in_iterations / 2)
		c.phase = CAL_BURN_A;
	else
		cal_burn_midpoint(S, g, cfg);
}

// what kind of step the chain needs next (PEND_NONE when it is finished)
APM_D int cal_next_kind(const DevState & S, int g) {
	const CalState & c = S.cal[g];
	switch (c.phase) {
	case CAL_BURN_A:
	case CAL_BURN_B:
	case CAL_FULL:
		return S.n_par;
	case CAL_SINGLE:
	case CAL_STEPS:
		return c.sub;
	default:
		return PEND_NONE;
	}
}

APM_D void cal_after_step(const DevState & S, int g, const CalibCfgDev & cfg) {
	CalState & c = S.cal[g];
	const int n = S.n_par;
	const int iter_readjust = cfg.iter_readjust > 0 ? cfg.iter_readjust : 200;
	const int no_rescaling_limit = cfg.no_rescaling_limit > 0 ? cfg.no_rescaling_limit : 15;
	switch (c.phase) {
	case CAL_STEPS: {
		chain_check_best(S, g);
		const u64 now = steps_counter(S, g, c.sub);
		if (S.alog != nullptr)
			S.alog[(size_t) c.steps_done * S.n_chains + S.g_base + g] = now != c.last_count ? 1 : 0;
		c.last_count = now;
		c.steps_done++;
		if (--c.steps_left == 0)
			cal_finish(S, g, 0);
		return;
	}
	case CAL_BURN_A:
	case CAL_BURN_B: {
		// blocks of 200 full steps, mcmc_check_best once per block (markov_chain.c:46-57,61-72)
		if (++c.sub < 200)
			return;
		c.sub = 0;
		c.iter += 200;
		chain_check_best(S, g);
		if (c.phase == CAL_BURN_A) {
			if (c.iter >= cfg.burn_in_iterations / 2)
				cal_burn_midpoint(S, g, cfg);
		} else if (c.iter >= cfg.burn_in_iterations) {
			for (int i = 0; i < n; i++)
				S.steps[(size_t) g * n + i] = c.saved_steps[i];
			cal_enter_orig(S, g, cfg);
		}
		return;
	}
	case CAL_SINGLE: {
		// for each parameter: markov_chain_step_for; mcmc_check_best (:1064-1067)
		chain_check_best(S, g);
		if (++c.sub < n)
			return;
		c.sub = 0;
		c.iter++;
		if (c.iter % (u64) iter_readjust != 0)
			return;
		// (:1069-1133) compare per-parameter acceptance rates with the target, rescale
		int rescaled = 0;
		for (int i = 0; i < n; i++) {
			double acc = (double) S.pacc[(size_t) g * n + i], rej = (double) S.prej[(size_t) g * n + i];
			double rate = acc / (rej + acc);
			double range = S.pmax[i] - S.pmin[i];
			double & step = S.steps[(size_t) g * n + i];
			if (rate > c.rat_limit + 0.05) {
				step = step / cfg.mul;
				if (rescaled == 0)
					rescaled = -1;
				if (step / range > 1) {
					step = 1 * range;
					if (rescaled == -1)
						rescaled = 0;
				}
				if (step / range > 10000) {
					cal_finish(S, g, 1);
					return;
				}
				if (rescaled == -1)
					rescaled = 1;
			}
			if (rate < c.rat_limit - 0.05) {
				step = step * cfg.mul;
				rescaled = 1;
			}
		}
		if (rescaled == 0)
			c.nchecks_without_rescaling++;
		c.rescaled = rescaled;
		chain_restart_from_best(S, g);
		chain_reset_accept_rejects(S, g);
		c.phase = CAL_FULL;
		return;
	}
	case CAL_FULL: {
		// ITER_READJUST x { markov_chain_step; mcmc_check_best } (:1135-1138)
		chain_check_best(S, g);
		if (++c.sub < iter_readjust)
			return;
		c.sub = 0;
		// calibration_progress.data rows (:1142-1147)
		if (S.progress != nullptr) {
			for (int i = 0; i < n; i++) {
				unsigned long long slot = atomicAdd(S.progress_n, 1ull);
				if ((long long) slot < S.progress_cap) {
					double acc = (double) S.pacc[(size_t) g * n + i], rej = (double) S.prej[(size_t) g * n + i];
					ProgressRow & row = S.progress[slot];
					row.chain = S.g_base + g;
					row.param = i;
					row.iter = c.iter;
					row.step_normalised = S.steps[(size_t) g * n + i] / (S.pmax[i] - S.pmin[i]);
					row.accept_rate = acc / (rej + acc);
				}
			}
		}
		// (:1148-1163)
		double delta = (double) S.accept[g] / (double) (S.accept[g] + S.reject[g])
				- cfg.desired_acceptance_rate;
		if (fabs(delta) < cfg.max_ar_deviation) {
			c.reached_perfection = 1;
		} else {
			c.reached_perfection = 0;
			if (delta < 0)
				c.rat_limit /= 0.99;
			else
				c.rat_limit *= 0.99;
		}
		if (c.nchecks_without_rescaling >= no_rescaling_limit && c.reached_perfection == 1
				&& c.rescaled == 0) {
			chain_reset_accept_rejects(S, g); // (:1177)
			cal_finish(S, g, 0);
			return;
		}
		if (c.iter > cfg.iter_limit) {
			cal_finish(S, g, 2);
			return;
		}
		c.phase = CAL_SINGLE;
		return;
	}
	default:
		return;
	}
}

} // namespace apm
