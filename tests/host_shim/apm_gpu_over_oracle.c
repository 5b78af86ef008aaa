/*
 * apm_gpu_over_oracle.c -- TEST INFRASTRUCTURE (tests/ only).
 *
 * Implements the C ABI of include/apemost_gpu.h on top of the CPU oracle
 * (oracle/apm_oracle.h), so that tests/test_host_cpu.py can link the UNCHANGED
 * host layer (apemost_b200/host/*.c) against the oracle in ORC_RNG_MT19937 mode
 * and compare the files the host layer writes, byte for byte, with the files the
 * unmodified reference wrote (tests/golden/*_phases.json) -- on a box without a
 * GPU.  This pins the host layer's phase logic and file formats.
 *
 * It is never part of the product: apemost_b200/host/Makefile links
 * libapemost_gpu.so, which has no CPU path.  The two APIs mirror each other
 * field for field (oracle/apm_oracle.h), which the static asserts below check.
 */
#include <stdlib.h>
#include <string.h>
#include "apemost_gpu.h"
#include "apm_oracle.h"

#define SAME_SIZE(a, b) typedef char same_size_##a[sizeof(a) == sizeof(b) ? 1 : -1]
SAME_SIZE(apm_gpu_chain_io, orc_chain_io);
SAME_SIZE(apm_gpu_trace_cfg, orc_trace_cfg);
SAME_SIZE(apm_gpu_calib_cfg, orc_calib_cfg);
SAME_SIZE(apm_gpu_calib_progress, orc_calib_progress);

struct apm_gpu {
	orc_engine * e;
};

int apm_gpu_create(apm_gpu ** handle, const apm_gpu_config * cfg) {
	orc_config oc;
	const char * rng = getenv("APM_TEST_ORACLE_RNG");
	apm_gpu * h = (apm_gpu *) calloc(1, sizeof(*h));
	memset(&oc, 0, sizeof(oc));
	oc.model_id = cfg->model_id;
	oc.n_ensembles = cfg->n_ensembles;
	oc.n_beta = cfg->n_beta;
	oc.n_par = cfg->n_par;
	oc.seed = cfg->seed;
	oc.proposal = cfg->proposal;
	oc.circular_mask = cfg->circular_mask;
	oc.quirks = cfg->quirks;
	oc.rng_kind = rng != NULL && strcmp(rng, "philox") == 0 ? ORC_RNG_PHILOX : ORC_RNG_MT19937;
	oc.chain_id_offset = cfg->chain_id_offset;
	oc.ensemble_id_offset = cfg->ensemble_id_offset;
	memcpy(oc.model_const, cfg->model_const, sizeof(oc.model_const));
	oc.n_threads = 1;
	if (orc_create(&h->e, &oc) != 0) {
		free(h);
		return APM_EINVAL;
	}
	/* -DRANDOMSWAP changes nothing but the position in the global MT19937 stream (one more draw
	 * per round); the product ABI has no such notion, so the test tells the oracle directly */
	if (getenv("APM_TEST_RANDOMSWAP") != NULL)
		orc_set_random_swap(h->e, 1);
	*handle = h;
	return APM_OK;
}
int apm_gpu_steps(apm_gpu * h, const unsigned char * select, int kind, long long n_steps, unsigned char * accepted) {
	return orc_steps(h->e, select, kind, n_steps, accepted);
}
int apm_gpu_host_uniform(apm_gpu * h, int g, double * u) {
	return orc_host_uniform(h->e, g, u);
}
int apm_gpu_set_adapt(apm_gpu * h, int enabled, double target_acceptance_rate) {
	return orc_set_adapt(h->e, enabled, target_acceptance_rate);
}
int apm_gpu_destroy(apm_gpu * h) {
	if (h != NULL) {
		orc_destroy(h->e);
		free(h);
	}
	return APM_OK;
}
const char * apm_gpu_last_error(const apm_gpu * h) {
	(void) h;
	return "oracle-backed test shim: call failed";
}
int apm_gpu_set_data(apm_gpu * h, const double * rowmajor, long long n_rows, int n_cols) {
	return orc_set_data(h->e, rowmajor, n_rows, n_cols);
}
int apm_gpu_set_bounds(apm_gpu * h, const double * lo, const double * hi) {
	return orc_set_bounds(h->e, lo, hi);
}
int apm_gpu_set_chains(apm_gpu * h, int first, int count, const apm_gpu_chain_io * in) {
	return orc_set_chains(h->e, first, count, (const orc_chain_io *) in);
}
int apm_gpu_get_chains(apm_gpu * h, int first, int count, apm_gpu_chain_io * out) {
	return orc_get_chains(h->e, first, count, (orc_chain_io *) out);
}
int apm_gpu_eval(apm_gpu * h, int n, const double * params, const double * beta, double * prob_out,
		double * prior_out) {
	return orc_eval(h->e, n, params, beta, prob_out, prior_out);
}
int apm_gpu_run(apm_gpu * h, long long n_rounds, int n_swap, const apm_gpu_trace_cfg * trace) {
	return orc_run(h->e, n_rounds, n_swap, (const orc_trace_cfg *) trace);
}
int apm_gpu_read_trace(apm_gpu * h, double * prob, double * dl, double * params, long long * n_prob_rows,
		long long * n_param_rows) {
	return orc_read_trace(h->e, prob, dl, params, n_prob_rows, n_param_rows);
}
int apm_gpu_calibrate(apm_gpu * h, const unsigned char * select, const apm_gpu_calib_cfg * cfg, int * status,
		apm_gpu_calib_progress * progress, long long progress_capacity, long long * n_progress) {
	return orc_calibrate(h->e, select, (const orc_calib_cfg *) cfg, status, (orc_calib_progress *) progress,
			progress_capacity, n_progress);
}
int apm_gpu_reset_stats(apm_gpu * h) {
	return orc_reset_stats(h->e);
}
int apm_gpu_get_stats(apm_gpu * h, unsigned long long * n, double * sum_dl, double * sum_params,
		double * sum_params_sq) {
	return orc_get_stats(h->e, n, sum_dl, sum_params, sum_params_sq);
}
int apm_gpu_set_marginals(apm_gpu * h, int which_chains, int n_bins, unsigned long long batch_size, int max_batches) {
	return orc_set_marginals(h->e, which_chains, n_bins, batch_size, max_batches);
}
int apm_gpu_get_marginals(apm_gpu * h, unsigned long long * counts, double * batch_means,
		unsigned long long * n_values, unsigned long long * n_batches) {
	return orc_get_marginals(h->e, counts, batch_means, n_values, n_batches);
}
