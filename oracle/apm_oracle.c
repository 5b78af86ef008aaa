/*
 * apm_oracle.c -- CPU oracle for the APEMoST hot path.  TEST INFRASTRUCTURE:
 * see apm_oracle.h for who may use it and how it is pinned to the reference.
 *
 * Plain C99, no GSL.  Compiled with -ffp-contract=off so that, like the
 * reference (ISO C mode, reference Makefile:9), no multiply-add is fused.
 * "ref:" comments name the reference file:line each block restates.
 */
#include "apm_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif

typedef unsigned long long u64;
typedef unsigned int u32;

/* ===================================================================== */
/* Models: calc_model of apps/<model>.c                                   */
/* ===================================================================== */

/* ref: apps/simplesin.c:12-38, SIGMA apps/simplesin.c:8-10 */
static void model_simplesin(const double * mc, const double * p, const double * d,
		long long n, int nc, double beta, double * prob) {
	const double sigma = mc[0] != 0 ? mc[0] : 0.5;
	const double amplitude = p[0], frequency = p[1], phase = p[2], offset = p[3];
	double square_sum = 0;
	long long i;
	for (i = 0; i < n; i++) {
		double x = d[i * nc + 0];
		double y = d[i * nc + 1];
		double m = amplitude * sin(2.0 * M_PI * (frequency * x + phase)) + offset;
		double deltay = m - y;
		square_sum += deltay * deltay;
	}
	*prob = beta * square_sum / (-2 * sigma * sigma);
}

/* ref: apps/simplesin5.c:15-41 with the two stale lines dropped:
 * loop bound = rows of the data table, no write to m->model */
static void model_simplesin5(const double * mc, const double * p, const double * d,
		long long n, int nc, double beta, double * prob) {
	const double sigma = mc[0] != 0 ? mc[0] : 0.5;
	const double param0 = p[0], param1 = p[1], param2 = p[2], param3 = p[3];
	double square_sum = 0;
	long long i;
	for (i = 0; i < n; i++) {
		double x = d[i * nc + 0];
		double f = param0 * sin(2.0 * M_PI * param1 * x + param2);
		double y = f + param3 - d[i * nc + 1];
		square_sum += y * y;
	}
	*prob = beta * square_sum / (-2 * sigma * sigma);
}

/* ref: apps/simplesin2.c:12-32 */
static void model_simplesin2(const double * mc, const double * p, const double * d,
		long long n, int nc, double beta, double * prob) {
	const double sigma = mc[0] != 0 ? mc[0] : 0.5;
	const double param0 = p[0], param1 = p[1];
	double square_sum = 0;
	long long i;
	for (i = 0; i < n; i++) {
		double x = d[i * nc + 0];
		double y = param0 * sin(2.0 * M_PI * (param1 * x + 0.3312)) - d[i * nc + 1];
		square_sum += y * y;
	}
	*prob = beta * square_sum / (-2 * sigma * sigma);
}

/* ref: apps/normal.c:8-34.  pow(1.0, i) == 1; i == 0 gives sigma = 0 and a NaN (or
 * -inf) candidate that `a > b` never takes. */
static void model_normal(const double * p, double beta, double * prob) {
	double x = p[0];
	double a, b = 0, sigma, pos, height;
	unsigned int i;
	for (i = 0; i < 10; i++) {
		pos = exp(i);
		height = 10 * pow(1.0, i);
		sigma = i;
		if (i % 2 == 0)
			a = -sigma * pow((x - pos) / sigma, 2) / 2 + height;
		else if (x > pos)
			a = -height * (x - pos) / sigma + height;
		else
			a = -height * (pos - x) / sigma + height;
		if (a > b)
			b = a;
	}
	*prob = beta * b;
}

/* ref: apps/pulse_vrot.c:12-65, HMIN :8-10 */
static void model_pulse_vrot(const double * mc, const double * p, int n_par,
		const double * d, long long n, int nc, double beta, double * prob_out,
		double * prior_out) {
	const double hmin = mc[0] != 0 ? mc[0] : 1e-6;
	double prior = 0, prob, lifetime, vrot, y, freq, distance, mode_freq, mode_height;
	unsigned int i;
	long long r;
	for (i = 3; i < (unsigned) n_par; i += 2)
		prior += log(p[i + 1] + hmin);
	i = (n_par - 3) / 2;
	prior = -prior / i;
	prob = p[1];
	lifetime = p[0];
	vrot = p[2];
	for (r = 0; r < n; r++) {
		y = 0;
		freq = d[r * nc + 0];
		mode_freq = p[3];
		mode_height = p[4];
		distance = mode_freq - freq;
		y += mode_height / (1 + pow(2 * M_PI * distance * lifetime, 2));
		mode_freq = p[5];
		mode_height = p[6];
		distance = mode_freq - freq + -1 * vrot;
		y += mode_height / (1 + pow(2 * M_PI * distance * lifetime, 2));
		distance = mode_freq - freq;
		y += mode_height / (1 + pow(2 * M_PI * distance * lifetime, 2));
		distance = mode_freq - freq + 1 * vrot;
		y += mode_height / (1 + pow(2 * M_PI * distance * lifetime, 2));
		prob += log(y) + d[r * nc + 1] / y;
	}
	*prior_out = prior;
	*prob_out = prior + -beta * prob;
}

/* ref: apps/pulse.c:12-56 */
static void model_pulse(const double * mc, const double * p, int n_par,
		const double * d, long long n, int nc, double beta, double * prob_out,
		double * prior_out) {
	const double hmin = mc[0] != 0 ? mc[0] : 1e-6;
	double prior = 0, prob, lifetime, y, freq, distance;
	unsigned int i, j;
	long long r;
	for (i = 2; i < (unsigned) n_par; i += 2)
		prior += log(p[i + 1] + hmin);
	i = (n_par - 2) / 2;
	prior = -prior / i;
	prob = p[1];
	lifetime = p[0];
	for (r = 0; r < n; r++) {
		y = 0;
		freq = d[r * nc + 0];
		for (j = 2; j < (unsigned) n_par; j += 2) {
			distance = p[j] - freq;
			y += p[j + 1] / (1 + pow(2 * M_PI * distance * lifetime, 2));
		}
		prob += log(y) + d[r * nc + 1] / y;
	}
	*prior_out = prior;
	*prob_out = prior + -beta * prob;
}

/* ref: apps/bernoulli_example.c:9-51 (SIGMA 2 :7) */
static void model_bernoulli(const double * p, int n_par, const double * d,
		long long n, int nc, double beta, double * prob_out, double * prior_out) {
	double prior = 0, prob = 0, eta_i, p_i, l_i;
	int j;
	long long i;
	for (j = 1; j < nc; j++)
		prior += -pow(p[j] / 2, 2) / 2;
	for (i = 0; i < n; i++) {
		eta_i = p[0];
		for (j = 1; j < n_par; j++)
			eta_i += d[i * nc + j] * p[j];
		if (eta_i > 0)
			p_i = 1 / (1 + exp(-eta_i));
		else
			p_i = exp(eta_i) / (1 + exp(eta_i));
		if (d[i * nc + 0] == 0)
			l_i = log(1 - p_i);
		else
			l_i = log(p_i);
		prob += l_i;
	}
	*prior_out = prior;
	*prob_out = prior + beta * prob;
}

/* The plugin contract (ref: src/mcmc.h:164, doc/manual.rst:126-168): calc_model
 * sets prob always and prior only if the model has one (*prior is left alone
 * otherwise, exactly like a model that never calls set_prior). */
void orc_calc_model(int model_id, const double * model_const, const double * params,
		int n_par, const double * data, long long n_rows, int n_cols, double beta,
		double * prob, double * prior) {
	static const double zeros[4] = { 0, 0, 0, 0 };
	const double * mc = model_const ? model_const : zeros;
	switch (model_id) {
	case ORC_MODEL_SIMPLESIN:
		model_simplesin(mc, params, data, n_rows, n_cols, beta, prob);
		break;
	case ORC_MODEL_SIMPLESIN5:
		model_simplesin5(mc, params, data, n_rows, n_cols, beta, prob);
		break;
	case ORC_MODEL_SIMPLESIN2:
		model_simplesin2(mc, params, data, n_rows, n_cols, beta, prob);
		break;
	case ORC_MODEL_NORMAL:
		model_normal(params, beta, prob);
		break;
	case ORC_MODEL_PULSE_VROT:
		model_pulse_vrot(mc, params, n_par, data, n_rows, n_cols, beta, prob, prior);
		break;
	case ORC_MODEL_PULSE:
		model_pulse(mc, params, n_par, data, n_rows, n_cols, beta, prob, prior);
		break;
	case ORC_MODEL_BERNOULLI:
		model_bernoulli(params, n_par, data, n_rows, n_cols, beta, prob, prior);
		break;
	default:
		*prob = NAN;
	}
}

/* ===================================================================== */
/* Small pinned helpers                                                   */
/* ===================================================================== */

/* ref: src/mcmc_internal.h:46-48 (macro mod_double) */
double orc_mod_double(double x, double div) {
	return x < 0 ? x - div * (int) (x / div - 1) : x - div * (int) (x / div);
}

/* ref: src/parallel_tempering_beta.c:66-69 (chebyshev_beta, the default
 * BETA_ALIGNMENT, parallel_tempering_beta.h:52-54) and :85-90 */
static double chebyshev_beta(unsigned i, unsigned n_beta, double beta_0) {
	return beta_0 + (1 - beta_0) / 2 * (1 - cos(i * M_PI / (n_beta - 1)));
}
double orc_get_chain_beta(unsigned i, unsigned n_beta, double beta_0) {
	if (n_beta == 1)
		return 1.0;
	return chebyshev_beta(n_beta - i - 1, n_beta, beta_0);
}

/* ref: src/parallel_tempering_beta.c:92-102, BETA_0_STEPWIDTH 1.0 */
double orc_calc_beta_0(int n_par, const double * pmin, const double * pmax,
		const double * steps, const double * stepwidth_factors) {
	double max = -INFINITY;
	int i;
	for (i = 0; i < n_par; i++) {
		double r = (pmax[i] - pmin[i]) * 1.0;
		r = r / steps[i];
		r = r / stepwidth_factors[i];
		if (r > max)
			max = r;
	}
	return pow(max, -0.5);
}

/* ref: src/analyse.c:50-93 (rectangle rule over the ladder; mean_dl[k] is the mean
 * of column 2 of prob-chain<k>.dump) */
double orc_evidence(int n_beta, const double * beta, const double * mean_dl) {
	double data_logprob = 0, previous_beta = 0;
	int j;
	for (j = n_beta - 1;; j--) {
		data_logprob += (mean_dl[j] / beta[j]) * (beta[j] - previous_beta);
		if (j == 0)
			break;
		previous_beta = beta[j];
	}
	return data_logprob;
}

/* ===================================================================== */
/* Random numbers                                                         */
/* ===================================================================== */

/* --- MT19937 with GSL's seeding and uniform convention (SURVEY.md App. E;
 *     ref use: src/mcmc.c:27-35, src/mcmc_gettersetter.c:286-309) -------- */
#define MT_N 624
#define MT_M 397
typedef struct {
	unsigned long mt[MT_N];
	int mti;
} mt_state;

static void mt_seed(mt_state * s, unsigned long seed) {
	int i;
	if (seed == 0)
		seed = 4357;
	s->mt[0] = seed & 0xffffffffUL;
	for (i = 1; i < MT_N; i++)
		s->mt[i] = (1812433253UL * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + i)
				& 0xffffffffUL;
	s->mti = MT_N;
}
static unsigned long mt_next(mt_state * s) {
	unsigned long y;
	unsigned long * mt = s->mt;
	if (s->mti >= MT_N) {
		int k;
		for (k = 0; k < MT_N; k++) {
			y = (mt[k] & 0x80000000UL) | (mt[(k + 1) % MT_N] & 0x7fffffffUL);
			mt[k] = mt[(k + MT_M) % MT_N] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfUL : 0);
		}
		s->mti = 0;
	}
	y = mt[s->mti++];
	y ^= y >> 11;
	y ^= (y << 7) & 0x9d2c5680UL;
	y ^= (y << 15) & 0xefc60000UL;
	y ^= y >> 18;
	return y & 0xffffffffUL;
}
static double mt_uniform(mt_state * s) {
	return mt_next(s) / 4294967296.0;
}
static double mt_uniform_pos(mt_state * s) {
	double x;
	do
		x = mt_uniform(s);
	while (x == 0);
	return x;
}

/* --- Philox4x32-10 (Salmon et al. 2011), the engine's counter RNG -------- */
void orc_philox4x32_10(const unsigned ctr[4], const unsigned key[2], unsigned out[4]) {
	u32 c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	u32 k0 = key[0], k1 = key[1];
	int r;
	for (r = 0; r < 10; r++) {
		u64 p0 = (u64) 0xD2511F53u * c0;
		u64 p1 = (u64) 0xCD9E8D57u * c2;
		u32 n0 = (u32) (p1 >> 32) ^ c1 ^ k0;
		u32 n1 = (u32) p1;
		u32 n2 = (u32) (p0 >> 32) ^ c3 ^ k1;
		u32 n3 = (u32) p0;
		c0 = n0;
		c1 = n1;
		c2 = n2;
		c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0;
	out[1] = c1;
	out[2] = c2;
	out[3] = c3;
}

#define PURPOSE_JUMP 0u
#define PURPOSE_ACCEPT 1u
#define PURPOSE_SWAP_PICK 2u
#define PURPOSE_SWAP_TEST 3u
#define PURPOSE_HOST 4u

/* Stream layout shared with the engine (apemost_b200/csrc/apm_rng.cuh):
 *   key     = (seed lo, seed hi)
 *   counter = (c0 = chain or ensemble id, c1 = step lo, c2 = step hi,
 *              c3 = purpose << 28 | idx << 20 | attempt & 0xfffff; attempt >> 20 is
 *              xor-ed into the top 12 bits of c2)
 *   u0 = ((w0 >> 5) * 2^26 + (w1 >> 6) + 0.5) / 2^53, u1 likewise from w2, w3:
 *   53-bit uniforms strictly inside (0, 1) (the one value that would round to 1.0 is clamped). */
void orc_philox_uniforms(u64 seed, unsigned c0, u64 step, unsigned purpose,
		unsigned idx, unsigned attempt, double * u0, double * u1) {
	unsigned ctr[4], key[2], w[4];
	ctr[0] = c0;
	ctr[1] = (u32) step;
	ctr[2] = (u32) (step >> 32) ^ ((attempt >> 20) << 20); /* redraws beyond 2^20 spill here */
	ctr[3] = (purpose << 28) | ((idx & 0xffu) << 20) | (attempt & 0xfffffu);
	key[0] = (u32) seed;
	key[1] = (u32) (seed >> 32);
	orc_philox4x32_10(ctr, key, w);
	/* integer + 0.5 rounds to even above 2^52: clamp the one value that would become 1.0 */
	*u0 = fmin(((double) (w[0] >> 5) * 67108864.0 + (double) (w[1] >> 6) + 0.5)
			* (1.0 / 9007199254740992.0), 0x1.fffffffffffffp-1);
	*u1 = fmin(((double) (w[2] >> 5) * 67108864.0 + (double) (w[3] >> 6) + 0.5)
			* (1.0 / 9007199254740992.0), 0x1.fffffffffffffp-1);
}

/* ===================================================================== */
/* Engine state                                                           */
/* ===================================================================== */

typedef struct {
	/* ref: src/mcmc_struct.h:30-106 */
	u64 accept, reject;
	double prob, prior, prob_best;
	double * params, *params_best, *steps;
	u64 * pacc, *prej;
	u64 n_iter;
	/* ref: src/parallel_tempering_beta.h:65-76 */
	double beta;
	u64 swapcount;
	/* counter-RNG position (PHILOX mode) */
	u64 rng_counter;
	/* on-device-accumulator twins */
	u64 stat_n;
	double stat_sum_dl;
	double * stat_sum_p, *stat_sum_p2;
} chain_t;

struct orc_engine {
	orc_config cfg;
	int n_chains;
	chain_t * chains;
	/* marginal statistics of the recorded chains (orc_set_marginals): ref src/analyse.c:115-247 */
	int marg_mode, marg_bins, marg_cap;
	u64 marg_batch;
	u64 * marg_counts;    /* [slots][n_par][bins] */
	double * marg_bsum;   /* [slots][n_par] */
	double * marg_means;  /* [slots][n_par][cap] */
	u64 * marg_n, *marg_nb; /* [slots] */
	double * pmin, *pmax;
	double * data;
	long long n_rows;
	int n_cols;
	mt_state mt;
	u64 * swap_round; /* per ensemble: PHILOX swap-stream position */
	u64 * host_draws; /* per chain: PHILOX host-stream position (orc_host_uniform) */
	int adapt;            /* -DADAPT, ref src/parallel_tempering.c:282-302 */
	double adapt_target;  /* TARGET_ACCEPTANCE_RATE */
	int random_swap;      /* -DRANDOMSWAP, ref src/parallel_tempering_interaction.c:130-131 */
	/* trace of the last run */
	double * tr_prob, *tr_dl, *tr_params;
	long long tr_prob_rows, tr_param_rows;
	int tr_dumped;
};

static void marginals_free(orc_engine * e);

static void chain_eval(orc_engine * e, chain_t * c) {
	/* the plugin call, ref: src/markov_chain.c:376 */
	orc_calc_model(e->cfg.model_id, e->cfg.model_const, c->params, e->cfg.n_par,
			e->data, e->n_rows, e->n_cols, c->beta, &c->prob, &c->prior);
}

/* ---- random draws of one chain -------------------------------------- */

/* ref: src/mcmc_gettersetter.c:290-306 (get_next_random_jump) with
 * gsl_ran_gaussian / gsl_ran_logistic / gsl_ran_flat in
 * MT19937 mode; in PHILOX mode the engine's own transforms of (u0, u1). */
static double next_jump(orc_engine * e, int g, chain_t * c, unsigned idx,
		unsigned attempt, double sigma) {
	if (e->cfg.rng_kind == ORC_RNG_MT19937) {
		if (e->cfg.proposal == 1) {
			double x;
			do
				x = mt_uniform_pos(&e->mt);
			while (x == 1);
			return sigma * log(x / (1 - x));
		} else if (e->cfg.proposal == 2) {
			double u = mt_uniform(&e->mt);
			return (-sigma) * (1 - u) + sigma * u;
		} else {
			double x, y, r2;
			do {
				x = -1 + 2 * mt_uniform_pos(&e->mt);
				y = -1 + 2 * mt_uniform_pos(&e->mt);
				r2 = x * x + y * y;
			} while (r2 > 1.0 || r2 == 0);
			return sigma * y * sqrt(-2.0 * log(r2) / r2);
		}
	} else {
		double u0, u1;
		orc_philox_uniforms(e->cfg.seed, (unsigned) (e->cfg.chain_id_offset + g),
				c->rng_counter, PURPOSE_JUMP, idx, attempt, &u0, &u1);
		if (e->cfg.proposal == 1)
			return sigma * log(u0 / (1 - u0));
		else if (e->cfg.proposal == 2)
			return (-sigma) * (1 - u0) + sigma * u0;
		else
			return sigma * (sqrt(-2.0 * log(u0)) * cos(2.0 * M_PI * u1));
	}
}

/* ref: src/mcmc_gettersetter.c:307-309 (get_next_alog_urandom).  GSL aborts on
 * log(0) (probability 2^-32 per draw); the oracle lets log(0) = -inf stand. */
static double next_alog_urandom(orc_engine * e, int g, chain_t * c) {
	if (e->cfg.rng_kind == ORC_RNG_MT19937)
		return log(mt_uniform(&e->mt));
	else {
		double u0, u1;
		orc_philox_uniforms(e->cfg.seed, (unsigned) (e->cfg.chain_id_offset + g),
				c->rng_counter, PURPOSE_ACCEPT, 0, 0, &u0, &u1);
		return log(u0);
	}
}

/* ---- proposal: ref src/markov_chain.c:226-277 ------------------------- */
static void do_step_for(orc_engine * e, int g, chain_t * c, unsigned i) {
	const double step = c->steps[i];
	const double old_value = c->params[i];
	const double max = e->pmax[i], min = e->pmin[i];
	double new_value;
	unsigned attempt = 0;
	if (e->cfg.circular_mask == 0) {
		/* CIRCULAR_PARAMS == 0: redraw until inside [min, max]  (:235-240) */
		do {
			new_value = old_value + next_jump(e, g, c, i, attempt++, step);
		} while (new_value > max || new_value < min);
	} else {
		/* :241-262 */
		new_value = old_value + next_jump(e, g, c, i, attempt++, step);
		if (new_value > max || new_value < min) {
			if (e->cfg.circular_mask & (1u << i)) {
				new_value = min + orc_mod_double(new_value - min, max - min);
			} else {
				do {
					new_value = old_value + next_jump(e, g, c, i, attempt++, step);
				} while (new_value > max || new_value < min);
			}
		}
	}
	c->params[i] = new_value;
}

/* ---- accept rule: ref src/markov_chain.c:282-311 ---------------------- */
static int check_accept(orc_engine * e, int g, chain_t * c, double prob_old) {
	double prob_new = c->prob;
	if (prob_new == prob_old)
		return 1;
	if (prob_new > prob_old)
		return 1;
	return next_alog_urandom(e, g, c) < (prob_new - prob_old) ? 1 : 0;
}

/* ---- markov_chain_step: ref src/markov_chain.c:369-386 ---------------- */
static void markov_chain_step(orc_engine * e, int g, chain_t * c) {
	const int n = e->cfg.n_par;
	double prob_old = c->prob;
	double prior_old = c->prior;
	double old_values[64];
	int i;
	memcpy(old_values, c->params, n * sizeof(double));
	for (i = 0; i < n; i++)
		do_step_for(e, g, c, (unsigned) i);
	chain_eval(e, c);
	if (check_accept(e, g, c, prob_old) == 1) {
		/* inc_params_accepts, ref: src/mcmc_gettersetter.c:98-103 */
		c->accept++;
		for (i = 0; i < n; i++)
			c->pacc[i]++;
	} else {
		/* revert() restores prob only (:313-315); prior keeps the rejected
		 * proposal's value unless the quirk is switched off */
		c->prob = prob_old;
		if (!(e->cfg.quirks & ORC_QUIRK_STALE_PRIOR_ON_REJECT))
			c->prior = prior_old;
		memcpy(c->params, old_values, n * sizeof(double));
		c->reject++;
		for (i = 0; i < n; i++)
			c->prej[i]++;
	}
	c->rng_counter++;
}

/* ---- markov_chain_step_for: ref src/markov_chain.c:317-333 ------------ */
static void markov_chain_step_for(orc_engine * e, int g, chain_t * c, unsigned index) {
	double prob_old = c->prob;
	double prior_old = c->prior;
	double old_value = c->params[index];
	do_step_for(e, g, c, index);
	chain_eval(e, c); /* every shipped calc_model_for just calls calc_model */
	if (check_accept(e, g, c, prob_old) == 1) {
		c->pacc[index]++;
	} else {
		c->prob = prob_old;
		if (!(e->cfg.quirks & ORC_QUIRK_STALE_PRIOR_ON_REJECT))
			c->prior = prior_old;
		c->params[index] = old_value;
		c->prej[index]++;
	}
	c->rng_counter++;
}

/* ---- mcmc_check_best: ref src/mcmc_calculate.c:35-41 ------------------ */
static void mcmc_check_best(orc_engine * e, chain_t * c) {
	if (c->prob > c->prob_best) {
		c->prob_best = c->prob;
		memcpy(c->params_best, c->params, e->cfg.n_par * sizeof(double));
	}
}

/* ---- restart_from_best: ref src/markov_chain.c:29-32 ------------------ */
static void restart_from_best(orc_engine * e, chain_t * c) {
	memcpy(c->params, c->params_best, e->cfg.n_par * sizeof(double));
	c->prob = c->prob_best;
}

/* ---- reset_accept_rejects: ref src/mcmc_gettersetter.c:119-127 -------- */
static void reset_accept_rejects(orc_engine * e, chain_t * c) {
	int i;
	for (i = 0; i < e->cfg.n_par; i++) {
		c->pacc[i] = 0;
		c->prej[i] = 0;
	}
	c->reject = 0;
	c->accept = 0;
}

/* ---- burn_in: ref src/markov_chain.c:34-79 ---------------------------- */
static void burn_in(orc_engine * e, int g, chain_t * c, u64 burn_in_iterations) {
	const int n = e->cfg.n_par;
	double original_steps[64];
	u64 iter, subiter;
	int i;
	memcpy(original_steps, c->steps, n * sizeof(double));
	for (i = 0; i < n; i++)
		c->steps[i] = (e->pmax[i] - e->pmin[i]) * 0.1;
	for (iter = 0; iter < burn_in_iterations / 2;) {
		for (subiter = 0; subiter < 200; subiter++)
			markov_chain_step(e, g, c);
		iter += subiter;
		mcmc_check_best(e, c);
	}
	restart_from_best(e, c);
	for (i = 0; i < n; i++)
		c->steps[i] *= 0.5;
	for (; iter < burn_in_iterations;) {
		for (subiter = 0; subiter < 200; subiter++)
			markov_chain_step(e, g, c);
		iter += subiter;
		mcmc_check_best(e, c);
	}
	memcpy(c->steps, original_steps, n * sizeof(double));
}

typedef struct {
	orc_calib_progress * rows;
	long long cap, n;
} progress_sink;

static void progress_push(progress_sink * s, int chain, int param, u64 iter,
		double stepn, double rate) {
	if (s == NULL)
		return;
#pragma omp critical(orc_progress)
	{
		if (s->rows != NULL && s->n < s->cap) {
			s->rows[s->n].chain = chain;
			s->rows[s->n].param = param;
			s->rows[s->n].iter = iter;
			s->rows[s->n].step_normalised = stepn;
			s->rows[s->n].accept_rate = rate;
		}
		s->n++;
	}
}

/* ---- markov_chain_calibrate_orig: ref src/markov_chain_calibrate.c:1039-1180.
 * Returns 0, 1 (step width too large, :1104-1110) or 2 (iteration limit,
 * :1169-1174) where the reference exit(1)s. */
static int calibrate_orig(orc_engine * e, int g, chain_t * c, const orc_calib_cfg * cfg,
		progress_sink * sink) {
	const int n = e->cfg.n_par;
	const int iter_readjust = cfg->iter_readjust > 0 ? cfg->iter_readjust : 200;
	const int no_rescaling_limit = cfg->no_rescaling_limit > 0 ? cfg->no_rescaling_limit : 15;
	const double mul = cfg->mul;
	double rat_limit = pow(cfg->desired_acceptance_rate, 1.0 / n);
	int reached_perfection = 0, nchecks_without_rescaling = 0, rescaled, i;
	u64 iter = 0, subiter;
	double delta_reject_accept_t;

	for (i = 0; i < n; i++)
		c->steps[i] *= cfg->adjust_step;
	reset_accept_rejects(e, c);

	while (1) {
		for (i = 0; i < n; i++) {
			markov_chain_step_for(e, g, c, (unsigned) i);
			mcmc_check_best(e, c);
		}
		iter++;
		if (iter % iter_readjust == 0) {
			rescaled = 0;
			for (i = 0; i < n; i++) {
				/* get_accept_rate, ref: src/mcmc_gettersetter.c:78-86 */
				double rate = (double) c->pacc[i] / ((double) c->prej[i] + (double) c->pacc[i]);
				double range = e->pmax[i] - e->pmin[i];
				if (rate > rat_limit + 0.05) {
					c->steps[i] = c->steps[i] / mul;
					if (rescaled == 0)
						rescaled = -1;
					if (c->steps[i] / range > 1) {
						c->steps[i] = 1 * range;
						if (rescaled == -1)
							rescaled = 0;
					}
					if (c->steps[i] / range > 10000)
						return 1;
					if (rescaled == -1)
						rescaled = 1;
				}
				if (rate < rat_limit - 0.05) {
					c->steps[i] = c->steps[i] * mul;
					rescaled = 1;
				}
			}
			if (rescaled == 0)
				nchecks_without_rescaling++;
			restart_from_best(e, c);
			reset_accept_rejects(e, c);
			for (subiter = 0; subiter < (u64) iter_readjust; subiter++) {
				markov_chain_step(e, g, c);
				mcmc_check_best(e, c);
			}
			for (i = 0; i < n; i++) {
				double rate = (double) c->pacc[i] / ((double) c->prej[i] + (double) c->pacc[i]);
				progress_push(sink, g, i, iter, c->steps[i] / (e->pmax[i] - e->pmin[i]), rate);
			}
			/* get_accept_rate_global, ref: src/mcmc_gettersetter.c:62-65; the target here
			 * is the TARGET_ACCEPTANCE_RATE macro (:1148-1149), which the callers also pass
			 * as desired_acceptance_rate (parallel_tempering.c:80,117) */
			delta_reject_accept_t = (double) c->accept / (double) (c->accept + c->reject)
					- cfg->desired_acceptance_rate;
			if (fabs(delta_reject_accept_t) < cfg->max_ar_deviation) {
				reached_perfection = 1;
			} else {
				reached_perfection = 0;
				if (delta_reject_accept_t < 0)
					rat_limit /= 0.99;
				else
					rat_limit *= 0.99;
			}
			if (nchecks_without_rescaling >= no_rescaling_limit
					&& reached_perfection == 1 && rescaled == 0)
				break;
			if (iter > cfg->iter_limit)
				return 2;
		}
	}
	reset_accept_rejects(e, c);
	return 0;
}

/* ---- markov_chain_calibrate: ref src/markov_chain_calibrate.c:1182-1204 */
static int markov_chain_calibrate(orc_engine * e, int g, chain_t * c,
		const orc_calib_cfg * cfg, progress_sink * sink) {
	burn_in(e, g, c, cfg->burn_in_iterations);
	if (cfg->skip_calibrate) /* SKIP_CALIBRATE_ALLCHAINS, ref: parallel_tempering.c:189-195 */
		return 0;
	return calibrate_orig(e, g, c, cfg, sink);
}

/* ---- swap: ref src/parallel_tempering_interaction.c:25-42,87-141 ------ */
static void tempering_interaction(orc_engine * e, int ens) {
	const int n_beta = e->cfg.n_beta;
	const int n = e->cfg.n_par;
	chain_t * ch = e->chains + (size_t) ens * n_beta;
	double u_pick, u_test, dummy, r, c;
	int a, b, i;
	if (n_beta == 1)
		return;
	if (e->cfg.rng_kind == ORC_RNG_MT19937) {
		/* parallel_tempering_decide_swap_random(chains, n_beta, 1) :47-65: one more uniform is
		 * drawn first and compared with 1.0 / n_swap = 1, which it is always below; the pair
		 * (a, (a + 1) % n_beta) is (a, a + 1) because a <= n_beta - 2.  Hence RANDOMSWAP differs
		 * from the default only by this draw. */
		if (e->random_swap)
			(void) mt_uniform(&e->mt);
		u_pick = mt_uniform(&e->mt);
	} else {
		orc_philox_uniforms(e->cfg.seed, (unsigned) (e->cfg.ensemble_id_offset + ens),
				e->swap_round[ens], PURPOSE_SWAP_PICK, 0, 0, &u_pick, &dummy);
	}
	/* parallel_tempering_decide_swap_now :87-97 */
	a = (int) (n_beta * 1000 * u_pick) % (n_beta - 1);
	b = a + 1;
	/* check_swap_probability :25-42 */
	if (e->cfg.quirks & ORC_QUIRK_STALE_PROB_ON_SWAP) {
		double a_prob = ch[a].prob, b_prob = ch[b].prob;
		double a_beta = ch[a].beta, b_beta = ch[b].beta;
		r = a_beta * b_prob / b_beta + b_beta * a_prob / a_beta - (a_prob + b_prob);
	} else {
		/* exact PT ratio on the tempered likelihood only */
		double la = (ch[a].prob - ch[a].prior) / ch[a].beta;
		double lb = (ch[b].prob - ch[b].prior) / ch[b].beta;
		r = (ch[a].beta - ch[b].beta) * (lb - la);
	}
	if (e->cfg.rng_kind == ORC_RNG_MT19937) {
		u_test = mt_uniform(&e->mt);
	} else {
		orc_philox_uniforms(e->cfg.seed, (unsigned) (e->cfg.ensemble_id_offset + ens),
				e->swap_round[ens], PURPOSE_SWAP_TEST, 0, 0, &u_test, &dummy);
		e->swap_round[ens]++;
	}
	c = log(u_test);
	if (r > c) {
		/* parallel_tempering_do_swap :99-123 */
		for (i = 0; i < n; i++) {
			double t = ch[a].params[i];
			ch[a].params[i] = ch[b].params[i];
			ch[b].params[i] = t;
		}
		if (!(e->cfg.quirks & ORC_QUIRK_STALE_PROB_ON_SWAP)) {
			double la = (ch[a].prob - ch[a].prior) / ch[a].beta;
			double lb = (ch[b].prob - ch[b].prior) / ch[b].beta;
			double pa = ch[a].prior, pb = ch[b].prior;
			ch[a].prior = pb;
			ch[b].prior = pa;
			ch[a].prob = pb + ch[a].beta * lb;
			ch[b].prob = pa + ch[b].beta * la;
		}
		r = ch[a].prob_best;
		if (r > ch[b].prob_best) {
			ch[b].prob_best = r;
			memcpy(ch[b].params_best, ch[a].params_best, n * sizeof(double));
		} else {
			r = ch[b].prob_best;
			ch[a].prob_best = r;
			memcpy(ch[a].params_best, ch[b].params_best, n * sizeof(double));
		}
		ch[a].swapcount++; /* inc_swapcount(chains[candidate]) :139 */
	}
}

/* ---- adapt() with -DADAPT: ref src/parallel_tempering.c:282-302 ---------
 * once per round, before the swap: a constant small rescaling of all step widths of a chain
 * from its per-parameter accept / reject counter sums (note: accepts / REJECTS) */
static void adapt_chain(orc_engine * e, chain_t * c) {
	const int n = e->cfg.n_par;
	u64 sa = 0, sr = 0;
	double ratio;
	int i;
	for (i = 0; i < n; i++) {
		sa += c->pacc[i];
		sr += c->prej[i];
	}
	if (sa + sr < 20000)
		return;
	ratio = sa * 1.0 / sr;
	if (ratio < e->adapt_target - 0.05) {
		for (i = 0; i < n; i++)
			c->steps[i] *= 0.99;
	} else if (ratio > e->adapt_target + 0.05) {
		for (i = 0; i < n; i++)
			c->steps[i] *= 1 / 0.99;
	}
	if (sa + sr > 100000)
		reset_accept_rejects(e, c);
}

/* ===================================================================== */
/* API                                                                    */
/* ===================================================================== */

int orc_set_adapt(orc_engine * e, int enabled, double target_acceptance_rate) {
	e->adapt = enabled;
	e->adapt_target = target_acceptance_rate;
	return 0;
}

int orc_set_random_swap(orc_engine * e, int enabled) {
	e->random_swap = enabled;
	return 0;
}

int orc_create(orc_engine ** out, const orc_config * cfg) {
	orc_engine * e;
	int g, n;
	if (out == NULL || cfg == NULL || cfg->n_par < 1 || cfg->n_par > 64
			|| cfg->n_beta < 1 || cfg->n_ensembles < 1)
		return -1;
	e = (orc_engine *) calloc(1, sizeof(orc_engine));
	e->cfg = *cfg;
	n = cfg->n_par;
	e->n_chains = cfg->n_ensembles * cfg->n_beta;
	e->chains = (chain_t *) calloc(e->n_chains, sizeof(chain_t));
	e->host_draws = (u64 *) calloc(e->n_chains, sizeof(u64));
	e->pmin = (double *) calloc(n, sizeof(double));
	e->pmax = (double *) calloc(n, sizeof(double));
	e->swap_round = (u64 *) calloc(cfg->n_ensembles, sizeof(u64));
	for (g = 0; g < e->n_chains; g++) {
		chain_t * c = e->chains + g;
		/* mcmc_init, ref: src/mcmc.c:37-78 */
		c->prob = -1e+10;
		c->prior = 0;
		c->prob_best = -1e+10;
		c->beta = 1.0;
		c->params = (double *) calloc(n, sizeof(double));
		c->params_best = (double *) calloc(n, sizeof(double));
		c->steps = (double *) calloc(n, sizeof(double));
		c->pacc = (u64 *) calloc(n, sizeof(u64));
		c->prej = (u64 *) calloc(n, sizeof(u64));
		c->stat_sum_p = (double *) calloc(n, sizeof(double));
		c->stat_sum_p2 = (double *) calloc(n, sizeof(double));
	}
	mt_seed(&e->mt, (unsigned long) cfg->seed);
	*out = e;
	return 0;
}

int orc_destroy(orc_engine * e) {
	int g;
	if (e == NULL)
		return 0;
	for (g = 0; g < e->n_chains; g++) {
		chain_t * c = e->chains + g;
		free(c->params);
		free(c->params_best);
		free(c->steps);
		free(c->pacc);
		free(c->prej);
		free(c->stat_sum_p);
		free(c->stat_sum_p2);
	}
	marginals_free(e);
	free(e->chains);
	free(e->pmin);
	free(e->pmax);
	free(e->data);
	free(e->swap_round);
	free(e->host_draws);
	free(e->tr_prob);
	free(e->tr_dl);
	free(e->tr_params);
	free(e);
	return 0;
}

void orc_mt_seed(orc_engine * e, unsigned long seed) {
	mt_seed(&e->mt, seed);
}
double orc_mt_uniform(orc_engine * e) {
	return mt_uniform(&e->mt);
}
/* mirrors apm_gpu_host_uniform: gsl_rng_uniform(get_random(m)) outside the stepping loops
 * (ref: src/markov_chain_calibrate.c:93-94,155).  MT19937 mode: the next number of the one global
 * stream, like the reference; PHILOX mode: the chain's host stream (purpose 4) */
int orc_host_uniform(orc_engine * e, int g, double * u) {
	double u1;
	if (e == NULL || u == NULL || g < 0 || g >= e->n_chains)
		return -1;
	if (e->cfg.rng_kind == ORC_RNG_MT19937) {
		*u = mt_uniform(&e->mt);
		return 0;
	}
	orc_philox_uniforms(e->cfg.seed, (unsigned) (e->cfg.chain_id_offset + g), e->host_draws[g]++, PURPOSE_HOST,
			0, 0, u, &u1);
	return 0;
}

int orc_set_data(orc_engine * e, const double * d, long long n_rows, int n_cols) {
	free(e->data);
	e->data = NULL;
	e->n_rows = n_rows;
	e->n_cols = n_cols;
	if (n_rows > 0 && n_cols > 0) {
		e->data = (double *) malloc((size_t) n_rows * n_cols * sizeof(double));
		memcpy(e->data, d, (size_t) n_rows * n_cols * sizeof(double));
	}
	return 0;
}

int orc_set_bounds(orc_engine * e, const double * pmin, const double * pmax) {
	memcpy(e->pmin, pmin, e->cfg.n_par * sizeof(double));
	memcpy(e->pmax, pmax, e->cfg.n_par * sizeof(double));
	return 0;
}

int orc_set_chains(orc_engine * e, int first, int count, const orc_chain_io * in) {
	const int n = e->cfg.n_par;
	int k;
	if (first < 0 || count < 0 || first + count > e->n_chains)
		return -1;
	for (k = 0; k < count; k++) {
		chain_t * c = e->chains + first + k;
		if (in->beta) {
			/* set_beta also zeroes swapcount, ref: src/parallel_tempering_beta.c:25-28 */
			c->beta = in->beta[k];
			c->swapcount = 0;
		}
		if (in->params)
			memcpy(c->params, in->params + (size_t) k * n, n * sizeof(double));
		if (in->steps)
			memcpy(c->steps, in->steps + (size_t) k * n, n * sizeof(double));
		if (in->prob)
			c->prob = in->prob[k];
		if (in->prior)
			c->prior = in->prior[k];
		if (in->prob_best)
			c->prob_best = in->prob_best[k];
		if (in->params_best)
			memcpy(c->params_best, in->params_best + (size_t) k * n, n * sizeof(double));
		if (in->accept)
			c->accept = in->accept[k];
		if (in->reject)
			c->reject = in->reject[k];
		if (in->params_accepts)
			memcpy(c->pacc, in->params_accepts + (size_t) k * n, n * sizeof(u64));
		if (in->params_rejects)
			memcpy(c->prej, in->params_rejects + (size_t) k * n, n * sizeof(u64));
		if (in->n_iter)
			c->n_iter = in->n_iter[k];
		if (in->swapcount)
			c->swapcount = in->swapcount[k];
		if (in->rng_counter)
			c->rng_counter = in->rng_counter[k];
	}
	return 0;
}

int orc_get_chains(orc_engine * e, int first, int count, orc_chain_io * out) {
	const int n = e->cfg.n_par;
	int k;
	if (first < 0 || count < 0 || first + count > e->n_chains)
		return -1;
	for (k = 0; k < count; k++) {
		chain_t * c = e->chains + first + k;
		if (out->beta)
			out->beta[k] = c->beta;
		if (out->params)
			memcpy(out->params + (size_t) k * n, c->params, n * sizeof(double));
		if (out->steps)
			memcpy(out->steps + (size_t) k * n, c->steps, n * sizeof(double));
		if (out->prob)
			out->prob[k] = c->prob;
		if (out->prior)
			out->prior[k] = c->prior;
		if (out->prob_best)
			out->prob_best[k] = c->prob_best;
		if (out->params_best)
			memcpy(out->params_best + (size_t) k * n, c->params_best, n * sizeof(double));
		if (out->accept)
			out->accept[k] = c->accept;
		if (out->reject)
			out->reject[k] = c->reject;
		if (out->params_accepts)
			memcpy(out->params_accepts + (size_t) k * n, c->pacc, n * sizeof(u64));
		if (out->params_rejects)
			memcpy(out->params_rejects + (size_t) k * n, c->prej, n * sizeof(u64));
		if (out->n_iter)
			out->n_iter[k] = c->n_iter;
		if (out->swapcount)
			out->swapcount[k] = c->swapcount;
		if (out->rng_counter)
			out->rng_counter[k] = c->rng_counter;
	}
	return 0;
}

/* ref: apps/eval_main.c:48-66 (beta given per vector; prior starts at 0 like a
 * freshly initialised chain, src/mcmc.c:48) */
int orc_eval(orc_engine * e, int n, const double * params, const double * beta,
		double * prob_out, double * prior_out) {
	int k;
#pragma omp parallel for schedule(dynamic) num_threads(e->cfg.n_threads > 0 ? e->cfg.n_threads : 1)
	for (k = 0; k < n; k++) {
		double prob = 0, prior = 0;
		orc_calc_model(e->cfg.model_id, e->cfg.model_const,
				params + (size_t) k * e->cfg.n_par, e->cfg.n_par, e->data, e->n_rows,
				e->n_cols, beta ? beta[k] : 1.0, &prob, &prior);
		prob_out[k] = prob;
		if (prior_out)
			prior_out[k] = prior;
	}
	return 0;
}

/* ---- what analyse derives from <name>-chain-<i>.prob.dump, accumulated step by step ----------
 * ref: calc_marginal_distribution src/analyse.c:159-247 (create_hist src/histogram.c:34-43 +
 * gsl_histogram_increment) and calc_mcmc_error src/analyse.c:115-142 */
static double marg_edge(int k, int nb, double lo, double hi) {
	/* gsl_histogram_set_ranges_uniform's expressions; create_hist widens the last edge */
	const double f1 = (double) (nb - k) / (double) nb, f2 = (double) k / (double) nb;
	double e = f1 * lo + f2 * hi;
	if (k == nb)
		e += (hi - lo) / 10000;
	return e;
}

static void marginals_add(orc_engine * e, int g, const double * params) {
	const int n = e->cfg.n_par, nb = e->marg_bins;
	int slot = -1, i;
	u64 seen, closed;
	if (e->marg_mode == 2)
		slot = g;
	else if (e->marg_mode == 1 && g % e->cfg.n_beta == 0)
		slot = g / e->cfg.n_beta;
	if (slot < 0)
		return;
	seen = e->marg_n[slot];
	closed = e->marg_nb[slot];
	for (i = 0; i < n; i++) {
		const double v = params[i], lo = e->pmin[i], hi = e->pmax[i], top = marg_edge(nb, nb, lo, hi);
		double * bs = &e->marg_bsum[(size_t) slot * n + i];
		if (v >= lo && v < top) {
			int b = 0;
			while (b < nb - 1 && v >= marg_edge(b + 1, nb, lo, hi)) /* the bin with edge[b] <= v < edge[b + 1] */
				b++;
			e->marg_counts[((size_t) slot * n + i) * nb + b]++;
		}
		/* n++; batchsum += v; if (n % batchsize == batchsize - 1) { batchmean = batchsum / batchsize; ... } */
		*bs += v;
		if (e->marg_batch > 0 && (seen + 1) % e->marg_batch == e->marg_batch - 1) {
			if (closed < (u64) e->marg_cap)
				e->marg_means[((size_t) slot * n + i) * e->marg_cap + closed] = *bs / (double) e->marg_batch;
			*bs = 0;
		}
	}
	e->marg_n[slot] = seen + 1;
	if (e->marg_batch > 0 && (seen + 1) % e->marg_batch == e->marg_batch - 1)
		e->marg_nb[slot] = closed + 1;
}

/* ---- run_sampler hot loop: ref src/parallel_tempering.c:392-409 -------- */
static void sampler_substep(orc_engine * e, int g, long long step_index,
		const orc_trace_cfg * tr) {
	chain_t * c = e->chains + g;
	const int n = e->cfg.n_par;
	int i;
	markov_chain_step(e, g, c);
	mcmc_check_best(e, c);
	/* mcmc_append_current_parameters, ref: src/mcmc_calculate.c:30-33 */
	c->n_iter++;
	/* the prob-chain line, ref: src/parallel_tempering.c:399-401 */
	if (tr->prob_every > 0 && e->tr_prob != NULL && step_index % tr->prob_every == 0) {
		long long row = step_index / tr->prob_every;
		e->tr_prob[row * e->n_chains + g] = c->prob;
		e->tr_dl[row * e->n_chains + g] = c->prob - c->prior;
	}
	if (e->tr_params != NULL) {
		int slot = -1;
		if (tr->params_chains == 2)
			slot = g;
		else if (tr->params_chains == 1 && g % e->cfg.n_beta == 0)
			slot = g / e->cfg.n_beta;
		if (slot >= 0)
			memcpy(e->tr_params + ((size_t) step_index * e->tr_dumped + slot) * n,
					c->params, n * sizeof(double));
	}
	c->stat_n++;
	c->stat_sum_dl += c->prob - c->prior;
	for (i = 0; i < n; i++) {
		c->stat_sum_p[i] += c->params[i];
		c->stat_sum_p2[i] += c->params[i] * c->params[i];
	}
	if (e->marg_mode)
		marginals_add(e, g, c->params);
}

int orc_run(orc_engine * e, long long n_rounds, int n_swap, const orc_trace_cfg * trace) {
	orc_trace_cfg tr = { 0, 0 };
	long long round, n_steps = n_rounds * n_swap;
	int nthreads = e->cfg.n_threads > 0 ? e->cfg.n_threads : 1;
	if (trace)
		tr = *trace;
	free(e->tr_prob);
	free(e->tr_dl);
	free(e->tr_params);
	e->tr_prob = e->tr_dl = e->tr_params = NULL;
	e->tr_prob_rows = e->tr_param_rows = 0;
	e->tr_dumped = tr.params_chains == 2 ? e->n_chains : (tr.params_chains == 1 ? e->cfg.n_ensembles : 0);
	if (tr.prob_every > 0) {
		e->tr_prob_rows = (n_steps + tr.prob_every - 1) / tr.prob_every;
		e->tr_prob = (double *) calloc((size_t) e->tr_prob_rows * e->n_chains + 1, sizeof(double));
		e->tr_dl = (double *) calloc((size_t) e->tr_prob_rows * e->n_chains + 1, sizeof(double));
	}
	if (e->tr_dumped > 0) {
		e->tr_param_rows = n_steps;
		e->tr_params = (double *) calloc((size_t) n_steps * e->tr_dumped * e->cfg.n_par + 1,
				sizeof(double));
	}
	if (e->cfg.rng_kind == ORC_RNG_MT19937)
		nthreads = 1; /* one global stream: the draw order is the result */
	for (round = 0; round < n_rounds; round++) {
		int g, ens;
		/* the reference's `omp parallel for` over chains with subiter private
		 *  */
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
		for (g = 0; g < e->n_chains; g++) {
			int subiter;
			for (subiter = 0; subiter < n_swap; subiter++)
				sampler_substep(e, g, round * n_swap + subiter, &tr);
		}
		if (e->adapt) /* adapt(chains, n_beta, n_swap), ref src/parallel_tempering.c:404 */
			for (g = 0; g < e->n_chains; g++)
				adapt_chain(e, e->chains + g);
		for (ens = 0; ens < e->cfg.n_ensembles; ens++)
			tempering_interaction(e, ens);
	}
	return 0;
}

int orc_read_trace(orc_engine * e, double * prob, double * dl, double * params,
		long long * n_prob_rows, long long * n_param_rows) {
	if (prob && e->tr_prob)
		memcpy(prob, e->tr_prob, (size_t) e->tr_prob_rows * e->n_chains * sizeof(double));
	if (dl && e->tr_dl)
		memcpy(dl, e->tr_dl, (size_t) e->tr_prob_rows * e->n_chains * sizeof(double));
	if (params && e->tr_params)
		memcpy(params, e->tr_params,
				(size_t) e->tr_param_rows * e->tr_dumped * e->cfg.n_par * sizeof(double));
	if (n_prob_rows)
		*n_prob_rows = e->tr_prob_rows;
	if (n_param_rows)
		*n_param_rows = e->tr_param_rows;
	return 0;
}

int orc_calibrate(orc_engine * e, const unsigned char * select, const orc_calib_cfg * cfg,
		int * status, orc_calib_progress * progress, long long progress_capacity,
		long long * n_progress) {
	progress_sink sink;
	int g, any_failed = 0;
	int nthreads = e->cfg.n_threads > 0 ? e->cfg.n_threads : 1;
	sink.rows = progress;
	sink.cap = progress_capacity;
	sink.n = 0;
	if (e->cfg.rng_kind == ORC_RNG_MT19937)
		nthreads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
	for (g = 0; g < e->n_chains; g++) {
		int st = -1;
		if (select == NULL || select[g]) {
			st = markov_chain_calibrate(e, g, e->chains + g, cfg, &sink);
			if (st != 0) {
#pragma omp atomic write
				any_failed = 1;
			}
		}
		if (status)
			status[g] = st;
	}
	if (n_progress)
		*n_progress = sink.n;
	return any_failed ? -6 : 0;
}

/* n_steps x { markov_chain_step_for(kind) or markov_chain_step; mcmc_check_best } with the accept
 * log assess_acceptance_rate keeps: ref src/markov_chain.c:143-172.  Chain after chain (each chain
 * does all its steps before the next starts), which in MT19937 mode is the reference's draw order
 * when one chain is selected. */
int orc_steps(orc_engine * e, const unsigned char * select, int kind, long long n_steps,
		unsigned char * accepted) {
	const int n = e->cfg.n_par;
	int g;
	long long i;
	if (kind < 0 || kind > n || n_steps < 0)
		return -1;
	if (accepted)
		memset(accepted, 0, (size_t) n_steps * e->n_chains);
	for (g = 0; g < e->n_chains; g++) {
		chain_t * c = e->chains + g;
		if (select != NULL && !select[g])
			continue;
		for (i = 0; i < n_steps; i++) {
			const u64 before = kind == n ? c->accept : c->pacc[kind];
			if (kind == n)
				markov_chain_step(e, g, c);
			else
				markov_chain_step_for(e, g, c, (unsigned) kind);
			mcmc_check_best(e, c);
			if (accepted)
				accepted[(size_t) i * e->n_chains + g] = (kind == n ? c->accept : c->pacc[kind]) != before;
		}
	}
	return 0;
}

static void marginals_free(orc_engine * e) {
	free(e->marg_counts);
	free(e->marg_bsum);
	free(e->marg_means);
	free(e->marg_n);
	free(e->marg_nb);
	e->marg_counts = NULL;
	e->marg_bsum = NULL;
	e->marg_means = NULL;
	e->marg_n = e->marg_nb = NULL;
	e->marg_mode = 0;
}

/* mirrors apm_gpu_set_marginals / apm_gpu_get_marginals */
int orc_set_marginals(orc_engine * e, int which_chains, int n_bins, u64 batch_size, int max_batches) {
	size_t slots, np;
	if (e == NULL || which_chains < 0 || which_chains > 2 || (which_chains && (n_bins < 1 || max_batches < 0)))
		return -1;
	marginals_free(e);
	if (which_chains == 0)
		return 0;
	slots = which_chains == 2 ? (size_t) e->n_chains : (size_t) e->cfg.n_ensembles;
	np = (size_t) e->cfg.n_par;
	e->marg_counts = (u64 *) calloc(slots * np * n_bins, sizeof(u64));
	e->marg_bsum = (double *) calloc(slots * np, sizeof(double));
	e->marg_means = (double *) calloc(slots * np * (max_batches > 0 ? max_batches : 1), sizeof(double));
	e->marg_n = (u64 *) calloc(slots, sizeof(u64));
	e->marg_nb = (u64 *) calloc(slots, sizeof(u64));
	e->marg_mode = which_chains;
	e->marg_bins = n_bins;
	e->marg_batch = batch_size;
	e->marg_cap = max_batches;
	return 0;
}

int orc_get_marginals(orc_engine * e, u64 * counts, double * batch_means, u64 * n_values, u64 * n_batches) {
	size_t slots, np;
	if (e == NULL || !e->marg_mode)
		return -5;
	slots = e->marg_mode == 2 ? (size_t) e->n_chains : (size_t) e->cfg.n_ensembles;
	np = (size_t) e->cfg.n_par;
	if (counts)
		memcpy(counts, e->marg_counts, slots * np * e->marg_bins * sizeof(u64));
	if (batch_means && e->marg_cap > 0)
		memcpy(batch_means, e->marg_means, slots * np * e->marg_cap * sizeof(double));
	if (n_values)
		memcpy(n_values, e->marg_n, slots * sizeof(u64));
	if (n_batches)
		memcpy(n_batches, e->marg_nb, slots * sizeof(u64));
	return 0;
}

int orc_reset_stats(orc_engine * e) {
	int g;
	for (g = 0; g < e->n_chains; g++) {
		chain_t * c = e->chains + g;
		c->stat_n = 0;
		c->stat_sum_dl = 0;
		memset(c->stat_sum_p, 0, e->cfg.n_par * sizeof(double));
		memset(c->stat_sum_p2, 0, e->cfg.n_par * sizeof(double));
	}
	return 0;
}

int orc_get_stats(orc_engine * e, u64 * n, double * sum_dl, double * sum_p, double * sum_p2) {
	int g;
	const int np = e->cfg.n_par;
	for (g = 0; g < e->n_chains; g++) {
		chain_t * c = e->chains + g;
		if (n)
			n[g] = c->stat_n;
		if (sum_dl)
			sum_dl[g] = c->stat_sum_dl;
		if (sum_p)
			memcpy(sum_p + (size_t) g * np, c->stat_sum_p, np * sizeof(double));
		if (sum_p2)
			memcpy(sum_p2 + (size_t) g * np, c->stat_sum_p2, np * sizeof(double));
	}
	return 0;
}
