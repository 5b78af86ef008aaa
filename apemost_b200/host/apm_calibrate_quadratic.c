/*
 * apm_calibrate_quadratic.c -- the reference's parabola calibrator (-DCALIBRATE_QUADRATIC) and the
 * per-parameter regression it hands over to, over the GPU engine:
 *
 *   markov_chain_calibrate_quadratic            reference src/markov_chain_calibrate.c:452-914
 *   markov_chain_calibrate_linear_regression    reference src/markov_chain_calibrate.c:239-450
 *
 * Every measurement is apm_assess_acceptance_rate (apm_assess.c -> apm_gpu_steps on the device);
 * what lives here is the host-side search.  The target is the n_par-th root of the wanted rate,
 * per parameter (single-parameter steps):
 *
 *   phase 1, per parameter: measure at the given width; guess a second width from the distance to
 *   the target; put a line through the two and measure where it hits the target; from then on keep
 *   the three latest (width, rate) pairs, put a parabola through them, solve it for the target and
 *   measure there -- until the three rates agree with the target, the parabola has no (finite)
 *   solution, the next point lies inside what is known already, or the parabola is not falling.
 *   On the way out the measurements far from the last suggested width are dropped and, if fewer
 *   than three remain, up to three more are taken around it.
 *
 *   phase 2, per parameter: weighted straight-line fit through the measurements kept, step to
 *   where it hits the target (at most 20 % outside the known range), measure, fit again, until a
 *   measurement lies within the fit's own noise of the suggestion or the table is full.
 *
 * Reference behaviour that is kept, because calibration_results and calibration_progress.data must
 * come out the same (the layout below is chosen so that it falls out by itself):
 *  - ONE running slot number is shared by all parameters' tables (:504-508: `n` is not per
 *    parameter), 30 slots in all; a table is therefore sparse, and the walk that drops far-away
 *    measurements stops at the first empty slot (:808) -- usually after the parameter's first point;
 *    running out of slots is fatal, as GSL's range check makes it in the reference;
 *  - the "rates agree with the target" test looks at the triple of the THIRD parameter whichever
 *    parameter is being worked on (min_column(acceptance_rates, 2), :598-607), and only from the
 *    pass numbered 2 on;
 *  - the pass counter is the same variable the drop-out walk counts in (:807-822), so "ITERATION"
 *    numbers jump;
 *  - "extrapolating" is declared when the next point lies INSIDE the known range (:757-766);
 *  - the triple's rates start at 1 (:478).
 */
#include "apm_session.h"

enum { SLOTS = 30 };

typedef struct {
	double width[SLOTS], rate[SLOTS], accuracy[SLOTS]; /* rate < 0: empty slot */
} measurements;

typedef struct {
	double x[3], v[3];  /* the three latest widths (normalised) and the rates measured there */
	int stage;          /* 0 nothing, 1 first width known, 2 second, 3.. parabola steps, > 100 done */
} triple;

typedef struct {
	apm_session * s;
	int g;
	mcmc * m;
	unsigned int n_par, steps_used, next_slot;
	double target;
	measurements * table; /* [n_par] */
	triple * tri;         /* [n_par] */
} quad_state;

static double fabs_(double x) {
	return x < 0 ? -x : x;
}
static double least(double a, double b) {
	return a < b ? a : b;
}
static double most(double a, double b) {
	return a > b ? a : b;
}
static int within(double x, double lo, double hi) {
	return x >= lo && x <= hi;
}
static double least3(const double * x) {
	double r = x[0];
	if (r > x[1])
		r = x[1];
	if (r > x[2])
		r = x[2];
	return r;
}
static double most3(const double * x) {
	double r = x[0];
	if (r < x[1])
		r = x[1];
	if (r < x[2])
		r = x[2];
	return r;
}

/* reference src/mcmc_gettersetter.c:155-158 */
static void set_width(mcmc * m, double normalised, unsigned int i) {
	set_steps_for(m, normalised * (get_params_max_for(m, i) - get_params_min_for(m, i)), i);
}

static unsigned int claim_slot(quad_state * q) {
	if (q->next_slot >= SLOTS) {
		fprintf(stderr, "calibration failed: more than %d acceptance rate measurements "
				"(the reference stops here with a GSL range error)\n", (int) SLOTS);
		abort();
	}
	return q->next_slot++;
}

/* measure parameter i at its current width with the phase-1 accuracy; file the result under `width` */
static double measure(quad_state * q, unsigned int i, double width, double accuracy_cap) {
	double rate, accuracy;
	unsigned int slot;
	q->steps_used += apm_assess_acceptance_rate(q->s, q->g, i, q->target, 0, accuracy_cap, &rate, &accuracy);
	slot = claim_slot(q);
	q->table[i].rate[slot] = rate;
	q->table[i].accuracy[slot] = accuracy;
	q->table[i].width[slot] = width;
	return rate;
}

/* a x^2 + b x + c through the triple (the reference's expressions, term for term: :634-681) */
static void parabola(const triple * t, double * a, double * b, double * c) {
	const double x0 = t->x[0], x1 = t->x[1], x2 = t->x[2];
	const double v0 = t->v[0], v1 = t->v[1], v2 = t->v[2];
	const double den = x0 * (x2 * x2 - x1 * x1) - x1 * (x2 * x2) + (x1 * x1) * x2 + (x0 * x0) * (x1 - x2);
	*a = -(v0 * (x2 - x1) - v1 * x2 + v2 * x1 + (v1 - v2) * x0) / den;
	*b = (v0 * (x2 * x2 - x1 * x1) - v1 * (x2 * x2) + v2 * (x1 * x1) + (v1 - v2) * (x0 * x0)) / den;
	*c = (v0 * ((x1 * x1) * x2 - x1 * (x2 * x2)) + x0 * (v1 * (x2 * x2) - v2 * (x1 * x1))
			+ (x0 * x0) * (v2 * x1 - v1 * x2)) / den;
}

/* which root of the parabola to go to: one between the known widths if there is one, else one in
 * about [0, 1], else the "minus" root (:711-745) */
static double choose_root(double a, double b, double root, double lo, double hi) {
	const double plus = (root - b) / (2 * a), minus = (-root - b) / (2 * a);
	if (within(plus, lo, hi))
		return plus;
	if (within(minus, lo, hi))
		return minus;
	if (within(plus, -0.02, 1.1))
		return plus;
	return minus;
}

/* leaving phase 1 for parameter i: forget what is far from `around`, top up to three measurements */
static unsigned int drop_out(quad_state * q, unsigned int i, double a, double b, double around, double accuracy_cap) {
	measurements * tb = &q->table[i];
	triple * t = &q->tri[i];
	const double slope = 2 * a * around + b;
	unsigned int kept = 0, k;
	printf("deleting points more than %f outside\n", fabs_(1. / slope * accuracy_cap * 10));
	for (k = 0;; k++) {
		if (k >= SLOTS) {
			fprintf(stderr, "calibration failed: walked off the table of measurements "
					"(the reference stops here with a GSL range error)\n");
			abort();
		}
		if (!(tb->rate[k] >= 0))
			break;
		if (fabs_(tb->width[k] - around) > fabs_(1. / slope * accuracy_cap * 10))
			tb->rate[k] = -1;
		else
			kept++;
	}
	if (kept < 3) {
		set_width(q->m, around, i);
		t->v[1] = measure(q, i, 0, accuracy_cap);
		tb->width[q->next_slot - 1] = get_steps_for_normalized(q->m, i);
		kept++;
	}
	if (kept < 3) {
		set_width(q->m, least(around - accuracy_cap * 3 / slope, 0.1 * around), i);
		t->v[1] = measure(q, i, 0, accuracy_cap);
		tb->width[q->next_slot - 1] = get_steps_for_normalized(q->m, i);
		set_width(q->m, most(around + accuracy_cap * 3 / slope, 1.0), i);
		t->v[1] = measure(q, i, 0, accuracy_cap);
		tb->width[q->next_slot - 1] = get_steps_for_normalized(q->m, i);
	}
	return kept;
}

/* ---- phase 2: straight-line fits, parameter by parameter ------------------------------------ */
static double weight_of(const measurements * tb, unsigned int j, double target) {
	return 1. / most(fabs_(tb->rate[j] - target), tb->accuracy[j]);
}

static void regress_each_parameter(quad_state * q, double max_ar_deviation, unsigned int step_budget) {
	const double floor_noise = max_ar_deviation * 2 / 3;
	unsigned int used = 0, i, j, l;
	int * settled = (int *) calloc(q->n_par, sizeof(int));
	FILE * plot = fopen(apm_out_path("calibration_progress.data"), "w");
	assert(plot != NULL && settled != NULL);
	for (j = 0; j < SLOTS && used < step_budget; j++)
		for (i = 0; i < q->n_par; i++)
			if (!(q->table[i].rate[j] < 0))
				fprintf(plot, "%d\t%d\t%f\t%f\t%f\n", i + 1, used, q->table[i].width[j], q->table[i].rate[j],
						q->table[i].accuracy[j]);
	for (l = 0; l < SLOTS && used < step_budget; l++) {
		for (i = 0; i < q->n_par; i++) {
			measurements * tb = &q->table[i];
			double x_bar = 0, y_bar = 0, wsum = 0, xy = 0, xx = 0, lo = 1, hi = 0;
			double k, d, sigma = 0, next, rate, accuracy;
			unsigned int n = 0, free_slot;
			if (settled[i])
				continue;
			/* the weighted line through what is in the table */
			for (j = 0; j < SLOTS; j++) {
				double w;
				if (tb->rate[j] < 0)
					continue;
				if (lo > tb->width[j])
					lo = tb->width[j];
				if (hi < tb->width[j])
					hi = tb->width[j];
				n++;
				w = weight_of(tb, j, q->target);
				x_bar += tb->width[j];
				y_bar += tb->rate[j] * w;
				wsum += w;
			}
			x_bar /= n;
			y_bar = y_bar / wsum;
			printf("  xbar = %f, ybar = %f, n=%d\n", x_bar, y_bar, n);
			for (j = 0; j < SLOTS; j++) {
				double w;
				if (tb->rate[j] < 0)
					continue;
				w = weight_of(tb, j, q->target);
				printf("  %f | %f | weight = %f\n", tb->width[j], tb->rate[j], w);
				xy += (tb->width[j] - x_bar) * (tb->rate[j] - y_bar) * w;
				xx += (tb->width[j] - x_bar) * (tb->width[j] - x_bar);
			}
			k = xy * n / wsum / xx;
			d = y_bar - k * x_bar;
			for (j = 0; j < SLOTS; j++) {
				double off;
				if (tb->rate[j] < 0)
					continue;
				off = k * tb->width[j] + d - tb->rate[j];
				sigma += weight_of(tb, j, q->target) * (off * off);
			}
			sigma = sqrt(sigma / wsum);

			/* where the line hits the target, not far outside what is known */
			next = (q->target - d) / k;
			if (next > hi + 0.2 * (hi - lo))
				next = hi + 0.2 * (hi - lo);
			if (next < lo - 0.2 * (hi - lo))
				next = lo - 0.2 * (hi - lo);
			if (next > 1)
				next = 1;
			if (next < 0)
				next = lo * 0.1;
			printf("%d: next stepwidth: %f\n", i, next);
			set_width(q->m, next, i);

			/* a measurement within the fit's own noise of the suggestion: nothing more to learn here */
			if (sigma < floor_noise)
				sigma = floor_noise;
			else
				sigma = sigma / 3;
			printf("%d: next stepwidth: %f +- %f\n", i, next, fabs_(sigma / k));
			if (n > 5)
				for (j = 0; j < SLOTS; j++) {
					if (tb->rate[j] < 0)
						continue;
					if (k < 0 && fabs_(tb->width[j] - next) < fabs_(sigma / k)) {
						printf("%d: best stepwidth possible reached.\n", i);
						settled[i] = 1;
						break;
					}
				}
			free_slot = SLOTS;
			for (j = 0; j < SLOTS; j++)
				if (tb->rate[j] < 0 && free_slot == SLOTS)
					free_slot = j;
			if (free_slot == SLOTS) {
				printf("%d: no space to store new points. should be enough.\n", i);
				settled[i] = 1;
			}
			if (!settled[i]) {
				used += apm_assess_acceptance_rate(q->s, q->g, i, q->target, sigma, sigma * 3, &rate, &accuracy);
				tb->width[free_slot] = next;
				tb->rate[free_slot] = rate;
				tb->accuracy[free_slot] = accuracy;
				fprintf(plot, "%d\t%d\t%f\t%f\t%f\t%f\n", i + 1, used, next, rate, accuracy, k);
				fflush(plot);
			}
		}
	}
	fclose(plot);
	free(settled);
}

/* ---- phase 1 ---------------------------------------------------------------------------------- */
void apm_calibrate_quadratic(apm_session * s, int g, double desired_acceptance_rate, const double max_ar_deviation,
		const unsigned int iter_limit) {
	const double accuracy_cap = 0.01;
	quad_state q;
	unsigned int i, j, pass;
	int someone_moved;
	q.s = s;
	q.g = g;
	q.m = s->chains[g];
	q.n_par = get_n_par(q.m);
	q.steps_used = 0;
	q.next_slot = 0;
	q.target = pow(desired_acceptance_rate, 1.0 / q.n_par);
	q.table = (measurements *) calloc(q.n_par, sizeof(measurements));
	q.tri = (triple *) calloc(q.n_par, sizeof(triple));
	assert(q.table != NULL && q.tri != NULL);
	for (i = 0; i < q.n_par; i++) {
		for (j = 0; j < SLOTS; j++)
			q.table[i].rate[j] = -1;
		for (j = 0; j < 3; j++)
			q.tri[i].v[j] = 1;
	}

	for (pass = 0, someone_moved = 1; someone_moved; pass++) {
		printf(" ==== ITERATION %d ==== \n", pass);
		someone_moved = 0;
		for (i = 0; i < q.n_par; i++) {
			triple * t = &q.tri[i];
			double rate, a, b, c, disc, next, lo, hi;
			if (t->stage == 0) {
				t->x[0] = get_steps_for_normalized(q.m, i);
				t->stage = 1;
			}
			if (t->stage == 1) {
				/* the given width; a first guess from how far off its rate is */
				rate = t->v[0] = measure(&q, i, t->x[0], accuracy_cap);
				t->x[1] = t->x[0] * (1 + 5 * (rate - q.target));
				if (t->x[1] < 0)
					t->x[1] = 0.05 * t->x[0];
				if (t->x[1] > 1)
					t->x[1] = 1;
				t->stage = 2;
			}
			if (t->stage == 2) {
				/* the guess; then where the line through both hits the target */
				set_width(q.m, t->x[1], i);
				t->v[1] = measure(&q, i, t->x[1], accuracy_cap);
				b = (t->v[0] - t->v[1]) / (t->x[0] - t->x[1]);
				c = t->v[1] - b * t->x[1];
				t->x[2] = (q.target - c) / b;
				if (t->x[2] < 0)
					t->x[2] = 0.05 * t->x[1];
				if (t->x[2] > 1)
					t->x[2] = 1;
				t->stage = 3;
			}
			if (!within(t->stage, 3, 99))
				continue;
			set_width(q.m, t->x[2], i);
			t->v[2] = measure(&q, i, t->x[2], accuracy_cap);
			if (pass > 1) {
				/* (the THIRD parameter's triple, whichever parameter this is) */
				if (q.n_par < 3) {
					fprintf(stderr, "calibration failed: fewer than 3 parameters "
							"(the reference stops here with a GSL range error)\n");
					abort();
				}
				printf("%d: a/r currently between [%f..%f] \n", i, least3(q.tri[2].v), most3(q.tri[2].v));
				if (most3(q.tri[2].v) < q.target + max_ar_deviation * 4
						&& least3(q.tri[2].v) > q.target - max_ar_deviation * 4) {
					printf("%d: a/r sufficient.\n", i);
					t->stage += 100;
					continue;
				}
				if (pass > 100)
					break;
			}
			parabola(t, &a, &b, &c);
			disc = 4 * a * (q.target - c) + b * b;
			if (disc < 0 || !(a - a == 0 && b - b == 0 && c - c == 0)) {
				printf(" polynomial has no solutions. \n");
				t->stage += 100;
				continue;
			}
			lo = least3(t->x);
			hi = most3(t->x);
			next = choose_root(a, b, sqrt(disc), lo, hi);
			if (t->stage > 4 && within((next - lo) / (hi - lo), -0.1, 1.1))
				t->stage += 100;
			if (t->stage > 2 && ((b + 2 * a * least(next, lo) > 0) || (b + 2 * a * most(next, hi) > 0)))
				t->stage += 100; /* not falling over the range of interest */
			/* the triple moves on */
			t->x[0] = t->x[1];
			t->x[1] = t->x[2];
			t->x[2] = next;
			if (t->x[2] <= 0)
				t->x[2] = 0.1 * t->x[1];
			if (t->x[2] > 1)
				t->x[2] = 1;
			next = t->x[2];
			t->v[0] = t->v[1];
			t->v[1] = t->v[2];
			t->v[2] = 0;
			t->stage += 1;
			someone_moved = next != -1;
			if (t->stage > 100)
				pass = drop_out(&q, i, a, b, next, accuracy_cap); /* (the walk counts in the pass counter, :807-838) */
		}
	}
	regress_each_parameter(&q, max_ar_deviation, iter_limit - q.steps_used);
	free(q.table);
	free(q.tri);
	/* the widths chosen last live in the host struct only */
	apm_session_push(s, g, 1);
}
