// apm_models.cuh -- __device__ counterparts of the reference's apps/<model>.c plugins.
//
// The reference binds one calc_model() per executable at link time
// (Makefile:54-55; contract src/mcmc.h:164,173, doc/manual.rst:126-168).  On the
// device a model is a struct with the static interface below; the kernels are
// templates over it, so a model costs nothing at run time:
//
//   NPAR        number of parameters (0 = decided at run time, n_par is passed in)
//   NCOLS       data columns the model reads (the row-major table may not be narrower);
//               -1 = one column per parameter (apps/bernoulli_example.c)
//   ROW_W       doubles per table row on the device: 2 (accum gets the row as x, y) or 4 (models
//               reading 3 or 4 columns; accum gets a double4, unused columns are zero)
//   HAS_DATA    false for models that never touch m->data (apps/normal.c)
//   HAS_PRIOR   whether calc_model calls set_prior()
//   LL_C, LL_U  register tile of the likelihood kernel: chains per work item x rows per
//               inner iteration (LL_C x LL_U independent row evaluations in flight per thread)
//   Prep        per-chain constants derived once per proposal (kept in registers)
//   prep()      fill Prep from the parameter vector
//   accum()      acc + (one data row's contribution to the model's running sum), exact
//                for every input (falls back to the math library outside sin_fast's range)
//   accum_fast() the same on the branch-free fast path (sin_fast), valid only when
//   fast_ok()    says so: fast_ok(q, xub) is evaluated once per (chain, thread, chunk) from
//                an upper bound xub >= |x| of the thread's rows, so the per-row code carries
//                no range check at all; where it says no, the kernel uses accum()
//   Acc, acc_zero(), acc_merge(), acc_value()   optional: the type of a thread's running sum when
//               a double is not the best carrier (pulse*: sum and product, one logarithm per thread
//               instead of one per row); accum / accum_fast then take and return an Acc
//   prior()     value passed to set_prior()
//   sum0()      initial value of the running sum (apps/pulse*.c start it at params[1])
//   finish()    the value passed to set_prob()
//
// Arithmetic follows each reference file's operation order; the argument of sin is
// rounded exactly like the reference (no fused multiply-add), because at |arg| ~ 1e4
// a half-ulp there is ~1e-12 absolute.  The whole library is compiled with
// -fmad=false: every fused operation below is an explicit fma().
//
// A user model: write apps/<name>.cuh defining `struct UserModel` with this
// interface and build with -DAPM_USER_MODEL_HEADER='"<name>.cuh"' (see
// apemost_b200/host/Makefile); it is then reachable as APM_MODEL_USER.
#pragma once

#include "apm_math.cuh"

namespace apm {

#define APM_MAX_PAR 16
#define APM_TWO_PI 6.283185307179586 /* (2.0 * M_PI) as the compiler folds it */
#ifndef APM_LL_C
#define APM_LL_C 8
#endif
#ifndef APM_LL_U
#define APM_LL_U 2
#endif

// ---- apps/simplesin.c:12-38 -----------------------------------------------------------
struct ModelSimplesin {
	static constexpr int LL_C = APM_LL_C, LL_U = APM_LL_U;
	static constexpr int NPAR = 4, NCOLS = 2;
	static constexpr bool HAS_DATA = true, HAS_PRIOR = false;
	static constexpr int ROW_W = 2;
	struct Prep {
		double amplitude, frequency, phase, offset;
	};
	APM_D static void prep(Prep & q, const double * p, int, const double *) {
		q.amplitude = p[0];
		q.frequency = p[1];
		q.phase = p[2];
		q.offset = p[3];
	}
	// y_model = amplitude * sin(2.0 * M_PI * (frequency * x + phase)) + offset   (:15)
	APM_D static double accum(double acc, const Prep & q, double x, double y) {
		double arg = mul_rn(APM_TWO_PI, add_rn(mul_rn(q.frequency, x), q.phase));
		double deltay = fma(q.amplitude, sin_full(arg), q.offset) - y;
		return fma(deltay, deltay, acc);
	}
	APM_D static double accum_fast(double acc, const Prep & q, double x, double y) {
		double arg = mul_rn(APM_TWO_PI, add_rn(mul_rn(q.frequency, x), q.phase));
		double deltay = fma(q.amplitude, sin_fast(arg), q.offset) - y;
		return fma(deltay, deltay, acc);
	}
	APM_D static bool fast_ok(const Prep & q, double xub) {
		return APM_TWO_PI * (fabs(q.frequency) * xub + fabs(q.phase)) < APM_SIN_FAST_BOUND;
	}
	APM_D static double sum0(const double *) { return 0.0; }
	APM_D static double prior(const double *, int, const double *) { return 0.0; }
	APM_D static double finish(double beta, double sum, double, const double *, const double * mc) {
		const double sigma = mc[0] != 0 ? mc[0] : 0.5; // SIGMA (:8-10)
		return beta * sum / (-2 * sigma * sigma);       // (:36)
	}
};

// ---- apps/simplesin5.c:15-41 (formula) ----------------------------------
struct ModelSimplesin5 {
	static constexpr int LL_C = APM_LL_C, LL_U = APM_LL_U;
	static constexpr int NPAR = 4, NCOLS = 2;
	static constexpr bool HAS_DATA = true, HAS_PRIOR = false;
	static constexpr int ROW_W = 2;
	struct Prep {
		double a, w, ph, off;
	};
	APM_D static void prep(Prep & q, const double * p, int, const double *) {
		q.a = p[0];
		q.w = mul_rn(APM_TWO_PI, p[1]); // 2.0 * M_PI * param1 associates to the left (:18)
		q.ph = p[2];
		q.off = p[3];
	}
	APM_D static double accum(double acc, const Prep & q, double x, double y) {
		double arg = add_rn(mul_rn(q.w, x), q.ph);
		double d = fma(q.a, sin_full(arg), q.off) - y; // (:35-36)
		return fma(d, d, acc);
	}
	APM_D static double accum_fast(double acc, const Prep & q, double x, double y) {
		double arg = add_rn(mul_rn(q.w, x), q.ph);
		double d = fma(q.a, sin_fast(arg), q.off) - y;
		return fma(d, d, acc);
	}
	APM_D static bool fast_ok(const Prep & q, double xub) {
		return fabs(q.w) * xub + fabs(q.ph) < APM_SIN_FAST_BOUND;
	}
	APM_D static double sum0(const double *) { return 0.0; }
	APM_D static double prior(const double *, int, const double *) { return 0.0; }
	APM_D static double finish(double beta, double sum, double, const double *, const double * mc) {
		const double sigma = mc[0] != 0 ? mc[0] : 0.5;
		return beta * sum / (-2 * sigma * sigma);
	}
};

// ---- apps/simplesin2.c:12-32 ------------------------------------------------------------
struct ModelSimplesin2 {
	static constexpr int LL_C = APM_LL_C, LL_U = APM_LL_U;
	static constexpr int NPAR = 2, NCOLS = 2;
	static constexpr bool HAS_DATA = true, HAS_PRIOR = false;
	static constexpr int ROW_W = 2;
	struct Prep {
		double a, f;
	};
	APM_D static void prep(Prep & q, const double * p, int, const double *) {
		q.a = p[0];
		q.f = p[1];
	}
	APM_D static double accum(double acc, const Prep & q, double x, double y) {
		double arg = mul_rn(APM_TWO_PI, add_rn(mul_rn(q.f, x), 0.3312));
		double d = fma(q.a, sin_full(arg), -y);
		return fma(d, d, acc);
	}
	APM_D static double accum_fast(double acc, const Prep & q, double x, double y) {
		double arg = mul_rn(APM_TWO_PI, add_rn(mul_rn(q.f, x), 0.3312));
		double d = fma(q.a, sin_fast(arg), -y);
		return fma(d, d, acc);
	}
	APM_D static bool fast_ok(const Prep & q, double xub) {
		return APM_TWO_PI * (fabs(q.f) * xub + 0.3312) < APM_SIN_FAST_BOUND;
	}
	APM_D static double sum0(const double *) { return 0.0; }
	APM_D static double prior(const double *, int, const double *) { return 0.0; }
	APM_D static double finish(double beta, double sum, double, const double *, const double * mc) {
		const double sigma = mc[0] != 0 ? mc[0] : 0.5;
		return beta * sum / (-2 * sigma * sigma);
	}
};

// ---- apps/normal.c:8-34 (data-free) -------------------------------------------------------
// pos = exp(i) (:20), i = 0..9: the correctly rounded values (what the C library returns), as
// constants -- ten calls of exp() per step were half of a step of this model
static __constant__ double NORMAL_EXP_I[10] = { 0x1.0000000000000p+0, 0x1.5bf0a8b145769p+1, 0x1.d8e64b8d4ddaep+2,
		0x1.415e5bf6fb106p+4, 0x1.b4c902e273a58p+5, 0x1.28d389970338fp+7, 0x1.936dc5690c08fp+8,
		0x1.122885aaeddaap+10, 0x1.749ea7d470c6ep+11, 0x1.fa7157c470f82p+12 };

// 1 / i, correctly rounded, for div_by_known (entry 0 is never used)
static __constant__ double NORMAL_INV_I[10] = { 0.0, 0x1.0000000000000p+0, 0x1.0000000000000p-1, 0x1.5555555555555p-2,
		0x1.0000000000000p-2, 0x1.999999999999ap-3, 0x1.5555555555555p-3, 0x1.2492492492492p-3, 0x1.0000000000000p-3,
		0x1.c71c71c71c71cp-4 };

struct ModelNormal {
	static constexpr int LL_C = 1, LL_U = 1;
	static constexpr int NPAR = 1, NCOLS = 0;
	static constexpr bool HAS_DATA = false, HAS_PRIOR = false;
	static constexpr int ROW_W = 2;
	struct Prep {
		double x;
	};
	APM_D static void prep(Prep & q, const double * p, int, const double *) { q.x = p[0]; }
	APM_D static double accum(double acc, const Prep &, double, double) { return acc; }
	APM_D static double accum_fast(double acc, const Prep &, double, double) { return acc; }
	APM_D static bool fast_ok(const Prep &, double) { return true; }
	APM_D static double sum0(const double *) { return 0.0; }
	APM_D static double prior(const double *, int, const double *) { return 0.0; }
	// Data-free models may split one evaluation over lanes (the kernels then give a chain several
	// lanes instead of one thread): LANE_TERMS independent terms, term(i, p), reduced by the
	// ORDER-INDEPENDENT reduce() from reduce_init(), and finish_reduced(beta, r) -- finish() below is
	// the same terms taken one after the other, so both give the same bits.
	static constexpr int LANE_TERMS = 10;
	// bump i of the loop (:19-33): pos = exp(i), height = 10 * pow(1.0, i) = 10, sigma = i.  One
	// division either way: even i: -sigma * pow((x - pos) / sigma, 2) / 2 + height, odd i:
	// -height * |x - pos| / sigma + height (the reference's two branches differ in the sign of the
	// difference only).  i = 0 divides by sigma = 0: its a is NaN (0 * inf) for every x and never
	// passes `a > b`, so it is handed out as NaN without the division.  The divisions by 1 .. 9 are
	// div_by_known (apm_math.cuh): the IEEE quotient in three instructions.
	APM_D static double term(int i, const double * p, const double *) {
		// (written for a short chain of dependent instructions, each step the reference's value bit
		// for bit: pos - x is -(x - pos) exactly, so |x - pos| is one subtraction with the sign dropped;
		// 1.0 * v is v; halving is exact, so (-sigma * q^2) / 2 is q^2 * (-sigma / 2))
		const double x = p[0];
		const double pos = NORMAL_EXP_I[i];
		const double height = 10.0;
		const double sigma = (double) i;
		const bool even = i % 2 == 0;
		const double v0 = x - pos;
		const double num = (even ? 1.0 : -height) * (even ? v0 : fabs(v0));
		if (i == 0)
			return __longlong_as_double(0x7ff8000000000000ll);
		const double quo = div_by_known(num, sigma, NORMAL_INV_I[i]); // = num / sigma, bit for bit
		return (even ? (quo * quo) * (-0.5 * sigma) : quo) + height;
	}
	APM_D static double reduce_init() { return 0.0; }                           // b = 0 (:11)
	APM_D static double reduce(double b, double a) { return a > b ? a : b; }    // if (a > b) b = a (:31-32)
	APM_D static double finish_reduced(double beta, double b) { return beta * b; } // (:34)
	APM_D static double finish(double beta, double, double, const double * p, const double * mc) {
		double b = reduce_init();
#pragma unroll
		for (int i = 0; i < LANE_TERMS; i++)
			b = reduce(b, term(i, p, mc));
		return finish_reduced(beta, b);
	}
};

// ---- apps/pulse_vrot.c:12-65 ----------------------------------------------------------------
#ifndef APM_PV_C
#define APM_PV_C 4 /* register tile of the likelihood kernel for pulse_vrot (overridable for sweeps) */
#endif
#ifndef APM_PV_U
#define APM_PV_U 1
#endif
struct ModelPulseVrot {
	static constexpr int LL_C = APM_PV_C, LL_U = APM_PV_U;
	static constexpr int NPAR = 7, NCOLS = 2;
	static constexpr bool HAS_DATA = true, HAS_PRIOR = true;
	static constexpr int ROW_W = 2;
	struct Prep {
		double lifetime, vrot, f1, h1, f2, h2;
	};
	APM_D static void prep(Prep & q, const double * p, int, const double *) {
		q.lifetime = p[0];
		q.vrot = p[2];
		q.f1 = p[3];
		q.h1 = p[4];
		q.f2 = p[5];
		q.h2 = p[6];
	}
	APM_D static double lorentz(double distance, double lifetime, double height) {
		// mode_height / (1 + pow(2 * M_PI * distance * lifetime, 2))   (:49)
		double t = (APM_TWO_PI * distance) * lifetime;
		return height / (1 + t * t);
	}
	// The running sum of a thread carries the product of the quotients r = 1 / y next to the sum of
	// d r: sum_i log r_i = log prod_i r_i, so the fast path multiplies (one DMUL, the exponent moved
	// to an integer counter every row so the product stays in [1, 2)) where it would take a
	// logarithm (25 FP64 instructions), and acc_value takes ONE logarithm per thread.
	struct Acc {
		double sum, prod;
		int e;
	};
	APM_D static Acc acc_zero() {
		Acc a;
		a.sum = 0.0;
		a.prod = 1.0;
		a.e = 0;
		return a;
	}
	APM_D static Acc acc_merge(const Acc & a, const Acc & b) {
		Acc c;
		c.sum = a.sum + b.sum;
		c.prod = a.prod * b.prod;
		c.e = a.e + b.e;
		return c;
	}
	APM_D static double acc_value(const Acc & a) {
		const double LN2_HI = 0x1.62e42fefa39efp-1, LN2_LO = 0x1.abc9e3b39803fp-56;
		const double e = (double) a.e;
		return a.sum - fma(e, LN2_HI, fma(e, LN2_LO, log_pos(a.prod)));
	}
	APM_D static Acc acc_add_row(const Acc & acc, double d, double r) {
		Acc o;
		o.sum = fma(d, r, acc.sum);
		const double pr = acc.prod * r;
		const int hi = hi32(pr);
		o.e = acc.e + ((hi >> 20) - 1023);
		o.prod = make_double((hi & 0x000fffff) | 0x3ff00000, lo32(pr));
		return o;
	}
	APM_D static Acc accum(const Acc & acc, const Prep & q, double freq, double d) {
		double y = 0;
		y += lorentz(q.f1 - freq, q.lifetime, q.h1);
		y += lorentz(q.f2 - freq + -1 * q.vrot, q.lifetime, q.h2);
		y += lorentz(q.f2 - freq, q.lifetime, q.h2);
		y += lorentz(q.f2 - freq + 1 * q.vrot, q.lifetime, q.h2);
		Acc o = acc;
		o.sum = acc.sum + (log(y) + d / y); // (:61)
		return o;
	}
	// The same row term with ONE division instead of five: y = P / Q over the common denominator
	// Q = d1 d2 d3 d4, d_k = 1 + t_k^2 (all terms positive: no cancellation), log y + d / y =
	// -log(Q / P) + d (Q / P).  A few ulps from the reference's sum of quotients, far inside
	// the 1e-12 parity bound; ~30 % fewer FP64 instructions.  fast_ok keeps Q far from overflow.
	APM_D static Acc accum_fast(const Acc & acc, const Prep & q, double freq, double d) {
		const double tw = APM_TWO_PI * q.lifetime;
		const double t1 = (q.f1 - freq) * tw, t2 = (q.f2 - freq + -1 * q.vrot) * tw;
		const double t3 = (q.f2 - freq) * tw, t4 = (q.f2 - freq + 1 * q.vrot) * tw;
		const double d1 = fma(t1, t1, 1.0), d2 = fma(t2, t2, 1.0), d3 = fma(t3, t3, 1.0), d4 = fma(t4, t4, 1.0);
		const double d34 = d3 * d4;
		const double Q = (d1 * d2) * d34;
		// P = h1 d2 d3 d4 + h2 d1 (d3 d4 + d2 d4 + d2 d3)
		const double P = fma(q.h1 * d2, d34, (q.h2 * d1) * fma(d2, d3 + d4, d34));
		// = 1 / y; the quotient without the library's slow-path branch, so that the chains (or rows)
		// a thread works on interleave (apm_math.cuh); its logarithm goes into the product
		return acc_add_row(acc, d, div_pos(Q, P));
	}
	APM_D static bool fast_ok(const Prep & q, double xub) {
		const double t = APM_TWO_PI * fabs(q.lifetime) * (fabs(q.f1) + fabs(q.f2) + fabs(q.vrot) + xub);
		// d_k < 1e60: 1 <= Q < 1e240; heights in (1e-30, 1e30): 1e-30 < P < 1e211
		return t < 1e30 && q.h1 > 1e-30 && q.h2 > 1e-30 && q.h1 < 1e30 && q.h2 < 1e30;
	}
	APM_D static double sum0(const double * p) { return p[1]; } // accumulator starts at params[1] (:34)
	APM_D static double prior(const double * p, int n_par, const double * mc) {
		const double hmin = mc[0] != 0 ? mc[0] : 1e-6; // HMIN (:8-10)
		double prior = 0;
		for (int i = 3; i < n_par; i += 2)
			prior += log(p[i + 1] + hmin);
		return -prior / (double) ((n_par - 3) / 2); // (:12-23)
	}
	APM_D static double finish(double beta, double sum, double prior, const double *, const double *) {
		return prior + -beta * sum; // (:64)
	}
};

// ---- apps/pulse.c:12-56 (n_par = 2 + 2k, k modes) -------------------------------------------
struct ModelPulse {
	static constexpr int LL_C = 2, LL_U = 1;
	static constexpr int NPAR = 0, NCOLS = 2;
	static constexpr bool HAS_DATA = true, HAS_PRIOR = true;
	static constexpr int ROW_W = 2;
	struct Prep {
		double lifetime;
		int n_modes;
		double f[(APM_MAX_PAR - 2) / 2], h[(APM_MAX_PAR - 2) / 2];
	};
	APM_D static void prep(Prep & q, const double * p, int n_par, const double *) {
		q.lifetime = p[0];
		q.n_modes = (n_par - 2) / 2;
#pragma unroll
		for (int j = 0; j < (APM_MAX_PAR - 2) / 2; j++) {
			if (j < q.n_modes) {
				q.f[j] = p[2 + 2 * j];
				q.h[j] = p[3 + 2 * j];
			}
		}
	}
	typedef ModelPulseVrot::Acc Acc; // sum of d r, product of r (see there)
	APM_D static Acc acc_zero() { return ModelPulseVrot::acc_zero(); }
	APM_D static Acc acc_merge(const Acc & a, const Acc & b) { return ModelPulseVrot::acc_merge(a, b); }
	APM_D static double acc_value(const Acc & a) { return ModelPulseVrot::acc_value(a); }
	APM_D static Acc accum(const Acc & acc, const Prep & q, double freq, double d) {
		double y = 0;
#pragma unroll
		for (int j = 0; j < (APM_MAX_PAR - 2) / 2; j++) {
			if (j < q.n_modes)
				y += ModelPulseVrot::lorentz(q.f[j] - freq, q.lifetime, q.h[j]);
		}
		Acc o = acc;
		o.sum = acc.sum + (log(y) + d / y);
		return o;
	}
	// one division for any number of modes: P_j / Q_j = sum_{i <= j} h_i / d_i by the recurrence
	// P <- P d_j + h_j Q, Q <- Q d_j (all terms positive), then -log(Q / P) + d (Q / P) as in pulse_vrot
	APM_D static Acc accum_fast(const Acc & acc, const Prep & q, double freq, double d) {
		const double tw = APM_TWO_PI * q.lifetime;
		double P = 0.0, Q = 1.0;
#pragma unroll
		for (int j = 0; j < (APM_MAX_PAR - 2) / 2; j++) {
			if (j < q.n_modes) {
				const double t = (q.f[j] - freq) * tw;
				const double dj = fma(t, t, 1.0);
				P = fma(P, dj, q.h[j] * Q);
				Q = Q * dj;
			}
		}
		return ModelPulseVrot::acc_add_row(acc, d, div_pos(Q, P));
	}
	APM_D static bool fast_ok(const Prep & q, double xub) {
		double fmax = 0.0;
		bool positive = true;
#pragma unroll
		for (int j = 0; j < (APM_MAX_PAR - 2) / 2; j++) {
			if (j < q.n_modes) {
				fmax = fmax > fabs(q.f[j]) ? fmax : fabs(q.f[j]);
				positive = positive && q.h[j] > 1e-30 && q.h[j] < 1e30;
			}
		}
		// d_j < 1e38: 1 <= Q < 1e266, 1e-30 < P < 7e258
		return APM_TWO_PI * fabs(q.lifetime) * (fmax + xub) < 1e19 && positive && q.n_modes > 0;
	}
	APM_D static double sum0(const double * p) { return p[1]; }
	APM_D static double prior(const double * p, int n_par, const double * mc) {
		const double hmin = mc[0] != 0 ? mc[0] : 1e-6;
		double prior = 0;
		for (int i = 2; i < n_par; i += 2)
			prior += log(p[i + 1] + hmin);
		return -prior / (double) ((n_par - 2) / 2);
	}
	APM_D static double finish(double beta, double sum, double prior, const double *, const double *) {
		return prior + -beta * sum;
	}
};

// ---- apps/bernoulli_example.c:9-51: logistic regression, column 0 = outcome (0 / 1), columns
// 1 .. n_par - 1 = regressors, parameter 0 = intercept.  On the device: up to 4 columns.
struct ModelBernoulli {
	static constexpr int LL_C = 4, LL_U = 2;
	static constexpr int NPAR = 0, NCOLS = -1;
	static constexpr bool HAS_DATA = true, HAS_PRIOR = true;
	static constexpr int ROW_W = 4;
	struct Prep {
		double p0, p1, p2, p3;
		int n;
	};
	APM_D static void prep(Prep & q, const double * p, int n_par, const double *) {
		q.n = n_par;
		q.p0 = p[0];
		q.p1 = n_par > 1 ? p[1] : 0.0;
		q.p2 = n_par > 2 ? p[2] : 0.0;
		q.p3 = n_par > 3 ? p[3] : 0.0;
	}
	APM_D static double accum(double acc, const Prep & q, const double4 & r) {
		// eta = [1, x] . params, accumulated term by term (:29-33)
		double eta = q.p0;
		if (q.n > 1)
			eta = add_rn(eta, mul_rn(r.y, q.p1));
		if (q.n > 2)
			eta = add_rn(eta, mul_rn(r.z, q.p2));
		if (q.n > 3)
			eta = add_rn(eta, mul_rn(r.w, q.p3));
		double p_i;
		if (eta > 0)
			p_i = 1 / (1 + exp(-eta));               // (:35-36)
		else
			p_i = exp(eta) / (1 + exp(eta));         // (:37-38)
		const double l_i = r.x == 0 ? log(1 - p_i) : log(p_i); // (:40-43)
		return acc + l_i;
	}
	APM_D static double accum_fast(double acc, const Prep & q, const double4 & r) { return accum(acc, q, r); }
	APM_D static bool fast_ok(const Prep &, double) { return true; }
	APM_D static double sum0(const double *) { return 0.0; }
	APM_D static double prior(const double * p, int n_par, const double * mc) {
		const double sigma = mc[0] != 0 ? mc[0] : 2.0; // SIGMA (:7)
		double prior = 0;
		for (int j = 1; j < n_par; j++) { // j < m->data->size2 = n_par (:20-22, assert :25)
			const double t = p[j] / sigma;
			prior += -(t * t) / 2;
		}
		return prior;
	}
	APM_D static double finish(double beta, double sum, double prior, const double *, const double *) {
		return prior + beta * sum; // (:46)
	}
};

#ifdef APM_USER_MODEL_HEADER
} // namespace apm
#include APM_USER_MODEL_HEADER
namespace apm {
#endif

} // namespace apm
