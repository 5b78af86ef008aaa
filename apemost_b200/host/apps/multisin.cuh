// multisin.cuh -- the __device__ counterpart of multisin.c (see there for the model), in the
// interface documented at the top of apemost_b200/csrc/apm_models.cuh.  Built into a private
// engine library by `make -C apemost_b200/host multisin.exe` (-DAPM_USER_MODEL_HEADER).
#pragma once

struct UserModel {
	static constexpr int MAX_SINES = (APM_MAX_PAR - 1) / 3;
	static constexpr int LL_C = 2, LL_U = 2; // register tile: 2 chains x 2 rows in flight
	static constexpr int NPAR = 0, NCOLS = 2; // 3K + 1 parameters, K decided by the params file
	static constexpr bool HAS_DATA = true, HAS_PRIOR = false;
	static constexpr int ROW_W = 2;             // table rows of 2 doubles: accum gets (x, y)
	struct Prep {
		int n_sines;
		double a[MAX_SINES], f[MAX_SINES], ph[MAX_SINES], offset;
	};
	APM_D static void prep(Prep & q, const double * p, int n_par, const double *) {
		q.n_sines = (n_par - 1) / 3;
		q.offset = p[n_par - 1];
#pragma unroll
		for (int k = 0; k < MAX_SINES; k++) {
			const bool on = k < q.n_sines;
			q.a[k] = on ? p[3 * k] : 0.0;
			q.f[k] = on ? p[3 * k + 1] : 0.0;
			q.ph[k] = on ? p[3 * k + 2] : 0.0;
		}
	}
	template<bool FAST>
	APM_D static double row(double acc, const Prep & q, double x, double y) {
		double model = 0.0;
#pragma unroll
		for (int k = 0; k < MAX_SINES; k++) {
			if (k < q.n_sines) {
				// the argument is rounded like the host code: 2 pi * (f * x + phi), no contraction
				const double arg = apm::mul_rn(APM_TWO_PI, apm::add_rn(apm::mul_rn(q.f[k], x), q.ph[k]));
				model = fma(q.a[k], FAST ? apm::sin_fast(arg) : apm::sin_full(arg), model);
			}
		}
		const double d = (model + q.offset) - y;
		return fma(d, d, acc);
	}
	APM_D static double accum(double acc, const Prep & q, double x, double y) { return row<false>(acc, q, x, y); }
	APM_D static double accum_fast(double acc, const Prep & q, double x, double y) { return row<true>(acc, q, x, y); }
	APM_D static bool fast_ok(const Prep & q, double xub) {
		bool ok = true;
#pragma unroll
		for (int k = 0; k < MAX_SINES; k++)
			ok = ok && APM_TWO_PI * (fabs(q.f[k]) * xub + fabs(q.ph[k])) < APM_SIN_FAST_BOUND;
		return ok;
	}
	APM_D static double sum0(const double *) { return 0.0; }
	APM_D static double prior(const double *, int, const double *) { return 0.0; }
	APM_D static double finish(double beta, double sum, double, const double *, const double * mc) {
		const double sigma = mc[0] != 0 ? mc[0] : 0.5; // SIGMA
		return beta * sum / (-2 * sigma * sigma);
	}
};
