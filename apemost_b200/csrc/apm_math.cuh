// apm_math.cuh -- numeric building blocks of the engine: the fp64 sine used by the
// likelihood kernels, the Philox4x32-10 counter RNG and the proposal transforms.
//
// Everything here is __host__ __device__ so that tests/test_math_cpu.py can compile
// the very same source with g++ -mfma (IEEE fma is deterministic, so the CPU run
// reproduces the device arithmetic bit for bit) and check it against 50-digit
// references without a GPU.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define APM_HD __host__ __device__ __forceinline__
#define APM_D __device__ __forceinline__
#else
#define APM_HD static inline
#define APM_D static inline
#endif

namespace apm {

// ---- bit access ----------------------------------------------------------------
APM_HD int hi32(double x) {
#if defined(__CUDA_ARCH__)
	return __double2hiint(x);
#else
	int64_t b;
	memcpy(&b, &x, 8);
	return (int) (b >> 32);
#endif
}
APM_HD int lo32(double x) {
#if defined(__CUDA_ARCH__)
	return __double2loint(x);
#else
	int64_t b;
	memcpy(&b, &x, 8);
	return (int) (b & 0xffffffff);
#endif
}
APM_HD double make_double(int hi, int lo) {
#if defined(__CUDA_ARCH__)
	return __hiloint2double(hi, lo);
#else
	int64_t b = ((int64_t) hi << 32) | (uint32_t) lo;
	double x;
	memcpy(&x, &b, 8);
	return x;
#endif
}
// strictly rounded (never contracted) multiply / add: used where the reference's
// own rounding sequence must be reproduced (the argument of sin)
APM_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
	return __dmul_rn(a, b);
#else
	volatile double r = a * b;
	return r;
#endif
}
APM_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
	return __dadd_rn(a, b);
#else
	volatile double r = a + b;
	return r;
#endif
}

// ---- sine ------------------------------------------------------------------------
// sin(x) for |x| < 2^30 in 13 FP64-pipe instructions and no table, select or
// conversion instruction (CUDA's sin() costs 14 FP64 + 2 F2I/I2F + 3 LDG + 6 FSEL and
// a slow-path call):
//   q = rint(x / pi) by the 1.5*2^52 magic-number add (1 DFMA + 1 DADD), its parity is
//   the low mantissa bit; r = x - q*pi by a two-term Cody-Waite reduction (2 DFMA; the
//   first is exact, the third term q*3e-33 is below 1e-23 for every q < 2^31);
//   sin(x) = (-1)^q sin(r), |r| <= pi/2: odd minimax polynomial of degree 15 in r
//   (Remez, weighted for absolute error; max error 1.1e-16 before rounding) evaluated
//   as r + (r*s)*P(s), s = r*r: 1 DMUL + 6 DFMA + 1 DMUL + 1 DFMA.
// Measured against 40-digit references (tests/test_math_cpu.py): max absolute error
// 2.7e-16 (rms 0.8e-16, unbiased) for |x| up to 1e9 -- inside the "few ulp" of the
// reference's own gsl_sf_sin.  Outside the fast range (or NaN/Inf) the
// caller falls back to the CUDA library sin().
#define APM_SIN_FAST_LIMIT 1073741824.0 /* 2^30 */
#define APM_SIN_FAST_BOUND 1.0e9          /* what the per-visit bound check compares with (slack for rounding) */

APM_HD double sin_fast(double x) {
	const double MAGIC = 6755399441055744.0; /* 1.5 * 2^52 */
	const double INV_PI = 0.3183098861837907;
	const double PI_HI = 3.141592653589793;
	const double PI_LO = 1.2246467991473532e-16;
	double t = fma(x, INV_PI, MAGIC);
	int q = lo32(t);
	double qd = t - MAGIC;
	double r = fma(qd, -PI_HI, x);
	r = fma(qd, -PI_LO, r);
	// (-1)^q: move q's parity into r's sign bit (sin is odd)
	r = make_double(hi32(r) ^ (q << 31), lo32(r));
	double s = r * r;
#if !defined(APM_SIN_DEGREE) || APM_SIN_DEGREE == 15
	/* degree 15 (default): polynomial error 1.1e-16, 13 FP64 instructions in total */
	double p = -0x1.9e96f0e4ab7e2p-41;
	p = fma(p, s, 0x1.60e23f9c870eep-33);
	p = fma(p, s, -0x1.ae6335183e8ccp-26);
	p = fma(p, s, 0x1.71de379039620p-19);
	p = fma(p, s, -0x1.a01a0198a4c74p-13);
	p = fma(p, s, 0x1.111111110723ap-7);
	p = fma(p, s, -0x1.5555555555421p-3);
#else
	/* degree 17 (-DAPM_SIN_DEGREE=17): polynomial error 2e-19, one more DFMA */
	double p = 0x1.87c623b020b36p-49;
	p = fma(p, s, -0x1.ae3f136452c88p-41);
	p = fma(p, s, 0x1.6123b9f483e33p-33);
	p = fma(p, s, -0x1.ae64547e37987p-26);
	p = fma(p, s, 0x1.71de3a51c5b7bp-19);
	p = fma(p, s, -0x1.a01a01a012713p-13);
	p = fma(p, s, 0x1.1111111111092p-7);
	p = fma(p, s, -0x1.5555555555555p-3);
#endif
	return fma(r * s, p, r);
}

// true when x is outside sin_fast's range (|x| >= 2^30, NaN, Inf): one LOP3 + one ISETP
APM_HD bool sin_fast_out_of_range(double x) {
	return (unsigned) (hi32(x) & 0x7fffffff) >= 0x41d00000u;
}

APM_HD double sin_full(double x) {
	if (fabs(x) < APM_SIN_FAST_LIMIT)
		return sin_fast(x);
	return sin(x); /* Payne-Hanek path of the math library; NaN/Inf land here too */
}

// ---- branch-free quotient and logarithm for positive normal arguments -------------------
// The math library's a / b and log(x) each carry a slow-path branch (denormals, infinities,
// huge quotients); inside a row loop those branches keep the compiler from interleaving the
// independent chains of a thread, and the FP64 pipe idles on every dependent chain (ncu on
// pulse_vrot's row term: 49 % pipe, `wait` the top stall; profiles/r01n_full_pulse_vrot.json).
// These versions have no branch and no call.  Domain -- the caller's fast_ok() vouches for it --:
// every argument positive, finite, in [1e-300, 1e300], and so is the quotient.
//   rcp_seed   : MUFU.RCP64H, at least 20 good bits
//   div_pos    : two Newton steps on the seed (error squared twice: below 2^-53 for any seed
//                better than 2^-14) and one correction of the quotient by its exact residual;
//                within 1 ulp of a / b.  1 MUFU + 7 FP64 instructions.
//   log_pos    : x = 2^e m, m in [sqrt(1/2), sqrt(2)); s = (m - 1) / (m + 1) by div_pos;
//                log m = 2 atanh s = 2 s + s^3 (2/3 + 2/5 s^2 + ... + 2/19 s^16), truncation
//                below 3e-17 relative for |s| <= 0.1716; + e ln2 in two parts.
//                Absolute error below 2.5e-16 max(1, |log x|) (tests/test_math_cpu.py).
//                1 MUFU + 24 FP64 instructions, exponent handling on the integer pipe.
APM_HD double rcp_seed(double b) {
#if defined(__CUDA_ARCH__)
	double y;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(b));
	return y;
#else
	return make_double(hi32(1.0 / b), 0); /* the host twin: same algorithm, its own 20-bit seed */
#endif
}
APM_HD double div_pos(double a, double b) {
	double y = rcp_seed(b);
	double e = fma(-b, y, 1.0);
	y = fma(y, e, y);
	e = fma(-b, y, 1.0);
	y = fma(y, e, y);
	double q = a * y;
	return fma(fma(-b, q, a), y, q);
}
// a / b, correctly rounded, for a divisor whose correctly rounded reciprocal y = rn(1 / b) is at hand:
// q0 = rn(a y), the residual a - b q0 is exact in one fma, q = rn(q0 + r y) is the IEEE quotient
// (Markstein's theorem; it needs b's mantissa not all ones -- true of the small integers this is
// used for).  3 FP64 instructions instead of the library division's ~45 with its slow-path branch,
// and the same bits: tests/test_math_cpu.py checks 1e7 quotients per divisor 1..9 against `/`.
APM_HD double div_by_known(double a, double b, double y) {
	const double q0 = a * y;
	const double r = fma(-b, q0, a);
	return fma(r, y, q0);
}
APM_HD double log_pos(double x) {
	const double LN2_HI = 0x1.62e42fefa39efp-1, LN2_LO = 0x1.abc9e3b39803fp-56;
	int hi = hi32(x);
	// mantissa above sqrt(2) (0x6a09e667f3bcd...): halve it, one more power of two
	int up = (hi & 0x000fffff) >= 0x0006a09f ? 1 : 0;
	double m = make_double((hi & 0x000fffff) | ((1023 - up) << 20), lo32(x));
	// the exponent as a double without a conversion instruction: the biased exponent in the low
	// word of 2^52, minus (2^52 + 1023)
	double ed = make_double(0x43300000, (hi >> 20) + up) - 4503599627371519.0;
	double s = div_pos(m - 1.0, m + 1.0);
	double z = s * s;
	double p = 2.0 / 19.0;
	p = fma(p, z, 2.0 / 17.0);
	p = fma(p, z, 2.0 / 15.0);
	p = fma(p, z, 2.0 / 13.0);
	p = fma(p, z, 2.0 / 11.0);
	p = fma(p, z, 2.0 / 9.0);
	p = fma(p, z, 2.0 / 7.0);
	p = fma(p, z, 2.0 / 5.0);
	p = fma(p, z, 2.0 / 3.0);
	double lm = fma(s * z, p, s + s);
	return fma(ed, LN2_HI, fma(ed, LN2_LO, lm));
}

// ---- Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011) -----------------------------
struct Philox4 {
	uint32_t w[4];
};

APM_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t) (((uint64_t) a * b) >> 32);
#endif
}

APM_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
		uint32_t k0, uint32_t k1) {
#pragma unroll
	for (int r = 0; r < 10; r++) {
		uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
		uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
		uint32_t n0 = h1 ^ c1 ^ k0;
		uint32_t n2 = h0 ^ c3 ^ k1;
		c0 = n0;
		c1 = l1;
		c2 = n2;
		c3 = l0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	Philox4 o;
	o.w[0] = c0;
	o.w[1] = c1;
	o.w[2] = c2;
	o.w[3] = c3;
	return o;
}

// Stream layout (shared, by construction, with the test oracle):
//   key     = (seed lo, seed hi)
//   counter = (id, step lo, step hi ^ (attempt >> 20) << 20, purpose << 28 | idx << 20 | attempt & 0xfffff)
//   id      = global chain id for per-chain draws, global ensemble id for swap draws
//   u0, u1  = two 53-bit uniforms strictly inside (0, 1)
enum {
	PURPOSE_JUMP = 0, PURPOSE_ACCEPT = 1, PURPOSE_SWAP_PICK = 2, PURPOSE_SWAP_TEST = 3,
	PURPOSE_HOST = 4 // apm_gpu_host_uniform: step = the chain's count of host draws
};

APM_HD void philox_uniforms(uint64_t seed, uint32_t id, uint64_t step, uint32_t purpose,
		uint32_t idx, uint32_t attempt, double & u0, double & u1) {
	// attempts beyond 2^20 (a proposal redrawn a million times: a step width far beyond the
	// parameter's range) spill into the top bits of the step's high word, so the stream of
	// redraws does not repeat before 2^32 attempts
	Philox4 o = philox4x32_10(id, (uint32_t) step, (uint32_t) (step >> 32) ^ ((attempt >> 20) << 20),
			(purpose << 28) | ((idx & 0xffu) << 20) | (attempt & 0xfffffu),
			(uint32_t) seed, (uint32_t) (seed >> 32));
	// integer + 0.5 is not representable above 2^52 and rounds to even: the largest value would
	// come out as exactly 1 (log(u / (1 - u)) = inf under PROPOSAL_LOGISTIC); clamp it below 1
	const double BELOW_ONE = 0x1.fffffffffffffp-1;
	u0 = fmin(((double) (o.w[0] >> 5) * 67108864.0 + (double) (o.w[1] >> 6) + 0.5)
			* (1.0 / 9007199254740992.0), BELOW_ONE);
	u1 = fmin(((double) (o.w[2] >> 5) * 67108864.0 + (double) (o.w[3] >> 6) + 0.5)
			* (1.0 / 9007199254740992.0), BELOW_ONE);
}

// the proposal jump: reference src/mcmc_gettersetter.c:290-306 (Gaussian default;
// PROPOSAL_LOGISTIC; PROPOSAL_UNIFORM = flat on (-sigma, sigma))
APM_HD double jump_from_uniforms(int proposal, double sigma, double u0, double u1) {
	if (proposal == 1)
		return sigma * log(u0 / (1 - u0));
	if (proposal == 2)
		return (-sigma) * (1 - u0) + sigma * u0;
	return sigma * (sqrt(-2.0 * log(u0)) * cos(2.0 * 3.14159265358979323846 * u1));
}

// The same jump in two halves, for draws made ahead of time: jump_unit() is everything that does
// not depend on the step width (it can be drawn steps ahead, by any thread: the draw depends on the
// chain's id and step counter only), jump_apply() the rest.  jump_apply(p, sigma, jump_unit(p, u0,
// u1)) performs the very operations of jump_from_uniforms(p, sigma, u0, u1) in the same order.
APM_HD double jump_unit(int proposal, double u0, double u1) {
	if (proposal == 1)
		return log(u0 / (1 - u0));
	if (proposal == 2)
		return u0;
	return sqrt(-2.0 * log(u0)) * cos(2.0 * 3.14159265358979323846 * u1);
}
// two at a time (the same operations on each: the same bits), written side by side so that the
// two chains of dependent instructions overlap
APM_HD void jump_unit2(int proposal, double u0a, double u1a, double u0b, double u1b, double & za, double & zb) {
	if (proposal == 1) {
		const double qa = u0a / (1 - u0a), qb = u0b / (1 - u0b);
		za = log(qa);
		zb = log(qb);
	} else if (proposal == 2) {
		za = u0a;
		zb = u0b;
	} else {
		const double la = log(u0a), lb = log(u0b);
		const double ca = cos(2.0 * 3.14159265358979323846 * u1a), cb = cos(2.0 * 3.14159265358979323846 * u1b);
		za = sqrt(-2.0 * la) * ca;
		zb = sqrt(-2.0 * lb) * cb;
	}
}
APM_HD double jump_apply(int proposal, double sigma, double z) {
	if (proposal == 2)
		return (-sigma) * (1 - z) + sigma * z;
	return sigma * z;
}

// reference src/mcmc_internal.h:46-48
APM_HD double mod_double(double x, double div) {
	return x < 0 ? x - div * (int) (x / div - 1) : x - div * (int) (x / div);
}

} // namespace apm
