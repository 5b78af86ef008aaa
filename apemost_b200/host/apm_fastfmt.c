/*
 * apm_fastfmt.c -- "%6e" without printf.
 *
 * prob-chain<k>.dump holds two "%6e" values per chain and iteration (reference
 * src/parallel_tempering.c:399-401).  At GPU step rates glibc's exact printf_fp -- arbitrary
 * precision arithmetic for every value -- is what `run` waits for, so the common case is done
 * here: seven significant digits from one 80-bit multiplication by a power of ten.  The result
 * must be the very bytes printf writes.  The scaled value carries a relative error below 2^-60,
 * i.e. an absolute error below 1e-11 on a number of magnitude 1e6..1e7; whenever its fractional
 * part is within 1e-6 of the rounding boundary .5 (and for zeros, subnormals, infinities, NaNs and
 * exponents of three digits) the function declines and the caller falls back to snprintf.
 * tests/test_host_cpu.py compares it with snprintf on tens of millions of values.
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#define P10_MIN (-340)
#define P10_MAX 340
static long double p10_table[P10_MAX - P10_MIN + 1];
static int p10_ready = 0;

static void p10_init(void) {
	int k;
	/* powl is not guaranteed correctly rounded; build the table by repeated exact-as-possible
	 * products from 10^0 outwards (each step adds at most half an ulp of the 64-bit mantissa) */
	p10_table[0 - P10_MIN] = 1.0L;
	for (k = 1; k <= P10_MAX; k++)
		p10_table[k - P10_MIN] = p10_table[k - 1 - P10_MIN] * 10.0L;
	for (k = -1; k >= P10_MIN; k--)
		p10_table[k - P10_MIN] = p10_table[k + 1 - P10_MIN] / 10.0L;
	p10_ready = 1;
}

/* writes the "%6e" text of v (no terminator) and returns its length, or 0 to decline */
int apm_format_e6(double v, char * out) {
	double a = fabs(v);
	int e10, tries, len = 0;
	long double m, fl, frac;
	long digits;
	if (!p10_ready)
		p10_init();
	if (!(a >= 1e-290 && a <= 1e290)) /* zero, subnormal neighbourhood, inf, nan, huge */
		return 0;
	e10 = (int) floor(ilogb(a) * 0.30102999566398120); /* within one of floor(log10(a)); fixed up below */
	for (tries = 0; tries < 3; tries++) {
		m = (long double) a * p10_table[6 - e10 - P10_MIN]; /* a * 10^(6 - e10), wanted in [1e6, 1e7) */
		if (m < 1e6L) {
			e10--;
			continue;
		}
		if (m >= 1e7L) {
			e10++;
			continue;
		}
		break;
	}
	if (tries == 3)
		return 0;
	fl = floorl(m);
	frac = m - fl;
	if (frac > 0.5L - 1e-6L && frac < 0.5L + 1e-6L)
		return 0; /* too close to a tie for this arithmetic: let printf decide */
	digits = (long) fl + (frac > 0.5L ? 1 : 0);
	if (digits >= 10000000L) { /* 9.9999995.. rounds up into the next decade */
		digits = 1000000L;
		e10++;
	}
	if (e10 > 99 || e10 < -99)
		return 0;
	if (v < 0 || (v == 0 && signbit(v)))
		out[len++] = '-';
	{
		char d[8];
		int i;
		for (i = 6; i >= 0; i--) {
			d[i] = (char) ('0' + digits % 10);
			digits /= 10;
		}
		out[len++] = d[0];
		out[len++] = '.';
		memcpy(out + len, d + 1, 6);
		len += 6;
	}
	out[len++] = 'e';
	out[len++] = e10 < 0 ? '-' : '+';
	if (e10 < 0)
		e10 = -e10;
	out[len++] = (char) ('0' + e10 / 10);
	out[len++] = (char) ('0' + e10 % 10);
	return len;
}

/* "%.15e" (DUMP_FORMAT: the parameter dumps, sixteen significant digits): the same idea with much
 * less room.  The scaled value a * 10^(15 - e10) has up to 54 bits before the point; with an exact
 * power of ten (10^k is exact in the 64-bit mantissa for 0 <= k <= 27, and 10^-k is then one
 * correctly rounded division) it carries an absolute error below 0.002, so everything whose
 * fractional part is not within 0.01 of .5 can be decided; the rest, and decimal exponents
 * outside [-12, 42], goes to snprintf.  Returns the length written, or 0 to decline. */
int apm_format_e15(double v, char * out) {
	static long double pos[28], neg[28];
	static int ready = 0;
	double a = fabs(v);
	int e10, tries, len = 0, k, i;
	long double m, fl, frac;
	unsigned long long digits;
	char d[17];
	if (!ready) {
		pos[0] = 1.0L;
		for (k = 1; k < 28; k++)
			pos[k] = pos[k - 1] * 10.0L; /* exact */
		for (k = 0; k < 28; k++)
			neg[k] = 1.0L / pos[k];      /* correctly rounded */
		ready = 1;
	}
	if (!(a >= 1e-12 && a < 1e42))
		return 0;
	e10 = (int) floor(ilogb(a) * 0.30102999566398120);
	for (tries = 0; tries < 3; tries++) {
		k = 15 - e10;
		if (k < -27 || k > 27)
			return 0;
		m = k >= 0 ? (long double) a * pos[k] : (long double) a * neg[-k];
		if (m < 1e15L) {
			e10--;
			continue;
		}
		if (m >= 1e16L) {
			e10++;
			continue;
		}
		break;
	}
	if (tries == 3)
		return 0;
	fl = floorl(m);
	frac = m - fl;
	if (frac > 0.49L && frac < 0.51L)
		return 0;
	digits = (unsigned long long) fl + (frac > 0.5L ? 1 : 0);
	if (digits >= 10000000000000000ULL) {
		digits = 1000000000000000ULL;
		e10++;
	}
	if (v < 0)
		out[len++] = '-';
	for (i = 15; i >= 0; i--) {
		d[i] = (char) ('0' + digits % 10);
		digits /= 10;
	}
	out[len++] = d[0];
	out[len++] = '.';
	memcpy(out + len, d + 1, 15);
	len += 15;
	out[len++] = 'e';
	out[len++] = e10 < 0 ? '-' : '+';
	if (e10 < 0)
		e10 = -e10;
	out[len++] = (char) ('0' + e10 / 10);
	out[len++] = (char) ('0' + e10 % 10);
	return len;
}

/* "%6e\t%6e\n" into buf (at least 64 bytes); returns the length */
int apm_format_prob_line(double prob, double dl, char * buf) {
	int n = apm_format_e6(prob, buf), m;
	if (n == 0)
		n = snprintf(buf, 32, "%6e", prob);
	buf[n++] = '\t';
	m = apm_format_e6(dl, buf + n);
	if (m == 0)
		m = snprintf(buf + n, 32, "%6e", dl);
	n += m;
	buf[n++] = '\n';
	return n;
}
