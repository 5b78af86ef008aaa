"""apm_math.cuh on the CPU: the header is __host__ __device__, and IEEE fma is deterministic, so
compiling it with g++ -mfma reproduces the device arithmetic of sin_fast / Philox / the proposal
transforms bit for bit; they are checked here against extended-precision references."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_binding import philox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "oracle", "_build")
SRC = r'''
#include "apm_math.cuh"
extern "C" {
void t_sin_fast(const double * x, double * y, long n) { for (long i = 0; i < n; i++) y[i] = apm::sin_fast(x[i]); }
void t_sin_full(const double * x, double * y, long n) { for (long i = 0; i < n; i++) y[i] = apm::sin_full(x[i]); }
void t_log_pos(const double * x, double * y, long n) { for (long i = 0; i < n; i++) y[i] = apm::log_pos(x[i]); }
void t_div_pos(const double * a, const double * b, double * y, long n) { for (long i = 0; i < n; i++) y[i] = apm::div_pos(a[i], b[i]); }
double t_mod_double(double x, double d) { return apm::mod_double(x, d); }
long t_div_by_known_mismatches(const double * a, long n, int b) {
	long bad = 0;
	const double d = (double) b, y = 1.0 / d;
	for (long i = 0; i < n; i++)
		bad += apm::div_by_known(a[i], d, y) != a[i] / d;
	return bad;
}
void t_philox(const unsigned * c, const unsigned * k, unsigned * o) {
	apm::Philox4 r = apm::philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
	for (int i = 0; i < 4; i++) o[i] = r.w[i];
}
void t_uniforms(unsigned long long seed, unsigned id, unsigned long long step, unsigned purpose, unsigned idx,
		unsigned attempt, double * u) { apm::philox_uniforms(seed, id, step, purpose, idx, attempt, u[0], u[1]); }
double t_jump(int proposal, double sigma, double u0, double u1) { return apm::jump_from_uniforms(proposal, sigma, u0, u1); }
}
'''


@pytest.fixture(scope="module")
def lib():
    os.makedirs(BUILD, exist_ok=True)
    src, so = os.path.join(BUILD, "math_test.cpp"), os.path.join(BUILD, "libmath_test.so")
    open(src, "w").write(SRC)
    subprocess.run(["g++", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-I",
                    os.path.join(ROOT, "apemost_b200", "csrc"), src, "-o", so], check=True)
    L = C.CDLL(so)
    L.t_mod_double.restype = C.c_double
    L.t_mod_double.argtypes = [C.c_double, C.c_double]
    L.t_jump.restype = C.c_double
    L.t_jump.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double]
    return L


def _apply(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    fn(x.ctypes.data_as(C.POINTER(C.c_double)), y.ctypes.data_as(C.POINTER(C.c_double)), C.c_long(x.size))
    return y


@pytest.mark.parametrize("scale", [1.0, 100.0, 6.3e4, 1e7, 1e9])
def test_sin_fast_absolute_error(lib, scale):
    """max |sin_fast(x) - sin(x)| over the fast range: a few 1e-16 (apm_math.cuh header claims 2.7e-16)"""
    rng = np.random.default_rng(int(scale) % 9973)
    x = rng.uniform(-scale, scale, 400_000)
    ref = np.sin(x.astype(np.longdouble))          # 64-bit mantissa, accurate argument reduction
    err = np.abs(_apply(lib.t_sin_fast, x).astype(np.longdouble) - ref).astype(np.float64)
    assert err.max() < 3.5e-16, err.max()
    assert abs(np.mean((_apply(lib.t_sin_fast, x).astype(np.longdouble) - ref).astype(np.float64))) < 2e-18


def test_log_pos_and_div_pos(lib):
    """the branch-free logarithm and quotient of pulse / pulse_vrot's row term on their domain
    (positive normal arguments, 1e-300 .. 1e300): log within 2.5e-16 max(1, |log x|) of an
    extended-precision logarithm, the quotient within 1 ulp of a / b"""
    rng = np.random.default_rng(31)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 300_000), rng.uniform(0.5, 2.0, 300_000),
                        1 + rng.normal(0, 1e-9, 1000), [1.0, 2.0, 0.5, np.sqrt(2), 1 / np.sqrt(2), 1e-300, 1e300]])
    ref = np.log(x.astype(np.longdouble))
    err = np.abs(_apply(lib.t_log_pos, x).astype(np.longdouble) - ref).astype(np.float64)
    assert (err <= 2.5e-16 * np.maximum(1.0, np.abs(ref.astype(np.float64)))).all(), err.max()
    assert _apply(lib.t_log_pos, np.array([1.0]))[0] == 0.0
    a, b = 10.0 ** rng.uniform(-140, 140, 300_000), 10.0 ** rng.uniform(-140, 140, 300_000)
    a[:1000], b[:1000] = rng.uniform(1, 2, 1000), rng.uniform(1, 2, 1000)
    q = np.empty_like(a)
    lib.t_div_pos(a.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double)),
                  q.ctypes.data_as(C.POINTER(C.c_double)), C.c_long(a.size))
    assert (np.abs(q - a / b) <= np.spacing(a / b)).all()
    assert (q == a / b).mean() > 0.99


def test_sin_fast_special_points(lib):
    x = np.array([0.0, -0.0, np.pi / 2, -np.pi / 2, np.pi, 1e-300, 2.0 ** 29])
    y = _apply(lib.t_sin_fast, x)
    assert y[0] == 0.0 and y[1] == 0.0
    np.testing.assert_allclose(y[2:4], [1.0, -1.0], rtol=0, atol=2e-16)
    assert abs(y[4] - 1.2246467991473532e-16) < 1e-31 and y[5] == 1e-300
    assert abs(y[6] - float(np.sin(np.longdouble(2.0 ** 29)))) < 3e-16


def test_sin_full_outside_the_fast_range(lib):
    x = np.array([2.0 ** 30, 1e15, -3e18, 1e300])
    np.testing.assert_allclose(_apply(lib.t_sin_full, x), np.sin(x), rtol=0, atol=2e-16)
    assert np.isnan(_apply(lib.t_sin_full, np.array([np.inf, np.nan]))).all()


def test_philox_matches_the_oracle_and_the_published_vectors(lib):
    for ctr, key in [((0, 0, 0, 0), (0, 0)), ((0xffffffff,) * 4, (0xffffffff,) * 2),
                     ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))]:
        c, k, o = (C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), (C.c_uint * 4)()
        lib.t_philox(c, k, o)
        assert list(o) == list(philox(ctr, key))
    # Random123 known-answer test, philox4x32-10
    c, k, o = (C.c_uint * 4)(0, 0, 0, 0), (C.c_uint * 2)(0, 0), (C.c_uint * 4)()
    lib.t_philox(c, k, o)
    assert list(o) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


def test_uniforms_are_strictly_inside_the_unit_interval(lib):
    u = (C.c_double * 2)()
    lo, hi = 1.0, 0.0
    for step in range(2000):
        lib.t_uniforms(C.c_ulonglong(12345), 7, C.c_ulonglong(step), 1, 3, 0, u)
        lo, hi = min(lo, u[0], u[1]), max(hi, u[0], u[1])
    assert 0.0 < lo < 0.01 and 0.99 < hi < 1.0


def test_proposal_transforms(lib):
    """reference src/mcmc_gettersetter.c:290-306: gaussian (Box-Muller), logistic, flat(-s, s)"""
    assert lib.t_jump(2, 0.5, 0.0, 0.0) == -0.5 and lib.t_jump(2, 0.5, 0.75, 0.0) == 0.25
    assert abs(lib.t_jump(1, 2.0, 0.75, 0.0) - 2.0 * np.log(3.0)) < 1e-15
    assert abs(lib.t_jump(0, 3.0, np.exp(-0.5), 0.0) - 3.0) < 1e-14  # sqrt(-2 ln u0) = 1, cos(0) = 1
    rng = np.random.default_rng(0)
    z = np.array([lib.t_jump(0, 1.0, a, b) for a, b in rng.uniform(1e-12, 1, (40000, 2))])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02


def test_mod_double_reference_values(lib):
    """the six values of the reference's tests/tests.c:152-160"""
    for x, d, want in [(5.0, 3.0, 2.0), (-1.0, 3.0, 2.0), (3.5, 1.0, 0.5), (-0.25, 1.0, 0.75)]:
        assert abs(lib.t_mod_double(x, d) - want) < 1e-12


def test_div_by_known_is_the_ieee_quotient(lib):
    """div_by_known (three fp64 instructions, used for apps/normal.c's divisions by sigma = 1 .. 9) gives
    the bits of `/`: wide-range values, differences x - pos, products -height * d, zero"""
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.normal(0, 1, 2_000_000) * 10.0 ** rng.uniform(-12, 12, 2_000_000),
                        rng.uniform(-1e4, 1e4, 4_000_000), -10.0 * rng.uniform(0, 1e4, 4_000_000), [0.0, -0.0]])
    a = np.ascontiguousarray(a)
    lib.t_div_by_known_mismatches.restype = C.c_long
    for b in range(1, 10):
        assert lib.t_div_by_known_mismatches(a.ctypes.data_as(C.POINTER(C.c_double)), C.c_long(a.size), C.c_int(b)) == 0

