"""const_math.cuh (exp / log / pow with constant-bank coefficients, TRIBS_CONST_MATH) compiled for
the host. The functions use IEEE operations only (explicit fma, one division), so what is measured
here is what the device computes."""
import ctypes as C
import os

import numpy as np
import pytest

PD = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def lib(built):
    return C.CDLL(os.path.join(os.path.dirname(__file__), "hostcheck", "libhostcheck.so"))


def _run(fn, *xs):
    xs = [np.ascontiguousarray(x, np.float64) for x in xs]
    y = np.empty_like(xs[0])
    fn(*[x.ctypes.data_as(PD) for x in xs], y.ctypes.data_as(PD), len(y))
    return y


def _ulps(y, ref):
    return np.abs(y - ref) / np.spacing(np.abs(ref))


def test_exp_within_one_ulp_of_glibc(lib):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-700, 700, 400_000), rng.uniform(-1, 1, 200_000), rng.normal(0, 1e-5, 50_000)])
    assert _ulps(_run(lib.hc_cm_exp, x), np.exp(x)).max() <= 1.0


def test_log_within_one_ulp_of_glibc(lib):
    rng = np.random.default_rng(2)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 400_000)), rng.uniform(0.5, 2, 400_000), 1 + rng.normal(0, 1e-6, 50_000)])
    assert _ulps(_run(lib.hc_cm_log, x), np.log(x)).max() <= 1.0


def test_special_arguments_take_the_library_path(lib):
    sp = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310, -np.inf, 701.0, -701.0, 710.0, -750.0])
    with np.errstate(all="ignore"):
        assert np.array_equal(_run(lib.hc_cm_exp, sp), np.exp(sp), equal_nan=True)
        assert np.array_equal(_run(lib.hc_cm_log, sp), np.log(sp), equal_nan=True)
        base = np.array([0.0, 0.0, -2.0, -2.0, 2.0, np.inf])
        ex = np.array([0.0, 2.0, 2.0, 0.5, -1075.0, -1.0])
        assert np.array_equal(_run(lib.hc_cm_pow, base, ex), np.power(base, ex), equal_nan=True)


def test_pow_in_the_range_the_soil_physics_uses(lib):
    """Brooks-Corey exponents: bases in (0, 1e4], |exponent| up to ~20; error ~ |y ln x| ulps."""
    rng = np.random.default_rng(3)
    x = np.exp(rng.uniform(-9, 9, 300_000))
    e = rng.uniform(-20, 20, 300_000)
    y, ref = _run(lib.hc_cm_pow, x, e), np.power(x, e)
    rel = np.abs(y - ref) / np.abs(ref)
    assert rel.max() < 5e-14
