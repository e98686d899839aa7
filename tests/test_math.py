"""include/slgpu_math.h on the host: within 1 ulp of libm (numpy), right special values, and the
Markstein division used by the step kernels equals IEEE division bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "libmath_shim.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"),
                           "-o", out, os.path.join(ROOT, "tests", "csrc", "math_shim.cpp")])
    return C.CDLL(out)


def _call(fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    o = np.empty_like(x)
    fn(x.ctypes.data_as(_dp), x.size, o.ctypes.data_as(_dp))
    return o


def _ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_exp_within_1ulp(shim):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-708, 709, 200000), rng.uniform(-30, 5, 200000), rng.normal(0, 1, 200000)])
    assert _ulps(_call(shim.v_exp, x), np.exp(x)).max() <= 1.0


def test_exp_special_values(shim):
    got = _call(shim.v_exp, [np.nan, np.inf, -np.inf, 710.0, -709.0, 0.0])
    assert np.isnan(got[0]) and got[1] == np.inf and got[2] == 0.0 and got[3] == np.inf and got[4] == 0.0 and got[5] == 1.0


def test_log_within_1ulp(shim):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(0, 1, 300000), np.exp(rng.uniform(-700, 700, 200000)), 1 + rng.normal(0, 1e-3, 100000),
                        (np.arange(1, 1000) + 0.5) * 2.0 ** -52])
    assert _ulps(_call(shim.v_log, x), np.log(x)).max() <= 1.0


def test_log_special_values(shim):
    got = _call(shim.v_log, [np.nan, np.inf, -1.0, 0.0, 5e-324, 1.0])
    assert np.isnan(got[0]) and got[1] == np.inf and np.isnan(got[2]) and got[3] == -np.inf and got[5] == 0.0
    assert got[4] == pytest.approx(np.log(5e-324), rel=1e-15)


def test_sincos2pi(shim):
    rng = np.random.default_rng(2)
    u = np.concatenate([rng.uniform(0, 1, 500000), np.arange(9) / 8.0])
    s, c = np.empty_like(u), np.empty_like(u)
    shim.v_sincos2pi(u.ctypes.data_as(_dp), u.size, s.ctypes.data_as(_dp), c.ctypes.data_as(_dp))
    th = np.longdouble(2) * np.longdouble("3.14159265358979323846264338327950288") * u.astype(np.longdouble)
    assert np.abs(s - np.sin(th).astype(np.float64)).max() <= 2.3e-16
    assert np.abs(c - np.cos(th).astype(np.float64)).max() <= 2.3e-16
    assert np.allclose(s * s + c * c, 1.0, atol=5e-16)


def test_division_by_shared_reciprocal_is_ieee_division(shim):
    rng = np.random.default_rng(3)
    n = rng.standard_normal(4_000_000) * np.exp(rng.uniform(-20, 20, 4_000_000))
    c = np.concatenate([rng.uniform(1, 2, 2_000_000) * np.exp2(rng.integers(-3, 4, 2_000_000)),
                        1.0 + np.arange(1_000_000) * 2.0 ** -52,              # divisors just above 1
                        2.0 - (1 + np.arange(1_000_000)) * 2.0 ** -52])       # significands of (almost) all ones
    o = np.empty_like(n)
    shim.v_div_by(n.ctypes.data_as(_dp), c.ctypes.data_as(_dp), n.size, o.ctypes.data_as(_dp))
    assert np.array_equal(o, n / c)
