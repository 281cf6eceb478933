"""CPU: the host-side scalar helpers of the native group driver (bnpc_host_*: counter-based uniforms,
Marsaglia-Tsang Gamma / Beta variates, scipy's truncated-normal ppf / logpdf restated with libm)
against scipy -- the draws of update_DP_alpha (libs/CRP.py:386-410) and of the error-rate moves
(libs/CRP_learning_errors.py:66-111) in production mode.  No GPU needed: these are host functions."""
import ctypes as C

import numpy as np
from scipy.stats import beta, gamma, kstest, truncnorm

from bnpc_b200 import _lib


def test_truncnorm_ppf_and_logpdf_match_scipy():
    L = _lib.lib()
    rng = np.random.default_rng(1)
    for _ in range(3000):
        cur = rng.uniform(1e-4, 0.9)
        sd = rng.choice([0.005, 0.01, 0.015, 0.05, 0.1, 0.15])
        u = rng.uniform()
        lo, hi = (0 - cur) / sd, (1 - cur) / sd
        got, want = L.host_truncnorm_ppf(u, lo, hi), truncnorm.ppf(u, lo, hi)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want))
        x = got * sd + cur
        np.testing.assert_allclose(L.host_truncnorm_logpdf(x, lo, hi, cur, sd), truncnorm.logpdf(x, lo, hi, cur, sd),
                                   rtol=1e-13, atol=1e-13)
    assert L.host_truncnorm_logpdf(2.0, -1.0, 1.0, 0.0, 1.0) == -np.inf


def test_host_variates_follow_their_distributions():
    L = _lib.lib()
    ctr = C.c_uint64(0)
    u = np.array([L.host_random(42, C.byref(ctr)) for _ in range(20000)])
    assert ctr.value == 20000 and 0 <= u.min() and u.max() < 1
    assert kstest(u, 'uniform').pvalue > 1e-3
    for shape in (0.3, 1.0, 2.5, 340.0):
        g = np.array([L.host_gamma(43, C.byref(ctr), shape) for _ in range(20000)])
        assert kstest(g, gamma(shape).cdf).pvalue > 1e-3, shape
    for a, b in ((0.25, 0.25), (1.0, 1.0), (317.2, 100000.0)):
        x = np.array([L.host_beta(44, C.byref(ctr), a, b) for _ in range(20000)])
        assert kstest(x, beta(a, b).cdf).pvalue > 1e-3, (a, b)


def test_host_stream_is_a_function_of_seed_and_counter():
    L = _lib.lib()
    a, b = C.c_uint64(5), C.c_uint64(5)
    assert L.host_random(7, C.byref(a)) == L.host_random(7, C.byref(b))
    c = C.c_uint64(5)
    assert L.host_random(8, C.byref(c)) != L.host_random(7, C.byref(C.c_uint64(5)))
