"""CPU: the decomposition of numpy-legacy `choice` / scipy `truncnorm.rvs` into primitive
draws (oracle/rng_tape.py, bnpc_b200/rng.py) consumes the MT19937 stream exactly like the
real functions, and the product-side tape reader agrees with the oracle-side one."""
import numpy as np
from scipy.stats import truncnorm

from bnpc_b200 import rng as prng
from oracle.rng_tape import LegacyRandom, Tape, TapeSource


def _after():
    return np.random.random(3)


def test_choice_matches_numpy_stream():
    cases = [
        dict(a=np.array([0.1, 0.25, 0.5]), size=17),
        dict(a=np.array([0.1, 0.25, 0.5]), size=(2, 9)),
        dict(a=np.array([0.1, 0.25, 0.5])),
        dict(a=7),
        dict(a=9, size=2, replace=False),
        dict(a=[0, 1], p=[0.75, 0.25]),
        dict(a=np.array([4, 8, 15, 16, 23]), p=np.array([.1, .2, .3, .25, .15])),
        dict(a=np.array([4, 8, 15, 16, 23]), p=np.array([.1, .2, .3, .25, .15]), size=2, replace=False),
        dict(a=np.array([4, 8]), p=np.array([.999, .001]), size=2, replace=False),
    ]
    for i, kw in enumerate(cases):
        for seed in range(5):
            np.random.seed(seed)
            want = np.random.choice(**kw)
            tail_want = _after()
            np.random.seed(seed)
            got = LegacyRandom().choice(**kw)
            tail_got = _after()
            np.testing.assert_array_equal(np.asarray(got), np.asarray(want), err_msg=f'case {i}')
            np.testing.assert_array_equal(tail_got, tail_want, err_msg=f'case {i}: stream position')


def test_truncnorm_rvs_matches_scipy_stream():
    old = np.float32([0.3, 1e-5, 0.99999, 0.5])
    sd = np.array([0.1, 0.25, 0.5, 0.1])
    a, b = (1e-5 - old) / sd, (1 - 1e-5 - old) / sd
    np.random.seed(3)
    want = truncnorm.rvs(a, b, loc=old, scale=sd, size=4)
    tail_want = _after()
    np.random.seed(3)
    got = LegacyRandom().truncnorm_rvs(a, b, loc=old, scale=sd, size=4)
    np.testing.assert_array_equal(got, want)
    np.testing.assert_array_equal(_after(), tail_want)
    np.random.seed(4)
    want = truncnorm.rvs(-2.0, 30.0, loc=0.2, scale=0.05)
    np.random.seed(4)
    got = LegacyRandom().truncnorm_rvs(-2.0, 30.0, loc=0.2, scale=0.05)
    assert got == want


def test_tape_roundtrip_and_product_reader():
    t = Tape()
    rec = LegacyRandom(record=t)
    np.random.seed(0)
    p = np.array([.2, .5, .3])
    first = [rec.choice(3, p=p), rec.choice(np.arange(3), p=p, size=2, replace=False),
             rec.choice(11, size=2, replace=False), rec.random(), rec.randint(0, 3, size=5),
             rec.beta(1.5, 2.0), rec.gamma(3.0, 0.5), rec.choice(5)]
    kinds, sizes, values = t.to_arrays()
    # oracle-side replay
    t2 = Tape.from_arrays(kinds, sizes, values)
    rep = LegacyRandom(source=TapeSource(t2))
    second = [rep.choice(3, p=p), rep.choice(np.arange(3), p=p, size=2, replace=False),
              rep.choice(11, size=2, replace=False), rep.random(), rep.randint(0, 3, size=5),
              rep.beta(1.5, 2.0), rep.gamma(3.0, 0.5), rep.choice(5)]
    for x, y in zip(first, second):
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    assert t2.exhausted()
    # product-side replay (host decisions only; no device needed)
    pr = prng.TapeRandom(prng.Tape(kinds, sizes, values))
    assert pr.pick_weighted(p) == first[0]
    assert list(pr.pick_two_weighted(p)) == list(first[1])
    assert list(pr.first_two_of_permutation(11)) == list(first[2])
    assert pr.random() == first[3]
    np.testing.assert_array_equal(pr.tape.take('int', 5), first[4])
    assert pr.beta(1.5, 2.0) == first[5]
    assert pr.gamma(3.0, 0.5) == first[6]
    assert pr.randint(5) == first[7]
    assert pr.tape.exhausted()
