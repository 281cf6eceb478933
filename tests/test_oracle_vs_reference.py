"""CPU, only where the reference checkout is mounted (/root/reference): the unmodified
reference and the oracle restatement, driven from the same numpy seed, produce identical
traces; and the recording RNG proxy consumes numpy's global stream exactly like the real
numpy/scipy calls."""
import numpy as np
import pytest

from oracle import ref_shim, rng_tape
from oracle.crp_oracle import (DEFAULT_MOVES, OracleCRP, OracleCRPLearnErrors, do_step, simulate,
                               snapshot)

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(),
                                reason='reference checkout not mounted')

LEARN_KW = dict(DP_alpha=[-1, -1], FP_mean=0.01, FP_sd=0.01, FN_mean=0.2, FN_sd=0.1)
FIXED_KW = dict(DP_alpha=[-1, -1], FN_error=0.2, FP_error=0.01)


def _run(kind, data, learning, pp, seed, steps, moves):
    np.random.seed(seed)
    if kind == 'oracle':
        rnd = rng_tape.LegacyRandom()
        cls = OracleCRPLearnErrors if learning else OracleCRP
        m = cls(data.copy(), param_beta=list(pp), rnd=rnd, **(LEARN_KW if learning else FIXED_KW))
        m.init()
        out = []
        for _ in range(steps):
            do_step(m, rnd, moves, learning)
            out.append(snapshot(m))
        return out
    ref = ref_shim.load_reference()
    if learning:
        m = ref.CRP_learning_errors.CRP_errors_learning(data.copy(), param_beta=list(pp), **LEARN_KW)
    else:
        m = ref.CRP.CRP(data.copy(), param_beta=list(pp), **FIXED_KW)
    out = []
    with ref_shim.ref_errstate():
        if kind == 'ref_patched':
            rnd = rng_tape.LegacyRandom()
            with ref_shim.patched_random(ref, rnd):
                m.init()
                for _ in range(steps):
                    do_step(m, rnd, moves, learning)
                    out.append(snapshot(m))
        else:
            m.init()
            for _ in range(steps):
                do_step(m, np.random, moves, learning)
                out.append(snapshot(m))
    return out


def _same(a, b):
    for i, (x, y) in enumerate(zip(a, b)):
        for k in x:
            assert np.array_equal(np.asarray(x[k]), np.asarray(y[k])), f'step {i + 1}: {k}'


@pytest.mark.parametrize('learning', [False, True])
@pytest.mark.parametrize('pp', [(0.25, 0.25), (1, 1)])
def test_same_seed_same_trace(learning, pp):
    data, _ = simulate(50, 30, k_true=4, seed=7)
    moves = dict(DEFAULT_MOVES, sm_prob=0.5)
    ref = _run('ref', data, learning, pp, 21, 25, moves)
    _same(ref, _run('ref_patched', data, learning, pp, 21, 25, moves))
    _same(ref, _run('oracle', data, learning, pp, 21, 25, moves))


def test_reference_chain_driver_order():
    """The step schedule used by the tests (oracle.crp_oracle.do_step) is the reference's
    Chain.do_step (libs/MCMC.py:320-342): run the reference's own Chain_steps and compare."""
    ref = ref_shim.load_reference(with_mcmc=True)
    data, _ = simulate(40, 20, k_true=3, seed=9)
    params = dict(sm_prob=0.33, dpa_prob=0.25, error_prob=0.25,
                  param_proposal_sd=np.array([0.1, 0.25, 0.5]), sm_ratios=[0.75, 0.25], sm_steps=3)
    with ref_shim.ref_errstate():
        np.random.seed(5)
        m1 = ref.CRP_learning_errors.CRP_errors_learning(data.copy(), param_beta=[0.25, 0.25], **LEARN_KW)
        m1.init()
        chain = ref.MCMC.Chain_steps(m1, 1, 20, 5, params, 0, False)
        chain.run()
        np.random.seed(5)
        m2 = ref.CRP_learning_errors.CRP_errors_learning(data.copy(), param_beta=[0.25, 0.25], **LEARN_KW)
        m2.init()
        ml = []
        for _ in range(20):
            do_step(m2, np.random, params, True)
            ml.append(m2.get_ll_full())
    np.testing.assert_array_equal(chain.results['ML'][1:], np.array(ml))
    np.testing.assert_array_equal(chain.results['assignments'][-1], m2.assignment)
