"""GPU: the native lockstep driver (bnpc_group_run, bnpc_b200/group.py) against the per-method
Python mirror of the model (bnpc_b200/engine.py, the path pinned to the oracle by the tape tests).

Both hosts draw from the same counter-based streams, so a chain must walk the SAME trajectory
under either: assignments, cluster lists and float32 theta identical after every step; the
float64 traces within 1e-10 relative (the two hosts evaluate log / truncnorm scalars with
different math libraries, 1 ulp apart).  A chain's trace must not depend on which other chains
share its launches: a group of n chains equals n groups of one, bit for bit.
"""
import copy

import numpy as np
import pytest

from oracle.crp_oracle import DEFAULT_MOVES, simulate

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')

LEARN = dict(DP_alpha=[-1, -1], FP_mean=0.01, FP_sd=0.01, FN_mean=0.2, FN_sd=0.1)
FIXED = dict(DP_alpha=[-1, -1], FN_error=0.2, FP_error=0.01)


def _moves(**kw):
    m = dict(DEFAULT_MOVES, **kw)
    m.setdefault('param_proposal_sd', np.array([0.1, 0.25, 0.5]))
    return m


def _chains(data, learning, pp, moves, steps, seeds, assign, burn_in=0, cls_attrs=None):
    from bnpc_b200.rng import PhiloxRandom
    from libs.MCMC import Chain_steps
    import libs.CRP as crp
    import libs.CRP_learning_errors as crple
    kw = dict(LEARN if learning else FIXED, param_beta=list(pp))
    proto = (crple.CRP_errors_learning if learning else crp.CRP)(data, **kw)
    out = []
    for i, seed in enumerate(seeds):
        m = copy.deepcopy(proto)
        m.device = 'cuda:0'
        m.rnd = PhiloxRandom(seed)
        for k, v in (cls_attrs or {}).items():
            setattr(m, k, v)
        m.init(assign=assign)
        out.append(Chain_steps(m, i + 1, steps, burn_in, moves, 0, False))
    return out


def _same(a, b, where, exact):
    ra, rb = a.results, b.results
    np.testing.assert_array_equal(ra['assignments'], rb['assignments'], err_msg=f'{where}: assignments')
    np.testing.assert_array_equal(ra['params'], rb['params'], err_msg=f'{where}: theta trace')
    assert list(a.model.cells_per_cluster.items()) == list(b.model.cells_per_cluster.items()), where
    for key in ('ML', 'MAP', 'DP_alpha', 'FN', 'FP'):
        if exact:
            np.testing.assert_array_equal(ra[key], rb[key], err_msg=f'{where}: {key}')
        else:
            np.testing.assert_allclose(ra[key], rb[key], rtol=1e-10, atol=0, err_msg=f'{where}: {key}')
    assert a.model.rnd.calls == b.model.rnd.calls and a.model.rnd.host_ctr.value == b.model.rnd.host_ctr.value, where


CASES = [
    # name, N, M, k_true, miss, learning, pp, init, steps, burn_in, moves, model attributes
    ('learn_default', 3000, 200, 6, 0.10, True, [0.25, 0.25], 'assign', 40, 0, _moves(), None),
    ('learn_burnin_smheavy', 1500, 96, 5, 0.10, True, [1, 1], 'assign', 40, 15, _moves(sm_prob=0.6, sm_steps=3), None),
    ('fixed_random_init', 700, 64, 4, 0.10, False, [0.25, 0.25], 'random', 25, 0, _moves(sm_prob=0.3), None),
    ('panel_wide', 5000, 50, 5, 0.30, True, [1, 1], 'assign', 20, 0, _moves(), dict(force_wide=True)),
    ('dense_rows', 1200, 150, 5, 0.10, True, [0.25, 0.25], 'assign', 15, 0, _moves(), dict(lean_enabled=False)),
]


@pytest.mark.parametrize('case', CASES, ids=[c[0] for c in CASES])
def test_native_group_equals_python_mirror(case):
    name, N, M, k_true, miss, learning, pp, init, steps, burn_in, moves, attrs = case
    from libs.MCMC import run_chains
    data, z = simulate(N, M, k_true=k_true, miss=miss, seed=11)
    assign = [int(v) for v in z] if init == 'assign' else None
    seeds = [101, 202, 303]
    py = _chains(data, learning, pp, moves, steps, seeds, assign, burn_in, attrs)
    for ch in py:
        ch.run_python()
    nat = _chains(data, learning, pp, moves, steps, seeds, assign, burn_in, attrs)
    run_chains(nat)
    for a, b in zip(nat, py):
        _same(a, b, f'{name} chain {a.no}', exact=False)
        assert a.results['burn_in'] == b.results['burn_in']
        np.testing.assert_array_equal(a.MH_counter, b.MH_counter)


def test_trace_is_independent_of_the_group():
    """a chain's trace depends on its seed only -- n chains in one group (batched
    launches) equal n groups of one chain, bit for bit"""
    from libs.MCMC import run_chains
    data, z = simulate(4000, 256, k_true=8, miss=0.1, seed=5)
    assign = [int(v) for v in z]
    moves = _moves(sm_prob=0.4)
    seeds = [7, 8, 9, 10, 11, 12, 13, 14, 15, 16]           # more than one batch of 8
    together = _chains(data, True, [0.25, 0.25], moves, 30, seeds, assign)
    run_chains(together)
    for i, seed in enumerate(seeds[:4]):
        alone = _chains(data, True, [0.25, 0.25], moves, 30, [seed], assign)
        run_chains(alone)
        _same(together[i], alone[0], f'seed {seed}', exact=True)


def test_group_extends_and_resumes():
    """lugsail-style extension (libs/MCMC.py:175-181): 20 steps, then 15 more, equal 35 at once"""
    from libs.MCMC import run_chains
    data, z = simulate(1500, 128, k_true=5, miss=0.1, seed=3)
    assign = [int(v) for v in z]
    moves = _moves()
    a = _chains(data, True, [0.25, 0.25], moves, 35, [41, 42], assign)
    run_chains(a)
    b = _chains(data, True, [0.25, 0.25], moves, 20, [41, 42], assign)
    run_chains(b)
    olds = [c.get_steps() for c in b]
    for c in b:
        c._extend_results(15, False)
        c.set_steps(15)
    run_chains(b, init_steps=olds[0] - 1)
    for x, y in zip(a, b):
        _same(x, y, f'chain {x.no}', exact=True)


def test_recorded_launches_merge_across_chains(monkeypatch):
    """the launches of a group step grow far slower than the number of chains: a step of 8 chains
    in one phase-synchronous group (BNPC_LOCKSTEP) costs one Gibbs sequence + one split-merge
    sequence + one parameter sequence, whoever drew what; the default asynchronous scheduler merges
    the chains that become ready together and still stays well below the launches of 8 single chains"""
    from bnpc_b200 import _lib
    import libs.MCMC as mcmc
    from libs.MCMC import run_chains
    monkeypatch.setattr(mcmc, 'GROUP_SIZE', 8)
    monkeypatch.setenv('BNPC_LOCKSTEP', '1')
    data, z = simulate(3000, 200, k_true=6, miss=0.1, seed=2)
    assign = [int(v) for v in z]
    moves = _moves()
    counts = {}
    for n in (1, 8):
        chains = _chains(data, True, [0.25, 0.25], moves, 30, list(range(50, 50 + n)), assign)
        before = _lib.launch_count()
        run_chains(chains)
        counts[n] = (_lib.launch_count() - before) / 30
    assert counts[8] < 4 * counts[1], counts            # 8 chains for far less than 8x the launches
    assert counts[8] / 8 < 25, counts                   # launches per chain-step
    monkeypatch.delenv('BNPC_LOCKSTEP')
    chains = _chains(data, True, [0.25, 0.25], moves, 30, list(range(50, 58)), assign)
    before = _lib.launch_count()
    run_chains(chains)
    assert (_lib.launch_count() - before) / 30 < 0.85 * 8 * counts[1], counts


def test_parallel_sweep_equals_single_sequencer():
    """The sweep walks the uncertain records with one warp per group of option-graph components;
    the single-sequencer route (`serial_sweep`) is its sequential definition.  Same seeds -> same
    traces, bit for bit (regression: a warp that saw a record posted for the exact path stopped
    walking although more of its own records lay in front of the posted one -- seed 15, step 10)."""
    from libs.MCMC import run_chains
    data, z = simulate(4000, 256, k_true=8, miss=0.1, seed=5)
    assign = [int(v) for v in z]
    moves = _moves(sm_prob=0.4)
    seeds = list(range(7, 17))
    par = _chains(data, True, [0.25, 0.25], moves, 30, seeds, assign)
    run_chains(par)
    ser = _chains(data, True, [0.25, 0.25], moves, 30, seeds, assign, 0, dict(serial_sweep=True))
    run_chains(ser)
    for a, b in zip(par, ser):
        _same(a, b, f'seed {seeds[a.no - 1]}', exact=True)


def test_lockstep_driver_equals_asynchronous_scheduler(monkeypatch):
    """the phase-synchronous driver (BNPC_LOCKSTEP=1) and the default asynchronous scheduler issue
    the same per-chain work in different interleavings: identical traces"""
    from libs.MCMC import run_chains
    data, z = simulate(2500, 160, k_true=6, miss=0.1, seed=4)
    assign = [int(v) for v in z]
    moves = _moves(sm_prob=0.4)
    seeds = [21, 22, 23, 24, 25]
    a = _chains(data, True, [0.25, 0.25], moves, 25, seeds, assign)
    run_chains(a)
    monkeypatch.setenv('BNPC_LOCKSTEP', '1')
    b = _chains(data, True, [0.25, 0.25], moves, 25, seeds, assign)
    run_chains(b)
    for x, y in zip(a, b):
        _same(x, y, f'chain {x.no}', exact=True)
