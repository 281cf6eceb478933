"""GPU: the posterior (MPEAR) estimator through libs.utils and the C ABI (bnpc_cocluster_counts,
bnpc_mpear_sums) against the fixtures generated from the reference and against the oracle."""
import numpy as np
import pytest

from helpers import EstimatorGolden, estimator_golden_names
from oracle import estimators_oracle as eo

pytestmark = pytest.mark.gpu
torch = pytest.importorskip('torch')


@pytest.mark.parametrize('name', estimator_golden_names())
def test_posterior_estimator_matches_reference_fixtures(name):
    import libs.utils as ut
    g = EstimatorGolden(name)
    z = g.z
    cat = eo.concat_chain_results(g.results)
    np.testing.assert_array_equal(ut.get_dist(cat['assignments']), z['dist'])       # bit-identical
    np.testing.assert_array_equal(ut._get_MPEAR(cat['assignments']), z['mpear'])
    post = ut.get_latents_posterior([dict(r) for r in g.results], g.data)[0]
    np.testing.assert_array_equal(post['assignment'], z['post_assignment'])
    np.testing.assert_allclose(post['genotypes'].T.values, z['post_genotypes'], rtol=1e-12, atol=0)


@pytest.mark.parametrize('shape', [(7, 2), (40, 65), (33, 200), (64, 1111)])
def test_pair_kernels_are_exact(shape):
    """counts and the three pair sums are integers: bit-exact against numpy, ragged tile edges
    included (N not a multiple of 64, more candidates than one staging round)."""
    import libs.utils as ut
    S, N = shape
    rng = np.random.default_rng(S * N)
    a = rng.integers(0, 6, (S, N))
    pc = ut._PairCounts(a)
    counts = pc.counts.cpu().numpy()
    np.testing.assert_array_equal(counts, eo.pair_counts(a))
    labels = np.stack([rng.integers(0, k, N) for k in list(range(1, 8)) + [11] * 30])
    total, same_pairs, same_counts = pc.sums(labels)
    iu = np.triu_indices(N, k=1)
    same = labels[:, iu[0]] == labels[:, iu[1]]
    assert total == int(counts.sum(dtype=np.int64))
    np.testing.assert_array_equal(same_pairs, same.sum(axis=1))
    np.testing.assert_array_equal(same_counts, (same * counts[None, :].astype(np.int64)).sum(axis=1))


def _scattered_start(z, k_true, seed=1):
    # tests/golden/make_golden_ari.py: the simulated partition with a quarter of the cells scattered
    rng = np.random.default_rng(seed)
    start = z.copy()
    scat = rng.random(z.size) < 0.25
    start[scat] = rng.integers(0, k_true + 3, scat.sum())
    return [int(v) for v in start]


@pytest.mark.parametrize('name', ['learn_2000x200', 'learn_1200x120_pp11'])
def test_estimator_ari_matches_the_reference_within_noise(name):
    """North-star gate: on seeded simulated data the posterior / MAP estimators recover the
    clusters with an ARI matching the reference's within noise.  tests/golden/ari_reference.json
    holds what the UNMODIFIED reference (its chains + its estimators) reaches on the same matrices
    from three seeds; the CUDA chains run from three seeds of their own.  Chains are stochastic and
    the streams differ, so the comparison is between the two samples of ARIs: the CUDA mean may not
    fall below the reference's mean by more than two of its standard deviations (+0.02), and the
    learned error rates must land in the reference's range."""
    import json
    import os
    import libs.utils as ut
    from libs.MCMC import Chain_steps, run_chains
    import libs.CRP_learning_errors as crple
    from bnpc_b200.rng import PhiloxRandom
    from helpers import GOLDEN_DIR
    from oracle.crp_oracle import simulate
    with open(os.path.join(GOLDEN_DIR, 'ari_reference.json')) as f:
        ref = json.load(f)[name]
    sc = ref['scenario']
    data, z = simulate(sc['n'], sc['m'], k_true=sc['k_true'], miss=sc['miss'], seed=sc['sim_seed'])
    start = _scattered_start(z, sc['k_true'])
    moves = dict(sm_prob=0.33, dpa_prob=0.5, error_prob=0.1, sm_ratios=[0.75, 0.25], sm_steps=3,
                 param_proposal_sd=np.array([0.1, 0.25, 0.5]))
    chains = []
    for seed in (11, 12, 13):
        m = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=sc['pp'], FP_mean=0.01, FP_sd=0.01,
                                      FN_mean=0.2, FN_sd=0.1, rnd=PhiloxRandom(seed), device='cuda:0')
        m.init(assign=start)
        chains.append(Chain_steps(m, len(chains) + 1, sc['steps'], sc['burn_in'], moves, 0, False))
    run_chains(chains)
    got = dict(ari_posterior=[], ari_map=[], fn=[])
    for ch in chains:
        res = ch.get_result()
        post = ut.get_latents_posterior([res], data)[0]
        point = ut.get_latents_point([res], 'MAP', data)[0]
        got['ari_posterior'].append(ut.get_ARI(post['assignment'], z))
        got['ari_map'].append(ut.get_ARI(point['assignment'], z))
        got['fn'].append(float(post['FN'][0]))
        assert post['genotypes'].shape == (sc['m'], sc['n'])
    for key in ('ari_posterior', 'ari_map'):
        want = np.array([r[key] for r in ref['runs']])
        assert np.mean(got[key]) >= want.mean() - 2 * want.std() - 0.02, (key, got[key], want.tolist())
        assert max(got[key]) >= want.min() and min(got[key]) <= want.max() + 1e-12, (key, got[key], want.tolist())
    fn_ref = [r['fn'] for r in ref['runs']]
    assert min(fn_ref) - 0.05 < np.mean(got['fn']) < max(fn_ref) + 0.05


def test_reference_example_c1():
    """BASELINE configuration 1: the reference's own example (example_data/data.csv: 100 cells x 100
    mutations, 5 simulated clusters -- example_data/data_params.txt) with the defaults of
    run_BnpC.py.  The unmodified reference, run three times for 1500 steps
    (tests/golden/make_c1.py -> c1_reference_posterior.json), ends all three runs with the same
    posterior estimate: 5 clusters of 23 / 22 / 21 / 20 / 14 cells, FN ~ 0.06.  Two CUDA chains
    through the public MCMC driver must find that partition and error rates in the same range.
    The data file is read from the reference checkout (the unmodified copy under baseline/_ref on
    the GPU box)."""
    import json
    import os
    import libs.CRP_learning_errors as crple
    import libs.dpmmIO as io
    import libs.utils as ut
    from libs.MCMC import MCMC
    from helpers import GOLDEN_DIR
    from oracle import ref_shim
    path = os.path.join(ref_shim.REF_ROOT, 'example_data', 'data.csv')
    if not os.path.isfile(path):
        pytest.skip('the reference example data is not available (oracle/fetch_ref.py ships it under baseline/_ref)')
    with open(os.path.join(GOLDEN_DIR, 'c1_reference_posterior.json')) as f:
        ref = json.load(f)
    data = io.load_data(path, transpose=True)
    assert data.shape == (100, 100)
    model = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=[0.25, 0.25], FP_mean=0.01, FP_sd=0.01,
                                      FN_mean=0.2, FN_sd=0.1)
    mcmc = MCMC(model, sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3)
    mcmc.run((ref['steps'], 0), 4, n=2, verbosity=0)
    results = mcmc.get_results()
    est = ut.get_latents_posterior(results, data, single_chains=False)[0]
    a = np.unique(est['assignment'], return_inverse=True)[1]
    assert sorted(np.bincount(a).tolist(), reverse=True) == ref['runs'][0]['sizes'] == [23, 22, 21, 20, 14]
    assert len(set(a)) == ref['clusters_simulated']
    for run in ref['runs']:
        assert ut.get_ARI(a, np.array(run['assignment'])) == 1.0
    fn_ref = [r['fn'] for r in ref['runs']]
    fp_ref = [r['fp'] for r in ref['runs']]
    assert min(fn_ref) - 0.03 < float(est['FN'][0]) < max(fn_ref) + 0.03
    assert float(est['FP'][0]) < max(fp_ref) + 0.005


def test_command_line_end_to_end(tmp_path):
    """run_BnpC.py on a generated matrix file: 2 chains, posterior + MAP estimators, text outputs
    in the reference's formats, ARI against the simulated clusters."""
    import pandas as pd
    import run_BnpC
    from oracle.crp_oracle import simulate
    data, z = simulate(400, 60, k_true=4, miss=0.1, seed=21)
    mat = np.where(np.isnan(data), 3, data).astype(int).T                   # file: mutations x cells
    path = tmp_path / 'data.csv'
    pd.DataFrame(mat, index=[f'mut{i}' for i in range(mat.shape[0])],
                 columns=[f'cell{j}' for j in range(mat.shape[1])]).to_csv(path, sep='\t')
    truth = tmp_path / 'truth.txt'
    truth.write_text(' '.join(str(int(v)) for v in z))
    out = tmp_path / 'out'
    args = run_BnpC.parse_args([str(path), '-n', '2', '-s', '150', '-e', 'posterior', 'MAP', '-o', str(out),
                                '--seed', '7', '-v', '0', '-np', '-tc', str(truth)])
    run_BnpC.main(args)
    files = sorted(p.name for p in out.iterdir())
    for want in ('ARI.txt', 'V_measure.txt', 'args.txt', 'assignment.txt', 'errors.txt',
                 'genotypes_MAP_mean.tsv', 'genotypes_posterior_mean.tsv', 'genotypes_cont_posterior_mean.tsv'):
        assert want in files, (want, files)
    ari = pd.read_csv(out / 'ARI.txt', sep='\t')
    # posterior over both chains recovers the clusters; the MAP estimate is ONE sample of a short
    # chain from a random start and may still have merged two clusters (the reference's does too)
    by_est = dict(zip(ari['estimator'], ari['ARI']))
    assert by_est['posterior'] > 0.9 and by_est['MAP'] > 0.5, ari
    geno = pd.read_csv(out / 'genotypes_posterior_mean.tsv', sep='\t', index_col=0)
    assert geno.shape == (60, 400) and list(geno.index[:2]) == ['mut0', 'mut1']


def test_lugsail_run_mode():
    """`-ls`: chains are extended by 200 steps until the lugsail PSRF of their ML traces falls
    below the cutoff (libs/MCMC.py:138-171 of the reference); burn-in = half of the steps run."""
    import libs.CRP_learning_errors as crple
    from libs.MCMC import MCMC
    from oracle.crp_oracle import simulate
    data, z = simulate(300, 50, k_true=3, miss=0.1, seed=8)
    model = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=[0.25, 0.25], FP_mean=0.01, FP_sd=0.01,
                                      FN_mean=0.2, FN_sd=0.1)
    mcmc = MCMC(model, sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3)
    mcmc.run((1.2, 0), 5, n=2, verbosity=0)
    results = mcmc.get_results()
    assert len(results) == 2
    for r in results:
        steps = r['ML'].size
        assert steps >= 10 and r['PSRF'][-1][1] <= 1.2 and r['PSRF_cutoff'] == 1.2
        assert r['burn_in'] == steps // 2 + 1 or r['burn_in'] == r['PSRF'][-1][0] // 2 + 1
        assert r['assignments'].shape == (steps, 300)
    # the recorded PSRF values are the reference's estimator (pinned against the reference in
    # tests/test_estimators_host.py) over the chains' ML traces at that point of the run
    import libs.utils as ut
    for steps_run, psrf in results[0]['PSRF']:
        want = ut.get_lugsail_batch_means_est([(r['ML'][:steps_run], steps_run // 2) for r in results])
        np.testing.assert_allclose(psrf, want, rtol=1e-12)
    assert [p for p in results[0]['PSRF']] == [p for p in results[1]['PSRF']]


def test_run_time_mode():
    """`-r`: Chain_time (libs/MCMC.py:395-440 of the reference) steps until the wall clock passes
    end_time; theta rows are kept once burn_in has passed; the traces are trimmed to the steps run."""
    from datetime import datetime, timedelta
    import libs.CRP_learning_errors as crple
    from libs.MCMC import MCMC
    from oracle.crp_oracle import simulate
    data, z = simulate(500, 64, k_true=4, miss=0.1, seed=9)
    model = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=[0.25, 0.25], FP_mean=0.01, FP_sd=0.01,
                                      FN_mean=0.2, FN_sd=0.1)
    mcmc = MCMC(model, sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3)
    t0 = datetime.now()
    mcmc.run((t0 + timedelta(seconds=3.0), t0 + timedelta(seconds=1.0)), 5, n=2, verbosity=0)
    assert 2.5 < (datetime.now() - t0).total_seconds() < 30
    results = mcmc.get_results()
    assert len(results) == 2
    for r in results:
        steps = r['ML'].size
        assert steps > 600                                   # more than the initial 500 rows: the traces grew
        assert r['assignments'].shape == (steps, 500) and r['FN'].size == steps
        assert 0 < r['burn_in'] < steps and r['params'].shape[0] == steps - r['burn_in']
        assert np.isfinite(r['ML']).all() and (r['ML'] < 0).all()
        k_last = np.unique(r['assignments'][-1]).size
        assert r['params'][-1, :k_last].min() > 0 and not r['params'][-1, k_last:].any()


def test_mpear_over_profiles_equals_mpear_over_cells(monkeypatch):
    """libs/utils.py::_get_MPEAR takes the profile route when cells share assignment profiles; on a
    matrix small enough for both, it must return the assignment of the all-cells route (scipy
    linkage + bnpc_mpear_sums over all pairs, the reference's arithmetic)."""
    import libs.utils as ut
    from test_estimators_host import _trace_with_shared_profiles
    rng = np.random.default_rng(17)
    for n, steps, k in ((900, 30, 5), (2500, 24, 8)):
        a = _trace_with_shared_profiles(rng, n, steps, k)
        got = ut._get_MPEAR(a)
        # the all-cells route: as if no two cells shared a profile
        monkeypatch.setattr(ut, '_unique_profiles', lambda x: (np.arange(x.shape[1]), np.arange(x.shape[1]),
                                                               np.ones(x.shape[1], dtype=np.int64)))
        want = ut._get_MPEAR(a)
        monkeypatch.undo()
        np.testing.assert_array_equal(got, want)


def test_posterior_estimator_at_100k_cells():
    """BASELINE config 3 asks for the posterior estimator at 100k cells: the reference's pair vector
    would hold 5e9 entries per sample.  Simulated traces of 100k cells x 120 samples (20 clusters,
    8 % of the cells wander): the estimator returns the simulated partition."""
    import libs.utils as ut
    rng = np.random.default_rng(5)
    n, steps, k, m = 100_000, 120, 20, 64
    z = rng.integers(0, k, n)
    moving = rng.random(n) < 0.08
    a = np.zeros((steps, n), dtype=np.int32)
    for s in range(steps):
        labels = rng.permutation(k + 3)[:k]
        row = labels[z]
        scat = moving & (rng.random(n) < 0.1)
        row[scat] = rng.choice(labels, scat.sum())
        a[s] = row
    params = rng.random((steps, k, m)).astype(np.float32)
    assign, geno = ut.get_mean_hierarchy_assignment(a, params)
    assert ut.get_ARI(assign, z) > 0.999
    assert geno.shape == (m, n)
