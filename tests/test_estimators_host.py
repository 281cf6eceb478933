"""CPU: host side of libs/utils.py (the reference's estimator API) against the oracle restatement
and the fixtures generated from the reference; the two CUDA kernels are replaced by numpy here."""
import numpy as np
import pytest

from helpers import EstimatorGolden, estimator_golden_names
from oracle import estimators_oracle as eo
from oracle import ref_shim
import libs.utils as ut


def _numpy_sums(counts, labels):
    n = labels.shape[1]
    iu = np.triu_indices(n, k=1)
    same = labels[:, iu[0]] == labels[:, iu[1]]
    return int(counts.sum()), same.sum(axis=1).astype(np.int64), (same * counts[None, :]).sum(axis=1).astype(np.int64)


@pytest.mark.parametrize('name', estimator_golden_names())
def test_scores_from_integer_sums_match_the_float_formula(name):
    g = EstimatorGolden(name)
    a = eo.concat_chain_results(g.results)['assignments']
    steps, cells = a.shape
    counts = eo.pair_counts(a)
    rng = np.random.default_rng(0)
    labels = np.stack([rng.integers(0, k, cells) for k in (2, 3, 5, 9)])
    total, same_pairs, same_counts = _numpy_sums(counts, labels)
    got = ut._mpear_scores(total, same_pairs, same_counts, steps, cells)
    sim = 1 - counts / steps
    want = [eo.calc_mpear(sim, c) for c in labels]
    np.testing.assert_allclose(got, want, rtol=1e-10)
    np.testing.assert_allclose([ut._calc_MPEAR(sim, c) for c in labels], want, rtol=1e-12)
    np.testing.assert_array_equal(ut._candidate_cluster_numbers(a), eo.cluster_number_range(a))


@pytest.mark.parametrize('name', estimator_golden_names())
def test_posterior_and_point_estimators_on_reference_fixtures(name, monkeypatch):
    g = EstimatorGolden(name)
    z = g.z
    # the MPEAR assignment of the fixture stands in for the device path (tested on the GPU)
    monkeypatch.setattr(ut, '_get_MPEAR', lambda a: z['mpear'].copy())
    post = ut.get_latents_posterior([dict(r) for r in g.results], g.data)[0]
    np.testing.assert_array_equal(post['assignment'], z['post_assignment'])
    np.testing.assert_allclose(post['genotypes'].T.values, z['post_genotypes'], rtol=1e-12, atol=0)
    got = [post['a'][0], post['a'][1], post['FN'][0], post['FN'][1], post['FP'][0], post['FP'][1],
           post['FN_geno'], post['FP_geno']]
    np.testing.assert_allclose(got, z['post_scalars'], rtol=1e-12)
    point = ut.get_latents_point([dict(r) for r in g.results], 'MAP', g.data)[0]
    assert point['step'] == int(z['point_step'])
    np.testing.assert_array_equal(point['assignment'], z['point_assignment'])
    np.testing.assert_array_equal(point['genotypes'].T.values, z['point_genotypes'])
    np.testing.assert_allclose([point['a'], point['FN'], point['FP'], point['FN_geno'], point['FP_geno']],
                               z['point_scalars'], rtol=1e-12)


def test_posterior_estimator_fails_loudly_without_gpu():
    torch = pytest.importorskip('torch')
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    with pytest.raises(RuntimeError, match='CUDA'):
        ut.get_dist(np.zeros((3, 5), dtype=int))


@pytest.mark.skipif(not ref_shim.reference_available(), reason='reference checkout not mounted')
def test_lugsail_psrf_matches_reference():
    ref = ref_shim.load_reference(with_mcmc=True)
    rng = np.random.default_rng(3)
    chains = [(np.cumsum(rng.normal(0, 1, 400)) * 0.05 + rng.normal(0, 1, 400), 100) for _ in range(4)]
    want = ref.utils.get_lugsail_batch_means_est(chains)
    np.testing.assert_allclose(ut.get_lugsail_batch_means_est(chains), want, rtol=1e-12)
    assert ut.get_lugsail_batch_means_est([(np.zeros(5), 0)]) == np.inf
    np.testing.assert_allclose(ut.get_cutoff_lugsail(0.1), ref.utils.get_cutoff_lugsail(0.1), rtol=1e-14)
