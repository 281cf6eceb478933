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


def _trace_with_shared_profiles(rng, n, steps, k, movers=0.15):
    """assignment samples in which most cells never leave their cluster (they share a profile)"""
    z = rng.integers(0, k, n)
    moving = rng.random(n) < movers
    out = np.zeros((steps, n), dtype=int)
    for s in range(steps):
        labels = rng.permutation(k + 2)[:k]                  # arbitrary ids, as cluster ids are
        a = labels[z]
        scat = moving & (rng.random(n) < 0.3)
        a[scat] = rng.choice(labels, scat.sum())
        out[s] = a
    return out


def test_weighted_ward_over_profiles_equals_scipy_over_cells():
    """The route of large matrices (libs/utils.py::_get_MPEAR): ward linkage of the DISTINCT
    assignment profiles, started from clusters of their multiplicities, cuts the cells exactly as
    scipy's linkage over all cells (the reference, libs/utils.py:100-116) -- same heights, same
    labels for every candidate number of clusters."""
    from scipy.cluster.hierarchy import cut_tree, linkage
    from scipy.spatial.distance import pdist, squareform
    import libs.utils as ut
    rng = np.random.default_rng(3)
    for trial in range(8):
        n, steps, k = int(rng.integers(60, 500)), int(rng.integers(5, 40)), int(rng.integers(2, 7))
        a = _trace_with_shared_profiles(rng, n, steps, k)
        d = sum(pdist(a[s][:, None], 'hamming') for s in range(steps)) / steps
        Z = linkage(d, 'ward')
        n_range = ut._candidate_cluster_numbers(a)
        want = cut_tree(Z, n_clusters=n_range)
        rep, inverse, weight = ut._unique_profiles(a)
        assert weight.sum() == n and (a[:, rep][:, inverse] == a).all() and rep.size < n
        du = sum(pdist(a[s][rep][:, None], 'hamming') for s in range(steps)) / steps
        Zw = ut._ward_linkage_weighted(squareform(du), weight)
        np.testing.assert_allclose(np.sort(Zw[:, 2]), Z[-(rep.size - 1):, 2], rtol=0, atol=1e-12)
        n_ok = n_range[n_range < rep.size]          # as _get_MPEAR (cut_tree mislabels n_clusters == points)
        got = cut_tree(Zw, n_clusters=n_ok)
        checked = 0
        for j, c in enumerate(n_ok):
            applied = rep.size - int(c)                      # merges below the cut
            if 0 < applied < rep.size - 1 and Zw[applied - 1, 2] >= Zw[applied, 2] - 1e-12:
                continue                                     # a tie at the cut: either order is a valid dendrogram
            np.testing.assert_array_equal(ut._canonical_labels(got[inverse][:, j]), want[:, j])
            checked += 1
        assert checked >= 1


def test_unique_profiles_handles_all_distinct_and_all_equal():
    import libs.utils as ut
    a = np.arange(12).reshape(3, 4)
    rep, inverse, weight = ut._unique_profiles(a)
    assert rep.tolist() == [0, 1, 2, 3] and inverse.tolist() == [0, 1, 2, 3] and weight.tolist() == [1, 1, 1, 1]
    b = np.zeros((5, 7), dtype=int)
    rep, inverse, weight = ut._unique_profiles(b)
    assert rep.tolist() == [6] and set(inverse.tolist()) == {0} and weight.tolist() == [7]
