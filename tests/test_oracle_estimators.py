"""CPU: the estimator restatement (oracle/estimators_oracle.py) against the golden fixtures
generated from the reference's libs/utils.py, and against the reference itself where it is
mounted (this container)."""
import numpy as np
import pytest

from helpers import EstimatorGolden, estimator_golden_names
from oracle import estimators_oracle as eo
from oracle import ref_shim


@pytest.mark.parametrize('name', estimator_golden_names())
def test_estimator_oracle_matches_reference_fixtures(name):
    g = EstimatorGolden(name)
    z = g.z
    cat = eo.concat_chain_results(g.results)
    np.testing.assert_array_equal(eo.get_dist(cat['assignments']), z['dist'])
    np.testing.assert_array_equal(eo.get_mpear_assignment(cat['assignments']), z['mpear'])
    post = eo.latents_posterior_chain(cat, g.data)
    np.testing.assert_array_equal(post['assignment'], z['post_assignment'])
    np.testing.assert_allclose(post['genotypes'], z['post_genotypes'], rtol=1e-12, atol=0)
    got = [post['a'][0], post['a'][1], post['FN'][0], post['FN'][1], post['FP'][0], post['FP'][1],
           post['FN_geno'], post['FP_geno']]
    np.testing.assert_allclose(got, z['post_scalars'], rtol=1e-12)
    best = g.results[int(np.argmax([np.max(r['MAP'][r['burn_in']:]) for r in g.results]))]
    point = eo.latents_point_chain(best, 'MAP', g.data)
    assert point['step'] == int(z['point_step'])
    np.testing.assert_array_equal(point['assignment'], z['point_assignment'])
    np.testing.assert_array_equal(point['genotypes'], z['point_genotypes'])
    np.testing.assert_allclose([point['a'], point['FN'], point['FP'], point['FN_geno'], point['FP_geno']],
                               z['point_scalars'], rtol=1e-12)


def test_fixtures_exist():
    assert len(estimator_golden_names()) >= 3


@pytest.mark.skipif(not ref_shim.reference_available(), reason='reference checkout not mounted')
def test_estimator_oracle_matches_reference_live():
    ref = ref_shim.load_reference(with_mcmc=True)
    rng = np.random.default_rng(5)
    steps, n = 25, 40
    z = rng.integers(0, 4, n)
    a = np.stack([rng.permutation(6)[z] for _ in range(steps)])
    flip = rng.random(a.shape) < 0.1
    a[flip] = rng.integers(0, 7, flip.sum())
    np.testing.assert_array_equal(eo.get_dist(a), ref.utils.get_dist(a))
    np.testing.assert_array_equal(eo.get_mpear_assignment(a), ref.utils._get_MPEAR(a))
    sim = 1 - eo.get_dist(a)
    c = rng.integers(0, 3, n)
    np.testing.assert_allclose(eo.calc_mpear(sim, c), ref.utils._calc_MPEAR(sim, c), rtol=1e-13)
