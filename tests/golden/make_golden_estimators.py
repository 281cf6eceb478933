#!/usr/bin/env python3
"""Generate tests/golden/estimators_*.npz FROM THE REFERENCE's libs/utils.py (run in the build
container, where the unmodified reference is mounted at /root/reference):

    python tests/golden/make_golden_estimators.py

Each fixture holds synthetic chain results (assignment / parameter / scalar traces of several
chains, in the layout of libs/MCMC.py:231-282) and what the reference's estimators make of them:
the co-clustering distance (libs/utils.py:90-97), the MPEAR assignment (:100-130), the posterior
estimator's genotypes and error rates (:148-241) and the MAP point estimator (:248-283).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402

SCENARIOS = [
    dict(name='estimators_two_chains', n=60, m=12, k=4, steps=30, burn_in=10, chains=2, noise=0.08, seed=11),
    dict(name='estimators_noisy', n=45, m=8, k=3, steps=40, burn_in=15, chains=1, noise=0.25, seed=12),
    dict(name='estimators_stable', n=70, m=10, k=5, steps=24, burn_in=8, chains=3, noise=0.0, seed=13),
]


def fake_chain(rng, n, m, k, steps, burn_in, noise, data):
    """A chain-like trace: cells follow a true partition, relabelled at random per sample with a
    fraction `noise` of cells scattered; params rows follow the SORTED live ids (MCMC.py:261-282)."""
    z = rng.integers(0, k, n)
    geno = rng.random((k, m))
    assignments = np.zeros((steps, n), dtype=int)
    kmax = 0
    rows = []
    for s in range(steps):
        labels = rng.permutation(k + 3)[:k]                 # arbitrary ids, as cluster ids are
        a = labels[z]
        scat = rng.random(n) < noise
        a[scat] = rng.choice(np.append(labels, k + 5), scat.sum())
        assignments[s] = a
        ids = np.unique(a)
        par = np.zeros((ids.size, m), dtype=np.float32)
        for r, cid in enumerate(ids):
            src = np.flatnonzero(labels == cid)
            par[r] = geno[src[0]] if src.size else rng.random(m)
            par[r] = np.clip(par[r] + rng.normal(0, 0.02, m), 1e-5, 1 - 1e-5)
        rows.append(par)
        kmax = max(kmax, ids.size)
    params = np.zeros((steps - burn_in, kmax, m), dtype=np.float32)
    for s in range(burn_in, steps):
        params[s - burn_in, :rows[s].shape[0]] = rows[s]
    return dict(assignments=assignments, params=params, burn_in=burn_in,
                ML=rng.normal(-100, 5, steps), MAP=rng.normal(-120, 5, steps),
                DP_alpha=rng.gamma(3, 1, steps), FN=rng.beta(2, 8, steps), FP=rng.beta(1, 50, steps))


def main():
    ref = ref_shim.load_reference(with_mcmc=True)
    ut = ref.utils
    warnings.simplefilter('ignore')
    for sc in SCENARIOS:
        rng = np.random.default_rng(sc['seed'])
        data = rng.integers(0, 2, (sc['n'], sc['m'])).astype(np.float64)
        data[rng.random(data.shape) < 0.1] = np.nan
        results = [fake_chain(rng, sc['n'], sc['m'], sc['k'], sc['steps'], sc['burn_in'], sc['noise'], data)
                   for _ in range(sc['chains'])]
        with np.errstate(all='ignore'):
            post = ut.get_latents_posterior([dict(r) for r in results], data, single_chains=False)[0]
            point = ut.get_latents_point([dict(r) for r in results], 'MAP', data, single_chains=False)[0]
            cat = ut._concat_chain_results([dict(r) for r in results])
            dist = ut.get_dist(cat['assignments'])
            mpear = ut._get_MPEAR(cat['assignments'])
        out = dict(data=np.where(np.isnan(data), -1, data).astype(np.int8), n_chains=sc['chains'],
                   dist=dist, mpear=np.asarray(mpear), post_assignment=np.asarray(post['assignment']),
                   post_genotypes=post['genotypes'].T.values.astype(np.float64),
                   post_scalars=np.array([post['a'][0], post['a'][1], post['FN'][0], post['FN'][1],
                                          post['FP'][0], post['FP'][1], post['FN_geno'], post['FP_geno']]),
                   point_step=point['step'], point_assignment=np.asarray(point['assignment']),
                   point_genotypes=point['genotypes'].T.values.astype(np.float64),
                   point_scalars=np.array([point['a'], point['FN'], point['FP'], point['FN_geno'],
                                           point['FP_geno']]))
        for c, r in enumerate(results):
            for key, v in r.items():
                out[f'chain{c}_{key}'] = np.asarray(v)
        np.savez_compressed(os.path.join(HERE, sc['name'] + '.npz'), **out)
        print(sc['name'], 'clusters', np.unique(mpear).size, 'dist pairs', dist.size)


if __name__ == '__main__':
    main()
