#!/usr/bin/env python3
"""Generate tests/golden/ari_reference.json FROM THE UNMODIFIED REFERENCE (build container only):

    python tests/golden/make_golden_ari.py

North-star gate "on seeded simulated data the posterior/MAP estimators must recover clusters with
ARI matching the reference within noise": the reference's own chains (libs/MCMC.py Chain_steps over
libs/CRP_learning_errors.py) are run from several numpy seeds on seeded simulated matrices, its own
estimators (libs/utils.py get_latents_posterior / get_latents_point) are applied to the traces and
the adjusted Rand index against the simulated partition is recorded per seed.  The GPU tests
(tests/test_gpu_estimators.py) run the CUDA chains on the SAME matrices with their own seeds and
demand that their ARI range overlaps the reference's.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from oracle.crp_oracle import simulate  # noqa: E402

SCENARIOS = [
    # the start is the simulated partition with a quarter of the cells scattered (the gate is about
    # the sampler + estimators in steady state, not about the collapse from a random start)
    dict(name='learn_2000x200', n=2000, m=200, k_true=8, miss=0.10, sim_seed=5, steps=160, burn_in=80,
         pp=[0.25, 0.25], seeds=[1, 2, 3]),
    dict(name='learn_1200x120_pp11', n=1200, m=120, k_true=5, miss=0.20, sim_seed=6, steps=120, burn_in=60,
         pp=[1, 1], seeds=[1, 2, 3]),
]
MOVES = dict(sm_prob=0.33, dpa_prob=0.5, error_prob=0.1, sm_ratios=[0.75, 0.25], sm_steps=3,
             param_proposal_sd=np.array([0.1, 0.25, 0.5]))
LEARN = dict(DP_alpha=[-1, -1], FP_mean=0.01, FP_sd=0.01, FN_mean=0.2, FN_sd=0.1)


def scattered_start(z, k_true, seed=1):
    rng = np.random.default_rng(seed)
    start = z.copy()
    scat = rng.random(z.size) < 0.25
    start[scat] = rng.integers(0, k_true + 3, scat.sum())
    return [int(v) for v in start]


def main():
    ref = ref_shim.load_reference(with_mcmc=True)
    out = {}
    for sc in SCENARIOS:
        data, z = simulate(sc['n'], sc['m'], k_true=sc['k_true'], miss=sc['miss'], seed=sc['sim_seed'])
        start = scattered_start(z, sc['k_true'])
        rows = []
        for seed in sc['seeds']:
            t0 = time.time()
            np.random.seed(seed)
            with ref_shim.ref_errstate():
                model = ref.CRP_learning_errors.CRP_errors_learning(data.copy(), param_beta=sc['pp'], **LEARN)
                model.init(assign=start)
                chain = ref.MCMC.Chain_steps(model, 1, sc['steps'], sc['burn_in'], dict(MOVES), 0, False)
                chain.run()
                res = chain.get_result()
                post = ref.utils.get_latents_posterior([res], data)[0]
                point = ref.utils.get_latents_point([res], 'MAP', data)[0]
            rows.append(dict(seed=seed, ari_posterior=float(ref.utils.get_ARI(post['assignment'], z)),
                             ari_map=float(ref.utils.get_ARI(point['assignment'], z)),
                             k_last=len(model.cells_per_cluster), fn=float(post['FN'][0]), fp=float(post['FP'][0])))
            print(sc['name'], rows[-1], f'{time.time() - t0:.0f}s', flush=True)
        out[sc['name']] = dict(scenario={k: v for k, v in sc.items()}, runs=rows)
    with open(os.path.join(HERE, 'ari_reference.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote ari_reference.json')


if __name__ == '__main__':
    main()
