#!/usr/bin/env python
"""BASELINE configuration 1: the reference's own example (example_data/data.csv, 100 cells x 100
mutations, 5 simulated clusters, example_data/data_params.txt) run through the UNMODIFIED reference
with its default model and move probabilities (run_BnpC.py defaults; 1500 steps instead of 5000 to
keep the generation short), three runs.  Writes tests/golden/c1_reference_posterior.json: the
posterior (MPEAR) estimate of every run -- assignment, cluster sizes, error rates -- which
tests/test_gpu_estimators.py::test_reference_example_c1 compares the CUDA chains with.

    python tests/golden/make_c1.py        (needs /root/reference; about 2.5 minutes)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
import libs.dpmmIO as io  # noqa: E402

STEPS = 1500


def main():
    path = os.path.join(ref_shim.REF_ROOT, 'example_data', 'data.csv')
    data = io.load_data(path, transpose=True)
    ref = ref_shim.load_reference(with_mcmc=True)
    runs = []
    for seed in (1, 2, 3):
        model = ref.CRP_learning_errors.CRP_errors_learning(
            data, DP_alpha=[-1, -1], param_beta=[0.25, 0.25], FP_mean=0.01, FP_sd=0.01, FN_mean=0.2, FN_sd=0.1)
        old = np.geterr()
        np.seterr(divide='raise', invalid='raise')
        try:
            mcmc = ref.MCMC.MCMC(model, sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25],
                                 sm_steps=3)
            mcmc.run((STEPS, 0), seed, n=1, verbosity=0, debug=True)       # debug: no process pool
        finally:
            np.seterr(**old)
        res = mcmc.get_results()
        est = ref.utils.get_latents_posterior(res, data, single_chains=False)[0]
        a = np.unique(est['assignment'], return_inverse=True)[1]
        runs.append(dict(seed=seed, assignment=[int(v) for v in a],
                         sizes=sorted(np.bincount(a).tolist(), reverse=True),
                         fn=float(est['FN'][0]), fp=float(est['FP'][0])))
        print(seed, runs[-1]['sizes'], runs[-1]['fn'], runs[-1]['fp'], flush=True)
    out = dict(source='example_data/data.csv of the reference, defaults of run_BnpC.py', steps=STEPS,
               clusters_simulated=5, runs=runs)
    with open(os.path.join(ROOT, 'tests', 'golden', 'c1_reference_posterior.json'), 'w') as f:
        json.dump(out, f)


if __name__ == '__main__':
    main()
