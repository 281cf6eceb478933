"""debug: where does a chain of a large group diverge from the same chain alone?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_group import _chains, _moves  # noqa: E402
from libs.MCMC import run_chains  # noqa: E402
from oracle.crp_oracle import simulate  # noqa: E402

n_group = int(sys.argv[1]) if len(sys.argv) > 1 else 10
attrs = dict(serial_sweep=True) if os.environ.get('DBG_SERIAL') else (dict(lean_enabled=False) if os.environ.get('DBG_DENSE') else None)
reps = int(os.environ.get('DBG_REPS', 2))
data, z = simulate(4000, 256, k_true=8, miss=0.1, seed=5)
assign = [int(v) for v in z]
moves = _moves(sm_prob=0.4)
seeds = list(range(7, 7 + n_group))
runs = []
for rep in range(reps):
    together = _chains(data, True, [0.25, 0.25], moves, 30, seeds, assign, 0, attrs)
    run_chains(together)
    runs.append(together)
n_bad = 0
for rep in range(1, reps):
    for i in range(n_group):
        a, b = runs[0][i].results, runs[rep][i].results
        same = np.array_equal(a['assignments'], b['assignments']) and np.array_equal(a['ML'], b['ML'])
        if not same:
            n_bad += 1
            bad = [s for s in range(31) if not np.array_equal(a['assignments'][s], b['assignments'][s])]
            print('NONDETERMINISTIC rep', rep, 'chain', i, 'first step', bad[:1])
print('nondeterministic chain-runs:', n_bad, 'of', (reps - 1) * n_group)
if os.environ.get('DBG_SKIP_ALONE'):
    sys.exit(0)
for i, seed in enumerate(seeds[:n_group]):
    alone = _chains(data, True, [0.25, 0.25], moves, 30, [seed], assign, 0, attrs)
    run_chains(alone)
    a, b = runs[0][i].results, alone[0].results
    bad = [s for s in range(31) if not np.array_equal(a['assignments'][s], b['assignments'][s])]
    badml = [s for s in range(31) if a['ML'][s] != b['ML'][s]]
    badp = [s for s in range(31) if not np.array_equal(a['params'][s], b['params'][s])] if a['params'].shape == b['params'].shape else 'shape'
    print(f'chain {i} seed {seed}: first bad assignment step {bad[:3]}, ML {badml[:3]}, params {badp[:3] if badp != "shape" else badp}',
          'n diff cells', (a['assignments'][bad[0]] != b['assignments'][bad[0]]).sum() if bad else 0)
