"""repro of the sweep nondeterminism: one chain (seed from argv) for a few steps, twice; prints whether
the two runs agree.  Used under compute-sanitizer racecheck (profiles/r2_racecheck_*.log)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_gpu_group import _chains, _moves  # noqa: E402
from libs.MCMC import run_chains  # noqa: E402
from oracle.crp_oracle import simulate  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 15
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 11
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
data, z = simulate(4000, 256, k_true=8, miss=0.1, seed=5)
assign = [int(v) for v in z]
moves = _moves(sm_prob=0.4)
runs = []
for rep in range(reps):
    ch = _chains(data, True, [0.25, 0.25], moves, steps, [seed], assign)
    run_chains(ch)
    runs.append(ch[0].results['assignments'].copy())
    print('rep', rep, 'K', len(ch[0].model.cells_per_cluster), ch[0].model.sweep_stats, flush=True)
for rep in range(1, reps):
    bad = [s for s in range(steps + 1) if not np.array_equal(runs[0][s], runs[rep][s])]
    print('rep', rep, 'first differing step', bad[:1])
