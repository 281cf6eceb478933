#!/usr/bin/env python3
"""Wall time (stream-synchronised) of every model method of one chain at a benchmark shape."""
import argparse
import os
import sys
import time
from collections import defaultdict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnpc_b200.synth import CONFIGS, make_matrix  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C3')
ap.add_argument('--steps', type=int, default=40)
args = ap.parse_args()

import torch  # noqa: E402
import libs.CRP as crp  # noqa: E402
import libs.CRP_learning_errors as crple  # noqa: E402
from bnpc_b200 import _lib  # noqa: E402
from bnpc_b200.rng import PhiloxRandom  # noqa: E402
from libs.MCMC import Chain_steps  # noqa: E402

cfg = CONFIGS[args.config]
data, z = make_matrix(cfg['cells'], cfg['muts'], cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
if cfg['learning']:
    m = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=list(cfg['pp']), FP_mean=0.01, FP_sd=0.01,
                                  FN_mean=0.2, FN_sd=0.1, rnd=PhiloxRandom(7), device='cuda:0')
else:
    m = crp.CRP(data, DP_alpha=[-1, -1], param_beta=list(cfg['pp']), FN_error=cfg['FN'], FP_error=cfg['FP'],
                rnd=PhiloxRandom(7), device='cuda:0')
m.init(assign=[int(v) for v in z])
moves = dict(sm_prob=cfg.get('sm_prob', 0.33), dpa_prob=0.25, error_prob=0.25 if cfg['learning'] else 0.0,
             sm_ratios=[0.75, 0.25], sm_steps=3, param_proposal_sd=np.array([0.1, 0.25, 0.5]))
ch = Chain_steps(m, 1, 3 * args.steps + 8, 0, moves, 0, False)

acc = defaultdict(list)
launches = defaultdict(list)


def wrap(obj, name):
    fn = getattr(obj, name)

    def timed(*a, **k):
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize()
        acc[name].append(time.perf_counter() - t0)
        launches[name].append(_lib.launch_count() - l0)
        return r
    setattr(obj, name, timed)


for name in ('update_assignments_Gibbs', 'update_assignments_split_merge', 'update_DP_alpha',
             'update_parameters', 'update_error_rates', 'get_ll_full', 'get_lprior_full'):
    if hasattr(m, name):
        wrap(m, name)

for i in range(3):
    ch.do_step()
    ch.update_results(1 + i, False)
acc.clear()
launches.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(args.steps):
    ch.do_step()
    ch.update_results(4 + i, False)
torch.cuda.synchronize()
total = time.perf_counter() - t0
print(f'{args.config}: {1e3 * total / args.steps:.3f} ms/step over {args.steps} steps (host traces), '
      f'K={len(m.cells_per_cluster)}')
for name, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print(f'  {name:34s} calls={len(v):4d} mean={1e3 * np.mean(v):8.3f} ms  max={1e3 * np.max(v):8.3f} ms '
          f'share={sum(v) / total:6.1%} launches/call={np.mean(launches[name]):.1f}')
print('  sweep stats', m.sweep_stats)
print('  sizes', sorted(m.cells_per_cluster.values()))
