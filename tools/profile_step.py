#!/usr/bin/env python3
"""Host-side profile of one chain at a benchmark shape (cProfile over N steps)."""
import argparse
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnpc_b200.synth import CONFIGS, make_matrix  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--config', default='C3')
ap.add_argument('--steps', type=int, default=30)
ap.add_argument('--top', type=int, default=35)
args = ap.parse_args()

import torch  # noqa: E402
import libs.CRP_learning_errors as crple  # noqa: E402
from bnpc_b200.rng import PhiloxRandom  # noqa: E402
from libs.MCMC import Chain_steps  # noqa: E402

cfg = CONFIGS[args.config]
data, z = make_matrix(cfg['cells'], cfg['muts'], cfg['k_true'], cfg['fn'], cfg['fp'], cfg['miss'], seed=0)
m = crple.CRP_errors_learning(data, DP_alpha=[-1, -1], param_beta=list(cfg['pp']), FP_mean=0.01, FP_sd=0.01,
                              FN_mean=0.2, FN_sd=0.1, rnd=PhiloxRandom(7), device='cuda:0')
m.init(assign=[int(v) for v in z])
moves = dict(sm_prob=0.33, dpa_prob=0.25, error_prob=0.25, sm_ratios=[0.75, 0.25], sm_steps=3,
             param_proposal_sd=np.array([0.1, 0.25, 0.5]))
ch = Chain_steps(m, 1, 3 * args.steps + 8, 0, moves, 0, False)
for i in range(3):
    ch.do_step()
    ch.update_results(1 + i, False)


def timed(name, fn, n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    pr.disable()
    dt = time.perf_counter() - t0
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(args.top)
    print(f'===== {name}: {1e3 * dt / n:.3f} ms/step')
    print(s.getvalue()[:9000])


def dev_step(i):
    ch.do_step()
    if i % 5 == 0:
        print('step', i, m.sweep_stats, 'K', len(m.cells_per_cluster))
    m.get_ll_full()
    m.get_lprior_full()


def host_step(i):
    ch.do_step()
    ch.update_results(4 + i, False)


timed('device-trace step', dev_step, args.steps)
timed('host-trace step', host_step, args.steps)
print('sweep stats', m.sweep_stats, 'K', len(m.cells_per_cluster), 'FN', m.FN, 'FP', m.FP, 'alpha', m.DP_a)
print('sizes', sorted(m.cells_per_cluster.values()))
